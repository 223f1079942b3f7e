// geometry.cu — the point-geometry kernels of the grasp-candidate path (sm_100a).
//
//   k_frames  : FrameEstimator::calculateLocalFrames  (frame_estimator.cpp:6-86, local_frame.cpp:14-41)
//   k_hands   : HandSearch::evalHands / HandSet::evalHands / FingerHand / Antipodal / Hand::construct
//               (hand_search.cpp:144-188, hand_set.cpp:31-116,235-261, finger_hand.cpp, antipodal.cpp:10-96,
//               hand.cpp:24-45) + GraspDetector::filterGraspsWorkspace/Direction (grasp_detector.cpp:334-456)
//   k_images  : ImageGenerator::createImages + Image{1,3,12,15}ChannelsStrategy (image_generator.cpp:17-99,
//               image_strategy.cpp:32-233, image_*_channels_strategy.cpp) + HandSet::calculateShadow in the
//               deterministic variant of include/gpd_b200_shadow.h
//
// Design (not a translation of the reference's per-sample OpenMP loops over Eigen temporaries):
//   * the two KdTreeFLANN builds are replaced by ONE uniform grid whose cells are ordered x-fastest, so
//     a row of cells is one contiguous segment of the cell-sorted point array: a radius search is a
//     handful of coalesced segment reads with FLANN's exact float32 predicate; no sorted result list
//     is ever materialised because every consumer is reformulated as an order-free reduction
//     (OR / min / max / arg-max / integer sums), SURVEY.md 9.6;
//   * one CTA owns one sample, its neighbourhood is staged once in shared memory and ONE WARP OWNS ONE
//     HAND POSE: all finger / deepen / closing-region / antipodal tests are warp-shuffle reductions;
//   * one CTA owns one grasp image: rasterisation into shared-memory tiles with 64-bit integer
//     atomics (arg-max key, fixed-point sum+count) => bit-reproducible, then dilate / min-max /
//     quantise in-tile.
// All float64 geometry is evaluated in the oracle's operation order; this file is compiled with
// -fmad=false so that no multiply-add is contracted (strict '<' on doubles).
#include <cfloat>
#include <cstdlib>
#include <cub/cub.cuh>

#include "common.cuh"
#include "grid.cuh"

namespace {

constexpr int NT_HANDS = 256;
#ifndef GPDB_NT_IMG
#define GPDB_NT_IMG 512
#endif
constexpr int NT_IMG = GPDB_NT_IMG;  // threads of k_images (one CTA per SM: 214 KB of shared memory)
constexpr int LRF_WARPS = 4;
constexpr int LRF_CAP_GLOBAL = 16384;  // last tier of k_frames: lists in global memory
constexpr int LRF_CAP = 1024;  // points of the r = nn_radius ball (dynamic shared memory: 2 x 8 B x LRF_CAP per warp)
constexpr int BOX_CAP = 2048;  // points inside one image box
constexpr int MAXPIX = 64 * 64;  // image_size <= 64

// o = frame^T v, frame column-major, summed left to right (PointList::transformToHandFrame, point_list.cpp:22-33)
__device__ __forceinline__ void to_frame(const double *F, double v0, double v1, double v2, double &o0, double &o1,
                                         double &o2) {
  o0 = (F[0] * v0 + F[1] * v1) + F[2] * v2;
  o1 = (F[3] * v0 + F[4] * v1) + F[5] * v2;
  o2 = (F[6] * v0 + F[7] * v1) + F[8] * v2;
}
__device__ __forceinline__ void mat3_mul(const double *A, const double *B, double *Cm) {
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) Cm[c * 3 + r] = (A[r] * B[c * 3] + A[3 + r] * B[c * 3 + 1]) + A[6 + r] * B[c * 3 + 2];
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

template <int NT>
struct SegScan {
  typedef cub::BlockScan<int, NT> Scan;
  typename Scan::TempStorage tmp;
  int start[NT];
  int prefix[NT + 1];
};
// fills seg.start/prefix for rows [row0, row0+NT) ; returns batch total. Contains __syncthreads.
template <int NT>
__device__ __forceinline__ int seg_batch(const DevParams &P, const int *cell_start, const SegRange &s, int row0,
                                         SegScan<NT> &seg) {
  int row = row0 + threadIdx.x, st = 0, len = 0;
  if (row < s.nrows) seg_row(P, cell_start, s, row, st, len);
  int excl, total;
  SegScan<NT>::Scan(seg.tmp).ExclusiveSum(len, excl, total);
  seg.start[threadIdx.x] = st;
  seg.prefix[threadIdx.x] = excl;
  if (threadIdx.x == 0) seg.prefix[NT] = total;
  __syncthreads();
  return total;
}
// Ball scan with exact load balance: the row bounds are fetched once (one thread per row), prefix-summed across
// the CTA, and every warp walks an equal share [T w / NW, T (w+1) / NW) of the flattened candidate range, row by row
// (rows are contiguous segments of the cell-sorted point array -> coalesced float4 loads, no per-candidate search).
// body(in_range, point, position in the cell-sorted array) is called by all 32 lanes together. Contains __syncthreads:
// call from uniform control flow.
template <int NT, class F>
__device__ __forceinline__ void scan_balanced(const DevParams &P, const DevCloud &cl, const SegRange &sr, SegScan<NT> &seg,
                                              F &&body) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = NT / 32;
  for (int row0 = 0; row0 < sr.nrows; row0 += NT) {
    __syncthreads();
    const int total = seg_batch<NT>(P, cl.cell_start, sr, row0, seg);  // fills seg.start / seg.prefix, syncs
    const int nr = min(NT, sr.nrows - row0);
    const int c_begin = (int)(((long long)total * warp) / NW), c_end = (int)(((long long)total * (warp + 1)) / NW);
    if (c_begin >= c_end) continue;
    int lo = 0, hi = nr;  // largest r with prefix[r] <= c_begin
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (seg.prefix[mid] <= c_begin) lo = mid; else hi = mid;
    }
    // software pipeline: the load of chunk i+1 is issued before chunk i is processed (the scan is latency bound:
    // ~10 short row segments per warp, each a dependent L2 round trip)
    int r = lo, c = c_begin;
    auto fetch = [&](bool &have, bool &in, float4 &p, int &where) {
      have = false;
      in = false;
      where = 0;
      p = make_float4(0.f, 0.f, 0.f, 0.f);
      while (c < c_end) {
        const int pe = (r + 1 < nr) ? seg.prefix[r + 1] : total;
        const int seg_end = min(pe, c_end);
        if (c >= seg_end) {  // row exhausted (or empty): next row
          r++;
          continue;
        }
        const int k = c + lane;
        in = k < seg_end;
        where = (seg.start[r] - seg.prefix[r]) + k;
        if (in) p = __ldg(cl.pts4 + where);
        c = min(c + 32, seg_end);
        have = true;
        return;
      }
    };
    bool have0, in0, have1, in1;
    float4 p0, p1;
    int w0, w1;
    fetch(have0, in0, p0, w0);
    while (have0) {
      fetch(have1, in1, p1, w1);
      body(in0, p0, w0);
      have0 = have1;
      in0 = in1;
      p0 = p1;
      w0 = w1;
    }
  }
}

template <int NT>
__device__ __forceinline__ int seg_lookup(const SegScan<NT> &seg, int c) {
  // largest r with prefix[r] <= c
  int lo = 0, hi = NT;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (seg.prefix[mid] <= c) lo = mid; else hi = mid;
  }
  return seg.start[lo] + (c - seg.prefix[lo]);
}

// ------------------------------------------------------------------------------------------------
// grid build
// ------------------------------------------------------------------------------------------------
__global__ void k_cell_ids(const DevParams *Pp, const float *xyz, int N, int *cid, int *idx, int *counts) {
  const DevParams &P = *Pp;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int c = (cell_of(P, xyz[3 * i + 2], 2) * P.dim[1] + cell_of(P, xyz[3 * i + 1], 1)) * P.dim[0] + cell_of(P, xyz[3 * i], 0);
  cid[i] = c;
  idx[i] = i;
  atomicAdd(counts + c + 1, 1);
}
__global__ void k_fill_sorted(const float *xyz, const int *idx_sorted, int N, float4 *pts4) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  int i = idx_sorted[k];
  pts4[k] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], __int_as_float(i));
}

// ------------------------------------------------------------------------------------------------
// 3x3 symmetric eigen-solver: Eigen::SelfAdjointEigenSolver<Matrix3d>::compute restated
// (same sequence of operations as oracle/gpd_oracle.cpp eigen3 -> identical bits).
// ------------------------------------------------------------------------------------------------
__device__ void make_givens(double p, double q, double &c, double &s) {
  if (q == 0.0) {
    c = p < 0.0 ? -1.0 : 1.0;
    s = 0.0;
  } else if (p == 0.0) {
    c = 0.0;
    s = q < 0.0 ? 1.0 : -1.0;
  } else if (fabs(p) > fabs(q)) {
    double t = q / p;
    double u = sqrt(1.0 + t * t);
    if (p < 0.0) u = -u;
    c = 1.0 / u;
    s = -t * c;
  } else {
    double t = p / q;
    double u = sqrt(1.0 + t * t);
    if (q < 0.0) u = -u;
    s = -1.0 / u;
    c = -t * s;
  }
}
__device__ double eigen_hypot(double x, double y) {
  double ax = fabs(x), ay = fabs(y);
  double p = fmax(ax, ay);
  if (p == 0.0) return 0.0;
  double qp = fmin(ax, ay) / p;
  return p * sqrt(1.0 + qp * qp);
}
// m: lower triangle used, column-major; eval ascending, Q column-major
__device__ void eigen3(const double *Min, double *eval, double *Q) {
  double m00 = Min[0], m10 = Min[1], m20 = Min[2], m11 = Min[4], m21 = Min[5], m22 = Min[8];
  double scale = fmax(fmax(fmax(fabs(m00), fabs(m10)), fmax(fabs(m20), fabs(m11))), fmax(fabs(m21), fabs(m22)));
  if (scale == 0.0) scale = 1.0;
  m00 /= scale; m10 /= scale; m20 /= scale; m11 /= scale; m21 /= scale; m22 /= scale;
  double diag[3], sub[2];
  const double tol = DBL_MIN;
  diag[0] = m00;
  double v1norm2 = m20 * m20;
  if (v1norm2 <= tol) {
    diag[1] = m11;
    diag[2] = m22;
    sub[0] = m10;
    sub[1] = m21;
    for (int i = 0; i < 9; i++) Q[i] = (i % 4 == 0) ? 1.0 : 0.0;
  } else {
    double beta = sqrt(m10 * m10 + v1norm2);
    double invBeta = 1.0 / beta;
    double m01 = m10 * invBeta;
    double m02 = m20 * invBeta;
    double q = 2.0 * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q;
    diag[2] = m22 - m02 * q;
    sub[0] = beta;
    sub[1] = m21 - m01 * q;
    Q[0] = 1; Q[1] = 0;   Q[2] = 0;
    Q[3] = 0; Q[4] = m01; Q[5] = m02;
    Q[6] = 0; Q[7] = m02; Q[8] = -m01;
  }
  const int n = 3;
  int end = n - 1, start = 0, iter = 0;
  const int maxIterations = 30;
  const double precision_inv = 1.0 / DBL_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i) {
      if (fabs(sub[i]) < DBL_MIN) {
        sub[i] = 0.0;
      } else {
        const double scaled = precision_inv * sub[i];
        if (scaled * scaled <= (fabs(diag[i]) + fabs(diag[i + 1]))) sub[i] = 0.0;
      }
    }
    while (end > 0 && sub[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIterations * n) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0) start--;
    double td = (diag[end - 1] - diag[end]) * 0.5;
    double e = sub[end - 1];
    double mu = diag[end];
    if (td == 0.0) {
      mu -= fabs(e);
    } else if (e != 0.0) {
      const double e2 = e * e;
      const double h = eigen_hypot(td, e);
      if (e2 == 0.0)
        mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
      else
        mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    double x = diag[start] - mu;
    double z = sub[start];
    for (int k = start; k < end && z != 0.0; ++k) {
      double c, s;
      make_givens(x, z, c, s);
      double sdk = s * diag[k] + c * sub[k];
      double dkp1 = s * sub[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      sub[k] = c * sdk - s * dkp1;
      if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
      x = sub[k];
      if (k < end - 1) {
        z = -s * sub[k + 1];
        sub[k + 1] = c * sub[k + 1];
      }
      for (int r = 0; r < 3; r++) {
        double xi = Q[k * 3 + r], yi = Q[(k + 1) * 3 + r];
        Q[k * 3 + r] = c * xi - s * yi;
        Q[(k + 1) * 3 + r] = s * xi + c * yi;
      }
    }
  }
  for (int i = 0; i < n - 1; ++i) {
    int k = 0;
    for (int j = 1; j < n - i; j++)
      if (diag[i + j] < diag[i + k]) k = j;
    if (k > 0) {
      double t = diag[i]; diag[i] = diag[k + i]; diag[k + i] = t;
      for (int r = 0; r < 3; r++) { t = Q[i * 3 + r]; Q[i * 3 + r] = Q[(k + i) * 3 + r]; Q[(k + i) * 3 + r] = t; }
    }
  }
  for (int i = 0; i < 3; i++) eval[i] = diag[i] * scale;
}

// ------------------------------------------------------------------------------------------------
// k_frames: one warp per sample. The r = nn_radius ball (~44 points) is gathered as 64-bit keys
// (dist bits << 32 | index), rank-sorted so that N*N^T and sum(n) are accumulated in exactly the
// (dist, index) order the reference's sorted radiusSearch yields -> bit-identical to the oracle.
// ------------------------------------------------------------------------------------------------
// Three tiers: tier 0 runs every sample with a small per-warp list (cap0 keys: 16 CTAs per SM instead of 3); samples whose
// ball does not fit are appended to `ovf` and re-run by tier 1 with LRF_CAP keys in shared memory; what still does not fit
// goes to `ovf2` and tier 2, whose lists live in a per-warp slice of global memory (`gkeys`, LRF_CAP_GLOBAL keys, a
// grid-stride loop over the list). Beyond that: err[0]. The reference has no limit (frame_estimator.cpp:6-86); a voxelised
// cloud never leaves tier 0 (at most ~155 voxels of 3 mm in a 1 cm ball).
__device__ __forceinline__ bool lrf_sample(const DevParams &P, const DevCloud &cl, const int *sidx, int i, unsigned long long *keys,
                                           unsigned long long *sorted, int cap, bool last, double *frames, uint8_t *valid, int *err,
                                           double *s_acc_w);

__global__ void __launch_bounds__(LRF_WARPS * 32) k_frames(const DevParams *Pp, DevCloud cl, const int *sidx, int n,
                                                            double *frames, uint8_t *valid, int *err, int cap, int *ovf,
                                                            int *ovf_count, int *ovf2, int *ovf2_count, unsigned long long *gkeys,
                                                            int tier) {
  const DevParams &P = *Pp;
  extern __shared__ __align__(16) unsigned char lrf_dyn[];
  __shared__ double s_acc[LRF_WARPS][9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (tier == 2) {
    const int gw = blockIdx.x * LRF_WARPS + warp, nw = gridDim.x * LRF_WARPS, cnt2 = *ovf2_count;
    unsigned long long *keys = gkeys + (size_t)gw * 2 * cap;
    for (int j = gw; j < cnt2; j += nw) {
      lrf_sample(P, cl, sidx, ovf2[j], keys, keys + cap, cap, true, frames, valid, err, s_acc[warp]);
      __syncwarp();
    }
    return;
  }
  int i = blockIdx.x * LRF_WARPS + warp;
  if (tier == 0) {
    if (i >= n) return;
  } else {
    if (i >= *ovf_count) return;
    i = ovf[i];
  }
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(lrf_dyn) + (size_t)warp * cap;
  unsigned long long *sorted = reinterpret_cast<unsigned long long *>(lrf_dyn) + (size_t)(LRF_WARPS + warp) * cap;
  if (!lrf_sample(P, cl, sidx, i, keys, sorted, cap, false, frames, valid, err, s_acc[warp]) && lane == 0) {
    if (tier == 0) ovf[atomicAdd(ovf_count, 1)] = i;  // re-run with the large list
    else ovf2[atomicAdd(ovf2_count, 1)] = i;          // ... with the global-memory list
  }
}

// one sample by one warp; false when the ball holds more than `cap` points and this is not the last tier (nothing written)
__device__ __forceinline__ bool lrf_sample(const DevParams &P, const DevCloud &cl, const int *sidx, int i, unsigned long long *keys,
                                           unsigned long long *sorted, int cap, bool last, double *frames, uint8_t *valid, int *err,
                                           double *s_acc_w) {
  const int lane = threadIdx.x & 31;
  const int si = sidx[i];
  double sp[3];
  sample_position(cl, si, sp);
  float q[3] = {(float)sp[0], (float)sp[1], (float)sp[2]};
  SegRange sr = seg_range(P, q, P.rf_lrf);
  int cnt = 0;
  for (int row = 0; row < sr.nrows; row++) {
    int st, len;
    seg_row(P, cl.cell_start, sr, row, st, len);
    for (int k0 = 0; k0 < len; k0 += 32) {
      int k = k0 + lane;
      bool hit = false;
      unsigned long long key = 0;
      if (k < len) {
        float4 p = __ldg(cl.pts4 + st + k);
        float d = l2_simple(q, p.x, p.y, p.z);
        hit = d < P.r2_lrf;
        key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      int pos = cnt + __popc(m & ((1u << lane) - 1));
      if (hit && pos < cap) keys[pos] = key;
      cnt += __popc(m);
    }
  }
  if (cnt > cap) {
    if (!last) return false;
    if (lane == 0) atomicAdd(err + 0, 1);
    cnt = cap;
  }
  __syncwarp();
  if (cnt == 0) {
    if (lane == 0) valid[i] = 0;
    if (lane < 9) frames[9 * (size_t)i + lane] = 0.0;
    return true;
  }
  for (int a = lane; a < cnt; a += 32) {
    unsigned long long ka = keys[a];
    int rank = 0;
    for (int b = 0; b < cnt; b++) rank += (keys[b] < ka);
    sorted[rank] = ka;
  }
  __syncwarp();
  // lanes 0..5: lower-triangle entries of M = N N^T ; lanes 6..8: sum of normals
  if (lane < 9) {
    const int rr[9] = {0, 1, 2, 1, 2, 2, 0, 1, 2}, cc[9] = {0, 0, 0, 1, 1, 2, 0, 0, 0};
    int r = rr[lane], c = cc[lane];
    double acc = 0.0;
    for (int a = 0; a < cnt; a++) {
      int idx = (int)(unsigned)(sorted[a] & 0xffffffffu);
      const double *nn = cl.nrm + 3 * (size_t)idx;
      if (lane < 6) acc += nn[r] * nn[c]; else acc += nn[r];
    }
    s_acc_w[lane] = acc;
  }
  __syncwarp();
  if (lane == 0) {
    const double *a = s_acc_w;
    double M[9] = {a[0], a[1], a[2], a[1], a[3], a[4], a[2], a[4], a[5]};
    double eval[3], evec[9];
    eigen3(M, eval, evec);
    int mn = 0, mx = 0;
    for (int k = 1; k < 3; k++) {
      if (eval[k] < eval[mn]) mn = k;
      if (eval[k] > eval[mx]) mx = k;
    }
    double curv[3] = {evec[mn * 3], evec[mn * 3 + 1], evec[mn * 3 + 2]};
    double normal[3] = {evec[mx * 3], evec[mx * 3 + 1], evec[mx * 3 + 2]};
    double nrm = sqrt((a[6] * a[6] + a[7] * a[7]) + a[8] * a[8]);
    double avg[3] = {a[6] / nrm, a[7] / nrm, a[8] / nrm};
    if ((avg[0] * normal[0] + avg[1] * normal[1]) + avg[2] * normal[2] < 0) {
      normal[0] *= -1.0; normal[1] *= -1.0; normal[2] *= -1.0;
    }
    double *f = frames + 9 * (size_t)i;
    f[0] = normal[0]; f[1] = normal[1]; f[2] = normal[2];
    f[3] = curv[1] * normal[2] - curv[2] * normal[1];
    f[4] = curv[2] * normal[0] - curv[0] * normal[2];
    f[5] = curv[0] * normal[1] - curv[1] * normal[0];
    f[6] = curv[0]; f[7] = curv[1]; f[8] = curv[2];
    valid[i] = 1;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// k_hands
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned slot_mask(const DevParams &P, const double *sfs, const double *sfsw, double y) {
  unsigned m = 0;
  const int F = 2 * P.nfp;
  if (P.slots_disjoint) {
    // slots within each half are disjoint, ascending and (nearly) equally spaced: estimate the slot index
    // arithmetically and test the estimate and its two neighbours with the exact table values
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const double *fs = sfs + half * P.nfp, *fsw = sfsw + half * P.nfp;
      const double tq = (y - fs[0]) * P.inv_slot_step;
      const double tf = floor(tq);
      if (tq - tf > 1e-9 && tf + 1.0 - tq > 1e-9) {
        // clearly inside one slot pitch: only that slot can contain y (one exact test)
        const int i = (int)tf;
        if (i >= 0 && i < P.nfp && y > fs[i] && y < fsw[i]) m |= 1u << (half * P.nfp + i);
      } else {
        // within rounding distance of a pitch boundary: test the estimate and both neighbours exactly
        const int e = min(max((int)tf, 1), P.nfp - 2 > 1 ? P.nfp - 2 : 1);
#pragma unroll
        for (int d = -1; d <= 1; d++) {
          const int i = e + d;
          if (i >= 0 && i < P.nfp && y > fs[i] && y < fsw[i]) m |= 1u << (half * P.nfp + i);
        }
      }
    }
  } else {
    for (int f = 0; f < F; f++)
      if (y > sfs[f] && y < sfsw[f]) m |= 1u << f;
  }
  return m;
}

constexpr int SURV_CAP = 1024;
struct HandsSmem {
  SegScan<NT_HANDS> seg;
  unsigned short surv[NT_HANDS / 32][SURV_CAP];  // per-warp closing-region member indices
  double fs[GPDB_MAX_SLOTS], fsw[GPDB_MAX_SLOTS];  // finger slot tables (copied from DevParams)
  int count;      // staged (slab) points
  int n_ball;     // all points of the r ball
  unsigned long long nb0_key;
  double T[9];    // frame * ROT_BINORMAL
  double frame[9];
  double sample[3];
  int cap;
};

// one CTA per sample; dynamic smem: float4 list[cap]
// tiers: in_list == nullptr: every sample, else the samples in_list[0 .. *in_count) (the overflow of the previous tier);
// samples whose slab does not fit `cap` go to out_list (next tier) or, in the last tier (out_list == nullptr), are an error.
// glist != nullptr: the staged neighbourhood lives in a per-CTA slice of global memory (last tier, any density up to cap).
__global__ void __launch_bounds__(NT_HANDS, 4) k_hands(const DevParams *Pp, DevCloud cl, const int *sidx, int n, int slot0,
                                                    const double *frames, const uint8_t *fvalid, gpdb_pose *poses,
                                                    uint8_t *flags, int cap, const int *in_list, const int *in_count,
                                                    int *out_list, int *out_count, float4 *glist, int *err) {
  const DevParams &P = *Pp;
  extern __shared__ __align__(16) unsigned char dyn[];
  float4 *list = glist ? glist + (size_t)blockIdx.x * cap : reinterpret_cast<float4 *>(dyn);
  __shared__ HandsSmem S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int work_n = in_list ? *in_count : n;
  if (tid < 2 * P.nfp) {
    S.fs[tid] = P.fs[tid];
    S.fsw[tid] = P.fsw[tid];
  }
  for (int w = blockIdx.x; w < work_n; w += gridDim.x) {
    const int i = in_list ? in_list[w] : w;
    const int si = sidx[i];
    __syncthreads();
    if (tid == 0) {
      S.count = 0;
      S.n_ball = 0;
      S.nb0_key = ~0ull;
    }
    if (tid < 9) S.frame[tid] = frames[9 * (size_t)i + tid];
    if (tid == 0) sample_position(cl, si, S.sample);
    __syncthreads();
    if (tid == 0) mat3_mul(S.frame, P.rotb, S.T);
    const bool fv = fvalid[i] != 0;
    if (!fv) {
      // no local frame (calculateFrame returned nullptr): no hand set for this sample
      for (int p = tid; p < P.P; p += NT_HANDS) {
        gpdb_pose *h = poses + (size_t)i * P.P + p;
        memset(h, 0, sizeof(gpdb_pose));
        h->sample_index = si;
        h->sample_slot = slot0 + i;
        h->pose_slot = (int16_t)p;
        h->finger_idx = -1;
        h->score = __int_as_float(0x7fc00000);
        flags[(size_t)i * P.P + p] = 0;
      }
      continue;
    }
    __syncthreads();
    // ---- stage the neighbourhood: ball r = nn_radius_hs, kept if inside the (slightly widened)
    // height slab when every rotation axis is the curvature axis (z is then pose independent)
    float q[3] = {(float)S.sample[0], (float)S.sample[1], (float)S.sample[2]};
    SegRange sr = seg_range(P, q, P.rf_hs);
    const double hz = P.hand_height * 1.001 + 1e-9;
    unsigned long long best = ~0ull;
    int nball = 0;
    scan_balanced<NT_HANDS>(P, cl, sr, S.seg, [&](bool in, const float4 &p, int) {
      bool keep = false;
      if (in) {
        float d = l2_simple(q, p.x, p.y, p.z);
        if (d < P.r2_hs) {
          nball++;
          unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
          best = key < best ? key : best;
          keep = true;
          if (P.all_axes_z) {
            double z0 = (S.T[6] * ((double)p.x - S.sample[0]) + S.T[7] * ((double)p.y - S.sample[1])) +
                        S.T[8] * ((double)p.z - S.sample[2]);
            keep = fabs(z0) < hz;
          }
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, keep);
      if (m) {
        int leader = __ffs(m) - 1, base = 0;
        if (lane == leader) base = atomicAdd(&S.count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        int pos = base + __popc(m & ((1u << lane) - 1));
        if (keep && pos < cap) list[pos] = p;
      }
    });
    nball = warp_sum(nball);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
      best = ob < best ? ob : best;
    }
    if (lane == 0) {
      atomicAdd(&S.n_ball, nball);
      atomicMin(&S.nb0_key, best);
    }
    __syncthreads();
    const int m = S.count;
    if (m > cap) {
      // does not fit this tier: defer to the next one (or report)
      if (tid == 0) {
        if (out_list) {
          int k = atomicAdd(out_count, 1);
          out_list[k] = i;
          if (!in_list) atomicAdd(err + 3, 1);
        } else {
          atomicAdd(err + 1, 1);
        }
      }
      if (!out_list) {
        for (int p = tid; p < P.P; p += NT_HANDS) flags[(size_t)i * P.P + p] = 0;
      }
      continue;
    }
    const int n_ball = S.n_ball;
    // nb0 = nearest neighbour (first of the sorted radius search): pads the cropped list
    // (PointList::cropByHandHeight quirk, point_list.cpp:44-55)
    const int nb0 = (int)(unsigned)(S.nb0_key & 0xffffffffu);
    double nbx = 0, nby = 0, nbz = 0;
    if (n_ball > 0) {
      nbx = (double)cl.xyz[3 * (size_t)nb0] - S.sample[0];
      nby = (double)cl.xyz[3 * (size_t)nb0 + 1] - S.sample[1];
      nbz = (double)cl.xyz[3 * (size_t)nb0 + 2] - S.sample[2];
    }
    // ---- one warp per pose
    for (int pose = warp; pose < P.P; pose += NT_HANDS / 32) {
      gpdb_pose *h = poses + (size_t)i * P.P + pose;
      double R[9];
      mat3_mul(S.T, P.rot[pose], R);
      const double hh = P.hand_height;
      uint8_t fl = 0;
      double top = 0, bottom = 0, center = 0, width = 0;
      int fidx = -1, fpi = -1;
      bool half = false, full = false;
      double x0, y0, z0;
      to_frame(R, nbx, nby, nbz, x0, y0, z0);
      if (n_ball > 0) {
        // pass A: crop + finger masks (FingerHand::evaluateFingers, finger_hand.cpp:26-73)
        const double b0 = P.init_bite, bot0 = P.init_bite - P.hand_depth;
        int k = 0;
        unsigned anyA = 0, anyB = 0, fmask = 0;
        for (int a = lane; a < m; a += 32) {
          float4 p = list[a];
          double x, y, z;
          to_frame(R, (double)p.x - S.sample[0], (double)p.y - S.sample[1], (double)p.z - S.sample[2], x, y, z);
          if (z > -1.0 * hh && z < hh) {
            k++;
            if (x < b0) {
              anyA = 1;
              if (x < bot0) anyB = 1;
              fmask |= slot_mask(P, S.fs, S.fsw, y);
            }
          }
        }
        k = warp_sum(k);
        const int npad = n_ball - k;
        if (npad > 0 && x0 < b0) {
          anyA = 1;
          if (x0 < bot0) anyB = 1;
          fmask |= slot_mask(P, S.fs, S.fsw, y0);
        }
        anyA = __reduce_or_sync(0xffffffffu, anyA);
        anyB = __reduce_or_sync(0xffffffffu, anyB);
        fmask = __reduce_or_sync(0xffffffffu, fmask);
        const unsigned allF = (2 * P.nfp >= 32) ? 0xffffffffu : ((1u << (2 * P.nfp)) - 1);
        unsigned freef = (anyA && !anyB) ? (~fmask & allF) : 0u;
        unsigned hand = freef & (freef >> P.nfp) & ((1u << P.nfp) - 1);  // evaluateHand (:75-81)
        if (hand) {
          // chooseMiddleHand (:89-105): hand_idx[ceil(m/2) - 1]
          int cntb = __popc(hand);
          int target = (cntb + 1) / 2;  // 1-based ordinal of the chosen set bit
          unsigned hm = hand;
          while (--target) hm &= hm - 1;
          fidx = __ffs(hm) - 1;
          // Hand::construct (hand.cpp:33-38): finger_placement_index_ = FIRST set bit of hand_; deepenHand
          // resets hand_ to the eroded index only (finger_hand.cpp:134-136), chooseMiddleHand does not
          fpi = P.deepen ? fidx : (__ffs(hand) - 1);
          top = b0;
          bottom = bot0;
          if (P.deepen) {
            // deepenHand (:107-139) as one min-reduction over the first failing step
            int jf = P.J;  // 0-based index of the first failing step; J = none fails
            const double sl0 = P.fs[fidx], sl1 = P.fsw[fidx], sr0 = P.fs[P.nfp + fidx], sr1 = P.fsw[P.nfp + fidx];
            const int J = P.J;
            auto visitB = [&](double x, double y) {
              if (J > 0 && x < P.botj[J - 1]) {
                int j = 0;
                while (!(x < P.botj[j])) j++;
                jf = min(jf, j);
              }
              if ((y > sl0 && y < sl1) || (y > sr0 && y < sr1)) {
                if (J > 0 && x < P.topj[J - 1]) {
                  int j = 0;
                  while (!(x < P.topj[j])) j++;
                  jf = min(jf, j);
                }
              }
            };
            for (int a = lane; a < m; a += 32) {
              float4 p = list[a];
              double x, y, z;
              to_frame(R, (double)p.x - S.sample[0], (double)p.y - S.sample[1], (double)p.z - S.sample[2], x, y, z);
              if (z > -1.0 * hh && z < hh) visitB(x, y);
            }
            if (npad > 0) visitB(x0, y0);
            jf = __reduce_min_sync(0xffffffffu, jf);
            if (jf > 0) {
              top = P.topj[jf - 1];
              bottom = P.botj[jf - 1];
            }
          }
          // computePointsInClosingRegion (:141-171)
          const double left = P.fsw[fidx], right = P.fs[P.nfp + fidx];
          center = 0.5 * (left + right);
          int cnt = 0;
          double mny = DBL_MAX, mxy = -DBL_MAX;
          // the closing-region members (~10 % of the slab) are remembered per warp so that the two Antipodal passes
          // below visit only them instead of re-transforming the whole slab
          unsigned short *surv = S.surv[warp];
          int nsurv = 0;
          for (int a0 = 0; a0 < m; a0 += 32) {
            const int a = a0 + lane;
            bool inr = false;
            if (a < m) {
              float4 p = list[a];
              double x, y, z;
              to_frame(R, (double)p.x - S.sample[0], (double)p.y - S.sample[1], (double)p.z - S.sample[2], x, y, z);
              if (z > -1.0 * hh && z < hh && x > bottom && x < top && y > left && y < right) {
                inr = true;
                cnt++;
                mny = fmin(mny, y);
                mxy = fmax(mxy, y);
              }
            }
            const unsigned mk = __ballot_sync(0xffffffffu, inr);
            const int pos = nsurv + __popc(mk & ((1u << lane) - 1));
            if (inr && pos < SURV_CAP) surv[pos] = (unsigned short)a;
            nsurv += __popc(mk);
          }
          const bool surv_ok = nsurv <= SURV_CAP && m <= 65535;
          __syncwarp();
          const bool nb_in = npad > 0 && x0 > bottom && x0 < top && y0 > left && y0 < right;
          cnt = warp_sum(cnt) + (nb_in ? npad : 0);
          mny = warp_min(mny);
          mxy = warp_max(mxy);
          if (nb_in) {
            mny = fmin(mny, y0);
            mxy = fmax(mxy, y0);
          }
          if (cnt > 0) {
            fl |= GPDB_POSE_VALID;
            width = mxy - mny;  // modifyCandidate (hand_set.cpp:245-247)
            // Antipodal::evaluateGrasp (antipodal.cpp:10-96), lateral = 1, forward = 0, vertical = 2
            const double min_x = mny + 0.003, max_x = mxy - 0.003;
            int cl_ = 0, cr_ = 0;
            double lmaxx = -DBL_MAX, lminx = DBL_MAX, lmaxz = -DBL_MAX, lminz = DBL_MAX;
            double rmaxx = -DBL_MAX, rminx = DBL_MAX, rmaxz = -DBL_MAX, rminz = DBL_MAX;
            auto visitD = [&](double x, double y, double z, int idx, int wgt) {
              const double *nn = cl.nrm + 3 * (size_t)idx;
              double n0, n1, n2;
              to_frame(R, nn[0], nn[1], nn[2], n0, n1, n2);
              double ldot = (0.0 * n0 + -1.0 * n1) + 0.0 * n2;
              double rdot = (0.0 * n0 + 1.0 * n1) + 0.0 * n2;
              if (ldot > P.cosf && y < min_x) {
                cl_ += wgt;
                lmaxx = fmax(lmaxx, x); lminx = fmin(lminx, x); lmaxz = fmax(lmaxz, z); lminz = fmin(lminz, z);
              }
              if (rdot > P.cosf && y > max_x) {
                cr_ += wgt;
                rmaxx = fmax(rmaxx, x); rminx = fmin(rminx, x); rmaxz = fmax(rmaxz, z); rminz = fmin(rminz, z);
              }
            };
            if (surv_ok) {
              for (int sidx_ = lane; sidx_ < nsurv; sidx_ += 32) {
                float4 p = list[surv[sidx_]];
                double x, y, z;
                to_frame(R, (double)p.x - S.sample[0], (double)p.y - S.sample[1], (double)p.z - S.sample[2], x, y, z);
                visitD(x, y, z, __float_as_int(p.w), 1);
              }
            } else {
              for (int a = lane; a < m; a += 32) {
                float4 p = list[a];
                double x, y, z;
                to_frame(R, (double)p.x - S.sample[0], (double)p.y - S.sample[1], (double)p.z - S.sample[2], x, y, z);
                if (z > -1.0 * hh && z < hh && x > bottom && x < top && y > left && y < right)
                  visitD(x, y, z, __float_as_int(p.w), 1);
              }
            }
            if (nb_in && lane == 0) visitD(x0, y0, z0, nb0, npad);
            cl_ = warp_sum(cl_);
            cr_ = warp_sum(cr_);
            half = cl_ > 0 || cr_ > 0;
            if (cl_ > 0 && cr_ > 0) {
              lmaxx = warp_max(lmaxx); lminx = warp_min(lminx); lmaxz = warp_max(lmaxz); lminz = warp_min(lminz);
              rmaxx = warp_max(rmaxx); rminx = warp_min(rminx); rmaxz = warp_max(rmaxz); rminz = warp_min(rminz);
              const double top_y = fmin(lmaxx, rmaxx), bot_y = fmax(lminx, rminx);
              const double top_z = fmin(lmaxz, rmaxz), bot_z = fmax(lminz, rminz);
              int nl = 0, nr = 0;
              auto visitE = [&](double x, double y, double z, int idx, int wgt) {
                const double *nn = cl.nrm + 3 * (size_t)idx;
                double n0, n1, n2;
                to_frame(R, nn[0], nn[1], nn[2], n0, n1, n2);
                double ldot = (0.0 * n0 + -1.0 * n1) + 0.0 * n2;
                double rdot = (0.0 * n0 + 1.0 * n1) + 0.0 * n2;
                bool inw = x >= bot_y && x <= top_y && z >= bot_z && z <= top_z;
                if (ldot > P.cosf && y < min_x && inw) nl += wgt;
                if (rdot > P.cosf && y > max_x && inw) nr += wgt;
              };
              if (surv_ok) {
                for (int sidx_ = lane; sidx_ < nsurv; sidx_ += 32) {
                  float4 p = list[surv[sidx_]];
                  double x, y, z;
                  to_frame(R, (double)p.x - S.sample[0], (double)p.y - S.sample[1], (double)p.z - S.sample[2], x, y, z);
                  visitE(x, y, z, __float_as_int(p.w), 1);
                }
              } else {
                for (int a = lane; a < m; a += 32) {
                  float4 p = list[a];
                  double x, y, z;
                  to_frame(R, (double)p.x - S.sample[0], (double)p.y - S.sample[1], (double)p.z - S.sample[2], x, y, z);
                  if (z > -1.0 * hh && z < hh && x > bottom && x < top && y > left && y < right)
                    visitE(x, y, z, __float_as_int(p.w), 1);
                }
              }
              if (nb_in && lane == 0) visitE(x0, y0, z0, nb0, npad);
              nl = warp_sum(nl);
              nr = warp_sum(nr);
              full = nl >= P.min_viable && nr >= P.min_viable;
            }
            if (half) fl |= GPDB_POSE_HALF;
            if (full) fl |= GPDB_POSE_FULL;
          }
        }
      }
      // ---- record (Hand::construct, hand.cpp:24-45) + A15 filters, by lane 0
      if (lane == 0) {
        gpdb_pose o;
        memset(&o, 0, sizeof(o));
        for (int r = 0; r < 3; r++) o.sample[r] = S.sample[r];
        for (int r = 0; r < 9; r++) o.frame[r] = R[r];
        o.sample_index = si;
        o.sample_slot = slot0 + i;
        o.pose_slot = (int16_t)pose;
        o.finger_idx = -1;
        o.score = __int_as_float(0x7fc00000);
        if (fl & GPDB_POSE_VALID) {
          o.top = top;
          o.bottom = bottom;
          o.center = center;
          o.width = width;
          o.finger_idx = (int16_t)fpi;
          o.half_antipodal = half;
          o.full_antipodal = full;
          for (int r = 0; r < 3; r++)
            o.position[r] = ((R[r] * bottom + R[3 + r] * center) + R[6 + r] * 0.0) + S.sample[r];
          // filterGraspsWorkspace (grasp_detector.cpp:334-398; right_top uses left_bottom, :362-363)
          const double half_width = 0.5 * P.hand_outer_diameter;
          bool ok = width >= P.min_ap && width <= P.max_ap;
          for (int r = 0; r < 3; r++) {
            double lb = o.position[r] + half_width * R[3 + r];
            double rb = o.position[r] - half_width * R[3 + r];
            double lt = lb + P.hand_depth * R[r];
            double rt = lb + P.hand_depth * R[r];
            double ap = o.position[r] - 0.05 * R[r];
            double mn = fmin(fmin(fmin(lb, rb), fmin(lt, rt)), ap);
            double mx = fmax(fmax(fmax(lb, rb), fmax(lt, rt)), ap);
            ok = ok && mn >= P.ws[2 * r] && mx <= P.ws[2 * r + 1];
          }
          if (ok && P.filt_dir) {  // filterGraspsDirection (:422-456)
            double dot = (P.dir[0] * R[0] + P.dir[1] * R[1]) + P.dir[2] * R[2];
            if (acos(dot) > P.thresh) ok = false;
          }
          if (ok) fl |= GPDB_POSE_FILTERED;
        }
        *h = o;
        flags[(size_t)i * P.P + pose] = fl;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// compaction of candidate poses (hands_out order of image_generator.cpp:91-98)
// ------------------------------------------------------------------------------------------------
__global__ void k_flag01(const uint8_t *flags, int n, int *f01) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f01[i] = (flags[i] & 3) == 3;
}
__global__ void k_scatter(const gpdb_pose *poses, const int *f01, const int *pos, int n, gpdb_pose *cand, int *count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (f01[i]) cand[pos[i]] = poses[i];
  if (i == n - 1) *count = pos[i] + f01[i];
}
__global__ void k_scatter_scores(const gpdb_pose *cand, const float *scores, int nc, int slot0, int P, float *pose_scores,
                                 gpdb_pose *cand_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  float s = scores[i];
  pose_scores[(size_t)(cand[i].sample_slot - slot0) * P + cand[i].pose_slot] = s;
  cand_out[i].score = s;
}

// selectGrasps (grasp_detector.cpp:405-420) on the device: sort key = descending score, stable in candidate order
__global__ void k_select_keys(const gpdb_pose *cand, int n, unsigned *keys, int *vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned u = __float_as_uint(cand[i].score);
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;  // ascending float order as unsigned
  keys[i] = ~u;                                // descending
  vals[i] = i;
}
__global__ void k_gather_poses(const gpdb_pose *cand, const int *order, int k, gpdb_pose *out) {
  // one warp per record: 176-byte pose = 44 words
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= k) return;
  const int *src = reinterpret_cast<const int *>(cand + order[w]);
  int *dst = reinterpret_cast<int *>(out + w);
  for (int j = lane; j < (int)(sizeof(gpdb_pose) / 4); j += 32) dst[j] = src[j];
}

// ------------------------------------------------------------------------------------------------
// HandSearch::reevaluateHypotheses (hand_search.cpp:66-134,190-228): the given hands are re-labelled against the installed
// cloud (GraspDetector::evalGroundTruth: a ground-truth mesh cloud). One WARP per hand; the r = nn_radius_hs ball is
// walked four times straight from the grid (L2-resident cloud) — finger test of the hand's own slot at its own depth,
// closing region, the two Antipodal passes — with the same order-free reductions and neighbour-0 padding weights as
// k_hands. A labelling path, not a throughput path: no shared-memory staging.
// ------------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void warp_scan_ball(const DevParams &P, const DevCloud &cl, const SegRange &sr, const float q[3], float r2,
                                               F &&body) {  // body(point) for every in-ball point, lanes in lockstep
  const int lane = threadIdx.x & 31;
  for (int j0 = 0; j0 < sr.nrows; j0 += 32) {
    const int myrow = j0 + lane;
    int st = 0, len = 0;
    if (myrow < sr.nrows) seg_row(P, cl.cell_start, sr, myrow, st, len);
    unsigned nonempty = __ballot_sync(0xffffffffu, len > 0);
    while (nonempty) {
      const int j = __ffs(nonempty) - 1;
      nonempty &= nonempty - 1;
      const int rs = __shfl_sync(0xffffffffu, st, j), rl = __shfl_sync(0xffffffffu, len, j);
      for (int k0 = 0; k0 < rl; k0 += 32) {
        const int k = k0 + lane;
        if (k < rl) {
          const float4 p = __ldg(cl.pts4 + rs + k);
          if (l2_simple(q, p.x, p.y, p.z) < r2) body(p);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(128) k_reeval(const DevParams *Pp, DevCloud cl, gpdb_pose *hands, int n, int *labels) {
  const DevParams &P = *Pp;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  gpdb_pose &h = hands[i];
  double R[9], smp[3];
#pragma unroll
  for (int r = 0; r < 9; r++) R[r] = h.frame[r];
#pragma unroll
  for (int r = 0; r < 3; r++) smp[r] = h.sample[r];
  const int idx = h.finger_idx;
  const double top = h.top, bottom = top - P.hand_depth, hh = P.hand_height;  // evaluateFingers(points, hand.getTop(), idx)
  const float q[3] = {(float)smp[0], (float)smp[1], (float)smp[2]};            // eigenVectorToPcl (:136-142)
  const SegRange sr = seg_range(P, q, P.rf_hs);
  int label = 0;
  bool half = false, full = false;
  if (idx >= 0 && idx < P.nfp) {
    const double s0 = P.fs[idx], s0w = P.fsw[idx], s1 = P.fs[P.nfp + idx], s1w = P.fsw[P.nfp + idx];
    // pass 1: neighbour 0, crop count, back-of-hand collision, the two finger gaps
    int nball = 0, k = 0;
    unsigned long long best = ~0ull;
    unsigned anyA = 0, anyB = 0, blocked = 0;
    auto finger_test = [&](double x, double y) {
      if (x < top) {
        anyA = 1;
        if (x < bottom) anyB = 1;
        if ((y > s0 && y < s0w) || (y > s1 && y < s1w)) blocked = 1;
      }
    };
    warp_scan_ball(P, cl, sr, q, P.r2_hs, [&](const float4 &p) {
      nball++;
      const unsigned long long key = ((unsigned long long)__float_as_uint(l2_simple(q, p.x, p.y, p.z)) << 32) | (unsigned)__float_as_int(p.w);
      best = key < best ? key : best;
      double x, y, z;
      to_frame(R, (double)p.x - smp[0], (double)p.y - smp[1], (double)p.z - smp[2], x, y, z);
      if (z > -1.0 * hh && z < hh) {
        k++;
        finger_test(x, y);
      }
    });
    nball = warp_sum(nball);
    k = warp_sum(k);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
      best = ob < best ? ob : best;
    }
    const int npad = nball - k;  // cropByHandHeight pads with copies of neighbour 0 (point_list.cpp:44-55)
    const int nb0 = (int)(unsigned)(best & 0xffffffffull);
    double x0 = 0, y0 = 0, z0 = 0;
    if (nball > 0)
      to_frame(R, (double)cl.xyz[3 * (size_t)nb0] - smp[0], (double)cl.xyz[3 * (size_t)nb0 + 1] - smp[1],
               (double)cl.xyz[3 * (size_t)nb0 + 2] - smp[2], x0, y0, z0);
    if (npad > 0 && lane == 0) finger_test(x0, y0);
    anyA = __reduce_or_sync(0xffffffffu, anyA);
    anyB = __reduce_or_sync(0xffffffffu, anyB);
    blocked = __reduce_or_sync(0xffffffffu, blocked);
    if (nball > 0 && anyA && !anyB && !blocked) {
      // pass 2: computePointsInClosingRegion (finger_hand.cpp:141-171)
      const double left = s0w, right = s1;
      int cnt = 0;
      double mny = DBL_MAX, mxy = -DBL_MAX;
      auto in_region = [&](double x, double y) { return x > bottom && x < top && y > left && y < right; };
      warp_scan_ball(P, cl, sr, q, P.r2_hs, [&](const float4 &p) {
        double x, y, z;
        to_frame(R, (double)p.x - smp[0], (double)p.y - smp[1], (double)p.z - smp[2], x, y, z);
        if (z > -1.0 * hh && z < hh && in_region(x, y)) {
          cnt++;
          mny = fmin(mny, y);
          mxy = fmax(mxy, y);
        }
      });
      const bool nb_in = npad > 0 && in_region(x0, y0);
      cnt = warp_sum(cnt) + (nb_in ? npad : 0);
      mny = warp_min(mny);
      mxy = warp_max(mxy);
      if (nb_in) {
        mny = fmin(mny, y0);
        mxy = fmax(mxy, y0);
      }
      if (cnt > 0) {
        // passes 3 and 4: Antipodal::evaluateGrasp (antipodal.cpp:10-96), as in k_hands
        const double min_x = mny + 0.003, max_x = mxy - 0.003;
        int cl_ = 0, cr_ = 0;
        double lmaxx = -DBL_MAX, lminx = DBL_MAX, lmaxz = -DBL_MAX, lminz = DBL_MAX;
        double rmaxx = -DBL_MAX, rminx = DBL_MAX, rmaxz = -DBL_MAX, rminz = DBL_MAX;
        auto dots = [&](int pidx, double &ldot, double &rdot) {
          const double *nn = cl.nrm + 3 * (size_t)pidx;
          double n0, n1, n2;
          to_frame(R, nn[0], nn[1], nn[2], n0, n1, n2);
          ldot = (0.0 * n0 + -1.0 * n1) + 0.0 * n2;
          rdot = (0.0 * n0 + 1.0 * n1) + 0.0 * n2;
        };
        auto visitD = [&](double x, double y, double z, int pidx, int wgt) {
          double ldot, rdot;
          dots(pidx, ldot, rdot);
          if (ldot > P.cosf && y < min_x) {
            cl_ += wgt;
            lmaxx = fmax(lmaxx, x); lminx = fmin(lminx, x); lmaxz = fmax(lmaxz, z); lminz = fmin(lminz, z);
          }
          if (rdot > P.cosf && y > max_x) {
            cr_ += wgt;
            rmaxx = fmax(rmaxx, x); rminx = fmin(rminx, x); rmaxz = fmax(rmaxz, z); rminz = fmin(rminz, z);
          }
        };
        warp_scan_ball(P, cl, sr, q, P.r2_hs, [&](const float4 &p) {
          double x, y, z;
          to_frame(R, (double)p.x - smp[0], (double)p.y - smp[1], (double)p.z - smp[2], x, y, z);
          if (z > -1.0 * hh && z < hh && in_region(x, y)) visitD(x, y, z, __float_as_int(p.w), 1);
        });
        if (nb_in && lane == 0) visitD(x0, y0, z0, nb0, npad);
        cl_ = warp_sum(cl_);
        cr_ = warp_sum(cr_);
        half = cl_ > 0 || cr_ > 0;
        if (cl_ > 0 && cr_ > 0) {
          lmaxx = warp_max(lmaxx); lminx = warp_min(lminx); lmaxz = warp_max(lmaxz); lminz = warp_min(lminz);
          rmaxx = warp_max(rmaxx); rminx = warp_min(rminx); rmaxz = warp_max(rmaxz); rminz = warp_min(rminz);
          const double top_y = fmin(lmaxx, rmaxx), bot_y = fmax(lminx, rminx);
          const double top_z = fmin(lmaxz, rmaxz), bot_z = fmax(lminz, rminz);
          int nl = 0, nr = 0;
          auto visitE = [&](double x, double y, double z, int pidx, int wgt) {
            double ldot, rdot;
            dots(pidx, ldot, rdot);
            const bool inw = x >= bot_y && x <= top_y && z >= bot_z && z <= top_z;
            if (ldot > P.cosf && y < min_x && inw) nl += wgt;
            if (rdot > P.cosf && y > max_x && inw) nr += wgt;
          };
          warp_scan_ball(P, cl, sr, q, P.r2_hs, [&](const float4 &p) {
            double x, y, z;
            to_frame(R, (double)p.x - smp[0], (double)p.y - smp[1], (double)p.z - smp[2], x, y, z);
            if (z > -1.0 * hh && z < hh && in_region(x, y)) visitE(x, y, z, __float_as_int(p.w), 1);
          });
          if (nb_in && lane == 0) visitE(x0, y0, z0, nb0, npad);
          nl = warp_sum(nl);
          nr = warp_sum(nr);
          full = nl >= P.min_viable && nr >= P.min_viable;
        }
        if (full) label = 1;
      }
    }
  }
  if (lane == 0) {
    h.half_antipodal = half ? 1 : 0;
    h.full_antipodal = full ? 1 : 0;
    labels[i] = label;
  }
}

// ------------------------------------------------------------------------------------------------
// Clustering::findClusters (clustering.cpp:5-105, remove_inliers = false): hand i becomes a cluster when at least
// min_inliers OTHER hands have an axis within 12 degrees, a position within 5 cm and an axis-orthogonal offset within
// 5 mm; cluster position = mean inlier position, score = lower bound of the 99 % confidence interval of the inlier
// scores (Welford update in index order, :62-70). One warp per hand: the lanes test 32 hands j at a time, lane 0 folds
// the inliers of the ballot IN INDEX ORDER, so the float64 running mean / variance are the reference's sequential ones.
// ------------------------------------------------------------------------------------------------
__global__ void k_clusters(const gpdb_pose *__restrict__ hands, int n, int min_inliers, double cos_thresh,
                           gpdb_pose *__restrict__ out, uint8_t *__restrict__ keep) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  const double AXIS_ALIGN_DIST_THRESH = 0.005, MAX_DIST_THRESH = 0.05;
  const double ai[3] = {hands[i].frame[6], hands[i].frame[7], hands[i].frame[8]};
  const double pi[3] = {hands[i].position[0], hands[i].position[1], hands[i].position[2]};
  double outer[3][3];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) outer[r][c] = ai[r] * ai[c];
  int num_inliers = 0;
  double pd[3] = {0.0, 0.0, 0.0}, mean = 0.0, sd = 0.0;
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int j = j0 + lane;
    bool inl = false;
    if (j < n && j != i) {
      const double aj[3] = {hands[j].frame[6], hands[j].frame[7], hands[j].frame[8]};
      const double axis_aligned = ai[0] * aj[0] + ai[1] * aj[1] + ai[2] * aj[2];
      const double d[3] = {pi[0] - hands[j].position[0], pi[1] - hands[j].position[1], pi[2] - hands[j].position[2]};
      double proj[3];
#pragma unroll
      for (int r = 0; r < 3; r++)
        proj[r] = ((r == 0 ? 1.0 : 0.0) - outer[r][0]) * d[0] + ((r == 1 ? 1.0 : 0.0) - outer[r][1]) * d[1] +
                  ((r == 2 ? 1.0 : 0.0) - outer[r][2]) * d[2];
      inl = fabs(axis_aligned) > cos_thresh && sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) <= MAX_DIST_THRESH &&
            sqrt(proj[0] * proj[0] + proj[1] * proj[1] + proj[2] * proj[2]) <= AXIS_ALIGN_DIST_THRESH;
    }
    unsigned m = __ballot_sync(0xffffffffu, inl);
    if (lane == 0)
      while (m) {
        const int jj = j0 + __ffs(m) - 1;
        m &= m - 1;
        num_inliers++;
#pragma unroll
        for (int r = 0; r < 3; r++) pd[r] += hands[jj].position[r];
        const double old_mean = mean, sj = (double)hands[jj].score;
        mean += (sj - mean) / (double)num_inliers;
        sd += (sj - mean) * (sj - old_mean);
      }
  }
  if (lane == 0) {
    gpdb_pose o = hands[i];
    uint8_t k = 0;
    if (num_inliers >= min_inliers) {
      const double dn = (double)num_inliers;
#pragma unroll
      for (int r = 0; r < 3; r++) pd[r] = pd[r] / dn - pi[r];
      sd /= dn;
      if (sd != 0) sd = sqrt(sd);
      const double conf_lb = mean - 2.576 * sd / sqrt((double)num_inliers);
#pragma unroll
      for (int r = 0; r < 3; r++) o.position[r] = pi[r] + pd[r];
      o.score = (float)conf_lb;
      k = 3;  // VALID | FILTERED: geo_compact keeps it
    }
    out[i] = o;
    keep[i] = k;
  }
}

// ------------------------------------------------------------------------------------------------
// k_images (round 2): one CTA per grasp image.
//
// Output stage. The reference post-processes every channel group as cv::dilate(3x3) -> cv::normalize(NORM_MINMAX over the
// group's channels) -> convertTo(CV_8U, 255) (image_strategy.cpp:145-153,179-187,222-230). The affine map + rounding is
// monotone non-decreasing, so it commutes with the max filter: the kernel quantises the OCCUPIED cells only (a few
// hundred per projection) into uint8 planes in shared memory and dilates the bytes afterwards with packed 4-pixel SIMD
// (__vmaxu4), while assembling the pixels. What the normalisation needs from the dilated float image is its max (= the
// max over the occupied cells: dilation moves maxima, it does not change them) and its min, which is 0 whenever the
// image has an all-empty 3x3 window (every value is >= 0 and the background is 0); the occupancy bitmap decides that
// with a few word operations. Images without such a window (fully covered: very dense clouds) take a general path that
// evaluates the min over the dilated float image explicitly (points_minmax_general / dilated_min), so the result is
// the reference's for every input.
//
// Pixel layout written to HBM ("P16"): one 16-byte group per pixel = channels 0..C-1 in the reference's order, bytes
// C..15 zero, pixels row-major — exactly the K-chunk the tcgen05 conv1 reads (lenet_tc.cu), so the classifier
// bulk-copies an image straight into its operand plane; k_p16_to_hwc produces the cv::Mat layout for callers that
// want the images themselves (gpdb_images, keep_images).
// ------------------------------------------------------------------------------------------------
struct ImgSmem {
  SegScan<NT_IMG> seg;
  gpdb_pose h;
  double red[NT_IMG / 32][4];
  double center[3];
  double sv[GPDB_MAX_CAMERAS][3];
  double svh[GPDB_MAX_CAMERAS][3];
  unsigned occf[MAXPIX / 32 + 2];  // occupancy of the image cells, flat (bit = pixel index), written by warp ballots
  unsigned lcgA[GPDB_MAX_NSP], lcgC[GPDB_MAX_NSP];  // LCG skip-ahead tables (DevParams), staged once per CTA
  int cam_or;
  int n_img;
  int box_n;
  int wl_n;
  int ball_n;                     // in-ball points recorded by scan 1 (positions in the cell-sorted array)
  int dl_n;                       // shadow draws that passed the window test (draw list)
  int bm_org[3], bm_dims[3];
  float fred[NT_IMG / 32][8];
};

__device__ __forceinline__ bool in_image_box(const DevParams &P, const gpdb_pose &h, double x, double y, double z) {
  const double half_od = P.vol_w / 2.0;
  return (x > h.bottom) && (x < h.bottom + P.vol_d) && (y > h.center - half_od) && (y < h.center + half_od) &&
         (z > -1.0 * P.vol_h) && (z < P.vol_h);
}
// unit coordinate + cell of one axis without the two float64 divisions of the reference formulas
// ((v - lo) / extent, floor(u / (1.0 / S))) in the common case: a reciprocal multiply gives u to ~2 ulp, and the cell is
// floor(u * S) unless u * S lies within 1e-9 of an integer — only then are the exact divisions evaluated, so the
// cell index is always the reference's. u feeds the per-cell MEAN only (float32 result, 2 ulp of float64 are invisible).
__device__ __forceinline__ void unit_axis(double v, double lo, double extent, double inv_extent, int S, double &u, int &cell) {
  u = (v - lo) * inv_extent;
  double q = u * (double)S;
  double fq = floor(q);
  const double fr = q - fq;  // in [0, 1): exact (Sterbenz) for q >= 1
  if (fr < 1e-9 || fr > 1.0 - 1e-9) {
    u = (v - lo) / extent;
    double cellsize = 1.0 / (double)S;
    fq = floor(u / cellsize);
  }
  cell = min((int)fq, S - 1);
}
__device__ __forceinline__ unsigned unit_q32(double u) {
  return __double2uint_rz(u * 4294967296.0);  // cvt.rzi.u32.f64 saturates: the same result as clamping to [0, 2^32 - 1] first
}

// mean of a cell from its packed accumulator (count << 48 | fixed-point sum, 32 fractional bits): sum / (count 2^32). The
// per-count reciprocals 1 / (c 2^32), c < RCP_N, are tabulated once per image in the (then idle) reduction scratch of the
// ball scan: one multiplication instead of a float64 division per occupied cell (count 1: an exact scaling either way).
constexpr int RCP_N = (NT_IMG / 32) * 4;
__device__ __forceinline__ double cell_mean(unsigned long long acc, const double *rcp) {
  const unsigned c = (unsigned)(acc >> 48);
  const double sum = (double)(acc & 0xffffffffffffull);
  return c < (unsigned)RCP_N ? sum * rcp[c] : sum / ((double)c * 4294967296.0);
}

__device__ __forceinline__ bool fully_covered(const unsigned *occf, int S);
// block-wide reduction of up to eight floats with max (use negated values for min). Contains two barriers. With `occf`
// non-null, warp 0 also evaluates fully_covered(occf) between the barriers (the occupancy words were written before the
// call) and every thread receives the verdict in *covered — one evaluation per CTA, no extra barrier.
template <int NT, int NV>
__device__ __forceinline__ void block_max(float (&v)[NV], float (*red)[8], const unsigned *occf = nullptr, int S = 0,
                                          bool *covered = nullptr) {
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int i = 0; i < NV; i++) red[threadIdx.x >> 5][i] = v[i];
  if (occf && threadIdx.x < 32) {
    const bool c = fully_covered(occf, S);
    if (threadIdx.x == 0) red[0][7] = c ? 1.0f : 0.0f;  // slot 7 is never used by a reduction (NV <= 6)
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = red[0][i];
  for (int w = 1; w < NT / 32; w++)
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = fmaxf(v[i], red[w][i]);
  if (covered) *covered = red[0][7] != 0.0f;
}

// cv::normalize(NORM_MINMAX, 0..1) + convertTo(CV_8U, 255) of one value, as OpenCV evaluates it: scale / shift in double,
// cast to float, one fused multiply-add, round-half-even, saturate.
struct Quant {
  float a, b;
  __device__ __forceinline__ Quant(float mn, float mx) {
    const double smin = (double)mn, smax = (double)mx;
    const double scale = (1.0 - 0.0) * (smax - smin > DBL_EPSILON ? 1.0 / (smax - smin) : 0.0);
    const double shift = 0.0 - smin * scale;
    a = (float)scale;
    b = (float)shift;
  }
  __device__ __forceinline__ unsigned operator()(float v) const {
    const float t = fmaf(v, a, b);
    const int q = __float2int_rn(t * 255.0f);
    return (unsigned)min(max(q, 0), 255);
  }
};

// true when the S x S occupancy has NO all-empty 3x3 window. occf: flat bitmap (bit = row * S + col) followed by two
// zero-readable words. One warp (32 rows per pass, a few word operations per row); block_max distributes the verdict.
__device__ __forceinline__ bool fully_covered(const unsigned *occf, int S) {
  const int lane = threadIdx.x & 31;
  const unsigned long long mask = S >= 64 ? ~0ull : ((1ull << S) - 1ull);
  auto row_bits = [&](int r) -> unsigned long long {
    if (r < 0 || r >= S) return 0ull;
    const int b = r * S, w = b >> 5, sh = b & 31;
    unsigned long long x = ((unsigned long long)occf[w] | ((unsigned long long)occf[w + 1] << 32)) >> sh;
    if (sh + S > 64) x |= (unsigned long long)occf[w + 2] << (64 - sh);
    return x & mask;
  };
  bool empty = false;
  for (int r = lane; r < S; r += 32) {
    const unsigned long long o = row_bits(r - 1) | row_bits(r) | row_bits(r + 1);
    const unsigned long long d = (o | (o << 1) | (o >> 1)) & mask;
    empty = empty || (d != mask);
  }
  return !__any_sync(0xffffffffu, empty);
}
// one bit per cell from a per-thread predicate over pix = tid + t * NT (warp-aligned): word (pix >> 5) of the flat bitmap
template <int NT>
__device__ __forceinline__ void occ_ballot(unsigned *occf, int t, bool occupied) {
  const unsigned word = __ballot_sync(0xffffffffu, occupied);
  if ((threadIdx.x & 31) == 0) occf[(threadIdx.x >> 5) + t * (NT / 32)] = word;
}

// min over the 3x3-dilated (border ignored) image of channel plane F[SS] (float, row-major): general path only
template <int NT>
__device__ __noinline__ float dilated_min(const float *F, int S) {
  float mn = FLT_MAX;
  for (int pix = threadIdx.x; pix < S * S; pix += NT) {
    const int r = pix / S, c = pix - r * S;
    float m = -FLT_MAX;
    for (int dr = -1; dr <= 1; dr++)
      for (int dc = -1; dc <= 1; dc++) {
        const int rr = r + dr, cc = c + dc;
        if (rr >= 0 && rr < S && cc >= 0 && cc < S) m = fmaxf(m, F[rr * S + cc]);
      }
    mn = fminf(mn, m);
  }
  return mn;
}

// optional phase timing (development aid, gpdb_debug_phase_cycles): thread 0 accumulates clock64() deltas
#define PHASE(i)                                                        \
  do {                                                                  \
    if (prof && threadIdx.x == 0) {                                     \
      long long now__ = clock64();                                      \
      if ((i) > 1) atomicAdd(prof + (i), (unsigned long long)(now__ - t_phase)); \
      t_phase = now__;                                                  \
    }                                                                   \
  } while (0)

// dynamic shared memory: planes C x S x RS bytes (RS = S rounded up to 4) | tiles 3 x 8 S S | box list 36 B x BOX_CAP
// (the shadow bitmaps + voxel list alias the box list; the shadow work list aliases the tiles)
// GL = true is the last tier: the box list (and the float images of the general min path) live in a per-CTA slice of global
// memory (`gl_base`, `gl_cap` points: L2-resident scratch) instead of shared memory, for clouds so dense that an image box
// holds more than BOX_CAP points (the reference has no limit; image_generator.cpp:54-64).
template <int S_T, bool GL>
__global__ void __launch_bounds__(NT_IMG, 1) k_images(const DevParams *Pp, DevCloud cl, const gpdb_pose *cand, int nc,
                                                      uint8_t *p16, const double *qtab, int *err, int plane_bytes,
                                                      int list_bytes, unsigned long long *prof, const int *work,
                                                      const int *work_n, unsigned char *gl_base, int gl_cap, int *ovf2,
                                                      int *ovf2_count) {
  // work != nullptr: only the images work[0 .. *work_n) (the overflow list of the previous tier); ovf2 != nullptr: images
  // whose box list overflows THIS tier are appended there (and redone by the next one) instead of being an error
  long long t_phase = 0;
  const DevParams &P = *Pp;
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ ImgSmem sm;
  const int S = S_T > 0 ? S_T : P.S, C = P.C, SS = S * S;
  const int RS = (S + 3) & ~3, RW = RS >> 2, PLB = S * RS;
  uint8_t *planes = dyn;
  unsigned long long *tileA = reinterpret_cast<unsigned long long *>(dyn + plane_bytes);
  unsigned long long *tileB = tileA + SS;
  unsigned long long *tileC = tileB + SS;
  unsigned char *lbase = reinterpret_cast<unsigned char *>(tileC + SS);
  const int BC = GL ? gl_cap : BOX_CAP;
  const int SS_ = (S_T > 0 ? S_T : P.S) * (S_T > 0 ? S_T : P.S);
  unsigned char *list_mem = GL ? gl_base + (size_t)blockIdx.x * ((size_t)gl_cap * 36 + (size_t)16 * SS_) : lbase;
  unsigned long long *bkeys = reinterpret_cast<unsigned long long *>(list_mem);
  unsigned *bq = reinterpret_cast<unsigned *>(bkeys + BC);        // [3][CAP]
  unsigned *bcell = bq + 3 * BC;                                  // packed 3 x 8 bit
  float *bnrm = reinterpret_cast<float *>(bcell + BC);            // [3][CAP]
  float *gF = reinterpret_cast<float *>(bnrm + 3 * BC);           // GL: 4 float images of the general min path
  unsigned *bitmap = reinterpret_cast<unsigned *>(lbase);         // aliases the list (shadow phase)
  // scan 1 records WHERE the in-ball points are (position in the cell-sorted array, 4 B each) in tile C, which nothing
  // else touches before the shadow phase: the shadow casting re-reads the neighbourhood as independent loads from that
  // list (one L2 round trip) instead of walking the grid again (cell bounds -> prefix scan -> search -> points).
  int *ball = reinterpret_cast<int *>(tileC);
  const int BALL_CAP = 2 * SS;  // 8 S S bytes / 4
  const int tid = threadIdx.x, lane = tid & 31;
  const int nproj = (C >= 12) ? 3 : 1;
  const int per = (C == 15) ? 5 : 4;
  const bool do_nrm = C != 1, do_dep = C == 1 || C >= 12;
  constexpr int PIXT = (MAXPIX + NT_IMG - 1) / NT_IMG;
  for (int k = tid; k < GPDB_MAX_NSP; k += NT_IMG) {
    sm.lcgA[k] = P.lcgA[k];
    sm.lcgC[k] = P.lcgC[k];
  }
  for (int k = tid; k < MAXPIX / 32 + 2; k += NT_IMG) sm.occf[k] = 0u;

  const int n_work = work ? *work_n : nc;
  for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    const int b = work ? work[wi] : wi;
    __syncthreads();
    {
      const int *src = reinterpret_cast<const int *>(cand + b);
      int *dst = reinterpret_cast<int *>(&sm.h);
      for (int k = tid; k < (int)(sizeof(gpdb_pose) / 4); k += NT_IMG) dst[k] = src[k];
    }
    if (tid == 0) {
      sm.cam_or = 0;
      sm.n_img = 0;
      sm.box_n = 0;
      sm.ball_n = 0;
    }
    for (int k = tid; k < plane_bytes >> 4; k += NT_IMG) reinterpret_cast<uint4 *>(planes)[k] = make_uint4(0, 0, 0, 0);
    for (int k = tid; k < SS; k += NT_IMG) reinterpret_cast<uint4 *>(tileA)[k] = make_uint4(0, 0, 0, 0);  // tiles A, B: clean
    __syncthreads();
    PHASE(1);   // image start
    const bool need_cam = (C == 15) && !P.all_seen;  // the camera set of the neighbourhood is only read by the shadow
    const gpdb_pose &h = sm.h;
    const double inv_d = 1.0 / P.vol_d, inv_w = 1.0 / P.vol_w, inv_h = 1.0 / (2.0 * P.vol_h);
    float q[3] = {(float)h.sample[0], (float)h.sample[1], (float)h.sample[2]};
    SegRange sr = seg_range(P, q, P.rf_img);
    // ---- ball scan 1: neighbourhood centre + camera set (HandSet::calculateShadow, hand_set.cpp:131-136)
    //      and the list of points inside the image box (ImageStrategy::transformToUnitImage)
    double sx = 0, sy = 0, sz = 0;
    int cnt = 0, cam_or = 0;
    scan_balanced<NT_IMG>(P, cl, sr, sm.seg, [&](bool in, const float4 &p, int where) {
      // the scan only APPENDS the raw box points (few lanes qualify: doing the per-point work here would run it at
      // ~6 % lane utilisation); unit coordinates, cells and normals are computed densely after the scan
      bool inb = false, inball = false;
      unsigned long long key = 0;
      if (in) {
        float d = l2_simple(q, p.x, p.y, p.z);
        if (d < P.r2_img) {
          inball = true;
          const int idx = __float_as_int(p.w);
          sx += (double)p.x;
          sy += (double)p.y;
          sz += (double)p.z;
          cnt++;
          if (need_cam) cam_or |= cl.cam[idx];
          double x, y, z;
          to_frame(h.frame, (double)p.x - h.sample[0], (double)p.y - h.sample[1], (double)p.z - h.sample[2], x, y, z);
          if (in_image_box(P, h, x, y, z)) {
            inb = true;
            key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
          }
        }
      }
      if (C == 15) {  // remember where the neighbourhood lives (warp-aggregated append)
        const unsigned mb = __ballot_sync(0xffffffffu, inball);
        if (mb) {
          const int leader = __ffs(mb) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(&sm.ball_n, __popc(mb));
          base = __shfl_sync(0xffffffffu, base, leader);
          const int pos = base + __popc(mb & ((1u << lane) - 1));
          if (inball && pos < BALL_CAP) ball[pos] = where;
        }
      }
      unsigned mk = __ballot_sync(0xffffffffu, inb);
      if (mk) {
        int leader = __ffs(mk) - 1, base = 0;
        if (lane == leader) base = atomicAdd(&sm.box_n, __popc(mk));
        base = __shfl_sync(0xffffffffu, base, leader);
        int pos = base + __popc(mk & ((1u << lane) - 1));
        if (inb && pos < BC) {
          bkeys[pos] = key;
          bq[pos] = __float_as_uint(p.x);  // raw coordinates, replaced by the fixed-point unit coordinates below
          bq[BC + pos] = __float_as_uint(p.y);
          bq[2 * BC + pos] = __float_as_uint(p.z);
        }
      }
    });
    // block reduce centre sums
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    cnt = warp_sum(cnt);
    cam_or = __reduce_or_sync(0xffffffffu, (unsigned)cam_or);
    __syncthreads();
    if (lane == 0) {
      sm.red[tid >> 5][0] = sx;
      sm.red[tid >> 5][1] = sy;
      sm.red[tid >> 5][2] = sz;
      atomicAdd(&sm.n_img, cnt);
      atomicOr(&sm.cam_or, cam_or);
    }
    __syncthreads();
    if (tid == 0) {
      double a0 = 0, a1 = 0, a2 = 0;
      for (int w = 0; w < NT_IMG / 32; w++) {
        a0 += sm.red[w][0];
        a1 += sm.red[w][1];
        a2 += sm.red[w][2];
      }
      double nn = (double)sm.n_img;
      if (!need_cam && sm.n_img > 0) sm.cam_or = (1 << P.K) - 1;  // every point is seen by every camera (set at upload)
      sm.center[0] = a0 / nn;
      sm.center[1] = a1 / nn;
      sm.center[2] = a2 / nn;
      if (sm.box_n > BC) {
        if (ovf2) ovf2[atomicAdd(ovf2_count, 1)] = b;  // redone by the next tier (larger list)
        else {
          atomicAdd(err + 2, 1);
          sm.box_n = BC;
        }
      }
    }
    __syncthreads();
    PHASE(2);  // scan 1 + reductions done
    if (sm.box_n > BC) continue;  // handed to the next tier (uniform; without a next tier box_n was clamped)
    const int bn = sm.box_n;
    double *rcp = &sm.red[0][0];  // the scan's reduction scratch is idle from here to the next image: reciprocal table (cell_mean)
    if (tid < RCP_N) rcp[tid] = 1.0 / ((double)tid * 4294967296.0);  // visible after the barrier that ends the box-point pass
    // dense pass over the box points: hand-frame coordinates -> unit cube, cell indices, |R^T n| (all lanes busy)
    for (int k = tid; k < bn; k += NT_IMG) {
      const double px = (double)__uint_as_float(bq[k]), py = (double)__uint_as_float(bq[BC + k]),
                   pz = (double)__uint_as_float(bq[2 * BC + k]);
      double x, y, z, u0, u1, u2;
      int c0, c1, c2;
      to_frame(h.frame, px - h.sample[0], py - h.sample[1], pz - h.sample[2], x, y, z);
      unit_axis(x, h.bottom, P.vol_d, inv_d, S, u0, c0);
      unit_axis(y, h.center - P.vol_w / 2.0, P.vol_w, inv_w, S, u1, c1);
      unit_axis(z, -P.vol_h, 2.0 * P.vol_h, inv_h, S, u2, c2);
      bq[k] = unit_q32(u0);
      bq[BC + k] = unit_q32(u1);
      bq[2 * BC + k] = unit_q32(u2);
      bcell[k] = (unsigned)c0 | ((unsigned)c1 << 8) | ((unsigned)c2 << 16);
      const double *nn = cl.nrm + 3 * (size_t)(unsigned)(bkeys[k] & 0xffffffffull);
      double n0, n1, n2;
      to_frame(h.frame, nn[0], nn[1], nn[2], n0, n1, n2);
      bnrm[k] = (float)fabs(n0);
      bnrm[BC + k] = (float)fabs(n1);
      bnrm[2 * BC + k] = (float)fabs(n2);
    }
    __syncthreads();

    // ---- points phase: per projection rasterise normals (arg-max key = last writer in (dist, index)
    // order, createNormalsImage :124-143) and depth (per-cell mean, createDepthImage :158-176)
    for (int pj = 0; pj < nproj; pj++) {
      // coordinate orders (x,y,z), (z,y,x), (z,x,y): rows cumulatively swapped {0<->2}, {1<->2}
      // (image_15_channels_strategy.cpp:57-64)
      const int a0 = (pj == 0) ? 0 : 2, a1 = (pj == 2) ? 0 : 1, a2 = (pj == 0) ? 2 : (pj == 1 ? 0 : 1);
      // tiles A and B are clean here (zeroed at image start; every projection wipes the cells it touched), occf likewise
      for (int k = tid; k < bn; k += NT_IMG) {
        const unsigned cc = bcell[k];
        const int row = S - 1 - (int)((cc >> (8 * a0)) & 255), col = (cc >> (8 * a1)) & 255;
        const int pix = row * S + col;
        atomicMax(tileA + pix, bkeys[k]);
        atomicAdd(tileB + pix, (1ull << 48) + (unsigned long long)bq[a2 * BC + k]);
        atomicOr(&sm.occf[pix >> 5], 1u << (pix & 31));
      }
      __syncthreads();
      // the winner of a cell (its point with the largest key) carries the cell's values: |n| of that point, 1 - mean depth
      auto cell_values = [&](int k, int &row, int &col, float &n0, float &n1, float &n2, float &dv) -> bool {
        const unsigned cc = bcell[k];
        row = S - 1 - (int)((cc >> (8 * a0)) & 255);
        col = (cc >> (8 * a1)) & 255;
        const int pix = row * S + col;
        if (tileA[pix] != bkeys[k]) return false;
        n0 = bnrm[k];
        n1 = bnrm[BC + k];
        n2 = bnrm[2 * BC + k];
        const float avg = (float)cell_mean(tileB[pix], rcp);
        dv = (float)(1.0 - (double)avg);
        return true;
      };
      float mxv[2] = {0.0f, 0.0f};  // max of the normals group / of the depth channel (all values are >= 0)
      for (int k = tid; k < bn; k += NT_IMG) {
        int row, col;
        float n0, n1, n2, dv;
        if (cell_values(k, row, col, n0, n1, n2, dv)) {
          mxv[0] = fmaxf(mxv[0], fmaxf(fmaxf(n0, n1), n2));
          mxv[1] = fmaxf(mxv[1], dv);
        }
      }
      bool covered_pj;
      block_max<NT_IMG, 2>(mxv, sm.fred, sm.occf, S, &covered_pj);  // (its barriers publish the occupancy words)
      float mnv[2] = {0.0f, 0.0f};  // min over the dilated image: 0 when an all-empty 3x3 window exists
      if (covered_pj) {
        if constexpr (GL) {
          // general path, global-list tier: the float images live in the CTA's global slice, the tiles stay intact and the
          // winners are simply re-derived (cell_values) — no per-thread value cache sized by the list capacity
          float *F = gF;
          for (int k = tid; k < SS; k += NT_IMG) reinterpret_cast<uint4 *>(F)[k] = make_uint4(0, 0, 0, 0);
          __syncthreads();
          for (int k = tid; k < bn; k += NT_IMG) {
            int row, col;
            float v[4];
            if (cell_values(k, row, col, v[0], v[1], v[2], v[3]))
#pragma unroll
              for (int c = 0; c < 4; c++) F[c * SS + row * S + col] = v[c];
          }
          __syncthreads();
          float neg[2];
          neg[0] = -fminf(fminf(dilated_min<NT_IMG>(F, S), dilated_min<NT_IMG>(F + SS, S)), dilated_min<NT_IMG>(F + 2 * SS, S));
          neg[1] = -dilated_min<NT_IMG>(F + 3 * SS, S);
          block_max<NT_IMG, 2>(neg, sm.fred);
          mnv[0] = -neg[0];
          mnv[1] = -neg[1];
          const Quant qn(mnv[0], mxv[0]), qd(mnv[1], mxv[1]);
          const unsigned bg_n = qn(0.0f) * 0x01010101u, bg_d = qd(0.0f) * 0x01010101u;
          const int cb = (C == 1) ? 0 : pj * per;
          for (int k = tid; k < (PLB >> 2); k += NT_IMG) {
            if (do_nrm)
#pragma unroll
              for (int c = 0; c < 3; c++) reinterpret_cast<unsigned *>(planes + (size_t)(cb + c) * PLB)[k] = bg_n;
            if (do_dep) reinterpret_cast<unsigned *>(planes + (size_t)(cb + (C == 1 ? 0 : 3)) * PLB)[k] = bg_d;
          }
          __syncthreads();
          for (int k = tid; k < bn; k += NT_IMG) {
            int row, col;
            float n0, n1, n2, dv;
            if (cell_values(k, row, col, n0, n1, n2, dv)) {
              const int o = row * RS + col;
              if (do_nrm) {
                planes[(size_t)(cb + 0) * PLB + o] = (uint8_t)qn(n0);
                planes[(size_t)(cb + 1) * PLB + o] = (uint8_t)qn(n1);
                planes[(size_t)(cb + 2) * PLB + o] = (uint8_t)qn(n2);
              }
              if (do_dep) planes[(size_t)(cb + (C == 1 ? 0 : 3)) * PLB + o] = (uint8_t)qd(dv);
            }
          }
          __syncthreads();
          for (int k = tid; k < SS; k += NT_IMG) reinterpret_cast<uint4 *>(tileA)[k] = make_uint4(0, 0, 0, 0);
          for (int k = tid; k < MAXPIX / 32; k += NT_IMG) sm.occf[k] = 0u;
        } else {
        // general path (no empty window): materialise the four float channel images over the (now dead) tiles and take the
          // min of their dilations. Winners keep their values in registers across the rewrite of the tiles.
          constexpr int JMAX = (BOX_CAP + NT_IMG - 1) / NT_IMG;
          float wv[JMAX][4];
          int wpix[JMAX];
  #pragma unroll
          for (int j = 0; j < JMAX; j++) {
            const int k = tid + j * NT_IMG;
            wpix[j] = -1;
            int row, col;
            if (k < bn && cell_values(k, row, col, wv[j][0], wv[j][1], wv[j][2], wv[j][3])) wpix[j] = row * S + col;
          }
          __syncthreads();
          float *F = reinterpret_cast<float *>(tileA);  // 4 x SS floats = tileA + tileB
          for (int k = tid; k < SS; k += NT_IMG) reinterpret_cast<uint4 *>(F)[k] = make_uint4(0, 0, 0, 0);
          __syncthreads();
  #pragma unroll
          for (int j = 0; j < JMAX; j++)
            if (wpix[j] >= 0)
  #pragma unroll
              for (int c = 0; c < 4; c++) F[c * SS + wpix[j]] = wv[j][c];
          __syncthreads();
          float neg[2];
          neg[0] = -fminf(fminf(dilated_min<NT_IMG>(F, S), dilated_min<NT_IMG>(F + SS, S)), dilated_min<NT_IMG>(F + 2 * SS, S));
          neg[1] = -dilated_min<NT_IMG>(F + 3 * SS, S);
          block_max<NT_IMG, 2>(neg, sm.fred);
          mnv[0] = -neg[0];
          mnv[1] = -neg[1];
          const Quant qn(mnv[0], mxv[0]), qd(mnv[1], mxv[1]);
          const unsigned bg_n = qn(0.0f) * 0x01010101u, bg_d = qd(0.0f) * 0x01010101u;  // background = quantised 0
          const int cb = (C == 1) ? 0 : pj * per;
          for (int k = tid; k < (PLB >> 2); k += NT_IMG) {
            if (do_nrm)
  #pragma unroll
              for (int c = 0; c < 3; c++) reinterpret_cast<unsigned *>(planes + (size_t)(cb + c) * PLB)[k] = bg_n;
            if (do_dep) reinterpret_cast<unsigned *>(planes + (size_t)(cb + (C == 1 ? 0 : 3)) * PLB)[k] = bg_d;
          }
          __syncthreads();
  #pragma unroll
          for (int j = 0; j < JMAX; j++)
            if (wpix[j] >= 0) {
              const int row = wpix[j] / S, col = wpix[j] - row * S, o = row * RS + col;
              if (do_nrm) {
                planes[(size_t)(cb + 0) * PLB + o] = (uint8_t)qn(wv[j][0]);
                planes[(size_t)(cb + 1) * PLB + o] = (uint8_t)qn(wv[j][1]);
                planes[(size_t)(cb + 2) * PLB + o] = (uint8_t)qn(wv[j][2]);
              }
              if (do_dep) planes[(size_t)(cb + (C == 1 ? 0 : 3)) * PLB + o] = (uint8_t)qd(wv[j][3]);
            }
          for (int k = tid; k < SS; k += NT_IMG) reinterpret_cast<uint4 *>(tileA)[k] = make_uint4(0, 0, 0, 0);  // float images -> clean
          for (int k = tid; k < MAXPIX / 32; k += NT_IMG) sm.occf[k] = 0u;
        }
      } else {
        const Quant qn(0.0f, mxv[0]), qd(0.0f, mxv[1]);
        const int cb = (C == 1) ? 0 : pj * per;
        for (int k = tid; k < bn; k += NT_IMG) {
          int row, col;
          float n0, n1, n2, dv;
          if (cell_values(k, row, col, n0, n1, n2, dv)) {
            const int o = row * RS + col;
            if (do_nrm) {
              planes[(size_t)(cb + 0) * PLB + o] = (uint8_t)qn(n0);
              planes[(size_t)(cb + 1) * PLB + o] = (uint8_t)qn(n1);
              planes[(size_t)(cb + 2) * PLB + o] = (uint8_t)qn(n2);
            }
            if (do_dep) planes[(size_t)(cb + (C == 1 ? 0 : 3)) * PLB + o] = (uint8_t)qd(dv);
          }
        }
        __syncthreads();  // every winner has read its cell: wipe the touched cells for the next projection
        for (int k = tid; k < bn; k += NT_IMG) {
          const unsigned cc = bcell[k];
          const int pix = (S - 1 - (int)((cc >> (8 * a0)) & 255)) * S + (int)((cc >> (8 * a1)) & 255);
          tileA[pix] = 0ull;
          tileB[pix] = 0ull;
          sm.occf[pix >> 5] = 0u;
        }
      }
      __syncthreads();
    }

    PHASE(3);  // points phase (3 projections) done
    // ---- shadow phase (15 channels): HandSet::calculateShadow, deterministic variant
    if (C == 15) {
      const int K = P.K;
      const int bmd = P.bm_dim;
      const int bm_words = 2 * bmd * bmd;  // rows of 64 bits along x (bm_dim <= 64)
      const double gmax = qtab[GPDB_QTAB_SIZE - 1];
      const double voxel = GPDB_SHADOW_VOXEL;
      if (tid == 0) {
        // AABB (in voxel indices) of the image box, widened by the largest jitter
        const double jmax = gmax * voxel * 0.3 + 1e-9;
        const double half_od = P.vol_w / 2.0;
        double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int cr = 0; cr < 8; cr++) {
          double cx = (cr & 1) ? h.bottom + P.vol_d : h.bottom;
          double cy = (cr & 2) ? h.center + half_od : h.center - half_od;
          double cz = (cr & 4) ? P.vol_h : -P.vol_h;
          for (int r = 0; r < 3; r++) {
            double wv = h.frame[r] * cx + h.frame[3 + r] * cy + h.frame[6 + r] * cz + h.sample[r];
            mn[r] = fmin(mn[r], wv);
            mx[r] = fmax(mx[r], wv);
          }
        }
        for (int r = 0; r < 3; r++) {
          int lo = (int)floor((mn[r] - jmax) * P.vox_mult) - 1;
          int hi = (int)floor((mx[r] + jmax) * P.vox_mult) + 1;
          sm.bm_org[r] = lo;
          sm.bm_dims[r] = min(hi - lo + 1, bmd);
        }
      }
      if (tid < K) {
        // shadow_vec = shadow_length * (center - view_point) / norm (hand_set.cpp:146-150)
        double s0 = sm.center[0] - P.vp[tid][0], s1 = sm.center[1] - P.vp[tid][1], s2 = sm.center[2] - P.vp[tid][2];
        double nn = sqrt((s0 * s0 + s1 * s1) + s2 * s2);
        sm.sv[tid][0] = P.shadow_length * s0 / nn;
        sm.sv[tid][1] = P.shadow_length * s1 / nn;
        sm.sv[tid][2] = P.shadow_length * s2 / nn;
        to_frame(h.frame, sm.sv[tid][0], sm.sv[tid][1], sm.sv[tid][2], sm.svh[tid][0], sm.svh[tid][1], sm.svh[tid][2]);
      }
      for (int k = tid; k < bm_words * K; k += NT_IMG) bitmap[k] = 0u;
      __syncthreads();
      PHASE(4);  // shadow setup done
      const int o0 = sm.bm_org[0], o1 = sm.bm_org[1], o2 = sm.bm_org[2];
      const int d0 = sm.bm_dims[0], d1 = sm.bm_dims[1], d2 = sm.bm_dims[2];
      const double mxu = 1.0 / 32767.0;
      // voxel -> jittered point -> hand frame -> inside the image box?
      auto voxel_point_in_box = [&](int v0, int v1, int v2, double &x, double &y, double &z) -> bool {
        double g = qtab[gpdb_voxel_hash(v0, v1, v2) & (GPDB_QTAB_SIZE - 1)];
        double jit = 1.0 * g * voxel * 0.3;
        double w0 = (double)v0 * voxel + jit, w1 = (double)v1 * voxel + jit, w2 = (double)v2 * voxel + jit;
        to_frame(h.frame, w0 - h.sample[0], w1 - h.sample[1], w2 - h.sample[2], x, y, z);
        return in_image_box(P, h, x, y, z);
      };
      const int cam_set = sm.cam_or;
      // Shadow casting is load-balanced in two steps per camera: (1) the ball scan appends the points whose
      // shadow segment can reach the bitmap to a work list (float4: x, y, z, LCG seed) living in the idle tile
      // region; (2) the (point, draw) pairs are spread evenly over all threads — draw t of a point comes from
      // the closed-form LCG skip-ahead seed_t = A^(t+1) seed_0 + C_(t+1) (mod 2^32).
      float4 *wl = reinterpret_cast<float4 *>(tileA);
      const int WL_CAP = (2 * SS * 8) / 20;  // tiles A + B (tile C holds the ball list until the casting is done)
      unsigned *wrange = reinterpret_cast<unsigned *>(wl + WL_CAP);
      // image box in the hand frame, widened by voxel truncation (<= 0.003 sqrt 3) + jitter (<= gmax 0.0009 sqrt 3)
      const double wm = 0.0105;
      const double bx_lo[3] = {h.bottom - wm, h.center - P.vol_w / 2.0 - wm, -P.vol_h - wm};
      const double bx_hi[3] = {h.bottom + P.vol_d + wm, h.center + P.vol_w / 2.0 + wm, P.vol_h + wm};
      // float32 copies for the pre-test: frame, sample, box widened by jitter (gmax 0.0009 sqrt 3) + 2e-5 slack
      float fR[9];
#pragma unroll
      for (int e = 0; e < 9; e++) fR[e] = (float)h.frame[e];
      const float fsx = (float)h.sample[0], fsy = (float)h.sample[1], fsz = (float)h.sample[2];
      const float jm = (float)(gmax * voxel * 0.3 * 1.7320508075688772 + 2e-5);
      const float fbx_lo[3] = {(float)h.bottom - jm, (float)(h.center - P.vol_w / 2.0) - jm, (float)(-P.vol_h) - jm};
      const float fbx_hi[3] = {(float)(h.bottom + P.vol_d) + jm, (float)(h.center + P.vol_w / 2.0) + jm, (float)P.vol_h + jm};
      // one shadow draw -> bit index in camera k's bitmap, or -1 (outside the bitmap AABB / fails the pre-test)
      auto draw_bit = [&](double px, double py, double pz, unsigned seed, int k) -> int {
        const double s0 = sm.sv[k][0], s1 = sm.sv[k][1], s2 = sm.sv[k][2];
        double u = (double)((seed >> 16) & 0x7FFFu) * mxu;
        int v0 = (int)((px + u * s0) * P.vox_mult);
        int v1 = (int)((py + u * s1) * P.vox_mult);
        int v2 = (int)((pz + u * s2) * P.vox_mult);
        int b0 = v0 - o0, b1 = v1 - o1, b2 = v2 - o2;
        if ((unsigned)b0 >= (unsigned)d0 || (unsigned)b1 >= (unsigned)d1 || (unsigned)b2 >= (unsigned)d2) return -1;
        // conservative float32 pre-test of the voxel's lattice point against the image box widened by the largest
        // jitter (+ rounding slack): rejects most out-of-box voxels for ~20 instructions. The exact float64 test
        // (jitter + frame transform, the oracle's operation order) runs once per UNIQUE surviving voxel below.
        const float wx = fmaf((float)v0, 0.003f, -fsx), wy = fmaf((float)v1, 0.003f, -fsy), wz = fmaf((float)v2, 0.003f, -fsz);
        const float hx = fmaf(fR[0], wx, fmaf(fR[1], wy, fR[2] * wz));
        const float hy = fmaf(fR[3], wx, fmaf(fR[4], wy, fR[5] * wz));
        const float hz = fmaf(fR[6], wx, fmaf(fR[7], wy, fR[8] * wz));
        if (hx < fbx_lo[0] || hx > fbx_hi[0] || hy < fbx_lo[1] || hy > fbx_hi[1] || hz < fbx_lo[2] || hz > fbx_hi[2]) return -1;
        return ((b2 * d1 + b1) << 6) + b0;
      };
      auto cast_draw = [&](double px, double py, double pz, unsigned seed, int k, unsigned *bm) {
        int bit = draw_bit(px, py, pz, seed, k);
        if (bit >= 0) atomicOr(bm + (bit >> 5), 1u << (bit & 31));
      };
      const int nball = sm.ball_n;
      bool ball_valid = nball <= BALL_CAP;
      for (int k = 0; k < K; k++) {
        if (!((cam_set >> k) & 1)) continue;  // camera_set(i) >= 1 (hand_set.cpp:141)
        unsigned *bm = bitmap + (size_t)k * bm_words;
        float cull_lo[3], cull_hi[3], cull_inv[3];
        bool cull_par[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const float dv = (float)sm.svh[k][a];
          cull_par[a] = fabsf(dv) < 1e-6f;  // moves the coordinate by < 1e-6 over the whole segment: below the margin
          cull_inv[a] = cull_par[a] ? 0.0f : 1.0f / dv;
          cull_lo[a] = (float)bx_lo[a] - 1e-5f;
          cull_hi[a] = (float)bx_hi[a] + 1e-5f;
        }
        __syncthreads();
        if (tid == 0) sm.wl_n = 0;
        __syncthreads();
        // conservative cull: clip the shadow segment p + u sv, u in [0,1], against the image box (hand frame)
        // widened by the largest displacement voxel truncation + jitter can add; only draws whose 15-bit LCG
        // value falls in [r0, r1] can produce a voxel point inside the box. float32 is enough here: its rounding
        // (~1e-7 of coordinates < 0.2 m, 3e-7 of t) is covered by the extra 1e-5 of box margin and by the +-1 of
        // slack on r0 / r1 (1 / 32767 = 3e-5); the draws themselves are evaluated in the reference's float64.
        auto cull = [&](const float4 &p, unsigned &rg) -> bool {
          const float wx = p.x - fsx, wy = p.y - fsy, wz = p.z - fsz;
          const float o3[3] = {fmaf(fR[0], wx, fmaf(fR[1], wy, fR[2] * wz)), fmaf(fR[3], wx, fmaf(fR[4], wy, fR[5] * wz)),
                               fmaf(fR[6], wx, fmaf(fR[7], wy, fR[8] * wz))};
          float tmin = 0.0f, tmax = 1.0f;
          bool hit = true;
#pragma unroll
          for (int a = 0; a < 3; a++) {
            if (cull_par[a]) {  // segment (numerically) parallel to the slab: inside or outside for every u
              hit = hit && o3[a] >= cull_lo[a] && o3[a] <= cull_hi[a];
            } else {
              const float t1 = (cull_lo[a] - o3[a]) * cull_inv[a], t2 = (cull_hi[a] - o3[a]) * cull_inv[a];
              tmin = fmaxf(tmin, fminf(t1, t2));
              tmax = fminf(tmax, fmaxf(t1, t2));
            }
          }
          if (!hit || tmin > tmax) return false;
          const int r0 = max((int)floorf(tmin * 32767.0f) - 1, 0), r1 = min((int)ceilf(tmax * 32767.0f) + 1, 32767);
          rg = (unsigned)r0 | ((unsigned)r1 << 16);
          return true;
        };
        // appends the point to the work list (one shared-memory atomic per warp); called by all 32 lanes together
        auto append = [&](bool ok, const float4 &p, unsigned rg) {
          const unsigned mk = __ballot_sync(0xffffffffu, ok);
          if (!mk) return;
          const int leader = __ffs(mk) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(&sm.wl_n, __popc(mk));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (!ok) return;
          const int pos = base + __popc(mk & ((1u << lane) - 1));
          const unsigned seed0 = gpdb_shadow_seed((unsigned)h.sample_index, (unsigned)__float_as_int(p.w), (unsigned)k);
          if (pos < WL_CAP) {
            wl[pos] = make_float4(p.x, p.y, p.z, __uint_as_float(seed0));
            wrange[pos] = rg;
          } else {  // work list full (very dense neighbourhood): cast this point's draws in place
            unsigned seed = seed0;
            const int r0 = (int)(rg & 0xFFFFu), r1 = (int)(rg >> 16);
            for (int t = 0; t < P.nsp; t++) {
              int r = (int)gpdb_fastrand(&seed);
              if (r >= r0 && r <= r1) cast_draw((double)p.x, (double)p.y, (double)p.z, seed, k, bm);
            }
          }
        };
        if (ball_valid) {  // the neighbourhood recorded by scan 1: independent loads, all lanes busy
          for (int i0 = 0; i0 < nball; i0 += NT_IMG) {
            const int i = i0 + tid;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            unsigned rg = 0;
            bool ok = false;
            if (i < nball) {
              p = __ldg(cl.pts4 + ball[i]);
              ok = cull(p, rg);
            }
            append(ok, p, rg);
          }
        } else {  // more in-ball points than the list holds: walk the grid again
          scan_balanced<NT_IMG>(P, cl, sr, sm.seg, [&](bool in, const float4 &p, int) {
            unsigned rg = 0;
            const bool ok = in && l2_simple(q, p.x, p.y, p.z) < P.r2_img && cull(p, rg);
            append(ok, p, rg);
          });
        }
        __syncthreads();
        const int nw = min(sm.wl_n, WL_CAP);
        if (prof && tid == 0) atomicAdd(prof + 9, (unsigned long long)nw);
        const int nsp = P.nsp;
        const unsigned nsp_magic = nsp > 1 ? (unsigned)((0x100000000ull + (unsigned)nsp - 1) / (unsigned)nsp) : 0u;  // ceil(2^32 / nsp)
        // Draws in two dense steps. (1) the cheap part of every (point, draw) pair — LCG skip-ahead + the window test of
        // the slab cull, which ~60 % of the draws fail — with the survivors compacted into a list in tile C (the ball list is
        // dead once the work list exists); (2) the float64 voxel arithmetic of the survivors only, all lanes busy (doing
        // it in place kept the 70-instruction body running at ~40 % lane utilisation).
        unsigned *dlist = reinterpret_cast<unsigned *>(tileC);  // item << 7 | t
        const int DL_CAP = 2 * SS;
        __syncthreads();  // the ball list has been consumed
        ball_valid = false;  // ... and is overwritten by the draw list: a further camera walks the grid again
        if (tid == 0) sm.dl_n = 0;
        __syncthreads();
        const int nd = nw * nsp;
        for (int w0 = 0; w0 < nd; w0 += NT_IMG) {
          const int w = w0 + tid;
          bool pass = false;
          unsigned code = 0, seed = 0;
          if (w < nd) {
            // w / nsp by multiply-high: exact for w < 2^32 / nsp (w < WL_CAP * nsp < 2^20)
            const int item = nsp > 1 ? (int)__umulhi((unsigned)w, nsp_magic) : w, t = w - item * nsp;
            seed = sm.lcgA[t] * __float_as_uint(wl[item].w) + sm.lcgC[t];
            const unsigned rg = wrange[item];
            const int r = (int)((seed >> 16) & 0x7FFFu);
            pass = r >= (int)(rg & 0xFFFFu) && r <= (int)(rg >> 16);
            code = ((unsigned)item << 7) | (unsigned)t;
          }
          const unsigned mk = __ballot_sync(0xffffffffu, pass);
          if (mk) {
            const int leader = __ffs(mk) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&sm.dl_n, __popc(mk));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (pass) {
              const int pos = base + __popc(mk & ((1u << lane) - 1));
              if (pos < DL_CAP) dlist[pos] = code;
              else {  // list full: evaluate in place
                const float4 e = wl[code >> 7];
                const int bit = draw_bit((double)e.x, (double)e.y, (double)e.z, seed, k);
                if (bit >= 0) atomicOr(bm + (bit >> 5), 1u << (bit & 31));
              }
            }
          }
        }
        __syncthreads();
        const int ndl = min(sm.dl_n, DL_CAP);
        if (prof && tid == 0) atomicAdd(prof + 10, (unsigned long long)sm.dl_n);
        auto list_bit = [&](int i) -> int {
          if (i >= ndl) return -1;
          const unsigned code = dlist[i];
          const float4 e = wl[code >> 7];
          const int t = (int)(code & 127u);
          const unsigned seed = sm.lcgA[t] * __float_as_uint(e.w) + sm.lcgC[t];
          return draw_bit((double)e.x, (double)e.y, (double)e.z, seed, k);
        };
        for (int i = tid; i < ndl; i += 2 * NT_IMG) {  // two independent float64 chains per thread
          const int bit_a = list_bit(i), bit_b = list_bit(i + NT_IMG);
          if (bit_a >= 0) atomicOr(bm + (bit_a >> 5), 1u << (bit_a & 31));
          if (bit_b >= 0) atomicOr(bm + (bit_b >> 5), 1u << (bit_b & 31));
        }
      }
      __syncthreads();
      PHASE(5);  // S1 (casting) done
      // set intersection over the cameras that see the neighbourhood, starting from camera 0's
      // set even when it is empty (hand_set.cpp:153-176)
      if (K > 1) {
        for (int wd = tid; wd < bm_words; wd += NT_IMG) {
          unsigned acc = bitmap[wd];
          for (int k = 1; k < K; k++)
            if ((cam_set >> k) & 1) acc &= bitmap[(size_t)k * bm_words + wd];
          bitmap[wd] = acc;
        }
      }
      for (int k = tid; k < (3 * SS) >> 1; k += NT_IMG) reinterpret_cast<uint4 *>(tileA)[k] = make_uint4(0, 0, 0, 0);
      __syncthreads();
      // compact the set bits into a list (behind bitmap 0, in the dead box-list region) so that the per-voxel work
      // is spread evenly: shadow voxels are spatially clustered, a thread-per-word loop would be badly unbalanced
      const int nbits = 64 * d1 * d2;
      unsigned *blist = bitmap + bm_words;
      const int BL_CAP = list_bytes / 4 - bm_words;
      auto eval_voxel = [&](unsigned packed) {  // b0 | b1 << 8 | b2 << 16
        int b0 = packed & 255, b1 = (packed >> 8) & 255, b2 = packed >> 16;
        double x, y, z;
        if (!voxel_point_in_box(b0 + o0, b1 + o1, b2 + o2, x, y, z)) return;
        double u[3];
        int cellv[3];
        unit_axis(x, h.bottom, P.vol_d, inv_d, S, u[0], cellv[0]);
        unit_axis(y, h.center - P.vol_w / 2.0, P.vol_w, inv_w, S, u[1], cellv[1]);
        unit_axis(z, -P.vol_h, 2.0 * P.vol_h, inv_h, S, u[2], cellv[2]);
#pragma unroll
        for (int pj = 0; pj < 3; pj++) {
          const int a0 = (pj == 0) ? 0 : 2, a1 = (pj == 2) ? 0 : 1, a2 = (pj == 0) ? 2 : (pj == 1 ? 0 : 1);
          const int row = S - 1 - cellv[a0], col = cellv[a1];
          atomicAdd(tileA + (size_t)pj * SS + row * S + col, (1ull << 48) + (unsigned long long)unit_q32(u[a2]));
        }
      };
      if (tid == 0) sm.wl_n = 0;
      __syncthreads();
      for (int wd0 = 0; wd0 * 32 < nbits; wd0 += NT_IMG) {
        const int wd = wd0 + tid;
        unsigned bits = (wd * 32 < nbits) ? bitmap[wd] : 0u;
        int cntb = __popc(bits);
        int incl = cntb;  // warp-aggregated reservation of list slots
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        int total = __shfl_sync(0xffffffffu, incl, 31), base = 0;
        if (lane == 31 && total) base = atomicAdd(&sm.wl_n, total);
        base = __shfl_sync(0xffffffffu, base, 31);
        int pos = base + incl - cntb;
        const int rowi = wd >> 1;  // (b2 * d1 + b1)
        const unsigned hi = ((unsigned)(rowi % d1) << 8) | ((unsigned)(rowi / d1) << 16) | ((unsigned)(wd & 1) << 5);
        while (bits) {
          int bi = __ffs(bits) - 1;
          bits &= bits - 1;
          if (pos < BL_CAP) blist[pos] = hi | (unsigned)bi;
          else eval_voxel(hi | (unsigned)bi);  // list full: evaluate in place
          pos++;
        }
      }
      __syncthreads();
      const int nset = min(sm.wl_n, BL_CAP);
      if (prof && tid == 0) {
        atomicAdd(prof + 11, (unsigned long long)sm.wl_n);
        atomicAdd(prof + 12, (unsigned long long)bn);
        atomicAdd(prof + 13, (unsigned long long)sm.n_img);
      }
      for (int i = tid; i < nset; i += NT_IMG) eval_voxel(blist[i]);
      __syncthreads();
      PHASE(6);  // S2 bitmap pass done
      // createShadowImage (image_strategy.cpp:193-233): mean per cell, value = max over occupied - mean on occupied cells
      for (int pj = 0; pj < 3; pj++) {
        const unsigned long long *tile = tileA + (size_t)pj * SS;
        uint8_t *plane = planes + (size_t)(pj * 5 + 4) * PLB;
        float avgr[PIXT];
        unsigned occm = 0;
        float mm[2] = {-FLT_MAX, -FLT_MAX};  // max avg, -(min avg) over the occupied cells
#pragma unroll
        for (int t = 0; t < PIXT; t++) {
          const int pix = tid + t * NT_IMG;
          avgr[t] = 0.0f;
          bool oc = false;
          if (pix < SS) {
            const unsigned long long acc = tile[pix];
            const unsigned cntc = (unsigned)(acc >> 48);
            if (cntc) {
              avgr[t] = (float)cell_mean(acc, rcp);
              occm |= 1u << t;
              oc = true;
              mm[0] = fmaxf(mm[0], avgr[t]);
              mm[1] = fmaxf(mm[1], -avgr[t]);
            }
          }
          if (t * NT_IMG < SS) occ_ballot<NT_IMG>(sm.occf, t, oc);
        }
        bool covered_pj;
        block_max<NT_IMG, 2>(mm, sm.fred, sm.occf, S, &covered_pj);  // (its barriers publish the occupancy words)
        const bool any = mm[0] != -FLT_MAX;
        const float maxf = any ? mm[0] : 0.0f;
        const float vmax = any ? maxf - (-mm[1]) : 0.0f;  // largest cell value = max avg - min avg
        float vmin = 0.0f;
        if (covered_pj) {  // general path: min over the dilated float image
          float *srcF = reinterpret_cast<float *>(tileA + (size_t)pj * SS);
          __syncthreads();
#pragma unroll
          for (int t = 0; t < PIXT; t++) {
            const int pix = tid + t * NT_IMG;
            if (pix < SS) srcF[pix] = ((occm >> t) & 1) ? (maxf - avgr[t]) : 0.0f;
          }
          __syncthreads();
          float neg[1] = {-dilated_min<NT_IMG>(srcF, S)};
          block_max<NT_IMG, 1>(neg, sm.fred);
          vmin = -neg[0];
        }
        const Quant qs(vmin, vmax);
        const unsigned bg = qs(0.0f);
#pragma unroll
        for (int t = 0; t < PIXT; t++) {
          const int pix = tid + t * NT_IMG;
          if (pix < SS) {
            const int row = pix / S, col = pix - row * S;
            const bool oc = (occm >> t) & 1;
            if (oc || bg) plane[row * RS + col] = (uint8_t)(oc ? qs(maxf - avgr[t]) : bg);
          }
        }
      }
      // the points phase of the next image expects a clean occupancy bitmap
      __syncthreads();
      for (int k = tid; k < MAXPIX / 32; k += NT_IMG) sm.occf[k] = 0u;
    }
    __syncthreads();
    PHASE(7);  // shadow images done
    // ---- dilate (3x3 max, border ignored) the quantised planes four pixels at a time and assemble the 16-byte pixels
    {
      uint4 *gout = reinterpret_cast<uint4 *>(p16) + (size_t)b * SS;
      for (int g = tid; g < S * RW; g += NT_IMG) {
        const int row = g / RW, c4 = g - row * RW;
        // 3x3 max of 4 pixels x 1 channel per word. Bytes are split into their even / odd 16-bit lanes (E = [b0, b2],
        // O = [b1, b3]) so that every max is ONE native VIMNMX3.U16x2 (the 8-bit SIMD max is emulated on sm_100): vertical
        // max of the three rows first (centre word: E and O; left word: only O, whose high lane is pixel -1; right word:
        // only E, whose low lane is pixel +4), then the horizontal neighbours by lane shifts.
        unsigned res[16];
        const bool up = row > 0, dn = row + 1 < S, lf = c4 > 0, rt = c4 + 1 < RW;
#pragma unroll
        for (int ch = 0; ch < 16; ch++) {
          res[ch] = 0u;
          if (ch < C) {
            const unsigned *W = reinterpret_cast<const unsigned *>(planes + (size_t)ch * PLB) + row * RW + c4;
            const unsigned m0 = W[0], u0 = up ? W[-RW] : 0u, d0 = dn ? W[RW] : 0u;
            const unsigned ml = lf ? W[-1] : 0u, ul = (up && lf) ? W[-RW - 1] : 0u, dl = (dn && lf) ? W[RW - 1] : 0u;
            const unsigned mr = rt ? W[1] : 0u, ur = (up && rt) ? W[-RW + 1] : 0u, dr = (dn && rt) ? W[RW + 1] : 0u;
            const unsigned E = __vimax3_u16x2(__byte_perm(u0, 0u, 0x4240), __byte_perm(m0, 0u, 0x4240), __byte_perm(d0, 0u, 0x4240));
            const unsigned O = __vimax3_u16x2(__byte_perm(u0, 0u, 0x4341), __byte_perm(m0, 0u, 0x4341), __byte_perm(d0, 0u, 0x4341));
            const unsigned LO = __vimax3_u16x2(__byte_perm(ul, 0u, 0x4341), __byte_perm(ml, 0u, 0x4341), __byte_perm(dl, 0u, 0x4341));
            const unsigned RE = __vimax3_u16x2(__byte_perm(ur, 0u, 0x4240), __byte_perm(mr, 0u, 0x4240), __byte_perm(dr, 0u, 0x4240));
            const unsigned En = __vimax3_u16x2(E, O, __byte_perm(LO, O, 0x5432));   // [max(p-1,p0,p1), max(p1,p2,p3)]
            const unsigned On = __vimax3_u16x2(O, E, __byte_perm(E, RE, 0x5432));   // [max(p0,p1,p2), max(p2,p3,p4)]
            res[ch] = __byte_perm(En, On, 0x6240);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          if (c4 * 4 + i < S) {
            const unsigned sel = (unsigned)i | ((unsigned)(4 + i) << 4);
            uint4 o;
            o.x = __byte_perm(__byte_perm(res[0], res[1], sel), __byte_perm(res[2], res[3], sel), 0x5410);
            o.y = __byte_perm(__byte_perm(res[4], res[5], sel), __byte_perm(res[6], res[7], sel), 0x5410);
            o.z = __byte_perm(__byte_perm(res[8], res[9], sel), __byte_perm(res[10], res[11], sel), 0x5410);
            o.w = __byte_perm(__byte_perm(res[12], res[13], sel), __byte_perm(res[14], res[15], sel), 0x5410);
            gout[row * S + c4 * 4 + i] = o;
          }
        }
      }
    }
    PHASE(8);  // assembled image stored
  }
}

// ------------------------------------------------------------------------------------------------
// k_images2: the fast path of the image stage — the same algorithm as k_images in 103 KB of shared memory and 64
// registers, so that TWO 512-thread CTAs share an SM (k_images needs 214 KB: one CTA per SM, 16 warps, issue slots idle
// behind barriers and dependent-issue stalls). What makes the footprint fit:
//   * the box list holds 1024 points (24 B each: key, unit coordinates, cells; the normal of a WINNER is re-read from the
//     cloud); images with more box points are appended to `ovf` and redone by k_images (2048 points, one CTA per SM);
//   * two 64-bit tiles instead of three: the shadow sums of projections 0 and 1 are taken first, the per-voxel result for
//     projection 2 (cell + fixed-point coordinate) is stashed next to the voxel code and summed in a second, cheap pass;
//   * the quantised POINT channels are scattered straight into the image's own (still unused) 57.6 KB of HBM / L2 as twelve
//     byte planes and read back into the dead tiles before the dilation; only the three shadow planes stay in shared memory.
// Requires image_size 60 and at most two cameras (otherwise k_images does all the work). Results are bit-identical to
// k_images (tests/test_gpu_parity.py::test_image_kernels_agree).
// ------------------------------------------------------------------------------------------------
constexpr int BOX_CAP2 = 1024;
struct Img2Smem {
  SegScan<NT_IMG> seg;
  gpdb_pose h;
  double red[NT_IMG / 32][4];
  double center[3];
  double sv[2][3];
  double svh[2][3];
  unsigned occf[MAXPIX / 32 + 2];
  unsigned lcgA[GPDB_MAX_NSP], lcgC[GPDB_MAX_NSP];
  int cam_or, n_img, box_n, wl_n, dl_n, st_n;
  int bm_org[3], bm_dims[3];
  float fred[NT_IMG / 32][8];
};

__global__ void __launch_bounds__(NT_IMG, 2) k_images2(const DevParams *Pp, DevCloud cl, const gpdb_pose *cand, int nc,
                                                       uint8_t *p16, const double *qtab, int *ovf, int *ovf_count,
                                                       unsigned long long *prof) {
  long long t_phase = 0;
  const DevParams &P = *Pp;
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ Img2Smem sm;
  constexpr int S = 60, SS = S * S, RW = S / 4, PLB = SS;  // 15 words per row, 3600-byte planes
  constexpr int PIXT = (SS + NT_IMG - 1) / NT_IMG;
  constexpr int JW = (BOX_CAP2 + NT_IMG - 1) / NT_IMG;  // box points per thread
  const int C = P.C;
  unsigned long long *tileA = reinterpret_cast<unsigned long long *>(dyn);
  unsigned long long *tileB = tileA + SS;
  uint8_t *splanes = reinterpret_cast<uint8_t *>(tileB + SS);       // 3 shadow planes
  unsigned char *lbase = splanes + 3 * PLB;
  constexpr int LIST_BYTES = BOX_CAP2 * 36;                         // 24 B per box point + room for the shadow phase
  unsigned long long *bkeys = reinterpret_cast<unsigned long long *>(lbase);
  unsigned *bq = reinterpret_cast<unsigned *>(bkeys + BOX_CAP2);    // [3][CAP]
  unsigned *bcell = bq + 3 * BOX_CAP2;                              // packed 3 x 8 bit
  unsigned *bitmap = reinterpret_cast<unsigned *>(lbase);           // shadow phase: aliases the (dead) box list
  uint8_t *tplanes = reinterpret_cast<uint8_t *>(tileA);            // final stage: the 12 point planes over the dead tiles
  const int tid = threadIdx.x, lane = tid & 31;
  const int nproj = (C >= 12) ? 3 : 1;
  const bool do_nrm = C != 1, do_dep = C == 1 || C >= 12;
  const int npp = (C == 15 || C == 12) ? 12 : C;  // point planes (C = 1: the depth plane is plane 0)
  for (int k = tid; k < GPDB_MAX_NSP; k += NT_IMG) {
    sm.lcgA[k] = P.lcgA[k];
    sm.lcgC[k] = P.lcgC[k];
  }
  for (int k = tid; k < MAXPIX / 32 + 2; k += NT_IMG) sm.occf[k] = 0u;

  for (int b = blockIdx.x; b < nc; b += gridDim.x) {
    __syncthreads();
    {
      const int *src = reinterpret_cast<const int *>(cand + b);
      int *dst = reinterpret_cast<int *>(&sm.h);
      for (int k = tid; k < (int)(sizeof(gpdb_pose) / 4); k += NT_IMG) dst[k] = src[k];
    }
    if (tid == 0) {
      sm.cam_or = 0;
      sm.n_img = 0;
      sm.box_n = 0;
    }
    uint8_t *gimg = p16 + (size_t)b * SS * 16;  // the image's own memory: first the point planes, finally the pixels
    for (int k = tid; k < (npp * PLB) >> 4; k += NT_IMG) reinterpret_cast<uint4 *>(gimg)[k] = make_uint4(0, 0, 0, 0);
    for (int k = tid; k < SS; k += NT_IMG) reinterpret_cast<uint4 *>(tileA)[k] = make_uint4(0, 0, 0, 0);  // tiles A, B
    for (int k = tid; k < (3 * PLB) >> 4; k += NT_IMG) reinterpret_cast<uint4 *>(splanes)[k] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    PHASE(1);
    const gpdb_pose &h = sm.h;
    const bool need_cam = (C == 15) && !P.all_seen;
    const double inv_d = 1.0 / P.vol_d, inv_w = 1.0 / P.vol_w, inv_h = 1.0 / (2.0 * P.vol_h);
    float q[3] = {(float)h.sample[0], (float)h.sample[1], (float)h.sample[2]};
    SegRange sr = seg_range(P, q, P.rf_img);
    // ---- ball scan 1: neighbourhood centre, camera set, raw box points
    double sx = 0, sy = 0, sz = 0;
    int cnt = 0, cam_or = 0;
    scan_balanced<NT_IMG>(P, cl, sr, sm.seg, [&](bool in, const float4 &p, int) {
      bool inb = false;
      unsigned long long key = 0;
      if (in) {
        float d = l2_simple(q, p.x, p.y, p.z);
        if (d < P.r2_img) {
          const int idx = __float_as_int(p.w);
          sx += (double)p.x;
          sy += (double)p.y;
          sz += (double)p.z;
          cnt++;
          if (need_cam) cam_or |= cl.cam[idx];
          double x, y, z;
          to_frame(h.frame, (double)p.x - h.sample[0], (double)p.y - h.sample[1], (double)p.z - h.sample[2], x, y, z);
          if (in_image_box(P, h, x, y, z)) {
            inb = true;
            key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
          }
        }
      }
      unsigned mk = __ballot_sync(0xffffffffu, inb);
      if (mk) {
        int leader = __ffs(mk) - 1, base = 0;
        if (lane == leader) base = atomicAdd(&sm.box_n, __popc(mk));
        base = __shfl_sync(0xffffffffu, base, leader);
        int pos = base + __popc(mk & ((1u << lane) - 1));
        if (inb && pos < BOX_CAP2) {
          bkeys[pos] = key;
          bq[pos] = __float_as_uint(p.x);
          bq[BOX_CAP2 + pos] = __float_as_uint(p.y);
          bq[2 * BOX_CAP2 + pos] = __float_as_uint(p.z);
        }
      }
    });
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    cnt = warp_sum(cnt);
    cam_or = __reduce_or_sync(0xffffffffu, (unsigned)cam_or);
    __syncthreads();
    if (lane == 0) {
      sm.red[tid >> 5][0] = sx;
      sm.red[tid >> 5][1] = sy;
      sm.red[tid >> 5][2] = sz;
      atomicAdd(&sm.n_img, cnt);
      atomicOr(&sm.cam_or, cam_or);
    }
    __syncthreads();
    if (tid == 0) {
      double a0 = 0, a1 = 0, a2 = 0;
      for (int w = 0; w < NT_IMG / 32; w++) {
        a0 += sm.red[w][0];
        a1 += sm.red[w][1];
        a2 += sm.red[w][2];
      }
      double nn = (double)sm.n_img;
      if (!need_cam && sm.n_img > 0) sm.cam_or = (1 << P.K) - 1;
      sm.center[0] = a0 / nn;
      sm.center[1] = a1 / nn;
      sm.center[2] = a2 / nn;
      if (sm.box_n > BOX_CAP2) ovf[atomicAdd(ovf_count, 1)] = b;  // redone by k_images (larger list)
    }
    __syncthreads();
    PHASE(2);
    if (sm.box_n > BOX_CAP2) continue;  // uniform
    const int bn = sm.box_n;
    double *rcp = &sm.red[0][0];  // the scan's reduction scratch is idle from here to the next image: reciprocal table (cell_mean)
    if (tid < RCP_N) rcp[tid] = 1.0 / ((double)tid * 4294967296.0);  // visible after the barrier that ends the box-point pass
    for (int k = tid; k < bn; k += NT_IMG) {
      const double px = (double)__uint_as_float(bq[k]), py = (double)__uint_as_float(bq[BOX_CAP2 + k]),
                   pz = (double)__uint_as_float(bq[2 * BOX_CAP2 + k]);
      double x, y, z, u0, u1, u2;
      int c0, c1, c2;
      to_frame(h.frame, px - h.sample[0], py - h.sample[1], pz - h.sample[2], x, y, z);
      unit_axis(x, h.bottom, P.vol_d, inv_d, S, u0, c0);
      unit_axis(y, h.center - P.vol_w / 2.0, P.vol_w, inv_w, S, u1, c1);
      unit_axis(z, -P.vol_h, 2.0 * P.vol_h, inv_h, S, u2, c2);
      bq[k] = unit_q32(u0);
      bq[BOX_CAP2 + k] = unit_q32(u1);
      bq[2 * BOX_CAP2 + k] = unit_q32(u2);
      bcell[k] = (unsigned)c0 | ((unsigned)c1 << 8) | ((unsigned)c2 << 16);
    }
    __syncthreads();

    // ---- points phase
    for (int pj = 0; pj < nproj; pj++) {
      const int a0 = (pj == 0) ? 0 : 2, a1 = (pj == 2) ? 0 : 1, a2 = (pj == 0) ? 2 : (pj == 1 ? 0 : 1);
      for (int k = tid; k < bn; k += NT_IMG) {
        const unsigned cc = bcell[k];
        const int pix = (S - 1 - (int)((cc >> (8 * a0)) & 255)) * S + (int)((cc >> (8 * a1)) & 255);
        atomicMax(tileA + pix, bkeys[k]);
        atomicAdd(tileB + pix, (1ull << 48) + (unsigned long long)bq[a2 * BOX_CAP2 + k]);
        atomicOr(&sm.occf[pix >> 5], 1u << (pix & 31));
      }
      __syncthreads();
      // winners (largest key of their cell) carry the cell's values; each thread owns <= JW box points
      float wv[JW][4];
      int wpix[JW];
      float mxv[2] = {0.0f, 0.0f};
#pragma unroll
      for (int j = 0; j < JW; j++) {
        const int k = tid + j * NT_IMG;
        wpix[j] = -1;
        if (k < bn) {
          const unsigned cc = bcell[k];
          const int pix = (S - 1 - (int)((cc >> (8 * a0)) & 255)) * S + (int)((cc >> (8 * a1)) & 255);
          const unsigned long long key = bkeys[k];
          if (tileA[pix] == key) {
            wpix[j] = pix;
            const double *nn = cl.nrm + 3 * (size_t)(unsigned)(key & 0xffffffffull);
            double n0, n1, n2;
            to_frame(h.frame, nn[0], nn[1], nn[2], n0, n1, n2);
            wv[j][0] = (float)fabs(n0);
            wv[j][1] = (float)fabs(n1);
            wv[j][2] = (float)fabs(n2);
            const float avg = (float)cell_mean(tileB[pix], rcp);
            wv[j][3] = (float)(1.0 - (double)avg);
            mxv[0] = fmaxf(mxv[0], fmaxf(fmaxf(wv[j][0], wv[j][1]), wv[j][2]));
            mxv[1] = fmaxf(mxv[1], wv[j][3]);
          }
        }
      }
      bool covered;
      block_max<NT_IMG, 2>(mxv, sm.fred, sm.occf, S, &covered);
      float mnv[2] = {0.0f, 0.0f};
      const int cb = (C == 1) ? 0 : pj * 4;  // first point plane of the projection
      if (covered) {  // general path: no all-empty 3x3 window -> min over the dilated float images
        float *F = reinterpret_cast<float *>(tileA);  // 4 x SS floats = tiles A + B (every winner holds its values)
        for (int k = tid; k < SS; k += NT_IMG) reinterpret_cast<uint4 *>(F)[k] = make_uint4(0, 0, 0, 0);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < JW; j++)
          if (wpix[j] >= 0)
#pragma unroll
            for (int c = 0; c < 4; c++) F[c * SS + wpix[j]] = wv[j][c];
        __syncthreads();
        float neg[2];
        neg[0] = -fminf(fminf(dilated_min<NT_IMG>(F, S), dilated_min<NT_IMG>(F + SS, S)), dilated_min<NT_IMG>(F + 2 * SS, S));
        neg[1] = -dilated_min<NT_IMG>(F + 3 * SS, S);
        block_max<NT_IMG, 2>(neg, sm.fred);
        mnv[0] = -neg[0];
        mnv[1] = -neg[1];
        const Quant qn(mnv[0], mxv[0]), qd(mnv[1], mxv[1]);
        const unsigned bg_n = qn(0.0f) * 0x01010101u, bg_d = qd(0.0f) * 0x01010101u;  // background = quantised 0
        for (int k = tid; k < (PLB >> 2); k += NT_IMG) {
          if (do_nrm)
#pragma unroll
            for (int c = 0; c < 3; c++) reinterpret_cast<unsigned *>(gimg + (size_t)(cb + c) * PLB)[k] = bg_n;
          if (do_dep) reinterpret_cast<unsigned *>(gimg + (size_t)(cb + (C == 1 ? 0 : 3)) * PLB)[k] = bg_d;
        }
        __syncthreads();  // the background is in place before the winners overwrite their cells
      }
      {
        const Quant qn(mnv[0], mxv[0]), qd(mnv[1], mxv[1]);
#pragma unroll
        for (int j = 0; j < JW; j++)
          if (wpix[j] >= 0) {
            if (do_nrm) {
              gimg[(size_t)(cb + 0) * PLB + wpix[j]] = (uint8_t)qn(wv[j][0]);
              gimg[(size_t)(cb + 1) * PLB + wpix[j]] = (uint8_t)qn(wv[j][1]);
              gimg[(size_t)(cb + 2) * PLB + wpix[j]] = (uint8_t)qn(wv[j][2]);
            }
            if (do_dep) gimg[(size_t)(cb + (C == 1 ? 0 : 3)) * PLB + wpix[j]] = (uint8_t)qd(wv[j][3]);
          }
      }
      // clean tiles / occupancy for the next projection
      if (covered) {
        for (int k = tid; k < SS; k += NT_IMG) reinterpret_cast<uint4 *>(tileA)[k] = make_uint4(0, 0, 0, 0);
        for (int k = tid; k < MAXPIX / 32; k += NT_IMG) sm.occf[k] = 0u;
      } else {
        __syncthreads();  // every winner has read its cell
        for (int k = tid; k < bn; k += NT_IMG) {
          const unsigned cc = bcell[k];
          const int pix = (S - 1 - (int)((cc >> (8 * a0)) & 255)) * S + (int)((cc >> (8 * a1)) & 255);
          tileA[pix] = 0ull;
          tileB[pix] = 0ull;
          sm.occf[pix >> 5] = 0u;
        }
      }
      __syncthreads();
    }
    PHASE(3);

    // ---- shadow phase (15 channels)
    if (C == 15) {
      const int K = P.K;
      const int bmd = P.bm_dim;
      const int bm_words = 2 * bmd * bmd;
      const double gmax = qtab[GPDB_QTAB_SIZE - 1];
      const double voxel = GPDB_SHADOW_VOXEL;
      if (tid == 0) {
        const double jmax = gmax * voxel * 0.3 + 1e-9;
        const double half_od = P.vol_w / 2.0;
        double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int cr = 0; cr < 8; cr++) {
          double cx = (cr & 1) ? h.bottom + P.vol_d : h.bottom;
          double cy = (cr & 2) ? h.center + half_od : h.center - half_od;
          double cz = (cr & 4) ? P.vol_h : -P.vol_h;
          for (int r = 0; r < 3; r++) {
            double wv2 = h.frame[r] * cx + h.frame[3 + r] * cy + h.frame[6 + r] * cz + h.sample[r];
            mn[r] = fmin(mn[r], wv2);
            mx[r] = fmax(mx[r], wv2);
          }
        }
        for (int r = 0; r < 3; r++) {
          int lo = (int)floor((mn[r] - jmax) * P.vox_mult) - 1;
          int hi = (int)floor((mx[r] + jmax) * P.vox_mult) + 1;
          sm.bm_org[r] = lo;
          sm.bm_dims[r] = min(hi - lo + 1, bmd);
        }
      }
      if (tid < K) {
        double s0 = sm.center[0] - P.vp[tid][0], s1 = sm.center[1] - P.vp[tid][1], s2 = sm.center[2] - P.vp[tid][2];
        double nn = sqrt((s0 * s0 + s1 * s1) + s2 * s2);
        sm.sv[tid][0] = P.shadow_length * s0 / nn;
        sm.sv[tid][1] = P.shadow_length * s1 / nn;
        sm.sv[tid][2] = P.shadow_length * s2 / nn;
        to_frame(h.frame, sm.sv[tid][0], sm.sv[tid][1], sm.sv[tid][2], sm.svh[tid][0], sm.svh[tid][1], sm.svh[tid][2]);
      }
      for (int k = tid; k < bm_words * K; k += NT_IMG) bitmap[k] = 0u;
      __syncthreads();
      PHASE(4);
      const int o0 = sm.bm_org[0], o1 = sm.bm_org[1], o2 = sm.bm_org[2];
      const int d0 = sm.bm_dims[0], d1 = sm.bm_dims[1], d2 = sm.bm_dims[2];
      const double mxu = 1.0 / 32767.0;
      auto voxel_point_in_box = [&](int v0, int v1, int v2, double &x, double &y, double &z) -> bool {
        double g = qtab[gpdb_voxel_hash(v0, v1, v2) & (GPDB_QTAB_SIZE - 1)];
        double jit = 1.0 * g * voxel * 0.3;
        double w0 = (double)v0 * voxel + jit, w1 = (double)v1 * voxel + jit, w2 = (double)v2 * voxel + jit;
        to_frame(h.frame, w0 - h.sample[0], w1 - h.sample[1], w2 - h.sample[2], x, y, z);
        return in_image_box(P, h, x, y, z);
      };
      const int cam_set = sm.cam_or;
      float4 *wl = reinterpret_cast<float4 *>(tileA);       // work list over tile A ...
      constexpr int WL_CAP = (SS * 8) / 20;
      unsigned *wrange = reinterpret_cast<unsigned *>(wl + WL_CAP);
      unsigned *dlist = reinterpret_cast<unsigned *>(tileB);  // ... draw list over tile B
      constexpr int DL_CAP = 2 * SS;
      const double wm = 0.0105;
      const double bx_lo[3] = {h.bottom - wm, h.center - P.vol_w / 2.0 - wm, -P.vol_h - wm};
      const double bx_hi[3] = {h.bottom + P.vol_d + wm, h.center + P.vol_w / 2.0 + wm, P.vol_h + wm};
      float fR[9];
#pragma unroll
      for (int e = 0; e < 9; e++) fR[e] = (float)h.frame[e];
      const float fsx = (float)h.sample[0], fsy = (float)h.sample[1], fsz = (float)h.sample[2];
      const float jm = (float)(gmax * voxel * 0.3 * 1.7320508075688772 + 2e-5);
      const float fbx_lo[3] = {(float)h.bottom - jm, (float)(h.center - P.vol_w / 2.0) - jm, (float)(-P.vol_h) - jm};
      const float fbx_hi[3] = {(float)(h.bottom + P.vol_d) + jm, (float)(h.center + P.vol_w / 2.0) + jm, (float)P.vol_h + jm};
      auto draw_bit = [&](double px, double py, double pz, unsigned seed, int k) -> int {
        const double s0 = sm.sv[k][0], s1 = sm.sv[k][1], s2 = sm.sv[k][2];
        double u = (double)((seed >> 16) & 0x7FFFu) * mxu;
        int v0 = (int)((px + u * s0) * P.vox_mult);
        int v1 = (int)((py + u * s1) * P.vox_mult);
        int v2 = (int)((pz + u * s2) * P.vox_mult);
        int b0 = v0 - o0, b1 = v1 - o1, b2 = v2 - o2;
        if ((unsigned)b0 >= (unsigned)d0 || (unsigned)b1 >= (unsigned)d1 || (unsigned)b2 >= (unsigned)d2) return -1;
        const float wx = fmaf((float)v0, 0.003f, -fsx), wy = fmaf((float)v1, 0.003f, -fsy), wz = fmaf((float)v2, 0.003f, -fsz);
        const float hx = fmaf(fR[0], wx, fmaf(fR[1], wy, fR[2] * wz));
        const float hy = fmaf(fR[3], wx, fmaf(fR[4], wy, fR[5] * wz));
        const float hz = fmaf(fR[6], wx, fmaf(fR[7], wy, fR[8] * wz));
        if (hx < fbx_lo[0] || hx > fbx_hi[0] || hy < fbx_lo[1] || hy > fbx_hi[1] || hz < fbx_lo[2] || hz > fbx_hi[2]) return -1;
        return ((b2 * d1 + b1) << 6) + b0;
      };
      for (int k = 0; k < K; k++) {
        if (!((cam_set >> k) & 1)) continue;
        unsigned *bm = bitmap + (size_t)k * bm_words;
        float cull_lo[3], cull_hi[3], cull_inv[3];
        bool cull_par[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const float dv = (float)sm.svh[k][a];
          cull_par[a] = fabsf(dv) < 1e-6f;
          cull_inv[a] = cull_par[a] ? 0.0f : 1.0f / dv;
          cull_lo[a] = (float)bx_lo[a] - 1e-5f;
          cull_hi[a] = (float)bx_hi[a] + 1e-5f;
        }
        __syncthreads();
        if (tid == 0) {
          sm.wl_n = 0;
          sm.dl_n = 0;
        }
        __syncthreads();
        auto cull = [&](const float4 &p, unsigned &rg) -> bool {
          const float wx = p.x - fsx, wy = p.y - fsy, wz = p.z - fsz;
          const float o3[3] = {fmaf(fR[0], wx, fmaf(fR[1], wy, fR[2] * wz)), fmaf(fR[3], wx, fmaf(fR[4], wy, fR[5] * wz)),
                               fmaf(fR[6], wx, fmaf(fR[7], wy, fR[8] * wz))};
          float tmin = 0.0f, tmax = 1.0f;
          bool hit = true;
#pragma unroll
          for (int a = 0; a < 3; a++) {
            if (cull_par[a]) {
              hit = hit && o3[a] >= cull_lo[a] && o3[a] <= cull_hi[a];
            } else {
              const float t1 = (cull_lo[a] - o3[a]) * cull_inv[a], t2 = (cull_hi[a] - o3[a]) * cull_inv[a];
              tmin = fmaxf(tmin, fminf(t1, t2));
              tmax = fminf(tmax, fmaxf(t1, t2));
            }
          }
          if (!hit || tmin > tmax) return false;
          const int r0 = max((int)floorf(tmin * 32767.0f) - 1, 0), r1 = min((int)ceilf(tmax * 32767.0f) + 1, 32767);
          rg = (unsigned)r0 | ((unsigned)r1 << 16);
          return true;
        };
        scan_balanced<NT_IMG>(P, cl, sr, sm.seg, [&](bool in, const float4 &p, int) {
          unsigned rg = 0;
          const bool ok = in && l2_simple(q, p.x, p.y, p.z) < P.r2_img && cull(p, rg);
          const unsigned mk = __ballot_sync(0xffffffffu, ok);
          if (!mk) return;
          const int leader = __ffs(mk) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(&sm.wl_n, __popc(mk));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (!ok) return;
          const int pos = base + __popc(mk & ((1u << lane) - 1));
          const unsigned seed0 = gpdb_shadow_seed((unsigned)h.sample_index, (unsigned)__float_as_int(p.w), (unsigned)k);
          if (pos < WL_CAP) {
            wl[pos] = make_float4(p.x, p.y, p.z, __uint_as_float(seed0));
            wrange[pos] = rg;
          } else {  // work list full: cast this point's draws in place
            unsigned seed = seed0;
            const int r0 = (int)(rg & 0xFFFFu), r1 = (int)(rg >> 16);
            for (int t = 0; t < P.nsp; t++) {
              int r = (int)gpdb_fastrand(&seed);
              if (r >= r0 && r <= r1) {
                int bit = draw_bit((double)p.x, (double)p.y, (double)p.z, seed, k);
                if (bit >= 0) atomicOr(bm + (bit >> 5), 1u << (bit & 31));
              }
            }
          }
        });
        __syncthreads();
        const int nw = min(sm.wl_n, WL_CAP);
        if (prof && tid == 0) atomicAdd(prof + 9, (unsigned long long)nw);
        const int nsp = P.nsp;
        const unsigned nsp_magic = nsp > 1 ? (unsigned)((0x100000000ull + (unsigned)nsp - 1) / (unsigned)nsp) : 0u;
        const int nd = nw * nsp;
        for (int w0 = 0; w0 < nd; w0 += NT_IMG) {  // window test of every draw, survivors compacted
          const int w = w0 + tid;
          bool pass = false;
          unsigned code = 0, seed = 0;
          if (w < nd) {
            const int item = nsp > 1 ? (int)__umulhi((unsigned)w, nsp_magic) : w, t = w - item * nsp;
            seed = sm.lcgA[t] * __float_as_uint(wl[item].w) + sm.lcgC[t];
            const unsigned rg = wrange[item];
            const int r = (int)((seed >> 16) & 0x7FFFu);
            pass = r >= (int)(rg & 0xFFFFu) && r <= (int)(rg >> 16);
            code = ((unsigned)item << 7) | (unsigned)t;
          }
          const unsigned mk = __ballot_sync(0xffffffffu, pass);
          if (mk) {
            const int leader = __ffs(mk) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&sm.dl_n, __popc(mk));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (pass) {
              const int pos = base + __popc(mk & ((1u << lane) - 1));
              if (pos < DL_CAP) dlist[pos] = code;
              else {
                const float4 e = wl[code >> 7];
                const int bit = draw_bit((double)e.x, (double)e.y, (double)e.z, seed, k);
                if (bit >= 0) atomicOr(bm + (bit >> 5), 1u << (bit & 31));
              }
            }
          }
        }
        __syncthreads();
        const int ndl = min(sm.dl_n, DL_CAP);
        if (prof && tid == 0) atomicAdd(prof + 10, (unsigned long long)sm.dl_n);
        auto list_bit = [&](int i) -> int {
          if (i >= ndl) return -1;
          const unsigned code = dlist[i];
          const float4 e = wl[code >> 7];
          const int t = (int)(code & 127u);
          const unsigned seed = sm.lcgA[t] * __float_as_uint(e.w) + sm.lcgC[t];
          return draw_bit((double)e.x, (double)e.y, (double)e.z, seed, k);
        };
        for (int i = tid; i < ndl; i += 2 * NT_IMG) {
          const int bit_a = list_bit(i), bit_b = list_bit(i + NT_IMG);
          if (bit_a >= 0) atomicOr(bm + (bit_a >> 5), 1u << (bit_a & 31));
          if (bit_b >= 0) atomicOr(bm + (bit_b >> 5), 1u << (bit_b & 31));
        }
      }
      __syncthreads();
      PHASE(5);
      if (K > 1) {  // intersection, starting from camera 0's set even when it is empty (hand_set.cpp:153-176)
        for (int wd = tid; wd < bm_words; wd += NT_IMG) {
          unsigned acc = bitmap[wd];
          for (int k = 1; k < K; k++)
            if ((cam_set >> k) & 1) acc &= bitmap[(size_t)k * bm_words + wd];
          bitmap[wd] = acc;
        }
      }
      for (int k = tid; k < SS; k += NT_IMG) reinterpret_cast<uint4 *>(tileA)[k] = make_uint4(0, 0, 0, 0);  // tiles A, B
      if (tid == 0) sm.st_n = 0;
      __syncthreads();
      // voxel list (8 B per voxel behind the bitmaps): .x = voxel code, later the stash for projection 2
      const int nbits = 64 * d1 * d2;
      uint2 *stash = reinterpret_cast<uint2 *>(bitmap + (((size_t)bm_words * (K > 1 ? K : 1) + 1) & ~(size_t)1));
      const int ST_CAP = (LIST_BYTES - (int)((reinterpret_cast<unsigned char *>(stash)) - lbase)) / 8;
      // one voxel: exact box test, then the sums of projections [pj_lo, pj_hi) (tile pj - pj_lo); returns the stash entry of
      // projection 2 (cell | 1 << 31, fixed-point coordinate), zero when the voxel point lies outside the box
      auto eval_voxel = [&](unsigned packed, int pj_lo, int pj_hi) -> uint2 {
        int b0 = packed & 255, b1 = (packed >> 8) & 255, b2 = packed >> 16;
        double x, y, z;
        if (!voxel_point_in_box(b0 + o0, b1 + o1, b2 + o2, x, y, z)) return make_uint2(0u, 0u);
        double u[3];
        int cellv[3];
        unit_axis(x, h.bottom, P.vol_d, inv_d, S, u[0], cellv[0]);
        unit_axis(y, h.center - P.vol_w / 2.0, P.vol_w, inv_w, S, u[1], cellv[1]);
        unit_axis(z, -P.vol_h, 2.0 * P.vol_h, inv_h, S, u[2], cellv[2]);
        uint2 st = make_uint2(0u, 0u);
#pragma unroll
        for (int pj = 0; pj < 3; pj++) {
          const int a0 = (pj == 0) ? 0 : 2, a1 = (pj == 2) ? 0 : 1, a2 = (pj == 0) ? 2 : (pj == 1 ? 0 : 1);
          const int pix = (S - 1 - cellv[a0]) * S + cellv[a1];
          const unsigned qv = unit_q32(u[a2]);
          if (pj >= pj_lo && pj < pj_hi) atomicAdd(tileA + (size_t)(pj - pj_lo) * SS + pix, (1ull << 48) + (unsigned long long)qv);
          if (pj == 2) st = make_uint2((unsigned)pix | 0x80000000u, qv);
        }
        return st;
      };
      for (int wd0 = 0; wd0 * 32 < nbits; wd0 += NT_IMG) {
        const int wd = wd0 + tid;
        unsigned bits = (wd * 32 < nbits) ? bitmap[wd] : 0u;
        int cntb = __popc(bits);
        int incl = cntb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        int total = __shfl_sync(0xffffffffu, incl, 31), base = 0;
        if (lane == 31 && total) base = atomicAdd(&sm.st_n, total);
        base = __shfl_sync(0xffffffffu, base, 31);
        int pos = base + incl - cntb;
        const int rowi = wd >> 1;
        const unsigned hi = ((unsigned)(rowi % d1) << 8) | ((unsigned)(rowi / d1) << 16) | ((unsigned)(wd & 1) << 5);
        while (bits) {
          int bi = __ffs(bits) - 1;
          bits &= bits - 1;
          if (pos < ST_CAP) stash[pos].x = hi | (unsigned)bi;
          else eval_voxel(hi | (unsigned)bi, 0, 2);  // list full: projections 0 and 1 in place (2: second walk below)
          pos++;
        }
      }
      __syncthreads();
      const int nset_all = sm.st_n, nset = min(nset_all, ST_CAP);
      if (prof && tid == 0) {
        atomicAdd(prof + 11, (unsigned long long)nset_all);
        atomicAdd(prof + 12, (unsigned long long)bn);
        atomicAdd(prof + 13, (unsigned long long)sm.n_img);
      }
      for (int i = tid; i < nset; i += NT_IMG) stash[i] = eval_voxel(stash[i].x, 0, 2);
      __syncthreads();
      PHASE(6);
      // createShadowImage (image_strategy.cpp:193-233) for one sum tile -> shadow plane pj
      auto shadow_channel = [&](int pj, unsigned long long *tile) {
        uint8_t *plane = splanes + (size_t)pj * PLB;
        float avgr[PIXT];
        unsigned occm = 0;
        float mm[2] = {-FLT_MAX, -FLT_MAX};
#pragma unroll
        for (int t = 0; t < PIXT; t++) {
          const int pix = tid + t * NT_IMG;
          avgr[t] = 0.0f;
          bool oc = false;
          if (pix < SS) {
            const unsigned long long acc = tile[pix];
            const unsigned cntc = (unsigned)(acc >> 48);
            if (cntc) {
              avgr[t] = (float)cell_mean(acc, rcp);
              occm |= 1u << t;
              oc = true;
              mm[0] = fmaxf(mm[0], avgr[t]);
              mm[1] = fmaxf(mm[1], -avgr[t]);
            }
          }
          occ_ballot<NT_IMG>(sm.occf, t, oc);
        }
        bool covered_pj;
        block_max<NT_IMG, 2>(mm, sm.fred, sm.occf, S, &covered_pj);
        const bool any = mm[0] != -FLT_MAX;
        const float maxf = any ? mm[0] : 0.0f;
        const float vmax = any ? maxf - (-mm[1]) : 0.0f;
        float vmin = 0.0f;
        if (covered_pj) {
          float *srcF = reinterpret_cast<float *>(tile);
          __syncthreads();
#pragma unroll
          for (int t = 0; t < PIXT; t++) {
            const int pix = tid + t * NT_IMG;
            if (pix < SS) srcF[pix] = ((occm >> t) & 1) ? (maxf - avgr[t]) : 0.0f;
          }
          __syncthreads();
          float neg[1] = {-dilated_min<NT_IMG>(srcF, S)};
          block_max<NT_IMG, 1>(neg, sm.fred);
          vmin = -neg[0];
        }
        const Quant qs(vmin, vmax);
        const unsigned bg = qs(0.0f);
#pragma unroll
        for (int t = 0; t < PIXT; t++) {
          const int pix = tid + t * NT_IMG;
          if (pix < SS) {
            const bool oc = (occm >> t) & 1;
            if (oc || bg) plane[pix] = (uint8_t)(oc ? qs(maxf - avgr[t]) : bg);
          }
        }
      };
      shadow_channel(0, tileA);
      shadow_channel(1, tileB);
      __syncthreads();
      for (int k = tid; k < SS / 2; k += NT_IMG) reinterpret_cast<uint4 *>(tileA)[k] = make_uint4(0, 0, 0, 0);  // tile A
      __syncthreads();
      // projection 2: from the stash, or — when the voxel list overflowed — by a second walk over the whole bitmap
      if (nset_all <= ST_CAP) {
        for (int i = tid; i < nset; i += NT_IMG) {
          const uint2 st = stash[i];
          if (st.x & 0x80000000u) atomicAdd(tileA + (st.x & 0x7fffffffu), (1ull << 48) + (unsigned long long)st.y);
        }
      } else {
        for (int wd = tid; wd * 32 < nbits; wd += NT_IMG) {
          unsigned bits = bitmap[wd];
          const int rowi = wd >> 1;
          const unsigned hi = ((unsigned)(rowi % d1) << 8) | ((unsigned)(rowi / d1) << 16) | ((unsigned)(wd & 1) << 5);
          while (bits) {
            int bi = __ffs(bits) - 1;
            bits &= bits - 1;
            eval_voxel(hi | (unsigned)bi, 2, 3);
          }
        }
      }
      __syncthreads();
      shadow_channel(2, tileA);
      __syncthreads();
      for (int k = tid; k < MAXPIX / 32; k += NT_IMG) sm.occf[k] = 0u;
    }
    __syncthreads();
    PHASE(7);
    // ---- the point planes come back from the image's memory into the dead tiles, then dilation + pixel assembly
    for (int k = tid; k < (npp * PLB) >> 4; k += NT_IMG)
      reinterpret_cast<uint4 *>(tplanes)[k] = __ldcg(reinterpret_cast<const uint4 *>(gimg) + k);  // written by this CTA: L2, not L1
    __syncthreads();
    {
      uint4 *gout = reinterpret_cast<uint4 *>(gimg);
      for (int g = tid; g < S * RW; g += NT_IMG) {
        const int row = g / RW, c4 = g - row * RW;
        unsigned res[16];
        const bool up = row > 0, dn = row + 1 < S, lf = c4 > 0, rt = c4 + 1 < RW;
#pragma unroll
        for (int ch = 0; ch < 16; ch++) {
          res[ch] = 0u;
          if (ch < C) {
            // reference channel ch -> its plane: point channels in the tile region, shadow channels in splanes
            const uint8_t *pl = (C == 15) ? ((ch % 5 == 4) ? splanes + (size_t)(ch / 5) * PLB : tplanes + (size_t)(4 * (ch / 5) + ch % 5) * PLB)
                                          : tplanes + (size_t)ch * PLB;
            const unsigned *W = reinterpret_cast<const unsigned *>(pl) + row * RW + c4;
            const unsigned m0 = W[0], u0 = up ? W[-RW] : 0u, d0w = dn ? W[RW] : 0u;
            const unsigned ml = lf ? W[-1] : 0u, ul = (up && lf) ? W[-RW - 1] : 0u, dl = (dn && lf) ? W[RW - 1] : 0u;
            const unsigned mr = rt ? W[1] : 0u, ur = (up && rt) ? W[-RW + 1] : 0u, dr = (dn && rt) ? W[RW + 1] : 0u;
            const unsigned E = __vimax3_u16x2(__byte_perm(u0, 0u, 0x4240), __byte_perm(m0, 0u, 0x4240), __byte_perm(d0w, 0u, 0x4240));
            const unsigned O = __vimax3_u16x2(__byte_perm(u0, 0u, 0x4341), __byte_perm(m0, 0u, 0x4341), __byte_perm(d0w, 0u, 0x4341));
            const unsigned LO = __vimax3_u16x2(__byte_perm(ul, 0u, 0x4341), __byte_perm(ml, 0u, 0x4341), __byte_perm(dl, 0u, 0x4341));
            const unsigned RE = __vimax3_u16x2(__byte_perm(ur, 0u, 0x4240), __byte_perm(mr, 0u, 0x4240), __byte_perm(dr, 0u, 0x4240));
            const unsigned En = __vimax3_u16x2(E, O, __byte_perm(LO, O, 0x5432));
            const unsigned On = __vimax3_u16x2(O, E, __byte_perm(E, RE, 0x5432));
            res[ch] = __byte_perm(En, On, 0x6240);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const unsigned sel = (unsigned)i | ((unsigned)(4 + i) << 4);
          uint4 o;
          o.x = __byte_perm(__byte_perm(res[0], res[1], sel), __byte_perm(res[2], res[3], sel), 0x5410);
          o.y = __byte_perm(__byte_perm(res[4], res[5], sel), __byte_perm(res[6], res[7], sel), 0x5410);
          o.z = __byte_perm(__byte_perm(res[8], res[9], sel), __byte_perm(res[10], res[11], sel), 0x5410);
          o.w = __byte_perm(__byte_perm(res[12], res[13], sel), __byte_perm(res[14], res[15], sel), 0x5410);
          gout[row * S + c4 * 4 + i] = o;
        }
      }
    }
    PHASE(8);
  }
}

// P16 (16-byte pixels) <-> HWC (the cv::Mat layout, C bytes per pixel)
__global__ void k_p16_to_hwc(const uint8_t *__restrict__ p16, size_t npix, int C, uint8_t *__restrict__ hwc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one output byte per thread
  if (i >= npix * (size_t)C) return;
  const size_t pix = i / C;
  hwc[i] = p16[pix * 16 + (i - pix * C)];
}
__global__ void k_hwc_to_p16(const uint8_t *__restrict__ hwc, size_t npix, int C, uint8_t *__restrict__ p16) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one pixel per thread
  if (i >= npix) return;
  unsigned w[4] = {0u, 0u, 0u, 0u};
  for (int c = 0; c < C; c++) w[c >> 2] |= (unsigned)hwc[i * C + c] << (8 * (c & 3));
  reinterpret_cast<uint4 *>(p16)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace

// ================================================================================================
// host launchers
// ================================================================================================
#define LAUNCH_CHECK()                                   \
  do {                                                   \
    ctx->launches++;                                     \
    cudaError_t e__ = cudaGetLastError();                \
    if (e__ != cudaSuccess) {                            \
      gpdb_set_error(ctx, GPDB_ERR_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return GPDB_ERR_CUDA;                              \
    }                                                    \
  } while (0)

int geo_build_grid(gpdb_ctx *ctx, const float lo[3], const float hi[3], int N) {
  DevParams &hp = ctx->hp;
  float cell = 0.02f;
  size_t ncell;
  for (;;) {
    double nc = 1;
    for (int a = 0; a < 3; a++) {
      hp.dim[a] = (int)floorf((hi[a] - lo[a]) / cell) + 2;
      nc *= hp.dim[a];
    }
    if (nc <= 48e6) { ncell = (size_t)nc; break; }
    cell *= 1.5f;
  }
  for (int a = 0; a < 3; a++) hp.lo[a] = lo[a];
  hp.inv_cell = 1.0f / cell;
  hp.N = N;
  CUDA_TRY(cudaMemcpyAsync(ctx->dp, &hp, sizeof(DevParams), cudaMemcpyHostToDevice, ctx->stream));
  if (ncell + 1 > ctx->cell_cap) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_cell_start);
    ctx->d_cell_start = nullptr;
    ctx->cell_cap = 0;
    CUDA_TRY(cudaMalloc(&ctx->d_cell_start, sizeof(int) * (ncell + 1 + ncell / 4)));
    ctx->cell_cap = ncell + 1 + ncell / 4;
  }
  CUDA_TRY(cudaMemsetAsync(ctx->d_cell_start, 0, sizeof(int) * (ncell + 1), ctx->stream));
  int *cid = (int *)gpdb_scratch(ctx, 0, sizeof(int) * (size_t)N * 4);
  if (!cid) return GPDB_ERR_CUDA;
  int *idx = cid + N, *cid2 = idx + N, *idx2 = cid2 + N;
  const int tb = 256, gb = (N + tb - 1) / tb;
  k_cell_ids<<<gb, tb, 0, ctx->stream>>>(ctx->dp, ctx->d_xyz, N, cid, idx, ctx->d_cell_start);
  LAUNCH_CHECK();
  size_t tmp_bytes = 0, tmp2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, cid, cid2, idx, idx2, N, 0, 32, ctx->stream);
  cub::DeviceScan::InclusiveSum(nullptr, tmp2, ctx->d_cell_start, ctx->d_cell_start, (int)(ncell + 1), ctx->stream);
  void *tmp = gpdb_scratch(ctx, 1, std::max(tmp_bytes, tmp2));
  if (!tmp) return GPDB_ERR_CUDA;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, cid, cid2, idx, idx2, N, 0, 32, ctx->stream));
  ctx->launches += 4;
  k_fill_sorted<<<gb, tb, 0, ctx->stream>>>(ctx->d_xyz, idx2, N, ctx->d_pts4);
  LAUNCH_CHECK();
  CUDA_TRY(cub::DeviceScan::InclusiveSum(tmp, tmp2, ctx->d_cell_start, ctx->d_cell_start, (int)(ncell + 1), ctx->stream));
  ctx->launches += 2;
  ctx->cloud.cell_start = ctx->d_cell_start;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GPDB_OK;
}

int geo_frames(gpdb_ctx *ctx, const int *d_sidx, int n, double *d_frames, uint8_t *d_valid) {
  if (n <= 0) return GPDB_OK;
  const int cap0 = 128;  // ~44 points at the default nn_radius on a 3 mm cloud
  const size_t smem0 = (size_t)2 * LRF_WARPS * cap0 * sizeof(unsigned long long);
  const size_t smem1 = (size_t)2 * LRF_WARPS * LRF_CAP * sizeof(unsigned long long);
  CUDA_TRY(cudaFuncSetAttribute(k_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  int *ovf = (int *)gpdb_scratch(ctx, 2, sizeof(int) * 2 * ((size_t)n + 1));
  const int g2 = 37;  // tier 2: 148 warps, 2 x 8 B x LRF_CAP_GLOBAL each (38 MB of scratch)
  unsigned long long *gkeys =
      (unsigned long long *)gpdb_scratch(ctx, 23, sizeof(unsigned long long) * 2 * LRF_CAP_GLOBAL * (size_t)g2 * LRF_WARPS);
  if (!ovf || !gkeys) return GPDB_ERR_CUDA;
  int *ovf_count = ovf + n, *ovf2 = ovf + n + 1, *ovf2_count = ovf2 + n;
  CUDA_TRY(cudaMemsetAsync(ovf_count, 0, sizeof(int), ctx->stream));
  CUDA_TRY(cudaMemsetAsync(ovf2_count, 0, sizeof(int), ctx->stream));
  const int grid = (n + LRF_WARPS - 1) / LRF_WARPS;
  k_frames<<<grid, LRF_WARPS * 32, smem0, ctx->stream>>>(ctx->dp, ctx->cloud, d_sidx, n, d_frames, d_valid, ctx->d_err, cap0, ovf,
                                                       ovf_count, ovf2, ovf2_count, nullptr, 0);
  LAUNCH_CHECK();
  // overflow tiers over the (usually empty) lists: the tier-1 grid is sized for the worst case (every sample overflowed) so
  // no host round trip is needed; warps beyond the list length exit at once
  k_frames<<<grid, LRF_WARPS * 32, smem1, ctx->stream>>>(ctx->dp, ctx->cloud, d_sidx, n, d_frames, d_valid, ctx->d_err, LRF_CAP, ovf,
                                                       ovf_count, ovf2, ovf2_count, nullptr, 1);
  LAUNCH_CHECK();
  k_frames<<<g2, LRF_WARPS * 32, 0, ctx->stream>>>(ctx->dp, ctx->cloud, d_sidx, n, d_frames, d_valid, ctx->d_err, LRF_CAP_GLOBAL, ovf,
                                                 ovf_count, ovf2, ovf2_count, gkeys, 2);
  LAUNCH_CHECK();
  return GPDB_OK;
}

static const int HANDS_CAP1 = 2176, HANDS_CAP2 = 12800;  // tier 1: 4 CTAs per SM (34 KB + 20 KB static each, <= 64 registers)
static const int HANDS_CAP3 = 131072;                    // last tier: neighbourhood staged in global memory (2 MB per CTA)

int geo_hands(gpdb_ctx *ctx, const int *d_sidx, int n, int slot0, const double *d_frames, const uint8_t *d_valid,
              gpdb_pose *d_poses, uint8_t *d_flags) {
  if (n <= 0) return GPDB_OK;
  // per call, not once per process: function attributes belong to the current device's context (one context per GPU)
  CUDA_TRY(cudaFuncSetAttribute(k_hands, cudaFuncAttributeMaxDynamicSharedMemorySize, HANDS_CAP2 * 16));
  int *ovf = (int *)gpdb_scratch(ctx, 2, sizeof(int) * 2 * ((size_t)n + 1));
  float4 *glist = (float4 *)gpdb_scratch(ctx, 21, sizeof(float4) * (size_t)HANDS_CAP3 * ctx->sm_count);
  if (!ovf || !glist) return GPDB_ERR_CUDA;
  int *ovf_count = ovf + n, *ovf2 = ovf + n + 1, *ovf2_count = ovf2 + n;
  CUDA_TRY(cudaMemsetAsync(ovf_count, 0, sizeof(int), ctx->stream));
  CUDA_TRY(cudaMemsetAsync(ovf2_count, 0, sizeof(int), ctx->stream));
  k_hands<<<n, NT_HANDS, HANDS_CAP1 * 16, ctx->stream>>>(ctx->dp, ctx->cloud, d_sidx, n, slot0, d_frames, d_valid, d_poses,
                                                         d_flags, HANDS_CAP1, nullptr, nullptr, ovf, ovf_count, nullptr, ctx->d_err);
  LAUNCH_CHECK();
  // large-tile pass over the samples whose neighbourhood did not fit tier 1 (persistent CTAs) ...
  k_hands<<<ctx->sm_count, NT_HANDS, HANDS_CAP2 * 16, ctx->stream>>>(ctx->dp, ctx->cloud, d_sidx, n, slot0, d_frames,
                                                                     d_valid, d_poses, d_flags, HANDS_CAP2, ovf, ovf_count,
                                                                     ovf2, ovf2_count, nullptr, ctx->d_err);
  LAUNCH_CHECK();
  // ... and the last tier over what did not fit that either: neighbourhood in global memory (usually an empty list)
  k_hands<<<ctx->sm_count, NT_HANDS, 16, ctx->stream>>>(ctx->dp, ctx->cloud, d_sidx, n, slot0, d_frames, d_valid, d_poses,
                                                        d_flags, HANDS_CAP3, ovf2, ovf2_count, nullptr, nullptr, glist, ctx->d_err);
  LAUNCH_CHECK();
  return GPDB_OK;
}

int geo_compact(gpdb_ctx *ctx, const gpdb_pose *d_poses, const uint8_t *d_flags, int n_poses, gpdb_pose *d_cand,
                int *d_count) {
  if (n_poses <= 0) {
    CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(int), ctx->stream));
    return GPDB_OK;
  }
  int *f01 = (int *)gpdb_scratch(ctx, 3, sizeof(int) * (size_t)n_poses * 2);
  if (!f01) return GPDB_ERR_CUDA;
  int *pos = f01 + n_poses;
  const int tb = 256, gb = (n_poses + tb - 1) / tb;
  k_flag01<<<gb, tb, 0, ctx->stream>>>(d_flags, n_poses, f01);
  LAUNCH_CHECK();
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, f01, pos, n_poses, ctx->stream);
  void *tmp = gpdb_scratch(ctx, 1, tmp_bytes);
  if (!tmp) return GPDB_ERR_CUDA;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, f01, pos, n_poses, ctx->stream));
  ctx->launches += 2;
  k_scatter<<<gb, tb, 0, ctx->stream>>>(d_poses, f01, pos, n_poses, d_cand, d_count);
  LAUNCH_CHECK();
  return GPDB_OK;
}

// `d_p16`: nc images of S*S 16-byte pixels (see k_images). Fast path: k_images2 (two CTAs per SM) over every image, then
// k_images over the images whose box list (1 024 points) overflowed, then its global-list instance over the images whose box
// holds more than 2 048 points (up to 32 768; the overflow lists are usually empty: those launches return at once);
// k_images alone when the geometry is outside the fast path's limits (image_size != 60, more than two cameras) or when
// GPD_B200_IMAGES_KERNEL=1 forces it (tests compare the kernels).
int geo_images(gpdb_ctx *ctx, const gpdb_pose *d_cand, int nc, uint8_t *d_p16) {
  if (nc <= 0) return GPDB_OK;
  const DevParams &hp = ctx->hp;
  const int S = hp.S, RS = (S + 3) & ~3;
  const size_t plane_bytes = ((size_t)hp.C * S * RS + 15) / 16 * 16;
  const size_t tiles = (size_t)3 * 8 * S * S;
  const size_t bm = (size_t)hp.K * (2 * (size_t)hp.bm_dim * hp.bm_dim) * 4;
  // shadow phase: the bitmaps alias the box list and the compacted voxel list follows them: room for >= 4 k voxels
  const size_t list_bytes = (std::max((size_t)BOX_CAP * 36, hp.C == 15 ? bm + 4096 * 4 : (size_t)0) + 15) / 16 * 16;
  const size_t smem = plane_bytes + tiles + list_bytes;
  if (smem > 219 * 1024) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "image geometry needs %zu B of shared memory per CTA (max 224256)", smem);
    return GPDB_ERR_INVALID;
  }
  const char *force = getenv("GPD_B200_IMAGES_KERNEL");
  const bool fast = S == 60 && (hp.C != 15 || (hp.K <= 2 && bm + 2048 <= (size_t)BOX_CAP2 * 36)) && !(force && force[0] == '1');
  const int *d_work = nullptr, *d_work_n = nullptr;
  if (fast) {
    int *ovf = (int *)gpdb_scratch(ctx, 22, sizeof(int) * ((size_t)nc + 1));  // not slot 2: the hand search of the next chunk
    if (!ovf) return GPDB_ERR_CUDA;                                            // may run concurrently on its own stream
    int *ovf_count = ovf + nc;
    CUDA_TRY(cudaMemsetAsync(ovf_count, 0, sizeof(int), ctx->stream));
    const size_t smem2 = (size_t)2 * 8 * S * S + (size_t)3 * S * S + (size_t)BOX_CAP2 * 36;
    CUDA_TRY(cudaFuncSetAttribute(k_images2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    k_images2<<<std::min(nc, ctx->sm_count * 64), NT_IMG, smem2, ctx->stream>>>(ctx->dp, ctx->cloud, d_cand, nc, d_p16, ctx->d_qtab, ovf,
                                                                               ovf_count, ctx->d_prof);
    LAUNCH_CHECK();
    d_work = ovf;
    d_work_n = ovf_count;
  }
  // general tier (2 048-point box list in shared memory): everything, or the overflow list of the fast path; its own
  // overflow goes to a second list ...
  int *ovf2 = (int *)gpdb_scratch(ctx, 20, sizeof(int) * ((size_t)nc + 1));
  const int gl_cap = 32768;
  const size_t gl_slice = (size_t)gl_cap * 36 + (size_t)16 * S * S;
  unsigned char *gl = (unsigned char *)gpdb_scratch(ctx, 19, gl_slice * (size_t)ctx->sm_count);
  if (!ovf2 || !gl) return GPDB_ERR_CUDA;
  int *ovf2_count = ovf2 + nc;
  CUDA_TRY(cudaMemsetAsync(ovf2_count, 0, sizeof(int), ctx->stream));
  const int grid = fast ? ctx->sm_count : std::min(nc, ctx->sm_count * 64);
  if (S == 60) {
    CUDA_TRY(cudaFuncSetAttribute(k_images<60, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(k_images<60, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_images<60, false><<<grid, NT_IMG, smem, ctx->stream>>>(ctx->dp, ctx->cloud, d_cand, nc, d_p16, ctx->d_qtab, ctx->d_err,
                                                             (int)plane_bytes, (int)list_bytes, ctx->d_prof, d_work, d_work_n,
                                                             nullptr, 0, ovf2, ovf2_count);
    LAUNCH_CHECK();
    // ... which the last tier redoes with the box list in global memory (32 768 points; beyond: GPDB_ERR_CAPACITY)
    k_images<60, true><<<ctx->sm_count, NT_IMG, smem, ctx->stream>>>(ctx->dp, ctx->cloud, d_cand, nc, d_p16, ctx->d_qtab, ctx->d_err,
                                                                     (int)plane_bytes, (int)list_bytes, ctx->d_prof, ovf2, ovf2_count,
                                                                     gl, gl_cap, nullptr, nullptr);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(k_images<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(k_images<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_images<0, false><<<grid, NT_IMG, smem, ctx->stream>>>(ctx->dp, ctx->cloud, d_cand, nc, d_p16, ctx->d_qtab, ctx->d_err,
                                                            (int)plane_bytes, (int)list_bytes, ctx->d_prof, d_work, d_work_n,
                                                            nullptr, 0, ovf2, ovf2_count);
    LAUNCH_CHECK();
    k_images<0, true><<<ctx->sm_count, NT_IMG, smem, ctx->stream>>>(ctx->dp, ctx->cloud, d_cand, nc, d_p16, ctx->d_qtab, ctx->d_err,
                                                                    (int)plane_bytes, (int)list_bytes, ctx->d_prof, ovf2, ovf2_count,
                                                                    gl, gl_cap, nullptr, nullptr);
  }
  LAUNCH_CHECK();
  return GPDB_OK;
}

int geo_p16_to_hwc(gpdb_ctx *ctx, const uint8_t *d_p16, int n, uint8_t *d_hwc) {
  if (n <= 0) return GPDB_OK;
  const size_t npix = (size_t)n * ctx->hp.S * ctx->hp.S, nb = npix * ctx->hp.C;
  k_p16_to_hwc<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(d_p16, npix, ctx->hp.C, d_hwc);
  LAUNCH_CHECK();
  return GPDB_OK;
}
int geo_hwc_to_p16(gpdb_ctx *ctx, const uint8_t *d_hwc, int n, uint8_t *d_p16) {
  if (n <= 0) return GPDB_OK;
  const size_t npix = (size_t)n * ctx->hp.S * ctx->hp.S;
  k_hwc_to_p16<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(d_hwc, npix, ctx->hp.C, d_p16);
  LAUNCH_CHECK();
  return GPDB_OK;
}

int geo_reeval(gpdb_ctx *ctx, gpdb_pose *d_hands, int n, int *d_labels) {
  if (n <= 0) return GPDB_OK;
  k_reeval<<<(n * 32 + 127) / 128, 128, 0, ctx->stream>>>(ctx->dp, ctx->cloud, d_hands, n, d_labels);
  LAUNCH_CHECK();
  return GPDB_OK;
}

int geo_clusters(gpdb_ctx *ctx, const gpdb_pose *d_hands, int n, int min_inliers, gpdb_pose *d_dense, uint8_t *d_keep) {
  if (n <= 0) return GPDB_OK;
  const double cos_thresh = std::cos(12.0 * M_PI / 180.0);  // AXIS_ALIGN_ANGLE_THRESH (clustering.cpp:9)
  k_clusters<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(d_hands, n, min_inliers, cos_thresh, d_dense, d_keep);
  LAUNCH_CHECK();
  return GPDB_OK;
}

int geo_select(gpdb_ctx *ctx, const gpdb_pose *d_cand, int n, int k, gpdb_pose *d_out) {
  if (n <= 0 || k <= 0) return GPDB_OK;
  unsigned *keys = (unsigned *)gpdb_scratch(ctx, 3, sizeof(unsigned) * (size_t)n * 4);
  if (!keys) return GPDB_ERR_CUDA;
  unsigned *keys2 = keys + n;
  int *vals = (int *)(keys2 + n), *vals2 = vals + n;
  k_select_keys<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_cand, n, keys, vals);
  LAUNCH_CHECK();
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, n, 0, 32, ctx->stream);
  void *tmp = gpdb_scratch(ctx, 1, tmp_bytes);
  if (!tmp) return GPDB_ERR_CUDA;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, n, 0, 32, ctx->stream));
  ctx->launches += 4;
  k_gather_poses<<<(k * 32 + 255) / 256, 256, 0, ctx->stream>>>(d_cand, vals2, k, d_out);
  LAUNCH_CHECK();
  return GPDB_OK;
}

int geo_scatter_scores(gpdb_ctx *ctx, const gpdb_pose *d_cand, const float *d_scores, int nc, int slot0, int P,
                       float *d_pose_scores, gpdb_pose *d_cand_out) {
  if (nc <= 0) return GPDB_OK;
  k_scatter_scores<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(d_cand, d_scores, nc, slot0, P, d_pose_scores, d_cand_out);
  LAUNCH_CHECK();
  return GPDB_OK;
}
