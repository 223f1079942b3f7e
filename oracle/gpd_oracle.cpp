/*
 * gpd_oracle.cpp — CPU ORACLE of the atenpas/gpd grasp-candidate hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load liboracle. The product
 * (libgpd_b200.so) never links, loads or calls anything in this directory.
 *
 * It is a dependency-free C++17 + OpenMP RESTATEMENT of the reference algorithm (the reference
 * itself needs PCL/Eigen/OpenCV, none of which exist in this image, so it cannot be compiled
 * here — see DESIGN.md "oracle"). It keeps the reference's stage structure, loop structure,
 * float64/float32 choices and quirks, each function citing the reference file:line it follows
 * (paths relative to /root/reference). Third-party arithmetic the reference delegates to
 * libraries absent from /root/reference is restated from the libraries' published algorithms:
 *   - Eigen (unpinned ">= 3.0"; restated from 3.3/3.4): SelfAdjointEigenSolver<Matrix3d>,
 *     AngleAxisd::toRotationMatrix, VectorXd::LinSpaced, dense products (fixed left-to-right
 *     summation order, no FMA contraction: compile with -ffp-contract=off);
 *   - PCL >= 1.9 / FLANN: KdTreeFLANN::radiusSearch = float32 L2_Simple distance, strict
 *     dist < (float)(r*r), results sorted by (dist, index);
 *   - OpenCV >= 3.4: dilate(3x3 rect, border ignored), normalize(NORM_MINMAX), minMaxLoc(mask),
 *     convertTo(CV_8U, 255) — these restatements ARE pinned against the real library through
 *     Python cv2 in tests/test_oracle.py (fixtures: tools/make_goldens.py);
 *   - the LeNet restatement is pinned against cv2.dnn running the reference's own
 *     .prototxt/.caffemodel (tests/golden/, tools/make_goldens.py).
 * PARITY STATUS: the reference ships no golden vectors or asserting tests for this path
 * (SURVEY.md section 4), so the geometry stages and the cloud preprocessing (PCL 1.9.1 normal
 * estimation restated from its published algorithm, see "Cloud preprocessing" below) are
 * "parity unpinned" against upstream binaries; what is pinned is listed above.
 *
 * The 15-channel shadow follows the deterministic variant specified in
 * include/gpd_b200_shadow.h (the reference's is racy and random by construction).
 */
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#include <random>
#endif

#include "../include/gpd_b200.h"
#include "../include/gpd_b200_shadow.h"

namespace {

double now_s() {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return (double)clock() / CLOCKS_PER_SEC;
#endif
}

// ------------------------------------------------------------------------------------------
// Cloud + neighbour search (A0). Reference: pcl::KdTreeFLANN::radiusSearch call sites
// frame_estimator.cpp:74, hand_search.cpp:178, image_generator.cpp:61.
// The search structure is a uniform grid; the PREDICATE and the ORDER are FLANN's.
// ------------------------------------------------------------------------------------------
struct Cloud {
  int N = 0, K = 0;
  std::vector<float> xyz;    // 3*N packed
  std::vector<double> nrm;   // 3 x N column-major
  std::vector<int32_t> cam;  // K x N column-major
  std::vector<double> vp;    // 3 x K column-major
  float cell = 0.02f;
  float lo[3] = {0, 0, 0};
  int dim[3] = {1, 1, 1};
  std::vector<int> cell_start;  // ncell + 1
  std::vector<int> order;       // point indices sorted by cell
  // Cloud::setSamples (cloud.cpp:662): arbitrary sample positions (3 x n, float64). A sample index >= N addresses
  // samples[index - N]; indices < N address the cloud points themselves (Cloud::getSampleIndices).
  std::vector<double> samples;
  void sample_position(int sample_index, double out[3]) const {
    if (sample_index < N) {
      for (int a = 0; a < 3; a++) out[a] = (double)xyz[3 * (size_t)sample_index + a];
    } else {
      for (int a = 0; a < 3; a++) out[a] = samples[3 * (size_t)(sample_index - N) + a];
    }
  }

  int cell_of(float v, int a) const {
    int c = (int)std::floor((v - lo[a]) / cell);
    return std::min(std::max(c, 0), dim[a] - 1);
  }
  void build() {
    float hi[3];
    for (int a = 0; a < 3; a++) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
    for (int i = 0; i < N; i++)
      for (int a = 0; a < 3; a++) {
        lo[a] = std::min(lo[a], xyz[3 * i + a]);
        hi[a] = std::max(hi[a], xyz[3 * i + a]);
      }
    if (N == 0) { lo[0] = lo[1] = lo[2] = 0; hi[0] = hi[1] = hi[2] = 0; }
    cell = 0.02f;
    for (;;) {
      double nc = 1;
      for (int a = 0; a < 3; a++) {
        dim[a] = std::max(1, (int)std::floor((hi[a] - lo[a]) / cell) + 1);
        nc *= dim[a];
      }
      if (nc <= 32e6) break;
      cell *= 2;
    }
    size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
    cell_start.assign(ncell + 1, 0);
    std::vector<int> cid(N);
    for (int i = 0; i < N; i++) {
      int c = (cell_of(xyz[3 * i + 2], 2) * dim[1] + cell_of(xyz[3 * i + 1], 1)) * dim[0] +
              cell_of(xyz[3 * i], 0);
      cid[i] = c;
      cell_start[c + 1]++;
    }
    for (size_t c = 0; c < ncell; c++) cell_start[c + 1] += cell_start[c];
    order.resize(N);
    std::vector<int> fill(cell_start.begin(), cell_start.end() - 1);
    for (int i = 0; i < N; i++) order[fill[cid[i]]++] = i;
  }
};

struct Nb {
  float d;
  int i;
};

// FLANN L2_Simple<float>: result += diff*diff over the 3 dims, float32, strict '<' against
// (float)(radius*radius) (PCL kdtree_flann.hpp radiusSearch casts radius*radius to float),
// sorted ascending by (dist, index).
void radius_search(const Cloud &c, const float q[3], double radius, std::vector<Nb> &out) {
  out.clear();
  const float r2 = (float)(radius * radius);
  const float rf = (float)radius * 1.0001f + 1e-6f;
  int c0[3], c1[3];
  for (int a = 0; a < 3; a++) {
    c0[a] = c.cell_of(q[a] - rf, a);
    c1[a] = c.cell_of(q[a] + rf, a);
  }
  for (int cz = c0[2]; cz <= c1[2]; cz++)
    for (int cy = c0[1]; cy <= c1[1]; cy++) {
      size_t row = ((size_t)cz * c.dim[1] + cy) * c.dim[0];
      int s = c.cell_start[row + c0[0]], e = c.cell_start[row + c1[0] + 1];
      for (int k = s; k < e; k++) {
        int i = c.order[k];
        float dx = q[0] - c.xyz[3 * i], dy = q[1] - c.xyz[3 * i + 1], dz = q[2] - c.xyz[3 * i + 2];
        float d = dx * dx;
        d += dy * dy;
        d += dz * dz;
        if (d < r2) out.push_back({d, i});
      }
    }
  std::sort(out.begin(), out.end(),
            [](const Nb &a, const Nb &b) { return a.d < b.d || (a.d == b.d && a.i < b.i); });
}

// ------------------------------------------------------------------------------------------
// 3x3 symmetric eigen-solver = Eigen::SelfAdjointEigenSolver<Matrix3d>::compute (iterative
// path), restated from Eigen 3.3/3.4 (Tridiagonalization.h tridiagonalization_inplace_selector
// <M,3,false>, SelfAdjointEigenSolver.h computeFromTridiagonal_impl / tridiagonal_qr_step,
// Jacobi.h makeGivens). Used by LocalFrame::findAverageNormalAxis (local_frame.cpp:18-21).
// M, evec: column-major 3x3. Only the lower triangle of M is read.
// ------------------------------------------------------------------------------------------
struct Givens {
  double c, s;
};
Givens make_givens(double p, double q) {
  Givens g;
  if (q == 0.0) {
    g.c = p < 0.0 ? -1.0 : 1.0;
    g.s = 0.0;
  } else if (p == 0.0) {
    g.c = 0.0;
    g.s = q < 0.0 ? 1.0 : -1.0;
  } else if (std::fabs(p) > std::fabs(q)) {
    double t = q / p;
    double u = std::sqrt(1.0 + t * t);
    if (p < 0.0) u = -u;
    g.c = 1.0 / u;
    g.s = -t * g.c;
  } else {
    double t = p / q;
    double u = std::sqrt(1.0 + t * t);
    if (q < 0.0) u = -u;
    g.s = -1.0 / u;
    g.c = -t * g.s;
  }
  return g;
}
double eigen_hypot(double x, double y) {  // Eigen numext::hypot (positive_real_hypot)
  double ax = std::fabs(x), ay = std::fabs(y);
  double p = std::max(ax, ay);
  if (p == 0.0) return 0.0;
  double qp = std::min(ax, ay) / p;
  return p * std::sqrt(1.0 + qp * qp);
}
#define Mx(r, c) m[(c)*3 + (r)]
void eigen3(const double *Min, double *eval, double *evec) {
  double m[9];
  // mat = lower triangle; scale = max |coeff| of it
  double scale = 0.0;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) {
      Mx(r, c) = (r >= c) ? Min[c * 3 + r] : 0.0;
      if (r >= c) scale = std::max(scale, std::fabs(Mx(r, c)));
    }
  if (scale == 0.0) scale = 1.0;
  for (int c = 0; c < 3; c++)
    for (int r = c; r < 3; r++) Mx(r, c) /= scale;
  double diag[3], sub[2];
  double Q[9];  // column-major
  const double tol = std::numeric_limits<double>::min();
  diag[0] = Mx(0, 0);
  double v1norm2 = Mx(2, 0) * Mx(2, 0);
  if (v1norm2 <= tol) {
    diag[1] = Mx(1, 1);
    diag[2] = Mx(2, 2);
    sub[0] = Mx(1, 0);
    sub[1] = Mx(2, 1);
    for (int i = 0; i < 9; i++) Q[i] = (i % 4 == 0) ? 1.0 : 0.0;
  } else {
    double beta = std::sqrt(Mx(1, 0) * Mx(1, 0) + v1norm2);
    double invBeta = 1.0 / beta;
    double m01 = Mx(1, 0) * invBeta;
    double m02 = Mx(2, 0) * invBeta;
    double q = 2.0 * m01 * Mx(2, 1) + m02 * (Mx(2, 2) - Mx(1, 1));
    diag[1] = Mx(1, 1) + m02 * q;
    diag[2] = Mx(2, 2) - m02 * q;
    sub[0] = beta;
    sub[1] = Mx(2, 1) - m01 * q;
    // Q << 1,0,0, 0,m01,m02, 0,m02,-m01 (row-wise fill) -> column-major storage
    Q[0] = 1; Q[1] = 0;   Q[2] = 0;
    Q[3] = 0; Q[4] = m01; Q[5] = m02;
    Q[6] = 0; Q[7] = m02; Q[8] = -m01;
  }
  const int n = 3;
  int end = n - 1, start = 0, iter = 0;
  const int maxIterations = 30;
  const double considerAsZero = std::numeric_limits<double>::min();
  const double precision_inv = 1.0 / std::numeric_limits<double>::epsilon();
  while (end > 0) {
    for (int i = start; i < end; ++i) {
      if (std::fabs(sub[i]) < considerAsZero) {
        sub[i] = 0.0;
      } else {
        const double scaled = precision_inv * sub[i];
        if (scaled * scaled <= (std::fabs(diag[i]) + std::fabs(diag[i + 1]))) sub[i] = 0.0;
      }
    }
    while (end > 0 && sub[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIterations * n) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0) start--;
    // tridiagonal_qr_step
    double td = (diag[end - 1] - diag[end]) * 0.5;
    double e = sub[end - 1];
    double mu = diag[end];
    if (td == 0.0) {
      mu -= std::fabs(e);
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
      Givens rot = make_givens(x, z);
      double sdk = rot.s * diag[k] + rot.c * sub[k];
      double dkp1 = rot.s * sub[k] + rot.c * diag[k + 1];
      diag[k] = rot.c * (rot.c * diag[k] - rot.s * sub[k]) - rot.s * (rot.c * sub[k] - rot.s * diag[k + 1]);
      diag[k + 1] = rot.s * sdk + rot.c * dkp1;
      sub[k] = rot.c * sdk - rot.s * dkp1;
      if (k > start) sub[k - 1] = rot.c * sub[k - 1] - rot.s * z;
      x = sub[k];
      if (k < end - 1) {
        z = -rot.s * sub[k + 1];
        sub[k + 1] = rot.c * sub[k + 1];
      }
      // Q = Q * G on columns k, k+1: x' = c x - s y ; y' = s x + c y
      for (int r = 0; r < 3; r++) {
        double xi = Q[k * 3 + r], yi = Q[(k + 1) * 3 + r];
        Q[k * 3 + r] = rot.c * xi - rot.s * yi;
        Q[(k + 1) * 3 + r] = rot.s * xi + rot.c * yi;
      }
    }
  }
  // selection sort ascending (first minimum), swapping columns
  for (int i = 0; i < n - 1; ++i) {
    int k = 0;
    for (int j = 1; j < n - i; j++)
      if (diag[i + j] < diag[i + k]) k = j;
    if (k > 0) {
      std::swap(diag[i], diag[k + i]);
      for (int r = 0; r < 3; r++) std::swap(Q[i * 3 + r], Q[(k + i) * 3 + r]);
    }
  }
  for (int i = 0; i < 3; i++) eval[i] = diag[i] * scale;
  for (int i = 0; i < 9; i++) evec[i] = Q[i];
}
#undef Mx

// Eigen::AngleAxisd(angle, unit axis).toRotationMatrix() (Eigen/src/Geometry/AngleAxis.h);
// out column-major. Used at hand_set.cpp:52-53,68-69.
void angle_axis_matrix(double angle, const double axis[3], double *R) {
  double s = std::sin(angle), c = std::cos(angle);
  double sin_axis[3] = {s * axis[0], s * axis[1], s * axis[2]};
  double cos1_axis[3] = {(1.0 - c) * axis[0], (1.0 - c) * axis[1], (1.0 - c) * axis[2]};
#define Rm(r, cc) R[(cc)*3 + (r)]
  double tmp;
  tmp = cos1_axis[0] * axis[1];
  Rm(0, 1) = tmp - sin_axis[2];
  Rm(1, 0) = tmp + sin_axis[2];
  tmp = cos1_axis[0] * axis[2];
  Rm(0, 2) = tmp + sin_axis[1];
  Rm(2, 0) = tmp - sin_axis[1];
  tmp = cos1_axis[1] * axis[2];
  Rm(1, 2) = tmp - sin_axis[0];
  Rm(2, 1) = tmp + sin_axis[0];
  Rm(0, 0) = cos1_axis[0] * axis[0] + c;
  Rm(1, 1) = cos1_axis[1] * axis[1] + c;
  Rm(2, 2) = cos1_axis[2] * axis[2] + c;
#undef Rm
}

// 3x3 product C = A*B, column-major, each entry summed left to right.
void mat3_mul(const double *A, const double *B, double *C) {
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++)
      C[c * 3 + r] = (A[0 * 3 + r] * B[c * 3 + 0] + A[1 * 3 + r] * B[c * 3 + 1]) + A[2 * 3 + r] * B[c * 3 + 2];
}

// Eigen 3.3 VectorXd::LinSpaced(size, low, high)(i) (NullaryFunctors.h linspaced_op_impl).
double linspaced(int size, double low, double high, int i) {
  int size1 = size == 1 ? 1 : size - 1;
  double step = size == 1 ? 0.0 : (high - low) / (double)(size - 1);
  bool flip = std::fabs(high) < std::fabs(low);
  if (flip) return (i == 0) ? low : (high - (double)(size1 - i) * step);
  return (i == size1) ? high : (low + (double)i * step);
}

// ------------------------------------------------------------------------------------------
// A1-A2: FrameEstimator::calculateFrame (frame_estimator.cpp:66-86) +
// LocalFrame::findAverageNormalAxis (local_frame.cpp:14-41).
// frame9 = normal | binormal | curvature_axis (3 column vectors).
// ------------------------------------------------------------------------------------------
bool calc_frame(const Cloud &c, const float q[3], double radius, double *frame9, std::vector<Nb> &nn) {
  radius_search(c, q, radius, nn);
  if (nn.empty()) return false;
  double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double sum[3] = {0, 0, 0};
  for (const Nb &nb : nn) {
    const double *n = &c.nrm[3 * (size_t)nb.i];
    for (int cc = 0; cc < 3; cc++)
      for (int r = 0; r < 3; r++) M[cc * 3 + r] += n[r] * n[cc];
    for (int r = 0; r < 3; r++) sum[r] += n[r];
  }
  double eval[3], evec[9];
  eigen3(M, eval, evec);
  int mn = 0, mx = 0;  // Eigen minCoeff/maxCoeff: first index on ties
  for (int i = 1; i < 3; i++) {
    if (eval[i] < eval[mn]) mn = i;
    if (eval[i] > eval[mx]) mx = i;
  }
  double curv[3] = {evec[mn * 3], evec[mn * 3 + 1], evec[mn * 3 + 2]};
  double normal[3] = {evec[mx * 3], evec[mx * 3 + 1], evec[mx * 3 + 2]};
  double nrm = std::sqrt((sum[0] * sum[0] + sum[1] * sum[1]) + sum[2] * sum[2]);
  double avg[3] = {sum[0] / nrm, sum[1] / nrm, sum[2] / nrm};
  if ((avg[0] * normal[0] + avg[1] * normal[1]) + avg[2] * normal[2] < 0) {
    normal[0] *= -1.0;
    normal[1] *= -1.0;
    normal[2] *= -1.0;
  }
  double bin[3] = {curv[1] * normal[2] - curv[2] * normal[1], curv[2] * normal[0] - curv[0] * normal[2],
                   curv[0] * normal[1] - curv[1] * normal[0]};
  for (int r = 0; r < 3; r++) {
    frame9[r] = normal[r];
    frame9[3 + r] = bin[r];
    frame9[6 + r] = curv[r];
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// util::PointList restated (point_list.cpp). Only what the path uses.
// ------------------------------------------------------------------------------------------
struct PointList {
  std::vector<double> p;  // 3 x n column-major
  std::vector<double> nr;
  std::vector<int> gidx;  // global cloud index per column
  int size() const { return (int)gidx.size(); }
};

// PointList::slice of the cloud by neighbour indices (hand_search.cpp:179; points are the
// float32 cloud cast to double, hand_search.cpp:160-161).
void slice_cloud(const Cloud &c, const std::vector<Nb> &nn, PointList &out) {
  int n = (int)nn.size();
  out.p.resize(3 * (size_t)n);
  out.nr.resize(3 * (size_t)n);
  out.gidx.resize(n);
  for (int j = 0; j < n; j++) {
    int i = nn[j].i;
    out.gidx[j] = i;
    for (int r = 0; r < 3; r++) {
      out.p[3 * (size_t)j + r] = (double)c.xyz[3 * (size_t)i + r];
      out.nr[3 * (size_t)j + r] = c.nrm[3 * (size_t)i + r];
    }
  }
}

// PointList::transformToHandFrame (point_list.cpp:22-33): R^T (p - centroid), R^T n,
// with `rot` = frame (column-major), i.e. rotation = frame^T.
inline void to_frame(const double *frame, const double *v, double *o) {
  for (int k = 0; k < 3; k++) o[k] = (frame[k * 3 + 0] * v[0] + frame[k * 3 + 1] * v[1]) + frame[k * 3 + 2] * v[2];
}
void transform_to_hand_frame(const PointList &in, const double *centroid, const double *frame, PointList &out) {
  int n = in.size();
  out.p.resize(3 * (size_t)n);
  out.nr.resize(3 * (size_t)n);
  out.gidx = in.gidx;
  for (int j = 0; j < n; j++) {
    double cen[3] = {in.p[3 * (size_t)j] - centroid[0], in.p[3 * (size_t)j + 1] - centroid[1],
                     in.p[3 * (size_t)j + 2] - centroid[2]};
    to_frame(frame, cen, &out.p[3 * (size_t)j]);
    to_frame(frame, &in.nr[3 * (size_t)j], &out.nr[3 * (size_t)j]);
  }
}

// PointList::cropByHandHeight (point_list.cpp:44-55) INCLUDING the quirk: `indices` has size()
// entries, value-initialised to 0 and never truncated to k, so the result is the k in-range
// columns followed by (n-k) copies of column 0.
void crop_by_hand_height(const PointList &in, double height, PointList &out) {
  int n = in.size();
  std::vector<int> indices(n, 0);
  int k = 0;
  for (int i = 0; i < n; i++) {
    double z = in.p[3 * (size_t)i + 2];
    if (z > -1.0 * height && z < height) indices[k++] = i;
  }
  out.p.resize(3 * (size_t)n);
  out.nr.resize(3 * (size_t)n);
  out.gidx.resize(n);
  for (int j = 0; j < n; j++) {
    int i = indices[j];
    out.gidx[j] = in.gidx[i];
    for (int r = 0; r < 3; r++) {
      out.p[3 * (size_t)j + r] = in.p[3 * (size_t)i + r];
      out.nr[3 * (size_t)j + r] = in.nr[3 * (size_t)i + r];
    }
  }
}

// ------------------------------------------------------------------------------------------
// A5: candidate::FingerHand restated (finger_hand.cpp). forward axis 0, lateral axis 1.
// ------------------------------------------------------------------------------------------
struct FingerHand {
  double finger_width, hand_depth;
  std::vector<double> fs;  // finger_spacing_ (2P)
  std::vector<char> fingers, hand;
  double bottom = 0, top = 0, left = 0, right = 0, center = 0;

  FingerHand(double fw, double od, double depth, int P) : finger_width(fw), hand_depth(depth) {
    // finger_hand.cpp:12-19
    fs.resize(2 * P);
    for (int i = 0; i < P; i++) {
      double h = linspaced(P, 0.0, od - fw, i);
      fs[i] = (h - od) + fw;
      fs[P + i] = h;
    }
    fingers.assign(2 * P, 0);
    hand.assign(P, 0);
  }
  bool gap_free(const PointList &pl, const std::vector<int> &idxs, int f) const {  // :173-184
    for (int i : idxs) {
      double x = pl.p[3 * (size_t)i + 1];
      if (x > fs[f] && x < fs[f] + finger_width) return false;
    }
    return true;
  }
  void evaluate_fingers(const PointList &pl, double bite, int idx = -1) {  // :26-73
    top = bite;
    bottom = bite - hand_depth;
    center = 0.0;
    std::fill(fingers.begin(), fingers.end(), 0);
    std::vector<int> cropped;
    int n = pl.size();
    for (int i = 0; i < n; i++) {
      double x = pl.p[3 * (size_t)i];
      if (x < bite) {
        if (x < bottom) return;
        cropped.push_back(i);
      }
    }
    if (cropped.empty()) return;
    int F = (int)fingers.size();
    if (idx == -1) {
      for (int i = 0; i < F; i++)
        if (gap_free(pl, cropped, i)) fingers[i] = 1;
    } else {
      if (gap_free(pl, cropped, idx)) fingers[idx] = 1;
      if (gap_free(pl, cropped, F / 2 + idx)) fingers[F / 2 + idx] = 1;
    }
  }
  void evaluate_hand() {  // :75-81
    int n = (int)fingers.size() / 2;
    for (int i = 0; i < n; i++) hand[i] = (fingers[i] && fingers[n + i]);
  }
  bool any_hand() const {
    for (char h : hand)
      if (h) return true;
    return false;
  }
  int choose_middle_hand() const {  // :89-105
    std::vector<int> hi;
    for (int i = 0; i < (int)hand.size(); i++)
      if (hand[i]) hi.push_back(i);
    if (hi.empty()) return -1;
    return hi[(int)std::ceil(hi.size() / 2.0) - 1];
  }
  int deepen_hand(const PointList &pl, double min_depth, double max_depth) {  // :107-139
    int e = choose_middle_hand();
    int opp = (int)fingers.size() / 2 + e;
    const double STEP = 0.005;
    FingerHand nh = *this;
    FingerHand last = nh;
    for (double depth = min_depth + STEP; depth <= max_depth; depth += STEP) {
      nh.evaluate_fingers(pl, depth, e);
      if (!nh.fingers[e] || !nh.fingers[opp]) break;
      hand[e] = 1;
      last = nh;
    }
    *this = last;
    std::fill(hand.begin(), hand.end(), 0);
    hand[e] = 1;
    return e;
  }
  std::vector<int> closing_region(const PointList &pl, int idx) {  // :141-171
    if (idx == -1)
      for (int i = 0; i < (int)hand.size(); i++)
        if (hand[i]) { idx = i; break; }
    left = fs[idx] + finger_width;
    right = fs[hand.size() + idx];
    center = 0.5 * (left + right);
    std::vector<int> out;
    int n = pl.size();
    for (int i = 0; i < n; i++) {
      double x = pl.p[3 * (size_t)i], y = pl.p[3 * (size_t)i + 1];
      if (x > bottom && x < top && y > left && y < right) out.push_back(i);
    }
    return out;
  }
};

// A7: Antipodal::evaluateGrasp (antipodal.cpp:10-96) with lateral=1, forward=0, vertical=2.
// Operates on the sliced closing-region list (pts/normals of `idx` columns of pl).
int antipodal_eval(const PointList &pl, const std::vector<int> &idx, double friction_coeff, int min_viable) {
  const double extremal_thresh = 0.003;  // hand_set.cpp:257
  int result = 0;
  double cosf = std::cos(friction_coeff * M_PI / 180.0);
  double mn = DBL_MAX, mx = -DBL_MAX;
  for (int i : idx) {
    double y = pl.p[3 * (size_t)i + 1];
    mn = std::min(mn, y);
    mx = std::max(mx, y);
  }
  double min_x = mn + extremal_thresh, max_x = mx - extremal_thresh;
  std::vector<int> L, R;
  for (int i : idx) {
    const double *nn = &pl.nr[3 * (size_t)i];
    double ldot = (0.0 * nn[0] + -1.0 * nn[1]) + 0.0 * nn[2];
    double rdot = (0.0 * nn[0] + 1.0 * nn[1]) + 0.0 * nn[2];
    double y = pl.p[3 * (size_t)i + 1];
    if (ldot > cosf && y < min_x) L.push_back(i);
    if (rdot > cosf && y > max_x) R.push_back(i);
  }
  if (!L.empty() || !R.empty()) result = 1;
  if (!L.empty() && !R.empty()) {
    auto ext = [&](const std::vector<int> &v, int a, bool mxq) {
      double e = mxq ? -DBL_MAX : DBL_MAX;
      for (int i : v) e = mxq ? std::max(e, pl.p[3 * (size_t)i + a]) : std::min(e, pl.p[3 * (size_t)i + a]);
      return e;
    };
    double top_y = std::min(ext(L, 0, true), ext(R, 0, true));
    double bot_y = std::max(ext(L, 0, false), ext(R, 0, false));
    double top_z = std::min(ext(L, 2, true), ext(R, 2, true));
    double bot_z = std::max(ext(L, 2, false), ext(R, 2, false));
    auto cnt = [&](const std::vector<int> &v) {
      int k = 0;
      for (int i : v) {
        double y = pl.p[3 * (size_t)i], z = pl.p[3 * (size_t)i + 2];
        if (y >= bot_y && y <= top_y && z >= bot_z && z <= top_z) k++;
      }
      return k;
    };
    if (cnt(L) >= min_viable && cnt(R) >= min_viable) result = 2;
  }
  return result;
}

struct Derived {
  int P;                       // poses per sample
  double nn_radius_hs;         // hand_search.cpp:13-17
  double img_radius;           // image_generator.cpp:43-46
  double shadow_length;        // image_15_channels_strategy.h:72-75
  std::vector<double> angles;  // hand_search.cpp:151-155
  double rot_binormal[9];      // hand_set.cpp:52-53
};
Derived derive(const gpdb_params &pr) {
  Derived d;
  d.P = pr.num_hand_axes * pr.num_orientations;
  d.nn_radius_hs = std::max(std::max(pr.hand_outer_diameter - pr.finger_width, pr.hand_depth), pr.hand_height / 2.0);
  d.img_radius = std::max(std::max(pr.volume_depth, pr.volume_height / 2.0), pr.volume_width);
  d.shadow_length = d.img_radius;
  d.angles.resize(pr.num_orientations);
  for (int i = 0; i < pr.num_orientations; i++)
    d.angles[i] = linspaced(pr.num_orientations + 1, -1.0 * M_PI / 2.0, M_PI / 2.0, i);
  const double uy[3] = {0, 1, 0};
  angle_axis_matrix(M_PI, uy, d.rot_binormal);
  return d;
}

void fill_pose_header(gpdb_pose &h, const double *sample, const double *frame_rot, int sample_index, int slot, int pose_slot) {
  std::memset(&h, 0, sizeof(h));
  for (int r = 0; r < 3; r++) h.sample[r] = sample[r];
  for (int r = 0; r < 9; r++) h.frame[r] = frame_rot[r];
  h.sample_index = sample_index;
  h.sample_slot = slot;
  h.pose_slot = (int16_t)pose_slot;
  h.finger_idx = -1;
  h.score = std::numeric_limits<float>::quiet_NaN();
}

// A3-A7: HandSearch::evalHands body for one frame (hand_search.cpp:172-182) =
// HandSet::evalHandSet (hand_set.cpp:31-47) + HandSet::evalHands (:49-116) +
// modifyCandidate/labelHypothesis (:235-261) + Hand::construct (hand.cpp:24-45).
void eval_hand_set(const Cloud &c, const gpdb_params &pr, const Derived &dv, int sample_index, int slot,
                   const double *lframe9, gpdb_pose *poses, uint8_t *flags) {
  // the sample stays float64 for the hand-frame transform; the radius search is run at its float32 image
  // (eigenVectorToPcl, hand_search.cpp:166-170)
  double sample[3];
  c.sample_position(sample_index, sample);
  float q[3] = {(float)sample[0], (float)sample[1], (float)sample[2]};
  for (int j = 0; j < dv.P; j++) flags[j] = 0;
  std::vector<Nb> nn;
  radius_search(c, q, dv.nn_radius_hs, nn);
  // frame_ << normal, binormal, curvature_axis (hand_set.cpp:39-40)
  const double *frame = lframe9;
  static const double AXES[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  if (nn.empty()) {
    for (int j = 0; j < dv.P; j++) fill_pose_header(poses[j], sample, frame, sample_index, slot, j);
    return;
  }
  PointList pl, plf, plc;
  slice_cloud(c, nn, pl);
  for (int a = 0; a < pr.num_hand_axes; a++) {
    int axis = pr.hand_axes[a];
    FingerHand fh(pr.finger_width, pr.hand_outer_diameter, pr.hand_depth, pr.num_finger_placements);
    for (int i = 0; i < pr.num_orientations; i++) {
      int slotp = a * pr.num_orientations + i;
      double rot[9], tmp[9], frame_rot[9];
      angle_axis_matrix(dv.angles[i], AXES[axis], rot);
      mat3_mul(frame, dv.rot_binormal, tmp);
      mat3_mul(tmp, rot, frame_rot);
      transform_to_hand_frame(pl, sample, frame_rot, plf);
      crop_by_hand_height(plf, pr.hand_height, plc);
      fh.evaluate_fingers(plc, pr.init_bite);
      fh.evaluate_hand();
      gpdb_pose &h = poses[slotp];
      fill_pose_header(h, sample, frame_rot, sample_index, slot, slotp);
      if (fh.any_hand()) {
        int fidx;
        if (pr.deepen_hand)
          fidx = fh.deepen_hand(plc, pr.init_bite, pr.hand_depth);
        else
          fidx = fh.choose_middle_hand();
        std::vector<int> closing = fh.closing_region(plc, fidx);
        if (closing.empty()) continue;
        flags[slotp] |= GPDB_POSE_VALID;
        // Hand::construct (hand.cpp:24-45)
        h.top = fh.top;
        h.bottom = fh.bottom;
        h.center = fh.center;
        double pb[3] = {h.bottom, fh.center, 0.0};
        for (int r = 0; r < 3; r++)
          h.position[r] = ((frame_rot[0 * 3 + r] * pb[0] + frame_rot[1 * 3 + r] * pb[1]) + frame_rot[2 * 3 + r] * pb[2]) + sample[r];
        int fpi = -1;
        for (int k = 0; k < (int)fh.hand.size(); k++)
          if (fh.hand[k]) { fpi = k; break; }
        h.finger_idx = (int16_t)fpi;
        double mn = DBL_MAX, mx = -DBL_MAX;
        for (int k : closing) {
          mn = std::min(mn, plc.p[3 * (size_t)k + 1]);
          mx = std::max(mx, plc.p[3 * (size_t)k + 1]);
        }
        h.width = mx - mn;
        int label = antipodal_eval(plc, closing, pr.friction_coeff, pr.min_viable);
        h.half_antipodal = (label == 1 || label == 2);
        h.full_antipodal = (label == 2);
        if (h.half_antipodal) flags[slotp] |= GPDB_POSE_HALF;
        if (h.full_antipodal) flags[slotp] |= GPDB_POSE_FULL;
      }
    }
  }
}

// A15: GraspDetector::filterGraspsWorkspace (grasp_detector.cpp:334-398, incl. the right_top
// quirk :362-363) and filterGraspsDirection (:422-456).
bool pose_passes_filters(const gpdb_params &pr, const gpdb_pose &h) {
  const double *approach = &h.frame[0], *binormal = &h.frame[3];
  double half_width = 0.5 * pr.hand_outer_diameter;
  double lb[3], rb[3], lt[3], rt[3], ap[3];
  for (int r = 0; r < 3; r++) {
    lb[r] = h.position[r] + half_width * binormal[r];
    rb[r] = h.position[r] - half_width * binormal[r];
    lt[r] = lb[r] + pr.hand_depth * approach[r];
    rt[r] = lb[r] + pr.hand_depth * approach[r];
    ap[r] = h.position[r] - 0.05 * approach[r];
  }
  bool ok = h.width >= pr.min_aperture && h.width <= pr.max_aperture;
  for (int r = 0; r < 3; r++) {
    double mn = std::min(std::min(std::min(lb[r], rb[r]), std::min(lt[r], rt[r])), ap[r]);
    double mx = std::max(std::max(std::max(lb[r], rb[r]), std::max(lt[r], rt[r])), ap[r]);
    ok = ok && mn >= pr.workspace_grasps[2 * r] && mx <= pr.workspace_grasps[2 * r + 1];
  }
  if (ok && pr.filter_approach_direction) {
    double dot = (pr.direction[0] * approach[0] + pr.direction[1] * approach[1]) + pr.direction[2] * approach[2];
    double angle = std::acos(dot);
    if (angle > pr.thresh_rad) ok = false;
  }
  return ok;
}

// ------------------------------------------------------------------------------------------
// OpenCV primitives restated for 60x60 float images (image_strategy.cpp:117-119,145-153,
// 179-187,215-230). Pinned against cv2 in tests/test_oracle_pins.py.
// ------------------------------------------------------------------------------------------
// cv::dilate with a 3x3 MORPH_RECT element, default border (BORDER_CONSTANT with
// morphologyDefaultBorderValue = -inf for dilation, i.e. the border is ignored).
void dilate3x3(const float *src, float *dst, int S, int ch) {
  for (int r = 0; r < S; r++)
    for (int cc = 0; cc < S; cc++)
      for (int k = 0; k < ch; k++) {
        float m = -FLT_MAX;
        for (int dr = -1; dr <= 1; dr++)
          for (int dc = -1; dc <= 1; dc++) {
            int rr = r + dr, c2 = cc + dc;
            if (rr < 0 || rr >= S || c2 < 0 || c2 >= S) continue;
            m = std::max(m, src[(rr * S + c2) * ch + k]);
          }
        dst[(r * S + cc) * ch + k] = m;
      }
}
// cv::normalize(img, img, 0, 1, NORM_MINMAX, CV_32F) then convertTo(CV_8U, 255.0).
// normalize: min/max over ALL channels; scale = 1/(max-min) if max-min > DBL_EPSILON else 0;
// shift = -min*scale; dst = src*(float)scale + (float)shift (convertScale casts to float).
// convertTo(CV_8U, 255): saturate_cast<uchar>(cvRound(v*255.f)) (round half to even).
void normalize_to_u8(const float *img, int n, uint8_t *out, int out_stride, int out_off, int ch) {
  double smin = DBL_MAX, smax = -DBL_MAX;
  for (int i = 0; i < n * ch; i++) {
    smin = std::min(smin, (double)img[i]);
    smax = std::max(smax, (double)img[i]);
  }
  double scale = (1.0 - 0.0) * (smax - smin > DBL_EPSILON ? 1.0 / (smax - smin) : 0.0);
  double shift = 0.0 - smin * scale;
  float a = (float)scale, b = (float)shift;
  for (int i = 0; i < n; i++)
    for (int k = 0; k < ch; k++) {
      float v = std::fmaf(img[i * ch + k], a, b);  // cv::convertScale 32f->32f uses v_fma on AVX2 builds (pinned vs cv2)
      float w = v * 255.0f;
      long rr = std::lrint((double)w);  // round half to even (default FE_TONEAREST)
      if (rr < 0) rr = 0;
      if (rr > 255) rr = 255;
      out[(size_t)i * out_stride + out_off + k] = (uint8_t)rr;
    }
}

struct ImgScratch {
  std::vector<float> a, b, avgs, counts;
  std::vector<uint8_t> nz;
};

// ImageStrategy::createNormalsImage (image_strategy.cpp:124-156). pts: unit coords [3 x m] after
// projection permutation; normals [3 x m]; writes 3 channels at out_off of an HWC image.
void create_normals_image(const std::vector<int> &cells, const std::vector<double> &normals, int S, uint8_t *out,
                          int C, int out_off, ImgScratch &s) {
  s.a.assign((size_t)S * S * 3, 0.0f);
  s.b.resize((size_t)S * S * 3);
  for (size_t i = 0; i < cells.size(); i++) {
    int idx = cells[i];
    int row = S - 1 - idx / S, col = idx % S;
    float *v = &s.a[(size_t)(row * S + col) * 3];
    const double *n = &normals[3 * i];
    float a0 = (float)std::fabs(n[0]), a1 = (float)std::fabs(n[1]), a2 = (float)std::fabs(n[2]);
    if (v[0] == 0 && v[1] == 0 && v[2] == 0) {
      v[0] = a0; v[1] = a1; v[2] = a2;
    } else {
      // v += (Vec3f(|n|) - v) * (1.0 / sqrt(v.v)): Vec3f * double -> each coeff float(float*double)
      double f = 1.0 / (double)std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      float d0 = (float)((double)(a0 - v[0]) * f), d1 = (float)((double)(a1 - v[1]) * f), d2 = (float)((double)(a2 - v[2]) * f);
      v[0] += d0; v[1] += d1; v[2] += d2;
    }
  }
  dilate3x3(s.a.data(), s.b.data(), S, 3);
  normalize_to_u8(s.b.data(), S * S, out, C, out_off, 3);
}

// ImageStrategy::createDepthImage (image_strategy.cpp:158-191).
void create_depth_image(const std::vector<int> &cells, const std::vector<double> &z, int S, uint8_t *out, int C,
                        int out_off, ImgScratch &s) {
  s.a.assign((size_t)S * S, 0.0f);
  s.b.resize((size_t)S * S);
  s.avgs.assign((size_t)S * S, 0.0f);
  s.counts.assign((size_t)S * S, 0.0f);
  for (size_t i = 0; i < cells.size(); i++) {
    int idx = cells[i];
    int row = S - 1 - idx / S, col = idx % S;
    s.counts[idx] += 1.0;
    s.avgs[idx] = (float)((double)s.avgs[idx] + (z[i] - (double)s.avgs[idx]) * (1.0 / (double)s.counts[idx]));
    s.a[(size_t)row * S + col] = (float)(1.0 - (double)s.avgs[idx]);
  }
  dilate3x3(s.a.data(), s.b.data(), S, 1);
  normalize_to_u8(s.b.data(), S * S, out, C, out_off, 1);
}

// ImageStrategy::createShadowImage (image_strategy.cpp:193-233).
void create_shadow_image(const std::vector<int> &cells, const std::vector<double> &z, int S, uint8_t *out, int C,
                         int out_off, ImgScratch &s) {
  s.a.assign((size_t)S * S, 0.0f);
  s.b.resize((size_t)S * S);
  s.counts.assign((size_t)S * S, 0.0f);
  s.nz.assign((size_t)S * S, 0);
  for (size_t i = 0; i < cells.size(); i++) {
    int idx = cells[i];
    int row = S - 1 - idx / S, col = idx % S;
    s.counts[idx] += 1.0;
    float &v = s.a[(size_t)row * S + col];
    v = (float)((double)v + (z[i] - (double)v) * (1.0 / (double)s.counts[idx]));
    s.nz[(size_t)row * S + col] = 1;
  }
  // cv::minMaxLoc with mask: max over masked pixels (0 when the mask is empty)
  double mx = 0.0;
  bool any = false;
  for (int i = 0; i < S * S; i++)
    if (s.nz[i]) {
      if (!any || (double)s.a[i] > mx) mx = (double)s.a[i];
      any = true;
    }
  for (int i = 0; i < S * S; i++) {
    float mi = s.nz[i] ? (float)mx : 0.0f;  // max_img.setTo(max, nonzero)
    s.a[i] = mi - s.a[i];
  }
  dilate3x3(s.a.data(), s.b.data(), S, 1);
  normalize_to_u8(s.b.data(), S * S, out, C, out_off, 1);
}

// ImageStrategy::findCellIndices (image_strategy.cpp:92-102): rows from coord 0, cols from coord 1.
void find_cells(const std::vector<double> &u, int m, int S, std::vector<int> &cells) {
  double cellsize = 1.0 / (double)S;
  cells.resize(m);
  for (int i = 0; i < m; i++) {
    int v = std::min((int)std::floor(u[3 * (size_t)i] / cellsize), S - 1);
    int h = std::min((int)std::floor(u[3 * (size_t)i + 1] / cellsize), S - 1);
    cells[i] = h + v * S;
  }
}

struct Vec3iHash {
  size_t operator()(const std::array<int, 3> &v) const { return gpdb_voxel_hash(v[0], v[1], v[2]); }
};
typedef std::unordered_set<std::array<int, 3>, Vec3iHash> VoxelSet;

// A9: HandSet::calculateShadow (hand_set.cpp:118-185) in the deterministic variant of
// include/gpd_b200_shadow.h. Returns shadow points [3 x m] in a defined (lexicographic) order.
void calculate_shadow(const Cloud &c, const PointList &pl, const double *qtab, double shadow_length, int sample_index,
                      std::vector<double> &shadow) {
  shadow.clear();
  const double voxel = GPDB_SHADOW_VOXEL;
  const int num_shadow_points = (int)std::floor(shadow_length / voxel);
  const int K = c.K, n = pl.size();
  std::vector<int> camera_set(K, 0);
  double center[3] = {0, 0, 0};
  for (int j = 0; j < n; j++) {
    for (int k = 0; k < K; k++) camera_set[k] += c.cam[(size_t)pl.gidx[j] * K + k];
    for (int r = 0; r < 3; r++) center[r] += pl.p[3 * (size_t)j + r];
  }
  for (int r = 0; r < 3; r++) center[r] /= (double)n;
  std::vector<VoxelSet> shadows(K);
  const double mult = 1.0 / voxel, mx = 1.0 / 32767.0;
  for (int k = 0; k < K; k++) {
    if (camera_set[k] < 1) continue;
    shadows[k].reserve((size_t)num_shadow_points * 10000);
    double sv[3] = {center[0] - c.vp[3 * k], center[1] - c.vp[3 * k + 1], center[2] - c.vp[3 * k + 2]};
    double nrm = std::sqrt((sv[0] * sv[0] + sv[1] * sv[1]) + sv[2] * sv[2]);
    for (int r = 0; r < 3; r++) sv[r] = shadow_length * sv[r] / nrm;
    for (int j = 0; j < n; j++) {
      uint32_t seed = gpdb_shadow_seed((uint32_t)sample_index, (uint32_t)pl.gidx[j], (uint32_t)k);
      for (int t = 0; t < num_shadow_points; t++) {
        double u = (double)gpdb_fastrand(&seed) * mx;
        std::array<int, 3> v;
        for (int r = 0; r < 3; r++) v[r] = (int)((pl.p[3 * (size_t)j + r] + u * sv[r]) * mult);
        shadows[k].insert(v);
      }
    }
  }
  VoxelSet all;
  if (K == 1) {
    all = std::move(shadows[0]);
  } else {
    all = shadows[0];
    for (int k = 1; k < K; k++) {
      if (camera_set[k] < 1) continue;
      VoxelSet nx;
      const VoxelSet &a = all.size() <= shadows[k].size() ? all : shadows[k];
      const VoxelSet &b = all.size() <= shadows[k].size() ? shadows[k] : all;
      for (const auto &v : a)
        if (b.find(v) != b.end()) nx.insert(v);
      all = std::move(nx);
    }
  }
  std::vector<std::array<int, 3>> vox(all.begin(), all.end());
  std::sort(vox.begin(), vox.end());
  shadow.resize(3 * vox.size());
  for (size_t i = 0; i < vox.size(); i++) {
    double g = qtab[gpdb_voxel_hash(vox[i][0], vox[i][1], vox[i][2]) & (GPDB_QTAB_SIZE - 1)];
    double jit = 1.0 * g * voxel * 0.3;  // Ones() * distr(gen) * voxel_grid_size * 0.3
    for (int r = 0; r < 3; r++) shadow[3 * i + r] = (double)vox[i][r] * voxel + jit;
  }
}

// A9 in the reference's LITERAL semantics (hand_set.cpp:118-233,263-266), for the statistical check of the deterministic
// variant: ONE sequential LCG stream shared by all points, cameras and hand sets (the static
// HandSet::seed_, initial value 0, hand_set.cpp:14; here the caller's `lcg_state`, advanced in place — what a
// single-threaded run of the reference does), draw i belongs to point i / num_shadow_points (:211-218), voxel set per
// camera, intersection as in the reference, then one std::normal_distribution<double>(0, 1) draw per voxel from a
// std::mt19937 (seeded from std::random_device upstream, :191-193; here from the caller's generator) added to x, y and z
// alike, in the set's iteration order. Irreproducible upstream by construction; two seeds give the reference's own
// noise floor.
void calculate_shadow_literal(const Cloud &c, const PointList &pl, double shadow_length, uint32_t &lcg_state,
                              std::mt19937 &gen, std::vector<double> &shadow) {
  shadow.clear();
  const double voxel = 0.003;
  const int num_shadow_points = (int)std::floor(shadow_length / voxel);
  const int K = c.K, n = pl.size();
  std::vector<int> camera_set(K, 0);
  double center[3] = {0, 0, 0};
  for (int j = 0; j < n; j++) {
    for (int k = 0; k < K; k++) camera_set[k] += c.cam[(size_t)pl.gidx[j] * K + k];
    for (int r = 0; r < 3; r++) center[r] += pl.p[3 * (size_t)j + r];
  }
  for (int r = 0; r < 3; r++) center[r] /= (double)n;
  std::vector<VoxelSet> shadows(K);
  const double mult = 1.0 / voxel, mx = 1.0 / 32767.0;
  for (int k = 0; k < K; k++) {
    if (camera_set[k] < 1) continue;
    double sv[3] = {center[0] - c.vp[3 * k], center[1] - c.vp[3 * k + 1], center[2] - c.vp[3 * k + 2]};
    double nrm = std::sqrt((sv[0] * sv[0] + sv[1] * sv[1]) + sv[2] * sv[2]);
    for (int r = 0; r < 3; r++) sv[r] = shadow_length * sv[r] / nrm;
    const int nd = n * num_shadow_points;
    for (int i = 0; i < nd; i++) {
      const int j = i / num_shadow_points;
      lcg_state = 214013u * lcg_state + 2531011u;  // HandSet::fastrand (two's-complement int upstream)
      double u = (double)((lcg_state >> 16) & 0x7FFFu) * mx;
      std::array<int, 3> v;
      for (int r = 0; r < 3; r++) v[r] = (int)((pl.p[3 * (size_t)j + r] + u * sv[r]) * mult);
      shadows[k].insert(v);
    }
  }
  VoxelSet all;
  if (K == 1) {
    all = std::move(shadows[0]);
  } else {
    all = shadows[0];
    for (int k = 1; k < K; k++) {
      if (camera_set[k] < 1) continue;
      VoxelSet nx;
      const VoxelSet &a = all.size() <= shadows[k].size() ? all : shadows[k];
      const VoxelSet &b = all.size() <= shadows[k].size() ? shadows[k] : all;
      for (const auto &v : a)
        if (b.find(v) != b.end()) nx.insert(v);
      all = std::move(nx);
    }
  }
  std::normal_distribution<double> distr{0.0, 1.0};
  shadow.resize(3 * all.size());
  size_t i = 0;
  for (const auto &v : all) {
    const double jit = 1.0 * distr(gen) * voxel * 0.3;
    for (int r = 0; r < 3; r++) shadow[3 * i + r] = (double)v[r] * voxel + jit;
    i++;
  }
}

// ImageStrategy::findPointsInUnitImage + transformPointsToUnitImage (image_strategy.cpp:53-90)
// applied to already hand-framed points `pf` [3 x n]; returns indices and unit coords.
void to_unit_image(const gpdb_params &pr, const gpdb_pose &h, const std::vector<double> &pf, int n,
                   std::vector<int> &idx, std::vector<double> &unit) {
  idx.clear();
  const double half_od = pr.volume_width / 2.0;
  for (int i = 0; i < n; i++) {
    double x = pf[3 * (size_t)i], y = pf[3 * (size_t)i + 1], z = pf[3 * (size_t)i + 2];
    if ((x > h.bottom) && (x < h.bottom + pr.volume_depth) && (y > h.center - half_od) && (y < h.center + half_od) &&
        (z > -1.0 * pr.volume_height) && (z < pr.volume_height))
      idx.push_back(i);
  }
  const double double_height = 2.0 * pr.volume_height;
  unit.resize(3 * idx.size());
  for (size_t k = 0; k < idx.size(); k++) {
    int i = idx[k];
    unit[3 * k] = (pf[3 * (size_t)i] - h.bottom) / pr.volume_depth;
    unit[3 * k + 1] = (pf[3 * (size_t)i + 1] - (h.center - half_od)) / pr.volume_width;
    unit[3 * k + 2] = (pf[3 * (size_t)i + 2] + pr.volume_height) / double_height;
  }
}

// A8, A10-A13: one grasp image. Image{1,3,12,15}ChannelsStrategy::createImage.
void create_image(const gpdb_params &pr, const gpdb_pose &h, const PointList &nnp, const std::vector<double> &shadow,
                  uint8_t *img, ImgScratch &s) {
  const int S = pr.image_size, C = pr.image_num_channels;
  std::memset(img, 0, (size_t)S * S * C);
  const int n = nnp.size();
  // transformToUnitImage (image_strategy.cpp:32-51)
  std::vector<double> pf(3 * (size_t)n), nf;
  for (int j = 0; j < n; j++) {
    double cen[3] = {nnp.p[3 * (size_t)j] - h.sample[0], nnp.p[3 * (size_t)j + 1] - h.sample[1],
                     nnp.p[3 * (size_t)j + 2] - h.sample[2]};
    to_frame(h.frame, cen, &pf[3 * (size_t)j]);
  }
  std::vector<int> idx;
  std::vector<double> unit;
  to_unit_image(pr, h, pf, n, idx, unit);
  int m = (int)idx.size();
  nf.resize(3 * (size_t)m);
  for (int k = 0; k < m; k++) to_frame(h.frame, &nnp.nr[3 * (size_t)idx[k]], &nf[3 * (size_t)k]);
  std::vector<int> cells;
  std::vector<double> zc(m);
  if (C == 3) {  // image_3_channels_strategy.cpp:25-41
    find_cells(unit, m, S, cells);
    create_normals_image(cells, nf, S, img, C, 0, s);
    return;
  }
  if (C == 1) {  // image_1_channels_strategy.cpp:25-48
    find_cells(unit, m, S, cells);
    for (int k = 0; k < m; k++) zc[k] = unit[3 * (size_t)k + 2];
    create_depth_image(cells, zc, S, img, C, 0, s);
    return;
  }
  // 12 / 15 channels: image_12_channels_strategy.cpp:35-86, image_15_channels_strategy.cpp:27-105
  std::vector<double> sunit;
  int ms = 0;
  if (C == 15) {
    int nsd = (int)shadow.size() / 3;
    std::vector<double> sf(3 * (size_t)nsd);
    for (int j = 0; j < nsd; j++) {
      double cen[3] = {shadow[3 * (size_t)j] - h.sample[0], shadow[3 * (size_t)j + 1] - h.sample[1],
                       shadow[3 * (size_t)j + 2] - h.sample[2]};
      to_frame(h.frame, cen, &sf[3 * (size_t)j]);
    }
    std::vector<int> sidx;
    to_unit_image(pr, h, sf, nsd, sidx, sunit);
    ms = (int)sidx.size();
  }
  const int per = (C == 15) ? 5 : 4;
  static const int swaps[3][2] = {{-1, -1}, {0, 2}, {1, 2}};
  std::vector<int> scells;
  std::vector<double> sz(ms);
  for (int pj = 0; pj < 3; pj++) {
    if (pj > 0) {
      for (int k = 0; k < m; k++) std::swap(unit[3 * (size_t)k + swaps[pj][0]], unit[3 * (size_t)k + swaps[pj][1]]);
      for (int k = 0; k < ms; k++) std::swap(sunit[3 * (size_t)k + swaps[pj][0]], sunit[3 * (size_t)k + swaps[pj][1]]);
    }
    find_cells(unit, m, S, cells);
    create_normals_image(cells, nf, S, img, C, pj * per, s);
    for (int k = 0; k < m; k++) zc[k] = unit[3 * (size_t)k + 2];
    create_depth_image(cells, zc, S, img, C, pj * per + 3, s);
    if (C == 15) {
      find_cells(sunit, ms, S, scells);
      for (int k = 0; k < ms; k++) sz[k] = sunit[3 * (size_t)k + 2];
      create_shadow_image(scells, sz, S, img, C, pj * per + 4, s);
    }
  }
}

// ------------------------------------------------------------------------------------------
// A14: net::EigenClassifier / ConvLayer / DenseLayer restated (eigen_classifier.cpp,
// conv_layer.cpp, dense_layer.cpp), parameterised in the channel count.
// ------------------------------------------------------------------------------------------
struct Weights {
  const float *c1w, *c1b, *c2w, *c2b, *i1w, *i1b, *i2w, *i2b;
};
struct NetScratch {
  std::vector<float> x, col, h1, p1, h2, p2, h3;
};
// ConvLayer::imageToColumns (conv_layer.cpp:62-98, Caffe im2col, stride 1, pad 0)
void im2col(const float *im, int ch, int H, int W, int k, float *col) {
  int oh = H - k + 1, ow = W - k + 1;
  for (int c = 0; c < ch; c++)
    for (int kr = 0; kr < k; kr++)
      for (int kc = 0; kc < k; kc++)
        for (int r = 0; r < oh; r++) {
          const float *src = im + ((size_t)c * H + (r + kr)) * W + kc;
          std::memcpy(col, src, sizeof(float) * ow);
          col += ow;
        }
}
// H = W * X + B (conv_layer.cpp:49-56): W [M x K] row-major, X [K x N] row-major
void gemm_bias(const float *W, const float *b, const float *X, float *H, int M, int K, int N) {
  for (int o = 0; o < M; o++) {
    float *h = H + (size_t)o * N;
    for (int j = 0; j < N; j++) h[j] = 0.0f;
    for (int k = 0; k < K; k++) {
      const float w = W[(size_t)o * K + k];
      const float *x = X + (size_t)k * N;
      for (int j = 0; j < N; j++) h[j] += w * x[j];
    }
    for (int j = 0; j < N; j++) h[j] += b[o];
  }
}
// EigenClassifier::poolForward 2x2 stride 2 (eigen_classifier.cpp:151-183)
void pool2(const float *X, int depth, int win, float *M) {
  int wout = win / 2;
  for (int i = 0; i < depth; i++)
    for (int r = 0; r < wout; r++)
      for (int cc = 0; cc < wout; cc++) {
        const float *p = X + (size_t)i * win * win + (size_t)(2 * r) * win + 2 * cc;
        M[(size_t)i * wout * wout + r * wout + cc] = std::max(std::max(p[0], p[1]), std::max(p[win], p[win + 1]));
      }
}
void lenet_forward(const gpdb_params &pr, const Weights &w, const uint8_t *img_hwc, float *logits2, NetScratch &s) {
  const int S = pr.image_size, C = pr.image_num_channels;
  const int o1 = S - 4, p1 = o1 / 2, o2 = p1 - 4, p2 = o2 / 2;
  // imageToArray (eigen_classifier.cpp:130-149): HWC uint8 -> CHW float, no scaling
  s.x.resize((size_t)C * S * S);
  for (int c = 0; c < C; c++)
    for (int r = 0; r < S; r++)
      for (int cc = 0; cc < S; cc++) s.x[((size_t)c * S + r) * S + cc] = (float)img_hwc[((size_t)r * S + cc) * C + c];
  s.col.resize((size_t)std::max(C * 25 * o1 * o1, 20 * 25 * o2 * o2));
  s.h1.resize((size_t)20 * o1 * o1);
  im2col(s.x.data(), C, S, S, 5, s.col.data());
  gemm_bias(w.c1w, w.c1b, s.col.data(), s.h1.data(), 20, C * 25, o1 * o1);
  if (pr.relu_after_conv)
    for (float &v : s.h1) v = std::max(v, 0.0f);
  s.p1.resize((size_t)20 * p1 * p1);
  pool2(s.h1.data(), 20, o1, s.p1.data());
  s.h2.resize((size_t)50 * o2 * o2);
  im2col(s.p1.data(), 20, p1, p1, 5, s.col.data());
  gemm_bias(w.c2w, w.c2b, s.col.data(), s.h2.data(), 50, 500, o2 * o2);
  if (pr.relu_after_conv)
    for (float &v : s.h2) v = std::max(v, 0.0f);
  s.p2.resize((size_t)50 * p2 * p2);
  pool2(s.h2.data(), 50, o2, s.p2.data());
  // flatten column-major: k = c + 50*j (eigen_classifier.cpp:103); dense1 W column-major (out,in)
  const int J = p2 * p2, KIN = 50 * J;
  s.h3.assign(500, 0.0f);
  for (int j = 0; j < J; j++)
    for (int c = 0; c < 50; c++) {
      const float xv = s.p2[(size_t)c * J + j];
      const float *wc = w.i1w + (size_t)(c + 50 * j) * 500;
      for (int o = 0; o < 500; o++) s.h3[o] += wc[o] * xv;
    }
  (void)KIN;
  for (int o = 0; o < 500; o++) s.h3[o] = std::max(s.h3[o] + w.i1b[o], 0.0f);  // + b, ReLU (:112)
  float y[2] = {0.0f, 0.0f};
  for (int k = 0; k < 500; k++) {
    y[0] += w.i2w[(size_t)k * 2] * s.h3[k];
    y[1] += w.i2w[(size_t)k * 2 + 1] * s.h3[k];
  }
  logits2[0] = y[0] + w.i2b[0];
  logits2[1] = y[1] + w.i2b[1];
}

// ------------------------------------------------------------------------------------------
// Cloud preprocessing: CandidatesGenerator::preprocessPointCloud
// (candidates_generator.cpp:14-37) = removeNans -> filterWorkspace -> voxelizeCloud ->
// calculateNormals(OMP) -> reverseNormals.
//
// PARITY UNPINNED: PCL (>= 1.9, unpinned, README.md:44) is absent from this image. The normal
// estimation restates the published algorithm of PCL 1.9.1 (the minimum version the reference
// names): pcl::NormalEstimationOMP::computeFeature -> computePointNormal ->
// computeMeanAndCovarianceMatrix (float32 single pass, common/impl/centroid.hpp) ->
// solvePlaneParameters -> pcl::eigen33 / computeRoots (common/impl/eigen.hpp, float32 closed
// form) -> flipNormalTowardsViewpoint (features/normal_3d.h), with the neighbours in the
// kd-tree's sorted (dist, index) order, which fixes the float32 summation order.
// ------------------------------------------------------------------------------------------

// pcl::computeRoots2 (common/impl/eigen.hpp)
void pcl_roots2(float b, float c, float *roots) {
  roots[0] = 0.0f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);  // Scalar (b * b - 4.0 * c): the subtraction is done in double
  if (d < 0.0f) d = 0.0f;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}
// pcl::computeRoots (common/impl/eigen.hpp), Scalar = float; m row-major symmetric
void pcl_roots(const float m[3][3], float *roots) {
  float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
             m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] -
             m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < FLT_EPSILON) {
    pcl_roots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = std::sqrt(3.0f);
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  float rho = std::sqrt(-a_over_3);
  float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
  float cos_theta = std::cos(theta);
  float sin_theta = std::sin(theta);
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0.0f) pcl_roots2(c2, c1, roots);
}
// pcl::eigen33 (mat, eigenvalue, eigenvector): smallest eigenvalue and its eigenvector
void pcl_eigen33_smallest(const float cov[3][3], float &eigenvalue, float *evec) {
  float scale = 0.0f;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) scale = std::max(scale, std::fabs(cov[r][c]));
  if (scale <= FLT_MIN) scale = 1.0f;
  float sm[3][3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) sm[r][c] = cov[r][c] / scale;
  float ev[3];
  pcl_roots(sm, ev);
  eigenvalue = ev[0] * scale;
  for (int d = 0; d < 3; d++) sm[d][d] -= ev[0];
  auto cross = [](const float *a, const float *b, float *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
  };
  float v1[3], v2[3], v3[3];
  cross(sm[0], sm[1], v1);
  cross(sm[0], sm[2], v2);
  cross(sm[1], sm[2], v3);
  auto sq = [](const float *v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; };
  float l1 = sq(v1), l2 = sq(v2), l3 = sq(v3);
  const float *best;
  float len;
  if (l1 >= l2 && l1 >= l3) { best = v1; len = l1; }
  else if (l2 >= l1 && l2 >= l3) { best = v2; len = l2; }
  else { best = v3; len = l3; }
  float sl = std::sqrt(len);
  for (int k = 0; k < 3; k++) evec[k] = best[k] / sl;
}

// Cloud::voxelizeCloud (cloud.cpp:286-348) LITERALLY, including the behaviour of
// std::set<Eigen::Vector4i, Cloud::UniqueVector4First3Comparator> (cloud.h:105-122). The comparator returns
// "a differs from b in one of the first three elements", which is not a strict weak ordering; with libstdc++'s
// red-black tree (bits/stl_tree.h _M_get_insert_unique_pos, tree.cc _Rb_tree_insert_and_rebalance) this means:
//   * the descent goes LEFT at every node that differs from the key and right at a node that equals it, so the
//     search path is the left spine of the tree down to the first equal node;
//   * an equal node on the left spine is recognised as a duplicate (the predecessor test finds it); if no node of
//     the left spine equals the key, the key is linked in as the new leftmost node, even when an equal key sits
//     elsewhere in the tree;
//   * iteration (in-order) therefore visits the keys newest first.
// Only the tree SHAPE matters for which old keys are still on the left spine, so the simulation keeps a real
// red-black tree with libstdc++'s insert fix-up (the new node is always a left child of a left child: only the
// "red uncle -> recolour" and "black uncle -> rotate right at the grandparent" cases occur).
// Used only to quantify the difference to the exact-set variant the product implements (tests/test_preprocess_oracle.py).
struct RbSim {
  struct Node { int parent, left, right; bool red; int item; };
  std::vector<Node> n;
  int root = -1, leftmost = -1;
  void rotate_right(int x) {
    const int y = n[x].left;
    n[x].left = n[y].right;
    if (n[y].right >= 0) n[n[y].right].parent = x;
    n[y].parent = n[x].parent;
    if (x == root) root = y;
    else if (x == n[n[x].parent].right) n[n[x].parent].right = y;
    else n[n[x].parent].left = y;
    n[y].right = x;
    n[x].parent = y;
  }
  void insert_leftmost(int item) {
    const int z = (int)n.size();
    n.push_back({leftmost, -1, -1, true, item});
    if (leftmost < 0) {
      root = z;
    } else {
      n[leftmost].left = z;
    }
    leftmost = z;
    int x = z;
    while (x != root && n[n[x].parent].red) {
      const int xp = n[x].parent, xpp = n[xp].parent;  // xp is red, so it is not the root: xpp exists
      const int y = n[xpp].right;                       // xp == n[xpp].left always (leftmost insertions)
      if (y >= 0 && n[y].red) {
        n[xp].red = false;
        n[y].red = false;
        n[xpp].red = true;
        x = xpp;
      } else {
        n[xp].red = false;
        n[xpp].red = true;
        rotate_right(xpp);
      }
    }
    n[root].red = false;
  }
};

struct PreOut {
  std::vector<float> xyz;
  std::vector<double> nrm;
  std::vector<int32_t> cam;
  std::vector<int32_t> src;  // index (into the raw cloud) of the point that represents each output point
};

// removeNans (cloud.cpp:154-164) + filterWorkspace (cloud.cpp:239-265): strict inequalities on float32 points
// against the double workspace bounds; cam_source / normals are filtered consistently.
// voxelizeCloud (cloud.cpp:286-348), exact-set variant (see include/gpd_b200.h gpdb_preprocess).
void preprocess_points(const float *xyz, const double *nrm_in, const int32_t *cam, int M, int K,
                       const gpdb_preprocess_params &pp, PreOut &o) {
  std::vector<int> keep;
  keep.reserve(M);
  for (int i = 0; i < M; i++) {
    const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
    if (!(std::isfinite(x) && std::isfinite(y) && std::isfinite(z))) continue;
    if (x > pp.workspace[0] && x < pp.workspace[1] && y > pp.workspace[2] && y < pp.workspace[3] && z > pp.workspace[4] &&
        z < pp.workspace[5])
      keep.push_back(i);
  }
  const int M1 = (int)keep.size();
  auto cam_of = [&](int i, int j) { return cam ? cam[(size_t)i * K + j] : 1; };
  if (!pp.voxelize || M1 == 0) {
    o.xyz.resize(3 * (size_t)M1);
    o.cam.resize((size_t)K * M1);
    o.src = keep;
    if (nrm_in) o.nrm.resize(3 * (size_t)M1);
    for (int k = 0; k < M1; k++) {
      const int i = keep[k];
      for (int a = 0; a < 3; a++) o.xyz[3 * (size_t)k + a] = xyz[3 * (size_t)i + a];
      for (int j = 0; j < K; j++) o.cam[(size_t)k * K + j] = cam_of(i, j);
      if (nrm_in)
        for (int a = 0; a < 3; a++) o.nrm[3 * (size_t)k + a] = nrm_in[3 * (size_t)i + a];
    }
    return;
  }
  const float cell = (float)pp.voxel_size;  // voxelizeCloud(float cell_size)
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};  // pcl::getMinMax3D
  for (int k = 0; k < M1; k++)
    for (int a = 0; a < 3; a++) mn[a] = std::min(mn[a], xyz[3 * (size_t)keep[k] + a]);
  struct V { long long key; int pos; };
  std::vector<V> v(M1);
  std::vector<int> vox(3 * (size_t)M1);
  for (int k = 0; k < M1; k++) {
    long long key = 0;
    for (int a = 0; a < 3; a++) {
      float t = (xyz[3 * (size_t)keep[k] + a] - mn[a]) / cell;  // (pt - min_pt) / cell_size, float32
      int c = (int)std::floor(t);                                // EigenUtils::floorVector
      vox[3 * (size_t)k + a] = c;
      key = (key << 21) | (long long)(c & 0x1FFFFF);
    }
    v[k] = {key, k};
  }
  std::stable_sort(v.begin(), v.end(), [](const V &a, const V &b) { return a.key < b.key; });
  struct G { int first, begin, end; };
  std::vector<G> groups;
  for (int s = 0; s < M1;) {
    int e = s;
    while (e < M1 && v[e].key == v[s].key) e++;
    groups.push_back({v[s].pos, s, e});
    s = e;
  }
  // iteration order of the reference's set when its de-duplication succeeds: every new voxel is linked in at the
  // leftmost position of the tree (comp(v, x) == "differs" sends the descent left), i.e. newest first
  std::sort(groups.begin(), groups.end(), [](const G &a, const G &b) { return a.first > b.first; });
  const int U = (int)groups.size();
  o.xyz.resize(3 * (size_t)U);
  o.cam.resize((size_t)K * U);
  o.src.resize(U);
  if (nrm_in) o.nrm.resize(3 * (size_t)U);
  for (int g = 0; g < U; g++) {
    const int k = groups[g].first, i = keep[k];
    o.src[g] = i;
    for (int a = 0; a < 3; a++) {
      float t = cell * (float)vox[3 * (size_t)k + a];  // min_pt + cell_size * v.cast<float>()
      o.xyz[3 * (size_t)g + a] = mn[a] + t;
    }
    for (int j = 0; j < K; j++) o.cam[(size_t)g * K + j] = (cam_of(i, j) == 1) ? 1 : 0;
    if (nrm_in) {
      double acc[3] = {0, 0, 0};
      for (int s = groups[g].begin; s < groups[g].end; s++)  // avg_normals.col(idx) += normals_.col(i), index order
        for (int a = 0; a < 3; a++) acc[a] += nrm_in[3 * (size_t)keep[v[s].pos] + a];
      const double cnt = (double)(groups[g].end - groups[g].begin);
      for (int a = 0; a < 3; a++) o.nrm[3 * (size_t)g + a] = acc[a] / cnt;
    }
  }
}

// Cloud::calculateNormalsOMP (cloud.cpp:497-535) + reverseNormals (cloud.cpp:573-604) on a built Cloud
// (c.xyz, c.cam, c.vp set, grid built); writes c.nrm.
void estimate_normals(Cloud &c, double radius) {
  const int N = c.N, K = c.K;
  c.nrm.assign(3 * (size_t)N, 0.0);
#pragma omp parallel
  {
    std::vector<Nb> nn;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < N; i++) {
      // convertCameraSourceMatrixToLists (cloud.cpp:606-621): the FIRST camera that sees the point
      int camera = -1;
      for (int j = 0; j < K; j++)
        if (c.cam[(size_t)i * K + j] == 1) { camera = j; break; }
      if (camera < 0) continue;  // the reference leaves such columns uninitialised; specified here as 0
      const float q[3] = {c.xyz[3 * (size_t)i], c.xyz[3 * (size_t)i + 1], c.xyz[3 * (size_t)i + 2]};
      radius_search(c, q, radius, nn);
      float n[3];
      if (nn.size() < 3) {  // computePointNormal returns false -> NaN normal (normal_3d_omp.hpp)
        n[0] = n[1] = n[2] = std::numeric_limits<float>::quiet_NaN();
      } else {
        float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (const Nb &b : nn) {
          const float x = c.xyz[3 * (size_t)b.i], y = c.xyz[3 * (size_t)b.i + 1], z = c.xyz[3 * (size_t)b.i + 2];
          accu[0] += x * x;
          accu[1] += x * y;
          accu[2] += x * z;
          accu[3] += y * y;
          accu[4] += y * z;
          accu[5] += z * z;
          accu[6] += x;
          accu[7] += y;
          accu[8] += z;
        }
        const float cnt = (float)nn.size();
        for (int k = 0; k < 9; k++) accu[k] /= cnt;
        float cov[3][3];
        cov[0][0] = accu[0] - accu[6] * accu[6];
        cov[0][1] = accu[1] - accu[6] * accu[7];
        cov[0][2] = accu[2] - accu[6] * accu[8];
        cov[1][1] = accu[3] - accu[7] * accu[7];
        cov[1][2] = accu[4] - accu[7] * accu[8];
        cov[2][2] = accu[5] - accu[8] * accu[8];
        cov[1][0] = cov[0][1];
        cov[2][0] = cov[0][2];
        cov[2][1] = cov[1][2];
        float ev;
        pcl_eigen33_smallest(cov, ev, n);
        // flipNormalTowardsViewpoint (features/normal_3d.h), float32; setViewPoint takes floats
        float vx = (float)c.vp[3 * camera] - q[0], vy = (float)c.vp[3 * camera + 1] - q[1], vz = (float)c.vp[3 * camera + 2] - q[2];
        float cos_theta = vx * n[0] + vy * n[1] + vz * n[2];
        if (cos_theta < 0) {
          n[0] *= -1;
          n[1] *= -1;
          n[2] *= -1;
        }
      }
      double nd[3] = {(double)n[0], (double)n[1], (double)n[2]};
      // reverseNormals (cloud.cpp:573-604)
      bool needs_reverse = true;
      for (int j = 0; j < K; j++)
        if (c.cam[(size_t)i * K + j] == 1) {
          double d0 = (double)q[0] - c.vp[3 * j], d1 = (double)q[1] - c.vp[3 * j + 1], d2 = (double)q[2] - c.vp[3 * j + 2];
          if (nd[0] * d0 + nd[1] * d1 + nd[2] * d2 < 0) {
            needs_reverse = false;
            break;
          }
        }
      if (needs_reverse)
        for (int a = 0; a < 3; a++) nd[a] *= -1.0;
      for (int a = 0; a < 3; a++) c.nrm[3 * (size_t)i + a] = nd[a];
    }
  }
}

std::vector<double> g_qtab;
const double *qtab() {
  if (g_qtab.empty()) {
    g_qtab.resize(GPDB_QTAB_SIZE);
    gpdb_build_qtab(g_qtab.data());
  }
  return g_qtab.data();
}

}  // namespace

// ==========================================================================================
// C interface for ctypes (tests / bench cpu_baseline only)
// ==========================================================================================
extern "C" {

void *gpdo_cloud_create(const float *xyz, const double *normals, const int32_t *cam, int32_t N, const double *vp,
                        int32_t K) {
  Cloud *c = new Cloud();
  c->N = N;
  c->K = K;
  c->xyz.assign(xyz, xyz + 3 * (size_t)N);
  c->nrm.assign(normals, normals + 3 * (size_t)N);
  if (cam)
    c->cam.assign(cam, cam + (size_t)K * N);
  else
    c->cam.assign((size_t)K * N, 1);
  c->vp.assign(vp, vp + 3 * (size_t)K);
  c->build();
  qtab();
  return c;
}
void gpdo_cloud_destroy(void *c) { delete (Cloud *)c; }
// Cloud::setSamples: sample indices N .. N + n - 1 address these positions afterwards
void gpdo_cloud_set_samples(void *c, const double *samples, int32_t n) {
  ((Cloud *)c)->samples.assign(samples, samples + 3 * (size_t)n);
}

int gpdo_radius_search(void *cloud, const float *q, double radius, int32_t *idx_out, float *dist_out, int32_t cap) {
  std::vector<Nb> nn;
  radius_search(*(Cloud *)cloud, q, radius, nn);
  int n = std::min((int)nn.size(), cap);
  for (int i = 0; i < n; i++) {
    idx_out[i] = nn[i].i;
    dist_out[i] = nn[i].d;
  }
  return (int)nn.size();
}

void gpdo_eigen3(const double *M, double *evals, double *evecs) { eigen3(M, evals, evecs); }

void gpdo_derived(const gpdb_params *pr, double *out /* [4+num_orient+9] */) {
  Derived d = derive(*pr);
  out[0] = d.P;
  out[1] = d.nn_radius_hs;
  out[2] = d.img_radius;
  out[3] = d.shadow_length;
  for (int i = 0; i < pr->num_orientations; i++) out[4 + i] = d.angles[i];
  for (int i = 0; i < 9; i++) out[4 + pr->num_orientations + i] = d.rot_binormal[i];
}

// FrameEstimator::calculateLocalFrames (frame_estimator.cpp:6-35), OpenMP over samples.
int gpdo_frames(void *cloud, const gpdb_params *pr, const int32_t *sidx, int32_t n, double *frames, uint8_t *valid,
                int32_t nthreads) {
  const Cloud &c = *(Cloud *)cloud;
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nn;
#pragma omp for schedule(dynamic, 16)
    for (int i = 0; i < n; i++) {
      double sp[3];
      c.sample_position(sidx[i], sp);
      float q[3] = {(float)sp[0], (float)sp[1], (float)sp[2]};  // frame_estimator.cpp:70-73
      valid[i] = calc_frame(c, q, pr->nn_radius, frames + 9 * (size_t)i, nn) ? 1 : 0;
      if (!valid[i])
        for (int r = 0; r < 9; r++) frames[9 * (size_t)i + r] = 0.0;
    }
  }
  return 0;
}

// HandSearch::evalHands (hand_search.cpp:144-188) + A15 filters. Dense outputs [n*P].
int gpdo_hand_search(void *cloud, const gpdb_params *pr, const int32_t *sidx, int32_t n, const double *frames,
                     const uint8_t *valid, gpdb_pose *poses, uint8_t *flags, int32_t nthreads) {
  const Cloud &c = *(Cloud *)cloud;
  Derived dv = derive(*pr);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
  for (int i = 0; i < n; i++) {
    gpdb_pose *ps = poses + (size_t)i * dv.P;
    uint8_t *fl = flags + (size_t)i * dv.P;
    if (!valid[i]) {
      std::memset(ps, 0, sizeof(gpdb_pose) * dv.P);
      for (int j = 0; j < dv.P; j++) {
        fl[j] = 0;
        ps[j].sample_index = sidx[i];
        ps[j].sample_slot = i;
        ps[j].pose_slot = (int16_t)j;
        ps[j].finger_idx = -1;
        ps[j].score = std::numeric_limits<float>::quiet_NaN();
      }
      continue;
    }
    eval_hand_set(c, *pr, dv, sidx[i], i, frames + 9 * (size_t)i, ps, fl);
    for (int j = 0; j < dv.P; j++)
      if ((fl[j] & GPDB_POSE_VALID) && pose_passes_filters(*pr, ps[j])) fl[j] |= GPDB_POSE_FILTERED;
  }
  return 0;
}

// ImageGenerator::createImages (image_generator.cpp:17-99). Consecutive poses with the same
// sample_slot form one hand set (neighbourhood + shadow computed once per set, as the
// reference does). images [n_poses * S*S*C] HWC.
int gpdo_images(void *cloud, const gpdb_params *pr, const gpdb_pose *poses, int32_t n_poses, uint8_t *images,
                int32_t nthreads) {
  const Cloud &c = *(Cloud *)cloud;
  Derived dv = derive(*pr);
  const size_t isz = (size_t)pr->image_size * pr->image_size * pr->image_num_channels;
  std::vector<int> set_start;
  for (int i = 0; i < n_poses; i++)
    if (i == 0 || poses[i].sample_slot != poses[i - 1].sample_slot || poses[i].sample_index != poses[i - 1].sample_index)
      set_start.push_back(i);
  set_start.push_back(n_poses);
  const double *qt = qtab();
  const int n_sets = (int)set_start.size() - 1;
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nn;
    PointList nnp;
    std::vector<double> shadow;
    ImgScratch s;
#pragma omp for schedule(dynamic, 2)
    for (int g = 0; g < n_sets; g++) {
      const gpdb_pose &h0 = poses[set_start[g]];
      float q[3] = {(float)h0.sample[0], (float)h0.sample[1], (float)h0.sample[2]};
      radius_search(c, q, dv.img_radius, nn);
      slice_cloud(c, nn, nnp);
      shadow.clear();
      if (pr->image_num_channels == 15 && !nn.empty())
        calculate_shadow(c, nnp, qt, dv.shadow_length, h0.sample_index, shadow);
      for (int i = set_start[g]; i < set_start[g + 1]; i++) create_image(*pr, poses[i], nnp, shadow, images + isz * i, s);
    }
  }
  return 0;
}

// gpdo_images with the shadow of the reference's literal semantics (calculate_shadow_literal): hand sets processed
// sequentially in order (one shared LCG stream starting at `lcg_seed`; one mt19937 seeded with mt_seed + set index per
// shadowVoxelsToPoints call). Used only by the statistical test of the deterministic variant.
int gpdo_images_literal_shadow(void *cloud, const gpdb_params *pr, const gpdb_pose *poses, int32_t n_poses, uint8_t *images,
                               uint32_t lcg_seed, uint32_t mt_seed) {
  const Cloud &c = *(Cloud *)cloud;
  Derived dv = derive(*pr);
  const size_t isz = (size_t)pr->image_size * pr->image_size * pr->image_num_channels;
  std::vector<Nb> nn;
  PointList nnp;
  std::vector<double> shadow;
  ImgScratch s;
  uint32_t lcg = lcg_seed;
  int set_no = 0;
  for (int i0 = 0; i0 < n_poses;) {
    int i1 = i0 + 1;
    while (i1 < n_poses && poses[i1].sample_slot == poses[i0].sample_slot && poses[i1].sample_index == poses[i0].sample_index) i1++;
    float q[3] = {(float)poses[i0].sample[0], (float)poses[i0].sample[1], (float)poses[i0].sample[2]};
    radius_search(c, q, dv.img_radius, nn);
    slice_cloud(c, nn, nnp);
    shadow.clear();
    if (pr->image_num_channels == 15 && !nn.empty()) {
      std::mt19937 gen(mt_seed + (uint32_t)set_no);
      calculate_shadow_literal(c, nnp, dv.shadow_length, lcg, gen, shadow);
    }
    for (int i = i0; i < i1; i++) create_image(*pr, poses[i], nnp, shadow, images + isz * i, s);
    i0 = i1;
    set_no++;
  }
  return 0;
}

// HandSearch::reevaluateHypotheses (hand_search.cpp:66-134) + reevaluateHypothesis (:190-209) + labelHypothesis (:211-228):
// the given hands are re-labelled against THIS cloud (the ground-truth mesh cloud of the data-generation path,
// GraspDetector::evalGroundTruth, grasp_detector.cpp:523-527): radius search r = nn_radius_hs around the hand's sample
// (eigenVectorToPcl: float32 image of the position), hand-frame transform with the hand's own frame, height crop (padding
// quirk), evaluateFingers(points, hand.top, hand.finger_idx) + evaluateHand(idx), closing region, Antipodal::evaluateGrasp.
// labels[i] = 1 for a full grasp, else 0; half / full flags are written back into the records.
int gpdo_reevaluate(void *cloud, const gpdb_params *pr, gpdb_pose *hands, int32_t n, int32_t *labels, int32_t nthreads) {
  const Cloud &c = *(Cloud *)cloud;
  Derived dv = derive(*pr);
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<Nb> nn;
    PointList pl, plf, plc;
#pragma omp for schedule(dynamic, 8)
    for (int i = 0; i < n; i++) {
      gpdb_pose &h = hands[i];
      labels[i] = 0;
      h.half_antipodal = 0;
      h.full_antipodal = 0;
      float q[3] = {(float)h.sample[0], (float)h.sample[1], (float)h.sample[2]};
      radius_search(c, q, dv.nn_radius_hs, nn);
      if (nn.empty() || h.finger_idx < 0 || h.finger_idx >= pr->num_finger_placements) continue;
      slice_cloud(c, nn, pl);
      transform_to_hand_frame(pl, h.sample, h.frame, plf);
      crop_by_hand_height(plf, pr->hand_height, plc);
      FingerHand fh(pr->finger_width, pr->hand_outer_diameter, pr->hand_depth, pr->num_finger_placements);
      fh.evaluate_fingers(plc, h.top, h.finger_idx);
      std::fill(fh.hand.begin(), fh.hand.end(), 0);  // evaluateHand(idx) (:83-87)
      fh.hand[h.finger_idx] = fh.fingers[h.finger_idx] && fh.fingers[pr->num_finger_placements + h.finger_idx];
      if (!fh.any_hand()) continue;
      std::vector<int> closing = fh.closing_region(plc, -1);
      if (closing.empty()) continue;
      int label = antipodal_eval(plc, closing, pr->friction_coeff, pr->min_viable);
      h.half_antipodal = (label == 1 || label == 2);
      h.full_antipodal = (label == 2);
      if (label == 2) labels[i] = 1;
    }
  }
  return n;
}

// EigenClassifier::classifyImages (eigen_classifier.cpp:59-79); race-free (per-thread scratch).
int gpdo_classify(const gpdb_params *pr, const float *const *wts /* 8 pointers */, const uint8_t *images, int32_t n,
                  float *scores, float *logits, int32_t nthreads) {
  Weights w = {wts[0], wts[1], wts[2], wts[3], wts[4], wts[5], wts[6], wts[7]};
  const size_t isz = (size_t)pr->image_size * pr->image_size * pr->image_num_channels;
#pragma omp parallel num_threads(nthreads)
  {
    NetScratch s;
#pragma omp for schedule(dynamic, 1)
    for (int i = 0; i < n; i++) {
      float y[2];
      lenet_forward(*pr, w, images + isz * i, y, s);
      scores[i] = y[1] - y[0];
      if (logits) {
        logits[2 * (size_t)i] = y[0];
        logits[2 * (size_t)i + 1] = y[1];
      }
    }
  }
  return 0;
}

// GraspDetector::detectGrasps steps 1-4 (grasp_detector.cpp:222-273) with the reference's three
// stage timers (:313-320). Returns n_candidates. out arrays allocated with malloc.
int gpdo_detect(void *cloud, const gpdb_params *pr, const float *const *wts, const int32_t *sidx, int32_t n,
                gpdb_result *out, int32_t nthreads, double *stage_seconds /* [4] cand, images, classify, total */) {
  Derived dv = derive(*pr);
  const int P = dv.P;
  std::memset(out, 0, sizeof(*out));
  out->n_samples = n;
  out->poses_per_sample = P;
  out->frame_valid = (uint8_t *)std::malloc((size_t)n + 1);
  out->frames = (double *)std::malloc(sizeof(double) * 9 * (size_t)n + 8);
  out->pose_flags = (uint8_t *)std::malloc((size_t)n * P + 1);
  out->pose_scores = (float *)std::malloc(sizeof(float) * (size_t)n * P + 4);
  std::vector<gpdb_pose> dense((size_t)n * P);
  double t0 = now_s();
  gpdo_frames(cloud, pr, sidx, n, out->frames, out->frame_valid, nthreads);
  gpdo_hand_search(cloud, pr, sidx, n, out->frames, out->frame_valid, dense.data(), out->pose_flags, nthreads);
  double t1 = now_s();
  int nc = 0;
  for (size_t i = 0; i < (size_t)n * P; i++) {
    out->pose_scores[i] = std::numeric_limits<float>::quiet_NaN();
    if ((out->pose_flags[i] & 3) == 3) nc++;
  }
  out->n_candidates = nc;
  out->n_total_candidates = nc;
  out->candidates = (gpdb_pose *)std::malloc(sizeof(gpdb_pose) * (size_t)nc + 8);
  int k = 0;
  for (size_t i = 0; i < (size_t)n * P; i++)
    if ((out->pose_flags[i] & 3) == 3) out->candidates[k++] = dense[i];
  const size_t isz = (size_t)pr->image_size * pr->image_size * pr->image_num_channels;
  uint8_t *images = (uint8_t *)std::malloc(isz * (size_t)nc + 8);
  gpdo_images(cloud, pr, out->candidates, nc, images, nthreads);
  double t2 = now_s();
  std::vector<float> scores(nc + 1);
  if (wts) gpdo_classify(pr, wts, images, nc, scores.data(), nullptr, nthreads);
  double t3 = now_s();
  for (int i = 0; i < nc; i++) {
    out->candidates[i].score = wts ? scores[i] : std::numeric_limits<float>::quiet_NaN();
    out->pose_scores[(size_t)out->candidates[i].sample_slot * P + out->candidates[i].pose_slot] = out->candidates[i].score;
  }
  if (pr->keep_images)
    out->images = images;
  else
    std::free(images);
  out->ms_candidates = (t1 - t0) * 1e3;
  out->ms_images = (t2 - t1) * 1e3;
  out->ms_classify = (t3 - t2) * 1e3;
  if (stage_seconds) {
    stage_seconds[0] = t1 - t0;
    stage_seconds[1] = t2 - t1;
    stage_seconds[2] = t3 - t2;
    stage_seconds[3] = t3 - t0;
  }
  return nc;
}

void gpdb_preprocess_params_default_o(gpdb_preprocess_params *p) {
  const double ws[6] = {-1, 1, -1, 1, -1, 1};
  for (int i = 0; i < 6; i++) p->workspace[i] = ws[i];
  p->voxel_size = 0.003;
  p->normals_radius = 0.03;
  p->voxelize = 1;
  p->estimate_normals = 1;
}

// CandidatesGenerator::preprocessPointCloud restated (see the preprocessing block above). Outputs are caller
// allocated for n_points entries; returns N'. ms_out[0..1]: seconds spent in voxelisation / normals.
int gpdo_preprocess(const float *xyz, const double *normals, const int32_t *cam, int32_t M, const double *vp, int32_t K,
                    const gpdb_preprocess_params *pp, float *xyz_out, double *nrm_out, int32_t *cam_out, int32_t *src_out,
                    double *sec_out, int32_t num_threads) {
#ifdef _OPENMP
  if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
  PreOut o;
  double t0 = now_s();
  preprocess_points(xyz, normals, cam, M, K, *pp, o);
  double t1 = now_s();
  const int N = (int)o.src.size();
  Cloud c;
  c.N = N;
  c.K = K;
  c.xyz = o.xyz;
  c.cam = o.cam;
  c.vp.assign(vp, vp + 3 * (size_t)K);
  if (pp->estimate_normals) {
    c.build();
    estimate_normals(c, pp->normals_radius);
  } else {
    c.nrm = o.nrm;
    if (c.nrm.size() != 3 * (size_t)N) return GPDB_ERR_INVALID;
  }
  double t2 = now_s();
  if (xyz_out) std::memcpy(xyz_out, c.xyz.data(), sizeof(float) * 3 * (size_t)N);
  if (nrm_out) std::memcpy(nrm_out, c.nrm.data(), sizeof(double) * 3 * (size_t)N);
  if (cam_out) std::memcpy(cam_out, c.cam.data(), sizeof(int32_t) * (size_t)K * N);
  if (src_out) std::memcpy(src_out, o.src.data(), sizeof(int32_t) * (size_t)N);
  if (sec_out) {
    sec_out[0] = t1 - t0;
    sec_out[1] = t2 - t1;
  }
  return N;
}
void gpdo_pcl_eigen33(const float *cov9 /* row-major */, float *eigenvalue, float *evec) {
  float m[3][3];
  for (int r = 0; r < 3; r++)
    for (int c2 = 0; c2 < 3; c2++) m[r][c2] = cov9[3 * r + c2];
  pcl_eigen33_smallest(m, *eigenvalue, evec);
}

// Literal voxel set of the reference (see RbSim): xyz = M finite points inside the workspace; src_out receives, in
// the reference's iteration order, the index of the point kept for each set element (may contain several elements
// per voxel); returns the number of elements. first_seen_dups_out (may be NULL): how many insertions were linked in
// although an equal voxel was already in the set.
int gpdo_voxelize_literal(const float *xyz, int32_t M, double voxel_size, int32_t *src_out, int32_t *missed_out) {
  const float cell = (float)voxel_size;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  for (int i = 0; i < M; i++)
    for (int a = 0; a < 3; a++) mn[a] = std::min(mn[a], xyz[3 * (size_t)i + a]);
  std::vector<long long> key(M);
  for (int i = 0; i < M; i++) {
    long long k = 0;
    for (int a = 0; a < 3; a++) k = (k << 21) | (long long)((int)std::floor((xyz[3 * (size_t)i + a] - mn[a]) / cell) & 0x1FFFFF);
    key[i] = k;
  }
  RbSim t;
  std::unordered_set<long long> seen;
  int missed = 0;
  for (int i = 0; i < M; i++) {
    bool dup = false;
    for (int x = t.root; x >= 0; x = t.n[x].left)  // the search path: the left spine, top down
      if (key[t.n[x].item] == key[i]) { dup = true; break; }
    if (!dup) {
      if (!seen.insert(key[i]).second) missed++;
      t.insert_leftmost(i);
    }
  }
  // in-order traversal = newest first: leftmost insertions only ever prepend
  int cnt = 0;
  std::vector<int> stack;
  int x = t.root;
  while (x >= 0 || !stack.empty()) {
    while (x >= 0) { stack.push_back(x); x = t.n[x].left; }
    x = stack.back();
    stack.pop_back();
    src_out[cnt++] = t.n[x].item;
    x = t.n[x].right;
  }
  if (missed_out) *missed_out = missed;
  return cnt;
}

void gpdo_free_result(gpdb_result *r) {
  std::free(r->frame_valid);
  std::free(r->frames);
  std::free(r->pose_flags);
  std::free(r->pose_scores);
  std::free(r->candidates);
  std::free(r->images);
  std::memset(r, 0, sizeof(*r));
}

// stand-alone pieces exposed so that the tests can pin them against cv2 / numpy
void gpdo_dilate_normalize_u8(const float *img /* S*S*ch HWC */, int32_t S, int32_t ch, uint8_t *out) {
  std::vector<float> d((size_t)S * S * ch);
  dilate3x3(img, d.data(), S, ch);
  normalize_to_u8(d.data(), S * S, out, ch, 0, ch);
}
// ConvLayer(width,height,depth,num_filters,spatial_extent,1,0)::forward (conv_layer.cpp:26-56) for the
// known-answer test of src/tests/test_conv_layer.cpp:9-40
void gpdo_conv_forward(const float *x, int32_t C, int32_t H, int32_t Wd, const float *w, const float *b, int32_t M,
                       int32_t k, float *out) {
  int oh = H - k + 1, ow = Wd - k + 1;
  std::vector<float> col((size_t)C * k * k * oh * ow);
  im2col(x, C, H, Wd, k, col.data());
  gemm_bias(w, b, col.data(), out, M, C * k * k, oh * ow);
}
void gpdo_angle_axis(double angle, const double *axis, double *R) { angle_axis_matrix(angle, axis, R); }
void gpdo_qtab(double *tab) { std::memcpy(tab, qtab(), sizeof(double) * GPDB_QTAB_SIZE); }
int gpdo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
