// api.cu — the C-ABI of libgpd_b200.so (include/gpd_b200.h): context, cloud upload, the chunked
// detect pipeline and the stage-level entry points. Host-side logic only; kernels live in
// geometry.cu / lenet_simt.cu / lenet_tc.cu. There is no CPU fallback anywhere in this library.
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

static char g_create_err[512] = "";

struct StageTimes {
  struct Span { int stage; cudaEvent_t b, e; };
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  std::vector<Span> spans;
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
  ~StageTimes() { for (cudaEvent_t e : pool) cudaEventDestroy(e); }
};
cudaEvent_t gpdb_st_begin(gpdb_ctx *ctx) {
  cudaEvent_t e = ctx->st->get();
  cudaEventRecord(e, ctx->stream);
  return e;
}
void gpdb_st_end(gpdb_ctx *ctx, int stage, cudaEvent_t begin) {
  cudaEvent_t e = ctx->st->get();
  cudaEventRecord(e, ctx->stream);
  ctx->st->spans.push_back({stage, begin, e});
}
static void st_collect(gpdb_ctx *ctx, double *ms) {
  for (auto &sp : ctx->st->spans) {
    float t = 0;
    if (cudaEventElapsedTime(&t, sp.b, sp.e) == cudaSuccess) ms[sp.stage] += t;
  }
  ctx->st->spans.clear();
  ctx->st->used = 0;
}



void gpdb_set_error(gpdb_ctx *ctx, int code, const char *fmt, ...) {
  char buf[480];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  char *dst = ctx ? ctx->err : g_create_err;
  snprintf(dst, 512, "gpd_b200 error %d: %s", code, buf);
  fprintf(stderr, "%s\n", dst);  // reference convention: errors are also printed
}

void *gpdb_scratch(gpdb_ctx *ctx, int slot, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (ctx->scratch_sz[slot] >= bytes) return ctx->scratch[slot];
  if (ctx->scratch[slot]) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->scratch[slot]);
    ctx->scratch[slot] = nullptr;
    ctx->scratch_sz[slot] = 0;
  }
  size_t want = bytes + bytes / 4 + 256;
  if (cudaMalloc(&ctx->scratch[slot], want) != cudaSuccess) {
    gpdb_set_error(ctx, GPDB_ERR_CUDA, "cudaMalloc(%zu) failed for scratch slot %d: %s", want, slot,
                   cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  ctx->scratch_sz[slot] = want;
  return ctx->scratch[slot];
}

// ---- host restatements of the Eigen helpers that define the rotation set --------------------------
// Eigen::AngleAxisd(angle, axis).toRotationMatrix(), column-major (hand_set.cpp:52-53,68-69)
static void angle_axis_matrix(double angle, const double axis[3], double *R) {
  double s = std::sin(angle), c = std::cos(angle);
  double sa[3] = {s * axis[0], s * axis[1], s * axis[2]};
  double c1[3] = {(1.0 - c) * axis[0], (1.0 - c) * axis[1], (1.0 - c) * axis[2]};
  double t;
  t = c1[0] * axis[1]; R[1 * 3 + 0] = t - sa[2]; R[0 * 3 + 1] = t + sa[2];
  t = c1[0] * axis[2]; R[2 * 3 + 0] = t + sa[1]; R[0 * 3 + 2] = t - sa[1];
  t = c1[1] * axis[2]; R[2 * 3 + 1] = t - sa[0]; R[1 * 3 + 2] = t + sa[0];
  R[0] = c1[0] * axis[0] + c;
  R[4] = c1[1] * axis[1] + c;
  R[8] = c1[2] * axis[2] + c;
}
// Eigen 3.3 VectorXd::LinSpaced(size, low, high)(i)
static double linspaced(int size, double low, double high, int i) {
  int size1 = size == 1 ? 1 : size - 1;
  double step = size == 1 ? 0.0 : (high - low) / (double)(size - 1);
  bool flip = std::fabs(high) < std::fabs(low);
  if (flip) return (i == 0) ? low : (high - (double)(size1 - i) * step);
  return (i == size1) ? high : (low + (double)i * step);
}

static int fill_dev_params(gpdb_ctx *ctx) {
  const gpdb_params &p = ctx->prm;
  DevParams &d = ctx->hp;
  memset(&d, 0, sizeof(d));
  if (p.num_orientations < 1 || p.num_orientations > GPDB_MAX_ORIENT || p.num_hand_axes < 1 ||
      p.num_hand_axes > GPDB_MAX_HAND_AXES || p.num_finger_placements < 1 || 2 * p.num_finger_placements > GPDB_MAX_SLOTS ||
      p.image_size < 8 || p.image_size > 64 ||
      !(p.image_num_channels == 1 || p.image_num_channels == 3 || p.image_num_channels == 12 || p.image_num_channels == 15)) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID,
                   "unsupported parameters (num_orientations 1..%d, hand_axes 1..3, num_finger_placements 1..%d, "
                   "image_size 8..64, image_num_channels 1/3/12/15)", GPDB_MAX_ORIENT, GPDB_MAX_SLOTS / 2);
    return GPDB_ERR_INVALID;
  }
  d.finger_width = p.finger_width;
  d.hand_outer_diameter = p.hand_outer_diameter;
  d.hand_depth = p.hand_depth;
  d.hand_height = p.hand_height;
  d.init_bite = p.init_bite;
  d.n_axes = p.num_hand_axes;
  d.n_orient = p.num_orientations;
  d.P = d.n_axes * d.n_orient;
  d.nfp = p.num_finger_placements;
  d.deepen = p.deepen_hand;
  d.all_axes_z = 1;
  for (int a = 0; a < d.n_axes; a++) {
    if (p.hand_axes[a] < 0 || p.hand_axes[a] > 2) {
      gpdb_set_error(ctx, GPDB_ERR_INVALID, "hand_axes[%d] = %d out of range 0..2", a, p.hand_axes[a]);
      return GPDB_ERR_INVALID;
    }
    d.axes[a] = p.hand_axes[a];
    if (p.hand_axes[a] != 2) d.all_axes_z = 0;
  }
  // FingerHand::FingerHand (finger_hand.cpp:6-24)
  for (int i = 0; i < d.nfp; i++) {
    double h = linspaced(d.nfp, 0.0, p.hand_outer_diameter - p.finger_width, i);
    d.fs[i] = (h - p.hand_outer_diameter) + p.finger_width;
    d.fs[d.nfp + i] = h;
  }
  for (int i = 0; i < 2 * d.nfp; i++) d.fsw[i] = d.fs[i] + p.finger_width;
  d.slots_disjoint = d.nfp >= 3 ? 1 : 0;  // the arithmetic slot lookup needs >= 3 equally spaced slots; else linear scan
  d.inv_slot_step = d.nfp >= 2 ? 1.0 / (d.fs[1] - d.fs[0]) : 0.0;
  for (int i = 0; i + 1 < d.nfp; i++)
    if (!(d.fsw[i] <= d.fs[i + 1]) || !(d.fsw[d.nfp + i] <= d.fs[d.nfp + i + 1])) d.slots_disjoint = 0;
  // deepenHand steps (finger_hand.cpp:118-121): repeated += 0.005 in double
  d.J = 0;
  for (double depth = p.init_bite + 0.005; depth <= p.hand_depth; depth += 0.005) {
    if (d.J >= GPDB_MAX_DEEPEN) {
      gpdb_set_error(ctx, GPDB_ERR_INVALID, "more than %d deepen steps", GPDB_MAX_DEEPEN);
      return GPDB_ERR_INVALID;
    }
    d.topj[d.J] = depth;
    d.botj[d.J] = depth - p.hand_depth;
    d.J++;
  }
  d.cosf = std::cos(p.friction_coeff * M_PI / 180.0);
  d.min_viable = p.min_viable;
  d.min_ap = p.min_aperture;
  d.max_ap = p.max_aperture;
  for (int i = 0; i < 6; i++) d.ws[i] = p.workspace_grasps[i];
  d.filt_dir = p.filter_approach_direction;
  for (int i = 0; i < 3; i++) d.dir[i] = p.direction[i];
  d.thresh = p.thresh_rad;
  const double uy[3] = {0, 1, 0};
  angle_axis_matrix(M_PI, uy, d.rotb);
  static const double AXES[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int a = 0; a < d.n_axes; a++)
    for (int i = 0; i < d.n_orient; i++) {
      // angles = LinSpaced(n+1, -pi/2, pi/2).head(n) (hand_search.cpp:151-155)
      double ang = linspaced(d.n_orient + 1, -1.0 * M_PI / 2.0, M_PI / 2.0, i);
      angle_axis_matrix(ang, AXES[d.axes[a]], d.rot[a * d.n_orient + i]);
    }
  d.vol_w = p.volume_width;
  d.vol_d = p.volume_depth;
  d.vol_h = p.volume_height;
  d.S = p.image_size;
  d.C = p.image_num_channels;
  // radii: hand_search.cpp:13-17, image_generator.cpp:43-46, image_15_channels_strategy.h:72-75
  double r_hs = std::max(std::max(p.hand_outer_diameter - p.finger_width, p.hand_depth), p.hand_height / 2.0);
  double r_img = std::max(std::max(p.volume_depth, p.volume_height / 2.0), p.volume_width);
  double r_lrf = p.nn_radius;
  d.r2_lrf = (float)(r_lrf * r_lrf);
  d.r2_hs = (float)(r_hs * r_hs);
  d.r2_img = (float)(r_img * r_img);
  d.rf_lrf = (float)r_lrf * 1.0001f + 1e-6f;
  d.rf_hs = (float)r_hs * 1.0001f + 1e-6f;
  d.rf_img = (float)r_img * 1.0001f + 1e-6f;
  d.shadow_length = r_img;
  d.vox_mult = 1.0 / GPDB_SHADOW_VOXEL;
  d.nsp = (int)std::floor(d.shadow_length / GPDB_SHADOW_VOXEL);
  if (d.nsp > GPDB_MAX_NSP) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "image volume too large: %d shadow draws per point (max %d)", d.nsp, GPDB_MAX_NSP);
    return GPDB_ERR_INVALID;
  }
  {  // closed-form skip-ahead of HandSet::fastrand (hand_set.cpp:263-266), mod 2^32
    unsigned A = 1u, Cc = 0u;
    for (int t = 0; t < GPDB_MAX_NSP; t++) {
      Cc = 214013u * Cc + 2531011u;
      A = 214013u * A;
      d.lcgA[t] = A;
      d.lcgC[t] = Cc;
    }
  }
  double diag = std::sqrt(p.volume_depth * p.volume_depth + p.volume_width * p.volume_width +
                          4.0 * p.volume_height * p.volume_height);
  d.bm_dim = (int)std::ceil((diag + 2.0 * 3.2 * GPDB_SHADOW_VOXEL * 0.3) / GPDB_SHADOW_VOXEL) + 4;
  if (d.bm_dim > 64 && d.C == 15) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "image volume too large for the shadow bitmap (%d > 64 voxels across)", d.bm_dim);
    return GPDB_ERR_INVALID;
  }
  d.relu_after_conv = p.relu_after_conv;
  d.K = 1;
  d.dim[0] = d.dim[1] = d.dim[2] = 1;
  d.inv_cell = 50.0f;
  return GPDB_OK;
}

extern "C" {

void gpdb_params_default(gpdb_params *p) {
  memset(p, 0, sizeof(*p));
  p->finger_width = 0.01;
  p->hand_outer_diameter = 0.12;
  p->hand_depth = 0.06;
  p->hand_height = 0.02;
  p->init_bite = 0.01;
  p->volume_width = 0.10;
  p->volume_depth = 0.06;
  p->volume_height = 0.02;
  p->image_size = 60;
  p->image_num_channels = 15;
  p->nn_radius = 0.01;
  p->num_orientations = 8;
  p->num_finger_placements = 10;
  p->num_hand_axes = 1;
  p->hand_axes[0] = 2;
  p->deepen_hand = 1;
  p->friction_coeff = 20.0;
  p->min_viable = 6;
  p->min_aperture = 0.0;
  p->max_aperture = 0.085;
  const double ws[6] = {-1, 1, -1, 1, -1, 1};
  for (int i = 0; i < 6; i++) p->workspace_grasps[i] = ws[i];
  p->filter_approach_direction = 0;
  p->direction[0] = 1.0;
  p->thresh_rad = 2.3;
}

const char *gpdb_build_info(void) {
  return "gpd_b200 v1, sm_100a, kernels: k_frames k_hands k_images k_normals (fp64 / PCL-order fp32, -fmad=false), lenet: "
         "conv1 tcgen05 kind::i8 implicit GEMM (uint8 image x 3 int8 weight digit planes, exact int32 in TMEM), conv2 tcgen05 "
         "f16 implicit GEMM + ip1 TMA-fed tcgen05 GEMM (fp16 hi/lo split operands, fp32 accumulate in TMEM), ip2 simt; "
         "lenet_impl=1 forces the simt-fp32 kernels";
}

const char *gpdb_last_error(const gpdb_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int gpdb_create(const gpdb_params *params, gpdb_ctx **ctx_out) {
  if (!params || !ctx_out) {
    gpdb_set_error(nullptr, GPDB_ERR_INVALID, "gpdb_create: null argument");
    return GPDB_ERR_INVALID;
  }
  *ctx_out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    gpdb_set_error(nullptr, GPDB_ERR_CUDA, "no CUDA device: libgpd_b200 has no CPU fallback");
    return GPDB_ERR_CUDA;
  }
  if (params->device < 0 || params->device >= ndev) {
    gpdb_set_error(nullptr, GPDB_ERR_INVALID, "device %d out of range (have %d)", params->device, ndev);
    return GPDB_ERR_INVALID;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, params->device) != cudaSuccess || prop.major != 10) {
    gpdb_set_error(nullptr, GPDB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                   params->device, prop.major, prop.minor);
    return GPDB_ERR_CUDA;
  }
  gpdb_ctx *ctx = new gpdb_ctx();
  memset(ctx, 0, sizeof(*ctx));
  ctx->st = new StageTimes();
  ctx->own_stream = true;
  ctx->prm = *params;
  ctx->device = params->device;
  ctx->sm_count = prop.multiProcessorCount;
  int rc = fill_dev_params(ctx);
  if (rc != GPDB_OK) {
    strncpy(g_create_err, ctx->err, sizeof(g_create_err) - 1);
    delete ctx->st;
    delete ctx;
    return rc;
  }
  bool ok = cudaSetDevice(ctx->device) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc(&ctx->dp, sizeof(DevParams)) == cudaSuccess &&
            cudaMalloc(&ctx->d_err, sizeof(int) * GPDB_NERR) == cudaSuccess &&
            cudaMalloc(&ctx->d_qtab, sizeof(double) * GPDB_QTAB_SIZE) == cudaSuccess;
  if (ok) {
    double qt[GPDB_QTAB_SIZE];
    gpdb_build_qtab(qt);
    ok = cudaMemcpy(ctx->d_qtab, qt, sizeof(qt), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(ctx->dp, &ctx->hp, sizeof(DevParams), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemset(ctx->d_err, 0, sizeof(int) * GPDB_NERR) == cudaSuccess;
  }
  for (int i = 0; ok && i < 8; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  ok = ok && gpdb_pipe_create(ctx) == GPDB_OK;
  ctx->overlap_hands = !(getenv("GPD_B200_OVERLAP") && getenv("GPD_B200_OVERLAP")[0] == '0');
  if (!ok) {
    gpdb_set_error(nullptr, GPDB_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    gpdb_destroy(ctx);
    return GPDB_ERR_CUDA;
  }
  *ctx_out = ctx;
  return GPDB_OK;
}

void gpdb_destroy(gpdb_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  gpdb_comm_destroy(ctx);
  gpdb_pipe_destroy(ctx);
  cudaFree(ctx->dp);
  cudaFree(ctx->d_err);
  cudaFree(ctx->d_prof);
  cudaFree(ctx->d_qtab);
  cudaFree(ctx->d_pts4);
  cudaFree(ctx->d_xyz);
  cudaFree(ctx->d_nrm);
  cudaFree(ctx->d_cam);
  cudaFree(ctx->d_src);
  cudaFree(ctx->d_samples);
  cudaFree(ctx->d_sel);
  cudaFree(ctx->d_cell_start);
  float *w[8] = {ctx->w.c1w, ctx->w.c1b, ctx->w.c2w, ctx->w.c2b, ctx->w.i1w, ctx->w.i1b, ctx->w.i2w, ctx->w.i2b};
  for (float *p : w) cudaFree(p);
  cudaFree(ctx->tc.b1);
  cudaFree(ctx->tc.b2);
  cudaFree(ctx->tc.b3);
  for (int i = 0; i < 24; i++) cudaFree(ctx->scratch[i]);
  for (int i = 0; i < 8; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx->st;
  delete ctx;
}

int gpdb_set_weights(gpdb_ctx *ctx, const float *conv1_w, const float *conv1_b, const float *conv2_w,
                     const float *conv2_b, const float *ip1_w, const float *ip1_b, const float *ip2_w,
                     const float *ip2_b) {
  if (!ctx) return GPDB_ERR_INVALID;
  const float *w[8] = {conv1_w, conv1_b, conv2_w, conv2_b, ip1_w, ip1_b, ip2_w, ip2_b};
  for (const float *p : w)
    if (!p) {
      gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_set_weights: null weight array");
      return GPDB_ERR_INVALID;
    }
  CUDA_TRY(cudaSetDevice(ctx->device));
  return lenet_upload(ctx, w);
}

// readBinaryFileIntoVector (eigen_classifier.cpp:185-205)
int gpdb_load_weights_dir(gpdb_ctx *ctx, const char *dir) {
  if (!ctx || !dir) return GPDB_ERR_INVALID;
  const int C = ctx->prm.image_num_channels;
  const char *names[8] = {"conv1_weights", "conv1_biases", "conv2_weights", "conv2_biases",
                          "ip1_weights",   "ip1_biases",   "ip2_weights",   "ip2_biases"};
  const size_t sizes[8] = {(size_t)20 * C * 25, 20, 50 * 20 * 25, 50, (size_t)500 * 7200, 500, 1000, 2};
  std::vector<std::vector<float>> bufs(8);
  for (int i = 0; i < 8; i++) {
    std::string path = std::string(dir) + names[i] + ".bin";
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
      gpdb_set_error(ctx, GPDB_ERR_IO, "Cannot open file: %s", path.c_str());
      return GPDB_ERR_IO;
    }
    bufs[i].resize(sizes[i]);
    size_t got = fread(bufs[i].data(), sizeof(float), sizes[i], f);
    char extra;
    bool more = fread(&extra, 1, 1, f) == 1;
    fclose(f);
    if (got != sizes[i] || more) {
      gpdb_set_error(ctx, GPDB_ERR_IO, "%s: expected %zu float32 values for %d channels", path.c_str(), sizes[i], C);
      return GPDB_ERR_IO;
    }
  }
  return gpdb_set_weights(ctx, bufs[0].data(), bufs[1].data(), bufs[2].data(), bufs[3].data(), bufs[4].data(),
                          bufs[5].data(), bufs[6].data(), bufs[7].data());
}

}  // extern "C"

// Cloud arrays are an arena: grown when a larger cloud arrives, never shrunk, so that repeated gpdb_set_cloud /
// gpdb_preprocess calls (one per camera frame) do not pay cudaMalloc / cudaFree.
int gpdb_cloud_reserve(gpdb_ctx *ctx, size_t n) {
  if (n <= ctx->cloud_cap) return GPDB_OK;
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->d_pts4); ctx->d_pts4 = nullptr;
  cudaFree(ctx->d_xyz); ctx->d_xyz = nullptr;
  cudaFree(ctx->d_nrm); ctx->d_nrm = nullptr;
  cudaFree(ctx->d_cam); ctx->d_cam = nullptr;
  cudaFree(ctx->d_src); ctx->d_src = nullptr;
  ctx->cloud_cap = 0;
  const size_t cap = n + n / 8 + 1024;
  CUDA_TRY(cudaMalloc(&ctx->d_pts4, sizeof(float4) * cap));
  CUDA_TRY(cudaMalloc(&ctx->d_xyz, sizeof(float) * 3 * cap));
  CUDA_TRY(cudaMalloc(&ctx->d_nrm, sizeof(double) * 3 * cap));
  CUDA_TRY(cudaMalloc(&ctx->d_cam, cap));
  CUDA_TRY(cudaMalloc(&ctx->d_src, sizeof(int) * cap));
  ctx->cloud_cap = cap;
  return GPDB_OK;
}

// The device arrays d_xyz / d_nrm / d_cam hold N points (uploaded by gpdb_set_cloud or received by ncclBroadcast): make
// them the context's cloud and build the neighbour grid (bounds by a device reduction).
int gpdb_install_device_cloud(gpdb_ctx *ctx, int N, int K, const double *view_points, int all_seen) {
  ctx->cloud_set = false;
  ctx->has_src = false;
  ctx->hp.K = K;
  ctx->hp.all_seen = all_seen;
  for (int k = 0; k < K; k++)
    for (int r = 0; r < 3; r++) ctx->hp.vp[k][r] = view_points[3 * k + r];
  ctx->N = N;
  ctx->K = K;
  ctx->cloud.pts4 = ctx->d_pts4;
  ctx->cloud.xyz = ctx->d_xyz;
  ctx->cloud.nrm = ctx->d_nrm;
  ctx->cloud.cam = ctx->d_cam;
  ctx->cloud.samples = nullptr;  // a new cloud drops the sample positions
  ctx->cloud.n_points = N;
  ctx->n_samples = 0;
  float lo[3], hi[3];
  int *d_bounds = (int *)gpdb_scratch(ctx, 4, sizeof(double) * 6 + sizeof(int) * 8);
  if (!d_bounds) return GPDB_ERR_CUDA;
  int rc = pre_bounds(ctx, ctx->d_xyz, N, d_bounds, lo, hi);
  if (rc == GPDB_OK) rc = geo_build_grid(ctx, lo, hi, N);
  if (rc != GPDB_OK) return rc;
  ctx->cloud_set = true;
  return GPDB_OK;
}

extern "C" {

int gpdb_set_cloud(gpdb_ctx *ctx, const float *xyz, const double *normals, const int32_t *cam_source, int32_t N,
                   const double *view_points, int32_t K) {
  if (!ctx) return GPDB_ERR_INVALID;
  if (!xyz || !normals || !view_points || N <= 0 || K <= 0 || K > GPDB_MAX_CAMERAS) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_set_cloud: need xyz, normals, view_points, N > 0, 1 <= cameras <= %d",
                   GPDB_MAX_CAMERAS);
    return GPDB_ERR_INVALID;
  }
  CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->cloud_set = false;
  cudaStreamSynchronize(ctx->stream);
  {
    int rc = gpdb_cloud_reserve(ctx, (size_t)N);
    if (rc != GPDB_OK) return rc;
  }
  ctx->has_src = false;
  std::vector<uint8_t> cam((size_t)N, (uint8_t)((1u << K) - 1));
  if (cam_source)
    for (int i = 0; i < N; i++) {
      uint8_t m = 0;
      for (int k = 0; k < K; k++)
        if (cam_source[(size_t)i * K + k] > 0) m |= (uint8_t)(1u << k);
      cam[i] = m;
    }
  CUDA_TRY(cudaMemcpyAsync(ctx->d_xyz, xyz, sizeof(float) * 3 * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_nrm, normals, sizeof(double) * 3 * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_cam, cam.data(), (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < N; i++)
    for (int a = 0; a < 3; a++)
      if (!std::isfinite(xyz[3 * (size_t)i + a])) {
        gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_set_cloud: point %d has a non-finite coordinate (run removeNans / gpdb_preprocess first)", i);
        return GPDB_ERR_INVALID;
      }
  bool all_seen = true;  // no per-point camera mask has to be read by the image kernel when every camera sees every point
  for (int i = 0; i < N && all_seen; i++) all_seen = cam[i] == (uint8_t)((1u << K) - 1);
  return gpdb_install_device_cloud(ctx, N, K, view_points, all_seen ? 1 : 0);
}

void gpdb_preprocess_params_default(gpdb_preprocess_params *p) {
  memset(p, 0, sizeof(*p));
  const double ws[6] = {-1, 1, -1, 1, -1, 1};
  for (int i = 0; i < 6; i++) p->workspace[i] = ws[i];
  p->voxel_size = 0.003;
  p->normals_radius = 0.03;
  p->voxelize = 1;
  p->estimate_normals = 1;
}

int gpdb_preprocess(gpdb_ctx *ctx, const float *xyz, const double *normals, const int32_t *cam_source, int32_t M,
                    const double *view_points, int32_t K, const gpdb_preprocess_params *pp) {
  if (!ctx) return GPDB_ERR_INVALID;
  if (!xyz || !view_points || !pp || M <= 0 || K <= 0 || K > GPDB_MAX_CAMERAS) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_preprocess: need xyz, view_points, params, n_points > 0, 1 <= cameras <= %d",
                   GPDB_MAX_CAMERAS);
    return GPDB_ERR_INVALID;
  }
  if (!pp->estimate_normals && !normals) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_preprocess: estimate_normals = 0 needs the caller's normals");
    return GPDB_ERR_INVALID;
  }
  if ((pp->voxelize && !(pp->voxel_size > 0.0)) || (pp->estimate_normals && !(pp->normals_radius > 0.0))) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_preprocess: voxel_size and normals_radius must be positive");
    return GPDB_ERR_INVALID;
  }
  CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->cloud_set = false;
  cudaEvent_t ev[6];
  for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
  auto drop_events = [&]() { for (auto &e : ev) cudaEventDestroy(e); };
  // ---- upload the raw cloud (camera source packed to one bit per camera, as gpdb_set_cloud does)
  std::vector<uint8_t> cam((size_t)M, (uint8_t)((1u << K) - 1));
  if (cam_source)
    for (int i = 0; i < M; i++) {
      uint8_t m = 0;
      for (int k = 0; k < K; k++)
        if (cam_source[(size_t)i * K + k] > 0) m |= (uint8_t)(1u << k);
      cam[i] = m;
    }
  const size_t raw_bytes = sizeof(float) * 3 * (size_t)M + (size_t)M + 16 + (normals ? sizeof(double) * 3 * (size_t)M : 0);
  unsigned char *raw = (unsigned char *)gpdb_scratch(ctx, 7, raw_bytes);
  if (!raw) { drop_events(); return GPDB_ERR_CUDA; }
  double *d_nrm_raw = normals ? (double *)raw : nullptr;
  float *d_xyz_raw = (float *)(raw + (normals ? sizeof(double) * 3 * (size_t)M : 0));
  uint8_t *d_cam_raw = (uint8_t *)(d_xyz_raw + 3 * (size_t)M);
  cudaEventRecord(ev[0], ctx->stream);
  CUDA_TRY(cudaMemcpyAsync(d_xyz_raw, xyz, sizeof(float) * 3 * (size_t)M, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(d_cam_raw, cam.data(), (size_t)M, cudaMemcpyHostToDevice, ctx->stream));
  if (normals) CUDA_TRY(cudaMemcpyAsync(d_nrm_raw, normals, sizeof(double) * 3 * (size_t)M, cudaMemcpyHostToDevice, ctx->stream));
  cudaEventRecord(ev[1], ctx->stream);
  // ---- removeNans + filterWorkspace + voxelizeCloud
  int N = 0;
  int rc = pre_filter_voxelize(ctx, d_xyz_raw, d_cam_raw, d_nrm_raw, M, *pp, &N, ev[2]);
  if (rc != GPDB_OK) { drop_events(); return rc; }
  cudaEventRecord(ev[3], ctx->stream);
  // ---- install as the context's cloud (the arrays were written in place)
  ctx->N = N;
  ctx->K = K;
  ctx->hp.all_seen = cam_source ? 0 : 1;  // without a camera-source matrix every point counts as seen by every camera
  ctx->hp.K = K;
  for (int k = 0; k < K; k++)
    for (int r = 0; r < 3; r++) ctx->hp.vp[k][r] = view_points[3 * k + r];
  memset(ctx->pre_ms, 0, sizeof(ctx->pre_ms));
  if (N == 0) { drop_events(); return 0; }
  ctx->has_src = true;
  ctx->cloud.pts4 = ctx->d_pts4;
  ctx->cloud.xyz = ctx->d_xyz;
  ctx->cloud.nrm = ctx->d_nrm;
  ctx->cloud.cam = ctx->d_cam;
  ctx->cloud.samples = nullptr;  // a new cloud drops the sample positions
  ctx->cloud.n_points = N;
  ctx->n_samples = 0;
  float lo[3], hi[3];
  int *d_bounds = (int *)gpdb_scratch(ctx, 4, sizeof(double) * 6 + sizeof(int) * 8);
  if (!d_bounds) { drop_events(); return GPDB_ERR_CUDA; }
  rc = pre_bounds(ctx, ctx->d_xyz, N, d_bounds, lo, hi);
  if (rc == GPDB_OK) rc = geo_build_grid(ctx, lo, hi, N);
  if (rc != GPDB_OK) { drop_events(); return rc; }
  cudaEventRecord(ev[4], ctx->stream);
  // ---- calculateNormalsOMP + reverseNormals
  if (pp->estimate_normals) {
    rc = pre_normals(ctx, pp->normals_radius);
    if (rc != GPDB_OK) { drop_events(); return rc; }
  }
  cudaEventRecord(ev[5], ctx->stream);
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  float t;
  const int a[5] = {0, 1, 2, 3, 4}, b[5] = {1, 2, 3, 4, 5};
  for (int i = 0; i < 5; i++)
    if (cudaEventElapsedTime(&t, ev[a[i]], ev[b[i]]) == cudaSuccess) ctx->pre_ms[i] = t;
  if (cudaEventElapsedTime(&t, ev[0], ev[5]) == cudaSuccess) ctx->pre_ms[5] = t;
  drop_events();
  ctx->cloud_set = true;
  return N;
}

int gpdb_set_samples(gpdb_ctx *ctx, const double *samples, int32_t n) {
  if (!ctx || n < 0 || (n > 0 && !samples)) return GPDB_ERR_INVALID;
  if (!ctx->cloud_set) {
    gpdb_set_error(ctx, GPDB_ERR_STATE, "no point cloud: call gpdb_set_cloud / gpdb_preprocess first");
    return GPDB_ERR_STATE;
  }
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->d_samples);
  ctx->d_samples = nullptr;
  ctx->n_samples = 0;
  ctx->cloud.samples = nullptr;
  if (n > 0) {
    CUDA_TRY(cudaMalloc(&ctx->d_samples, sizeof(double) * 3 * (size_t)n));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_samples, samples, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->cloud.samples = ctx->d_samples;
    ctx->n_samples = n;
  }
  return ctx->N;  // the first sample index that addresses samples[0]
}

int gpdb_get_cloud(gpdb_ctx *ctx, float *xyz_out, double *normals_out, int32_t *cam_source_out) {
  if (!ctx) return GPDB_ERR_INVALID;
  if (!ctx->cloud_set) {
    gpdb_set_error(ctx, GPDB_ERR_STATE, "no point cloud: call gpdb_set_cloud / gpdb_preprocess first");
    return GPDB_ERR_STATE;
  }
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t N = (size_t)ctx->N;
  if (xyz_out) CUDA_TRY(cudaMemcpyAsync(xyz_out, ctx->d_xyz, sizeof(float) * 3 * N, cudaMemcpyDeviceToHost, ctx->stream));
  if (normals_out) CUDA_TRY(cudaMemcpyAsync(normals_out, ctx->d_nrm, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, ctx->stream));
  if (cam_source_out) {
    int *d = (int *)gpdb_scratch(ctx, 5, sizeof(int) * N * ctx->K);
    if (!d) return GPDB_ERR_CUDA;
    int rc = pre_cam_expand(ctx, d);
    if (rc != GPDB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(cam_source_out, d, sizeof(int) * N * ctx->K, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return ctx->N;
}

int gpdb_get_cloud_source_index(gpdb_ctx *ctx, int32_t *src_out) {
  if (!ctx || !src_out) return GPDB_ERR_INVALID;
  if (!ctx->cloud_set || !ctx->has_src) {
    gpdb_set_error(ctx, GPDB_ERR_STATE, "no preprocessed cloud: call gpdb_preprocess first");
    return GPDB_ERR_STATE;
  }
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaMemcpyAsync(src_out, ctx->d_src, sizeof(int) * (size_t)ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return ctx->N;
}

int gpdb_preprocess_timings(const gpdb_ctx *ctx, double ms_out[6]) {
  if (!ctx || !ms_out) return GPDB_ERR_INVALID;
  for (int i = 0; i < 6; i++) ms_out[i] = ctx->pre_ms[i];
  return GPDB_OK;
}

}  // extern "C"

// ---- pipeline ------------------------------------------------------------------------------------

// Page-locked host memory of one gpdb_result. The device writes the result arrays straight into it (no pageable
// bounce, no second copy), chunk by chunk on the copy stream while the next chunk computes; gpdb_free_result hands it
// back to the context for the next call. Reference-counted: a result may outlive its context.
struct HostArena {
  std::atomic<int> refs{1};        // the owning context + an outstanding result
  std::atomic<bool> in_use{false};
  void *buf[4] = {nullptr, nullptr, nullptr, nullptr};  // 0: per-sample / per-pose arrays, 1: candidate records, 2: images,
  size_t cap[4] = {0, 0, 0, 0};                         // 3: extra (gathered per-pose arrays of gpdb_detect_sharded)
};
static void arena_unref(HostArena *a) {
  if (a->refs.fetch_sub(1) == 1) {
    for (void *b : a->buf)
      if (b) cudaFreeHost(b);
    delete a;
  }
}
// grows buffer `which` to at least `bytes`, keeping the first `keep` bytes (the caller has drained the copies into it)
static bool arena_reserve(HostArena *a, int which, size_t bytes, size_t keep) {
  if (a->cap[which] >= bytes) return true;
  const size_t want = bytes + bytes / 2 + 4096;
  void *nb = nullptr;
  if (cudaHostAlloc(&nb, want, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (a->buf[which]) {
    if (keep) memcpy(nb, a->buf[which], keep);
    cudaFreeHost(a->buf[which]);
  }
  a->buf[which] = nb;
  a->cap[which] = want;
  return true;
}

struct PipeState {
  cudaStream_t copy = nullptr;  // D2H of finished chunks + the per-chunk candidate count, concurrent with compute
  cudaStream_t hands = nullptr; // hand search + compaction of the chunks ahead, concurrent with images / LeNet of the current one
  cudaEvent_t ev_compact[2], ev_count[2], ev_scored[2], ev_copied[2], ev_frames, ev_consumed[2], ev_hands_done;
  int *h_count = nullptr;       // pinned [2]
  std::vector<HostArena *> arenas;
  bool ok = false;
};
int gpdb_pipe_create(gpdb_ctx *ctx) {
  PipeState *ps = new PipeState();
  ctx->pipe = ps;
  bool ok = cudaStreamCreateWithFlags(&ps->copy, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ps->hands, cudaStreamNonBlocking) == cudaSuccess &&
            cudaHostAlloc((void **)&ps->h_count, 2 * sizeof(int), cudaHostAllocDefault) == cudaSuccess;
  cudaEvent_t *evs[12] = {&ps->ev_compact[0], &ps->ev_compact[1], &ps->ev_count[0], &ps->ev_count[1], &ps->ev_scored[0],
                          &ps->ev_scored[1], &ps->ev_copied[0], &ps->ev_copied[1], &ps->ev_frames, &ps->ev_consumed[0],
                          &ps->ev_consumed[1], &ps->ev_hands_done};
  for (cudaEvent_t *e : evs) {
    *e = nullptr;
    ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
  }
  ps->ok = ok;
  return ok ? GPDB_OK : GPDB_ERR_CUDA;
}
void gpdb_pipe_destroy(gpdb_ctx *ctx) {
  PipeState *ps = ctx->pipe;
  if (!ps) return;
  if (ps->copy) {
    cudaStreamSynchronize(ps->copy);
    cudaStreamDestroy(ps->copy);
  }
  if (ps->hands) {
    cudaStreamSynchronize(ps->hands);
    cudaStreamDestroy(ps->hands);
  }
  cudaEvent_t evs[12] = {ps->ev_compact[0], ps->ev_compact[1], ps->ev_count[0], ps->ev_count[1], ps->ev_scored[0],
                         ps->ev_scored[1], ps->ev_copied[0], ps->ev_copied[1], ps->ev_frames, ps->ev_consumed[0],
                         ps->ev_consumed[1], ps->ev_hands_done};
  for (cudaEvent_t e : evs)
    if (e) cudaEventDestroy(e);
  if (ps->h_count) cudaFreeHost(ps->h_count);
  for (HostArena *a : ps->arenas) arena_unref(a);
  delete ps;
  ctx->pipe = nullptr;
}
// pinned memory that lives and dies with a result (gpdb_detect_sharded: the gathered per-pose arrays of all ranks)
void *gpdb_result_extra(gpdb_result *r, size_t bytes) {
  if (!r || !r->owner_) return nullptr;
  HostArena *a = (HostArena *)r->owner_;
  return arena_reserve(a, 3, bytes, 0) ? a->buf[3] : nullptr;
}

static HostArena *arena_acquire(gpdb_ctx *ctx) {
  PipeState &ps = *ctx->pipe;
  for (HostArena *a : ps.arenas) {
    bool expect = false;
    if (a->in_use.compare_exchange_strong(expect, true)) {
      a->refs.fetch_add(1);
      return a;
    }
  }
  if (ps.arenas.size() >= 8) {  // results that were never freed: do not pin host memory without bound
    gpdb_set_error(ctx, GPDB_ERR_STATE, "8 results of this context are outstanding: release them with gpdb_free_result");
    return nullptr;
  }
  HostArena *a = new HostArena();
  a->in_use = true;
  a->refs = 2;
  ps.arenas.push_back(a);
  return a;
}

int gpdb_check_state(gpdb_ctx *ctx, bool need_cloud, bool need_weights) {
  if (!ctx) return GPDB_ERR_INVALID;
  if (need_cloud && !ctx->cloud_set) {
    gpdb_set_error(ctx, GPDB_ERR_STATE, "no point cloud: call gpdb_set_cloud first");
    return GPDB_ERR_STATE;
  }
  if (need_weights && !ctx->w.set) {
    gpdb_set_error(ctx, GPDB_ERR_STATE, "no classifier weights: call gpdb_load_weights_dir / gpdb_set_weights first");
    return GPDB_ERR_STATE;
  }
  CUDA_TRY(cudaSetDevice(ctx->device));
  return GPDB_OK;
}

namespace {

int check_device_errors(gpdb_ctx *ctx) {
  int e[GPDB_NERR];
  CUDA_TRY(cudaMemcpyAsync(e, ctx->d_err, sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (e[0] || e[1] || e[2]) {
    CUDA_TRY(cudaMemsetAsync(ctx->d_err, 0, sizeof(e), ctx->stream));
    gpdb_set_error(ctx, GPDB_ERR_CAPACITY,
                   "neighbourhood exceeded an on-chip tile (frame ball: %d samples, hand-search ball: %d samples, "
                   "image box: %d images): the cloud is denser than the supported %d / %d / %d points",
                   e[0], e[1], e[2], 16384, 131072, 32768);
    return GPDB_ERR_CAPACITY;
  }
  return GPDB_OK;
}

}  // namespace

// The chunked device pipeline behind gpdb_detect / gpdb_hand_search / gpdb_detect_resident / gpdb_detect_sharded.
//   resident == false: sample_idx is a HOST array, every result is copied back to the host (out)
//   resident == true : sample_idx, flags_ext, scores_ext are DEVICE arrays; nothing but the per-chunk
//                      candidate count crosses PCIe (out receives counts and timings only)
// select_k >= 0 (gpdb_detect_select): the classified candidates of all chunks stay on the device, the select_k best are
// sorted out there and only they are copied back; no per-sample / per-pose array is returned.
//
// Stream schedule (H = hand search + compaction of a chunk, I/L/S = images, LeNet, score scatter):
//   compute: F  H0  H1  I0 L0 S0  H2  I1 L1 S1  ...      copy:  frames | n0 | n1 | cand0 flags0 scores0 | n2 | cand1 ...
// The candidate count of chunk i is read back on the copy stream while H(i+1) runs, so the device never waits for the
// host; every chunk is ONE k_images / LeNet launch sized to its candidate count (no tail launches), and the results of
// chunk i go to the pinned arena while chunk i+1 computes.
int gpdb_run_pipeline(gpdb_ctx *ctx, const int32_t *sample_idx, int32_t n, gpdb_result *out, bool with_images_and_scores,
                      bool resident, uint8_t *flags_ext, float *scores_ext, int select_k, int slot_base) {
  const bool selecting = select_k >= 0;
  memset(out, 0, sizeof(*out));
  PipeState &ps = *ctx->pipe;
  const int P = ctx->hp.P, S = ctx->hp.S, C = ctx->hp.C;
  const size_t isz = (size_t)S * S * C;    // one image in the cv::Mat layout (what the caller receives)
  const size_t psz = (size_t)S * S * 16;   // one image in the device layout (16-byte pixels, see k_images)
  out->n_samples = n;
  out->poses_per_sample = P;
  if (!resident)
    for (int i = 0; i < n; i++)
      if (sample_idx[i] < 0 || sample_idx[i] >= ctx->N + ctx->n_samples) {
        gpdb_set_error(ctx, GPDB_ERR_INVALID, "sample index %d at position %d outside the cloud (N = %d, + %d sample positions)",
                       sample_idx[i], i, ctx->N, ctx->n_samples);
        return GPDB_ERR_INVALID;
      }
  const int64_t launches0 = ctx->launches;
  double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int chunk = ctx->prm.chunk_samples > 0 ? ctx->prm.chunk_samples : 16384;
  const int batch_cap = ctx->prm.batch_size > 0 ? ctx->prm.batch_size : 32768;  // images per k_images / LeNet launch
  const bool keep = with_images_and_scores && ctx->prm.keep_images && !resident && !selecting;
  const bool to_host = !resident && !selecting;  // per-sample / per-pose arrays + all candidate records go to the host
  const size_t nP = (size_t)n * P;
  const int cmax = std::min(chunk, std::max(n, 1));
  int *d_sidx = resident ? const_cast<int *>(sample_idx) : (int *)gpdb_scratch(ctx, 7, sizeof(int) * (size_t)n);
  double *d_frames = (double *)gpdb_scratch(ctx, 8, sizeof(double) * 9 * (size_t)n);
  uint8_t *d_valid = (uint8_t *)gpdb_scratch(ctx, 9, (size_t)n);
  uint8_t *d_flags = resident ? flags_ext : (uint8_t *)gpdb_scratch(ctx, 10, nP);
  float *d_pscores = resident ? scores_ext : (float *)gpdb_scratch(ctx, 11, sizeof(float) * nP);
  gpdb_pose *d_poses = (gpdb_pose *)gpdb_scratch(ctx, 12, sizeof(gpdb_pose) * (size_t)cmax * P);
  gpdb_pose *d_cand2 = (gpdb_pose *)gpdb_scratch(ctx, 13, 2 * sizeof(gpdb_pose) * (size_t)cmax * P);  // double-buffered
  int *d_count = (int *)gpdb_scratch(ctx, 14, 64);
  if (!d_sidx || !d_frames || !d_valid || !d_flags || !d_pscores || !d_poses || !d_cand2 || !d_count) return GPDB_ERR_CUDA;
  gpdb_pose *d_cand[2] = {d_cand2, d_cand2 + (size_t)cmax * P};

  HostArena *ar = nullptr;
  int rc = GPDB_OK;
  // every exit after this point goes through finish(): drains both streams, clears the device error counters, collects
  // the stage timers and releases the arena on failure, so that a failed call leaves no state behind
  auto finish = [&](int code) -> int {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ps.hands);
    cudaStreamSynchronize(ps.copy);
    if (code >= 0) {
      int e = check_device_errors(ctx);
      if (e != GPDB_OK) code = e;
    } else {
      cudaMemsetAsync(ctx->d_err, 0, sizeof(int) * GPDB_NERR, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
    }
    st_collect(ctx, ms);
    for (int i = 0; i < 8; i++) ctx->last_ms[i] = ms[i];
    if (code < 0) {
      if (ar) {
        ar->in_use = false;
        arena_unref(ar);
      }
      memset(out, 0, sizeof(*out));
    }
    return code;
  };
#define PIPE_TRY(expr)                                \
  do {                                                \
    if ((rc = (expr)) != GPDB_OK) return finish(rc);  \
  } while (0)
#define PIPE_CUDA(expr)                                                                                           \
  do {                                                                                                            \
    cudaError_t e__ = (expr);                                                                                     \
    if (e__ != cudaSuccess) {                                                                                     \
      gpdb_set_error(ctx, GPDB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));   \
      return finish(GPDB_ERR_CUDA);                                                                               \
    }                                                                                                             \
  } while (0)

  // host-side layout of the fixed-size arrays inside arena buffer 0
  const size_t off_valid = 0, off_frames = ((size_t)n + 63) / 64 * 64, off_flags = off_frames + sizeof(double) * 9 * (size_t)n,
               off_scores = (off_flags + nP + 63) / 64 * 64, fixed_bytes = off_scores + sizeof(float) * nP + 64;
  if (!resident) {
    ar = arena_acquire(ctx);
    if (!ar) return GPDB_ERR_STATE;
    const size_t guess = std::max((size_t)1024, nP / 8);  // grown on demand
    if (!arena_reserve(ar, 0, to_host ? fixed_bytes : 64, 0) ||
        !arena_reserve(ar, 1, sizeof(gpdb_pose) * (selecting ? (size_t)std::max(select_k, 1) : guess), 0)) {
      gpdb_set_error(ctx, GPDB_ERR_CUDA, "cudaHostAlloc of the result arena failed");
      return finish(GPDB_ERR_CUDA);
    }
  }
  uint8_t *h_fixed = ar ? (uint8_t *)ar->buf[0] : nullptr;

  ctx->st->spans.clear();
  ctx->st->used = 0;
  cudaEvent_t t_all = gpdb_st_begin(ctx);
  if (n > 0) {
    if (!resident)
      PIPE_CUDA(cudaMemcpyAsync(d_sidx, sample_idx, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    PIPE_CUDA(cudaMemsetAsync(d_pscores, 0xFF, sizeof(float) * nP, ctx->stream));  // 0xFFFFFFFF = NaN
  }
  cudaEvent_t t0 = gpdb_st_begin(ctx);
  PIPE_TRY(geo_frames(ctx, d_sidx, n, d_frames, d_valid));
  gpdb_st_end(ctx, 0, t0);
  // The hand search of the chunks AHEAD runs on its own stream: its CTAs (54 KB, 64 registers) fill what the tensor-core
  // kernels of the current chunk leave idle (conv1: one 162 KB CTA per SM at 42 % issue utilisation).
  // gpdb_set_overlap(ctx, 0) / GPD_B200_OVERLAP=0 put everything on one stream (exclusive stage timers, A/B).
  const bool overlap = ctx->overlap_hands;
  cudaStream_t const main_stream = ctx->stream, hs = overlap ? ps.hands : ctx->stream;
  PIPE_CUDA(cudaEventRecord(ps.ev_frames, main_stream));
  if (overlap) PIPE_CUDA(cudaStreamWaitEvent(hs, ps.ev_frames, 0));
  if (to_host && n > 0) {
    PIPE_CUDA(cudaStreamWaitEvent(ps.copy, ps.ev_frames, 0));
    PIPE_CUDA(cudaMemcpyAsync(h_fixed + off_valid, d_valid, (size_t)n, cudaMemcpyDeviceToHost, ps.copy));
    PIPE_CUDA(cudaMemcpyAsync(h_fixed + off_frames, d_frames, sizeof(double) * 9 * (size_t)n, cudaMemcpyDeviceToHost, ps.copy));
  }
  const int nchunks = (n + chunk - 1) / chunk;
  int total_nc = 0;
  size_t img_host = 0;  // bytes of images already placed in arena buffer 2
  bool cand_busy[2] = {false, false};      // a D2H copy out of d_cand[b] has been issued (ev_copied[b] marks its end)
  bool cand_consumed[2] = {false, false};  // ev_consumed[b] marks the end of the main stream's reads of d_cand[b]
  auto launch_hands = [&](int ci) -> int {
    const int c0 = ci * chunk, nn = std::min(chunk, n - c0), b = ci & 1;
    if (cand_busy[b]) CUDA_TRY(cudaStreamWaitEvent(hs, ps.ev_copied[b], 0));  // chunk ci-2 has left d_cand[b] for the host
    if (overlap && cand_consumed[b]) CUDA_TRY(cudaStreamWaitEvent(hs, ps.ev_consumed[b], 0));  // ... and images / scatter read it
    ctx->stream = hs;  // the launchers (and the stage timers) use the context's current stream
    cudaEvent_t t1 = gpdb_st_begin(ctx);
    int r = geo_hands(ctx, d_sidx + c0, nn, c0 + slot_base, d_frames + 9 * (size_t)c0, d_valid + c0, d_poses,
                      d_flags + (size_t)c0 * P);
    if (r == GPDB_OK) r = geo_compact(ctx, d_poses, d_flags + (size_t)c0 * P, nn * P, d_cand[b], d_count + b);
    if (r == GPDB_OK) gpdb_st_end(ctx, 1, t1);
    ctx->stream = main_stream;
    if (r != GPDB_OK) return r;
    CUDA_TRY(cudaEventRecord(ps.ev_compact[b], hs));
    CUDA_TRY(cudaStreamWaitEvent(ps.copy, ps.ev_compact[b], 0));
    CUDA_TRY(cudaMemcpyAsync(ps.h_count + b, d_count + b, sizeof(int), cudaMemcpyDeviceToHost, ps.copy));
    CUDA_TRY(cudaEventRecord(ps.ev_count[b], ps.copy));
    return GPDB_OK;
  };
  if (nchunks > 0) PIPE_TRY(launch_hands(0));
  for (int ci = 0; ci < nchunks; ci++) {
    const int c0 = ci * chunk, nn = std::min(chunk, n - c0), b = ci & 1;
    if (ci + 1 < nchunks) PIPE_TRY(launch_hands(ci + 1));  // queued BEHIND which the host now waits for chunk ci's count
    PIPE_CUDA(cudaEventSynchronize(ps.ev_count[b]));
    const int nc = ps.h_count[b];
    if (overlap) PIPE_CUDA(cudaStreamWaitEvent(main_stream, ps.ev_compact[b], 0));  // d_cand[b], flags of chunk ci are ready
    total_nc += nc;
    uint8_t *d_img = nullptr;  // keep_images: the chunk's images in the cv::Mat layout
    if (with_images_and_scores && nc > 0) {
      float *d_scores = (float *)gpdb_scratch(ctx, 15, sizeof(float) * (size_t)nc);
      const int ib = std::min(nc, batch_cap);
      uint8_t *d_p16 = (uint8_t *)gpdb_scratch(ctx, 0, psz * (size_t)ib);
      if (keep) d_img = (uint8_t *)gpdb_scratch(ctx, 16, isz * (size_t)nc);
      if (!d_scores || !d_p16 || (keep && !d_img)) return finish(GPDB_ERR_CUDA);
      if (keep && ci > 0) PIPE_CUDA(cudaStreamWaitEvent(ctx->stream, ps.ev_copied[b ^ 1], 0));  // d_img is being read
      for (int b0 = 0; b0 < nc; b0 += batch_cap) {
        const int bn = std::min(batch_cap, nc - b0);
        cudaEvent_t t2 = gpdb_st_begin(ctx);
        PIPE_TRY(geo_images(ctx, d_cand[b] + b0, bn, d_p16));
        if (keep) PIPE_TRY(geo_p16_to_hwc(ctx, d_p16, bn, d_img + isz * (size_t)b0));
        gpdb_st_end(ctx, 2, t2);
        cudaEvent_t t3 = gpdb_st_begin(ctx);
        PIPE_TRY(lenet_forward(ctx, d_p16, bn, d_scores + b0, nullptr));
        gpdb_st_end(ctx, 3, t3);
      }
      PIPE_TRY(geo_scatter_scores(ctx, d_cand[b], d_scores, nc, c0 + slot_base, P, d_pscores + (size_t)c0 * P, d_cand[b]));
      if (keep) {
        PIPE_CUDA(cudaStreamSynchronize(ps.copy));  // growing moves the buffer: earlier image copies must have landed
        if (!arena_reserve(ar, 2, img_host + isz * (size_t)nc, img_host)) {
          gpdb_set_error(ctx, GPDB_ERR_CUDA, "cudaHostAlloc of %zu B for the images failed", img_host + isz * (size_t)nc);
          return finish(GPDB_ERR_CUDA);
        }
      }
    }
    if (nc > 0 && selecting) {  // keep the chunk's scored candidates on the device
      if ((size_t)total_nc > ctx->sel_cap) {
        const size_t cap = std::max((size_t)total_nc * 2, (size_t)65536);
        gpdb_pose *grown = nullptr;
        PIPE_CUDA(cudaMalloc(&grown, sizeof(gpdb_pose) * cap));
        if (ctx->d_sel && total_nc > nc)
          cudaMemcpyAsync(grown, ctx->d_sel, sizeof(gpdb_pose) * (size_t)(total_nc - nc), cudaMemcpyDeviceToDevice, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->d_sel);
        ctx->d_sel = grown;
        ctx->sel_cap = cap;
      }
      PIPE_CUDA(cudaMemcpyAsync(ctx->d_sel + (total_nc - nc), d_cand[b], sizeof(gpdb_pose) * (size_t)nc, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (overlap) {
      PIPE_CUDA(cudaEventRecord(ps.ev_consumed[b], main_stream));
      cand_consumed[b] = true;
    }
    if (to_host) {  // this chunk's results leave for the pinned arena while the next chunk computes
      if (sizeof(gpdb_pose) * (size_t)total_nc > ar->cap[1]) {
        PIPE_CUDA(cudaStreamSynchronize(ps.copy));
        if (!arena_reserve(ar, 1, sizeof(gpdb_pose) * (size_t)total_nc, sizeof(gpdb_pose) * (size_t)(total_nc - nc))) {
          gpdb_set_error(ctx, GPDB_ERR_CUDA, "cudaHostAlloc of the candidate arena failed");
          return finish(GPDB_ERR_CUDA);
        }
      }
      PIPE_CUDA(cudaEventRecord(ps.ev_scored[b], ctx->stream));
      PIPE_CUDA(cudaStreamWaitEvent(ps.copy, ps.ev_scored[b], 0));
      if (nc > 0)
        PIPE_CUDA(cudaMemcpyAsync((gpdb_pose *)ar->buf[1] + (total_nc - nc), d_cand[b], sizeof(gpdb_pose) * (size_t)nc,
                                  cudaMemcpyDeviceToHost, ps.copy));
      PIPE_CUDA(cudaMemcpyAsync(h_fixed + off_flags + (size_t)c0 * P, d_flags + (size_t)c0 * P, (size_t)nn * P,
                                cudaMemcpyDeviceToHost, ps.copy));
      PIPE_CUDA(cudaMemcpyAsync(h_fixed + off_scores + sizeof(float) * (size_t)c0 * P, d_pscores + (size_t)c0 * P,
                                sizeof(float) * (size_t)nn * P, cudaMemcpyDeviceToHost, ps.copy));
      if (keep && d_img) {
        PIPE_CUDA(cudaMemcpyAsync((uint8_t *)ar->buf[2] + img_host, d_img, isz * (size_t)nc,
                                  cudaMemcpyDeviceToHost, ps.copy));
        img_host += isz * (size_t)nc;
      }
      PIPE_CUDA(cudaEventRecord(ps.ev_copied[b], ps.copy));
      cand_busy[b] = true;
    }
  }
  int n_sel = 0;
  if (selecting) {
    n_sel = std::min(select_k, total_nc);
    if (n_sel > 0) {
      gpdb_pose *d_top = (gpdb_pose *)gpdb_scratch(ctx, 12, sizeof(gpdb_pose) * (size_t)std::max(n_sel, cmax * P));
      if (!d_top) return finish(GPDB_ERR_CUDA);
      PIPE_TRY(geo_select(ctx, ctx->d_sel, total_nc, n_sel, d_top));
      PIPE_CUDA(cudaMemcpyAsync(ar->buf[1], d_top, sizeof(gpdb_pose) * (size_t)n_sel, cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  gpdb_st_end(ctx, 4, t_all);
  rc = finish(total_nc);
  if (rc < 0) return rc;
#undef PIPE_TRY
#undef PIPE_CUDA
  out->n_candidates = selecting ? n_sel : total_nc;
  out->n_total_candidates = total_nc;
  if (ar) {
    out->owner_ = ar;
    out->candidates = (gpdb_pose *)ar->buf[1];
    if (to_host) {
      out->frame_valid = h_fixed + off_valid;
      out->frames = (double *)(h_fixed + off_frames);
      out->pose_flags = h_fixed + off_flags;
      out->pose_scores = (float *)(h_fixed + off_scores);
      if (keep) out->images = (uint8_t *)ar->buf[2];
    }
  }
  out->ms_candidates = ms[0] + ms[1];
  out->ms_images = ms[2];
  out->ms_classify = ms[3];
  out->kernel_launches = ctx->launches - launches0;
  return out->n_candidates;
}

static int run_pipeline(gpdb_ctx *ctx, const int32_t *sample_idx, int32_t n, gpdb_result *out, bool with_images_and_scores,
                        bool resident, uint8_t *flags_ext, float *scores_ext, int select_k = -1) {
  return gpdb_run_pipeline(ctx, sample_idx, n, out, with_images_and_scores, resident, flags_ext, scores_ext, select_k, 0);
}
static int check_state(gpdb_ctx *ctx, bool need_cloud, bool need_weights) { return gpdb_check_state(ctx, need_cloud, need_weights); }

extern "C" {

int gpdb_detect(gpdb_ctx *ctx, const int32_t *sample_idx, int32_t n, gpdb_result *out) {
  int rc = check_state(ctx, true, true);
  if (rc != GPDB_OK) return rc;
  if (!out || (n > 0 && !sample_idx) || n < 0) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_detect: bad arguments");
    return GPDB_ERR_INVALID;
  }
  return run_pipeline(ctx, sample_idx, n, out, true, false, nullptr, nullptr);
}

int gpdb_detect_select(gpdb_ctx *ctx, const int32_t *sample_idx, int32_t n, int32_t num_selected, gpdb_result *out) {
  int rc = check_state(ctx, true, true);
  if (rc != GPDB_OK) return rc;
  if (!out || (n > 0 && !sample_idx) || n < 0 || num_selected < 0) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_detect_select: bad arguments");
    return GPDB_ERR_INVALID;
  }
  return run_pipeline(ctx, sample_idx, n, out, true, false, nullptr, nullptr, num_selected);
}

int gpdb_detect_resident(gpdb_ctx *ctx, const int32_t *d_sample_idx, int32_t n, uint8_t *d_flags_out,
                         float *d_scores_out, gpdb_result *stats) {
  int rc = check_state(ctx, true, true);
  if (rc != GPDB_OK) return rc;
  if (!stats || n < 0 || (n > 0 && (!d_sample_idx || !d_flags_out || !d_scores_out))) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_detect_resident: bad arguments");
    return GPDB_ERR_INVALID;
  }
  return run_pipeline(ctx, d_sample_idx, n, stats, true, true, d_flags_out, d_scores_out);
}

int gpdb_set_overlap(gpdb_ctx *ctx, int32_t enable) {
  if (!ctx) return GPDB_ERR_INVALID;
  ctx->overlap_hands = enable != 0;
  return GPDB_OK;
}

int gpdb_set_stream(gpdb_ctx *ctx, void *cuda_stream) {
  if (!ctx) return GPDB_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return GPDB_OK;
}

int gpdb_hand_search(gpdb_ctx *ctx, const int32_t *sample_idx, int32_t n, gpdb_result *out) {
  int rc = check_state(ctx, true, false);
  if (rc != GPDB_OK) return rc;
  if (!out || (n > 0 && !sample_idx) || n < 0) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_hand_search: bad arguments");
    return GPDB_ERR_INVALID;
  }
  return run_pipeline(ctx, sample_idx, n, out, false, false, nullptr, nullptr);
}

int gpdb_frames(gpdb_ctx *ctx, const int32_t *sample_idx, int32_t n, double *frames_out, uint8_t *valid_out) {
  int rc = check_state(ctx, true, false);
  if (rc != GPDB_OK) return rc;
  if (n < 0 || (n > 0 && (!sample_idx || !frames_out || !valid_out))) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_frames: bad arguments");
    return GPDB_ERR_INVALID;
  }
  if (n == 0) return 0;
  for (int i = 0; i < n; i++)
    if (sample_idx[i] < 0 || sample_idx[i] >= ctx->N) {
      gpdb_set_error(ctx, GPDB_ERR_INVALID, "sample index %d outside the cloud (N = %d)", sample_idx[i], ctx->N);
      return GPDB_ERR_INVALID;
    }
  int *d_sidx = (int *)gpdb_scratch(ctx, 7, sizeof(int) * (size_t)n);
  double *d_frames = (double *)gpdb_scratch(ctx, 8, sizeof(double) * 9 * (size_t)n);
  uint8_t *d_valid = (uint8_t *)gpdb_scratch(ctx, 9, (size_t)n);
  if (!d_sidx || !d_frames || !d_valid) return GPDB_ERR_CUDA;
  CUDA_TRY(cudaMemcpyAsync(d_sidx, sample_idx, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = geo_frames(ctx, d_sidx, n, d_frames, d_valid)) != GPDB_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(frames_out, d_frames, sizeof(double) * 9 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(valid_out, d_valid, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  if ((rc = check_device_errors(ctx)) != GPDB_OK) return rc;
  return n;
}

int gpdb_images(gpdb_ctx *ctx, const gpdb_pose *poses, int32_t n, uint8_t *images_out) {
  int rc = check_state(ctx, true, false);
  if (rc != GPDB_OK) return rc;
  if (n < 0 || (n > 0 && (!poses || !images_out))) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_images: bad arguments");
    return GPDB_ERR_INVALID;
  }
  const size_t isz = (size_t)ctx->hp.S * ctx->hp.S * ctx->hp.C, psz = (size_t)ctx->hp.S * ctx->hp.S * 16;
  const int batch = 8192;
  for (int b0 = 0; b0 < n; b0 += batch) {
    const int bn = std::min(batch, n - b0);
    gpdb_pose *d_cand = (gpdb_pose *)gpdb_scratch(ctx, 13, sizeof(gpdb_pose) * (size_t)bn);
    uint8_t *d_p16 = (uint8_t *)gpdb_scratch(ctx, 0, psz * (size_t)bn);
    uint8_t *d_img = (uint8_t *)gpdb_scratch(ctx, 16, isz * (size_t)bn);
    if (!d_cand || !d_p16 || !d_img) return GPDB_ERR_CUDA;
    CUDA_TRY(cudaMemcpyAsync(d_cand, poses + b0, sizeof(gpdb_pose) * (size_t)bn, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = geo_images(ctx, d_cand, bn, d_p16)) != GPDB_OK) return rc;
    if ((rc = geo_p16_to_hwc(ctx, d_p16, bn, d_img)) != GPDB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(images_out + isz * (size_t)b0, d_img, isz * (size_t)bn, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  if ((rc = check_device_errors(ctx)) != GPDB_OK) return rc;
  return n;
}

int gpdb_classify(gpdb_ctx *ctx, const uint8_t *images_hwc, int32_t n, float *scores_out, float *logits_out) {
  int rc = check_state(ctx, false, true);
  if (rc != GPDB_OK) return rc;
  if (n < 0 || (n > 0 && (!images_hwc || !scores_out))) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_classify: bad arguments");
    return GPDB_ERR_INVALID;
  }
  const size_t isz = (size_t)ctx->hp.S * ctx->hp.S * ctx->hp.C, psz = (size_t)ctx->hp.S * ctx->hp.S * 16;
  const int batch = ctx->prm.batch_size > 0 ? ctx->prm.batch_size : 8192;
  for (int b0 = 0; b0 < n; b0 += batch) {
    const int bn = std::min(batch, n - b0);
    uint8_t *d_img = (uint8_t *)gpdb_scratch(ctx, 16, isz * (size_t)bn);
    uint8_t *d_p16 = (uint8_t *)gpdb_scratch(ctx, 0, psz * (size_t)bn);
    float *d_scores = (float *)gpdb_scratch(ctx, 15, sizeof(float) * (size_t)bn * 3);
    if (!d_img || !d_p16 || !d_scores) return GPDB_ERR_CUDA;
    float *d_logits = d_scores + bn;
    CUDA_TRY(cudaMemcpyAsync(d_img, images_hwc + isz * (size_t)b0, isz * (size_t)bn, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = geo_hwc_to_p16(ctx, d_img, bn, d_p16)) != GPDB_OK) return rc;  // cv::Mat bytes -> 16-byte pixels
    if ((rc = lenet_forward(ctx, d_p16, bn, d_scores, d_logits)) != GPDB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(scores_out + b0, d_scores, sizeof(float) * (size_t)bn, cudaMemcpyDeviceToHost, ctx->stream));
    if (logits_out)
      CUDA_TRY(cudaMemcpyAsync(logits_out + 2 * (size_t)b0, d_logits, sizeof(float) * 2 * (size_t)bn,
                               cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  return n;
}

int gpdb_reevaluate(gpdb_ctx *ctx, gpdb_pose *hands, int32_t n, int32_t *labels_out) {
  int rc = check_state(ctx, true, false);
  if (rc != GPDB_OK) return rc;
  if (n < 0 || (n > 0 && (!hands || !labels_out))) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_reevaluate: bad arguments");
    return GPDB_ERR_INVALID;
  }
  if (n == 0) return 0;
  gpdb_pose *d_h = (gpdb_pose *)gpdb_scratch(ctx, 17, sizeof(gpdb_pose) * (size_t)n);
  int *d_l = (int *)gpdb_scratch(ctx, 18, sizeof(int) * (size_t)n);
  if (!d_h || !d_l) return GPDB_ERR_CUDA;
  CUDA_TRY(cudaMemcpyAsync(d_h, hands, sizeof(gpdb_pose) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = geo_reeval(ctx, d_h, n, d_l)) != GPDB_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(hands, d_h, sizeof(gpdb_pose) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(labels_out, d_l, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return n;
}

int gpdb_find_clusters(gpdb_ctx *ctx, const gpdb_pose *hands, int32_t n, int32_t min_inliers, gpdb_pose *clusters_out) {
  if (!ctx) return GPDB_ERR_INVALID;
  if (n < 0 || (n > 0 && (!hands || !clusters_out))) {
    gpdb_set_error(ctx, GPDB_ERR_INVALID, "gpdb_find_clusters: bad arguments");
    return GPDB_ERR_INVALID;
  }
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(ctx->device));
  gpdb_pose *d_in = (gpdb_pose *)gpdb_scratch(ctx, 17, sizeof(gpdb_pose) * (size_t)n * 3);
  uint8_t *d_keep = (uint8_t *)gpdb_scratch(ctx, 18, (size_t)n);
  int *d_count = (int *)gpdb_scratch(ctx, 14, 64);
  if (!d_in || !d_keep || !d_count) return GPDB_ERR_CUDA;
  gpdb_pose *d_dense = d_in + n, *d_out = d_dense + n;
  CUDA_TRY(cudaMemcpyAsync(d_in, hands, sizeof(gpdb_pose) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  int rc;
  if ((rc = geo_clusters(ctx, d_in, n, min_inliers, d_dense, d_keep)) != GPDB_OK) return rc;
  if ((rc = geo_compact(ctx, d_dense, d_keep, n, d_out, d_count + 2)) != GPDB_OK) return rc;  // order of i kept
  int nc = 0;
  CUDA_TRY(cudaMemcpyAsync(&nc, d_count + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (nc > 0) {
    CUDA_TRY(cudaMemcpyAsync(clusters_out, d_out, sizeof(gpdb_pose) * (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  return nc;
}

void gpdb_free_result(gpdb_result *r) {
  if (!r) return;
  if (r->owner_) {  // the arrays live in a pinned arena of the context that produced them: hand it back
    HostArena *a = (HostArena *)r->owner_;
    a->in_use = false;
    arena_unref(a);
  } else {
    free(r->frame_valid);
    free(r->frames);
    free(r->pose_flags);
    free(r->pose_scores);
    free(r->candidates);
    free(r->images);
  }
  memset(r, 0, sizeof(*r));
}

int gpdb_debug_phase_cycles(gpdb_ctx *ctx, int enable, uint64_t cycles_out[16]) {
  if (!ctx) return GPDB_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (ctx->d_prof && cycles_out)
    CUDA_TRY(cudaMemcpy(cycles_out, ctx->d_prof, sizeof(uint64_t) * 16, cudaMemcpyDeviceToHost));
  if (enable && !ctx->d_prof) CUDA_TRY(cudaMalloc(&ctx->d_prof, sizeof(uint64_t) * 16));
  if (enable) CUDA_TRY(cudaMemset(ctx->d_prof, 0, sizeof(uint64_t) * 16));
  if (!enable && ctx->d_prof) {
    cudaFree(ctx->d_prof);
    ctx->d_prof = nullptr;
  }
  return GPDB_OK;
}

int gpdb_last_timings(const gpdb_ctx *ctx, double ms_out[8]) {
  if (!ctx || !ms_out) return GPDB_ERR_INVALID;
  for (int i = 0; i < 8; i++) ms_out[i] = ctx->last_ms[i];
  return GPDB_OK;
}

}  // extern "C"
