// common.cuh — shared declarations of libgpd_b200.so (sm_100a only).
//
// HBM layout of one context (see DESIGN.md "data layout"):
//   cloud   pts4   float4[N]  points SORTED BY GRID CELL: x,y,z + original index bits  (4.8 MB @300k)
//           xyz    float[3N]  points by original index (nb0 lookups)
//           nrm    double[3N] normals by original index, 3xN column-major as the ABI gives them
//           cam    uint8[N]   bit k set when camera k sees the point
//           cell_start int[ncell+1]   uniform grid, x fastest: a run of cells along x is ONE
//                                     contiguous segment of pts4
//   per chunk of samples: frames double[9n], dense pose records gpdb_pose[n*P], flags uint8[n*P],
//           compact candidate list gpdb_pose[nc], images uint8[nc*S*S*C] (HWC, the cv::Mat layout),
//           LeNet activations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gpd_b200.h"
#include "../../include/gpd_b200_shadow.h"

#define GPDB_MAX_ORIENT 32
#define GPDB_MAX_SLOTS 32   // 2 * num_finger_placements
#define GPDB_MAX_DEEPEN 64  // deepen steps
#define GPDB_MAX_NSP 128    // shadow draws per point

// Everything the kernels need, resident in global memory (uniform, L1/L2-cached loads).
struct DevParams {
  // hand geometry / search (cfg/hand_geometry.cfg, grasp_detector.cpp:67-86)
  double finger_width, hand_outer_diameter, hand_depth, hand_height, init_bite;
  int P, n_axes, n_orient, nfp;
  int axes[GPDB_MAX_HAND_AXES];
  int deepen, slots_disjoint, all_axes_z;
  double inv_slot_step;        // 1 / spacing of the finger slots (index estimate in slot_mask)
  double fs[GPDB_MAX_SLOTS];   // FingerHand::finger_spacing_ (finger_hand.cpp:12-19)
  double fsw[GPDB_MAX_SLOTS];  // fs + finger_width
  int J;                       // deepen steps d_j = init_bite + j*0.005 accumulated in double (finger_hand.cpp:120-121)
  double topj[GPDB_MAX_DEEPEN], botj[GPDB_MAX_DEEPEN];
  double cosf;                 // cos(friction_coeff * pi / 180) (antipodal.cpp:28)
  int min_viable;
  // filters (grasp_detector.cpp:334-456)
  double min_ap, max_ap, ws[6];
  int filt_dir;
  double dir[3], thresh;
  // rotations (hand_set.cpp:52-53,68-69): rotb = AngleAxis(pi, UnitY); rot[a*n_orient+i]
  double rotb[9];
  double rot[GPDB_MAX_HAND_AXES * GPDB_MAX_ORIENT][9];
  // image geometry (cfg/image_geometry_*.cfg)
  double vol_w, vol_d, vol_h;
  int S, C;
  // shadow (hand_set.cpp:118-233, include/gpd_b200_shadow.h)
  double shadow_length, vox_mult;
  int nsp;                     // num_shadow_points
  int bm_dim;                  // bitmap edge (voxels)
  unsigned lcgA[GPDB_MAX_NSP], lcgC[GPDB_MAX_NSP];  // LCG skip-ahead: seed after t+1 steps = lcgA[t]*seed0 + lcgC[t]
  // radii: float32 predicates (dist < r2) and search extents
  float r2_lrf, r2_hs, r2_img;
  float rf_lrf, rf_hs, rf_img;
  // grid
  float lo[3], inv_cell;
  int dim[3];
  // cloud
  int N, K;
  int all_seen;                // every point is seen by every camera (cam mask complete): the per-point masks need not be read
  double vp[GPDB_MAX_CAMERAS][3];
  // LeNet
  int relu_after_conv;
};

struct DevCloud {
  const float4 *pts4;
  const float *xyz;
  const double *nrm;
  const uint8_t *cam;
  const int *cell_start;
  // Cloud::setSamples (gpdb_set_samples): float64 positions addressed by sample indices >= n_points
  const double *samples;
  int n_points;
};

// device error counters: [0] LRF capacity, [1] hand-search capacity (final tier), [2] image box list,
// [3] hand-search tier-1 overflow count (informational), [4] normal-estimation capacity (final tier)
#define GPDB_NERR 8

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      gpdb_set_error(ctx, GPDB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,         \
                     cudaGetErrorString(e__));                                                \
      return GPDB_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

struct LenetWeights {  // device pointers, layouts documented in lenet_simt.cu
  float *c1w, *c1b, *c2w, *c2b, *i1w, *i1b, *i2w, *i2b;
  int C;
  bool set;
};

struct C1Affine {  // conv1 epilogue: per-filter weight scale and bias (kernel parameter = constant bank)
  float scale[20], bias[20];
};
struct LenetTc {  // tensor-core (tcgen05) weight blobs, lenet_tc.cu
  void *b1, *b2, *b3;
  C1Affine c1_aff;
  int npl, nch1;
  float w2_scale, a2_scale, w3_scale, x3_scale;
  bool ready;
};

struct StageTimes;  // api.cu
struct PipeState;   // api.cu: copy stream, events and the pinned result arenas of the chunk pipeline
struct CommState;   // comm.cu: NCCL communicator (multi-GPU sharding)

struct gpdb_ctx {
  gpdb_params prm;
  DevParams hp;       // host copy
  DevParams *dp;      // device copy
  int device;
  cudaStream_t stream;
  bool own_stream;
  bool overlap_hands;  // hand search of the chunks ahead on its own stream (gpdb_set_overlap)
  StageTimes *st;
  PipeState *pipe;
  CommState *comm;
  int sm_count;
  char err[512];
  // cloud
  DevCloud cloud;
  float4 *d_pts4;
  float *d_xyz;
  double *d_nrm;
  uint8_t *d_cam;
  int *d_cell_start;
  size_t cloud_cap;       // capacity (points) of pts4 / xyz / nrm / cam / src: grown, never shrunk
  size_t cell_cap;        // capacity (ints) of cell_start
  int N, K;
  bool cloud_set;
  double *d_qtab;
  // weights
  LenetWeights w;
  LenetTc tc;
  // scratch (grown on demand)
  void *scratch[24];
  size_t scratch_sz[24];
  int *d_err;
  unsigned long long *d_prof;  // optional phase counters (gpdb_debug_phase_cycles), nullptr = off
  int64_t launches;
  double last_ms[8];
  double pre_ms[6];   // gpdb_preprocess stage timings
  gpdb_pose *d_sel;   // gpdb_detect_select: all classified candidates of a call (grown on demand)
  size_t sel_cap;
  double *d_samples;  // gpdb_set_samples positions (3 x n_samples), or nullptr
  int n_samples;
  int *d_src;         // raw index of each processed point (valid after gpdb_preprocess: has_src)
  bool has_src;
  cudaEvent_t ev[8];
};

void gpdb_set_error(gpdb_ctx *ctx, int code, const char *fmt, ...);
// device-side stage timers (CUDA events on the context stream). stages: 0 frames, 1 hand search +
// compaction, 2 images, 3 LeNet, 4 whole call, 5 conv1, 6 conv2, 7 ip1+ip2
cudaEvent_t gpdb_st_begin(gpdb_ctx *ctx);
void gpdb_st_end(gpdb_ctx *ctx, int stage, cudaEvent_t begin);
void *gpdb_scratch(gpdb_ctx *ctx, int slot, size_t bytes);  // returns nullptr on failure (error set)

// api.cu
int gpdb_pipe_create(gpdb_ctx *ctx);
void gpdb_pipe_destroy(gpdb_ctx *ctx);
int gpdb_check_state(gpdb_ctx *ctx, bool need_cloud, bool need_weights);
void *gpdb_result_extra(gpdb_result *r, size_t bytes);  // pinned host memory owned by the result (freed with it)
// the chunked device pipeline (see api.cu); slot_base is added to every sample_slot (rank offset of a sharded call)
int gpdb_run_pipeline(gpdb_ctx *ctx, const int32_t *sample_idx, int32_t n, gpdb_result *out, bool with_images_and_scores,
                      bool resident, uint8_t *flags_ext, float *scores_ext, int select_k, int slot_base);
// installs the cloud whose device arrays d_xyz / d_nrm / d_cam already hold N points (grid bounds by device reduction)
int gpdb_install_device_cloud(gpdb_ctx *ctx, int N, int K, const double *view_points, int all_seen);

// geometry.cu
// builds the neighbour grid over ctx->d_xyz (N points) whose per-axis bounds are lo / hi
int geo_build_grid(gpdb_ctx *ctx, const float lo[3], const float hi[3], int N);
int geo_frames(gpdb_ctx *ctx, const int *d_sidx, int n, double *d_frames, uint8_t *d_valid);
int geo_hands(gpdb_ctx *ctx, const int *d_sidx, int n, int slot0, const double *d_frames, const uint8_t *d_valid,
              gpdb_pose *d_poses, uint8_t *d_flags);
// compacts poses with VALID|FILTERED into d_cand (in (sample,pose) order); *d_count receives the count
int geo_compact(gpdb_ctx *ctx, const gpdb_pose *d_poses, const uint8_t *d_flags, int n_poses, gpdb_pose *d_cand,
                int *d_count);
// grasp images in the P16 layout: S*S pixels of 16 bytes (channels 0..C-1, zero padded) per image = conv1's operand
int geo_images(gpdb_ctx *ctx, const gpdb_pose *d_cand, int nc, uint8_t *d_p16);
int geo_p16_to_hwc(gpdb_ctx *ctx, const uint8_t *d_p16, int n, uint8_t *d_hwc);  // -> cv::Mat layout (C bytes per pixel)
int geo_hwc_to_p16(gpdb_ctx *ctx, const uint8_t *d_hwc, int n, uint8_t *d_p16);
int geo_scatter_scores(gpdb_ctx *ctx, const gpdb_pose *d_cand, const float *d_scores, int nc, int slot0, int P,
                       float *d_pose_scores, gpdb_pose *d_cand_out);

// HandSearch::reevaluateHypotheses: labels + half / full flags of the given hands against the installed cloud
int geo_reeval(gpdb_ctx *ctx, gpdb_pose *d_hands, int n, int *d_labels);
// Clustering::findClusters (remove_inliers = false): dense per-hand cluster records + keep flags (3 = cluster), for geo_compact
int geo_clusters(gpdb_ctx *ctx, const gpdb_pose *d_hands, int n, int min_inliers, gpdb_pose *d_dense, uint8_t *d_keep);
// the k highest-scoring of the n candidate records (scores filled), descending, stable -> d_out[k]
int geo_select(gpdb_ctx *ctx, const gpdb_pose *d_cand, int n, int k, gpdb_pose *d_out);

// preprocess.cu (cloud preprocessing)
// (re)allocates the context's cloud arrays for at least n points (api.cu)
int gpdb_cloud_reserve(gpdb_ctx *ctx, size_t n);
int pre_bounds(gpdb_ctx *ctx, const float *d_xyz, int n, int *d_bounds, float lo[3], float hi[3]);
// filters + voxelises the raw arrays into the context's cloud arrays (reserved inside); *n_out = processed points
int pre_filter_voxelize(gpdb_ctx *ctx, const float *d_xyz_raw, const uint8_t *d_cam_raw, const double *d_nrm_raw, int M,
                        const gpdb_preprocess_params &pp, int *n_out, cudaEvent_t ev_filter_done);
int pre_normals(gpdb_ctx *ctx, double radius);
int pre_cam_expand(gpdb_ctx *ctx, int *d_out);

// lenet_simt.cu
int lenet_upload(gpdb_ctx *ctx, const float *const w[8]);
int lenet_forward(gpdb_ctx *ctx, const uint8_t *d_images, int n, float *d_scores, float *d_logits);

// lenet_tc.cu (tcgen05 conv1 / conv2)
int lenet_tc_upload(gpdb_ctx *ctx, const float *const w[8]);
struct __half;
int lenet_tc_forward(gpdb_ctx *ctx, const uint8_t *d_images, int n, float *p1, __half *xc, float *h3);
size_t lenet_tc_xc_bytes(int n);
