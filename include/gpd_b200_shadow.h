/*
 * gpd_b200_shadow.h — SPECIFICATION of the deterministic occlusion ("shadow") variant used by
 * the 15-channel grasp images (GPDB_SHADOW_DETERMINISTIC).
 *
 * Why a variant is needed: the reference's HandSet::calculateShadow (hand_set.cpp:118-233) is
 * irreproducible by construction — a static LCG seed shared and raced across OpenMP threads
 * (hand_set.cpp:14,263-266 called from image_generator.cpp:83-89), Gaussian jitter from
 * std::random_device (hand_set.cpp:191-199) and hash-set iteration order. Both the CPU oracle
 * and the CUDA kernels implement THIS definition, which keeps every other semantic of the
 * reference (same LCG constants and output bits, same voxel arithmetic, same set semantics):
 *
 *  1. draws: for hand set with sample index s, cloud point i, camera c the stream is re-seeded
 *     seed0 = gpdb_mix32(s*0x9E3779B1 + i*0x85EBCA77 + c*0xC2B2AE3D) and then advanced with the
 *     reference recurrence seed = 214013*seed + 2531011 (mod 2^32), value = (seed>>16)&0x7FFF
 *     (hand_set.cpp:263-266), u_j = value_j * (1.0/32767.0) for j = 0..num_shadow_points-1
 *     (hand_set.cpp:212-224).
 *  2. voxel = trunc((p + u*shadow_vec) * (1.0/0.003)) per component in float64, evaluated as
 *     round(round(p + round(u*sv)) * mult) with no FMA contraction (hand_set.cpp:219-223).
 *  3. the voxel SET per camera, and the intersection over the cameras that see >= 1 point of
 *     the neighbourhood, starting from camera 0's set even when it is empty
 *     (hand_set.cpp:140-176).
 *  4. voxel -> point: v*0.003 + g(v)*0.003*0.3 added to x, y and z alike
 *     (hand_set.cpp:196-199), with g(v) = GPDB_QTAB[gpdb_voxel_hash(v) & 1023], QTAB[k] = the
 *     standard-normal quantile at (k+0.5)/1024 (a counter-based Gaussian keyed on the voxel).
 *
 * Everything here is integer or IEEE float64 arithmetic with a fixed evaluation order, so the
 * oracle and the kernels agree bit for bit.
 */
#ifndef GPD_B200_SHADOW_H_
#define GPD_B200_SHADOW_H_

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GPDB_HD __host__ __device__ __forceinline__
#else
#define GPDB_HD static inline
#endif

#define GPDB_SHADOW_VOXEL 0.003 /* hand_set.cpp:121 */
#define GPDB_QTAB_SIZE 1024

GPDB_HD uint32_t gpdb_mix32(uint32_t h) { /* murmur3 finaliser */
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

GPDB_HD uint32_t gpdb_shadow_seed(uint32_t sample_index, uint32_t point_index, uint32_t cam) {
  return gpdb_mix32(sample_index * 0x9E3779B1u + point_index * 0x85EBCA77u + cam * 0xC2B2AE3Du);
}

/* one step of HandSet::fastrand (hand_set.cpp:263-266); returns the 15-bit value */
GPDB_HD uint32_t gpdb_fastrand(uint32_t *seed) {
  *seed = 214013u * (*seed) + 2531011u;
  return (*seed >> 16) & 0x7FFFu;
}

GPDB_HD uint32_t gpdb_voxel_hash(int32_t vx, int32_t vy, int32_t vz) {
  return gpdb_mix32((uint32_t)vx * 73856093u ^ (uint32_t)vy * 19349663u ^ (uint32_t)vz * 83492791u);
}

/* Standard-normal quantile (P. J. Acklam's rational approximation, |rel err| < 1.2e-9). Host only:
 * the table is built once on the host and uploaded, so the device never evaluates log(). */
static inline double gpdb_norm_quantile(double p) {
  static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                              1.383577518672690e+02,  -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                              6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                              -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
  static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                              3.754408661907416e+00};
  const double plow = 0.02425, phigh = 1.0 - 0.02425;
  if (p < plow) {
    double q = sqrt(-2.0 * log(p));
    return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
  }
  if (p > phigh) {
    double q = sqrt(-2.0 * log(1.0 - p));
    return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
  }
  double q = p - 0.5, r = q * q;
  return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
         (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
}

static inline void gpdb_build_qtab(double *tab /* [GPDB_QTAB_SIZE] */) {
  for (int k = 0; k < GPDB_QTAB_SIZE; k++) tab[k] = gpdb_norm_quantile((k + 0.5) / GPDB_QTAB_SIZE);
}

#endif /* GPD_B200_SHADOW_H_ */
