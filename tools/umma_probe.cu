// umma_probe.cu — hardware probe of the tcgen05 descriptor conventions lenet_tc.cu relies on (development aid).
//   T1: plain K-major no-swizzle GEMM 128 x N x K (bf16)        -> validates smem/instr descriptors, TMEM ld
//   T2: "Hankel" A operand: row stride 16 B, LBO = 16 B (rows and K-chunks overlap in memory) -> the im2col-free
//       convolution addressing
//   T3: fp16, N = 96
//   T4: per-instruction LBO between two distant K-chunks
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../gpd_b200/csrc/umma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct MmaOp { uint32_t a_off, a_lbo, b_off, b_lbo; };  // byte offsets into the A / B smem regions

// generic probe: copies a_bytes / b_bytes of raw operand memory to smem, issues nops MMAs, dumps D[128][N]
__global__ void __launch_bounds__(128) k_probe(const uint8_t *a_raw, int a_bytes, const uint8_t *b_raw, int b_bytes,
                                               const MmaOp *ops, int nops, int N, int fmt, uint32_t a_sbo, uint32_t b_sbo,
                                               float *D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  uint8_t *sa = smem, *sb = smem + ((a_bytes + 1023) / 1024) * 1024;
  for (int i = threadIdx.x; i < a_bytes / 16; i += 128) reinterpret_cast<uint4 *>(sa)[i] = reinterpret_cast<const uint4 *>(a_raw)[i];
  for (int i = threadIdx.x; i < b_bytes / 16; i += 128) reinterpret_cast<uint4 *>(sb)[i] = reinterpret_cast<const uint4 *>(b_raw)[i];
  if (threadIdx.x == 0) {
    umma::mbar_init(&mbar, 1);
    umma::fence_mbar_init();
  }
  if (threadIdx.x < 32) umma::tmem_alloc(&tmem_base, 128);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma::instr_desc(128, N, fmt);
    for (int i = 0; i < nops; i++) {
      uint64_t da = umma::smem_desc(umma::smem_u32(sa) + ops[i].a_off, ops[i].a_lbo, a_sbo);
      uint64_t db = umma::smem_desc(umma::smem_u32(sb) + ops[i].b_off, ops[i].b_lbo, b_sbo);
      umma::mma_f16(tb, da, db, idesc, i > 0);
    }
    umma::commit(&mbar);
  }
  umma::mbar_wait(&mbar, 0);
  umma::fence_after_sync();
  const int warp = threadIdx.x >> 5, row = threadIdx.x;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    umma::tmem_ld_wait();
    for (int j = 0; j < 16; j++) D[row * N + c0 + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tb, 128);
}

static float rnd() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

template <class T> static float tofloat(T x);
template <> float tofloat(__nv_bfloat16 x) { return __bfloat162float(x); }
template <> float tofloat(__half x) { return __half2float(x); }

// canonical K-major layout: element (r,k) at (r/8)*sbo + (r%8)*16 + (k/8)*lbo + (k%8)*2
template <class T>
static void put(std::vector<uint8_t> &buf, int r, int k, uint32_t lbo, uint32_t sbo, T v) {
  size_t off = (size_t)(r / 8) * sbo + (r % 8) * 16 + (size_t)(k / 8) * lbo + (k % 8) * 2;
  if (off + 2 > buf.size()) buf.resize(off + 2);
  *reinterpret_cast<T *>(&buf[off]) = v;
}

template <class T>
static int run(const char *name, std::vector<uint8_t> a, std::vector<uint8_t> b, std::vector<MmaOp> ops, int N, int fmt,
               uint32_t a_sbo, uint32_t b_sbo, const std::vector<double> &ref) {
  a.resize((a.size() + 1023) / 1024 * 1024);
  b.resize((b.size() + 1023) / 1024 * 1024);
  uint8_t *da, *db;
  MmaOp *dops;
  float *dD;
  CK(cudaMalloc(&da, a.size())); CK(cudaMalloc(&db, b.size())); CK(cudaMalloc(&dops, sizeof(MmaOp) * ops.size()));
  CK(cudaMalloc(&dD, sizeof(float) * 128 * N));
  CK(cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dops, ops.data(), sizeof(MmaOp) * ops.size(), cudaMemcpyHostToDevice));
  size_t smem = a.size() + b.size() + 1024;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_probe<<<1, 128, smem>>>(da, (int)a.size(), db, (int)b.size(), dops, (int)ops.size(), N, fmt, a_sbo, b_sbo, dD);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * N);
  CK(cudaMemcpy(D.data(), dD, sizeof(float) * D.size(), cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < D.size(); i++) { maxerr = fmax(maxerr, fabs(D[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
  printf("%-40s max|err| = %.3e (max|ref| = %.3e) %s\n", name, maxerr, maxref, maxerr <= 1e-4 * maxref ? "OK" : "MISMATCH");
  if (maxerr > 1e-4 * maxref) {
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 4; c++) printf("  D[%d][%d]=%.4f ref=%.4f", r, c, D[r * N + c], ref[r * N + c]); printf("\n"); }
  }
  cudaFree(da); cudaFree(db); cudaFree(dops); cudaFree(dD);
  return maxerr <= 1e-4 * maxref ? 0 : 1;
}

int main() {
  srand(1);
  int fails = 0;
  {  // T1: plain GEMM bf16, N = 32, K = 64
    const int N = 32, K = 64;
    const uint32_t a_lbo = 128 * 16, b_lbo = N * 16, sbo = 128;
    std::vector<uint8_t> a, b;
    std::vector<float> A(128 * K), B(N * K);
    for (int r = 0; r < 128; r++) for (int k = 0; k < K; k++) { __nv_bfloat16 v = __float2bfloat16(rnd()); A[r * K + k] = tofloat(v); put(a, r, k, a_lbo, sbo, v); }
    for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) { __nv_bfloat16 v = __float2bfloat16(rnd()); B[n * K + k] = tofloat(v); put(b, n, k, b_lbo, sbo, v); }
    std::vector<double> ref(128 * N, 0.0);
    for (int r = 0; r < 128; r++) for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) ref[r * N + n] += (double)A[r * K + k] * B[n * K + k];
    std::vector<MmaOp> ops;
    for (int i = 0; i < K / 16; i++) ops.push_back({(uint32_t)(2 * i) * a_lbo, a_lbo, (uint32_t)(2 * i) * b_lbo, b_lbo});
    fails += run<__nv_bfloat16>("T1 plain bf16 128x32x64", a, b, ops, N, umma::BF16, sbo, sbo, ref);
  }
  {  // T2: Hankel A (row stride 16 B, LBO 16 B), bf16, N = 32, K = 80
    const int N = 32, K = 80, NV = 128 + K / 8;  // 16-byte groups
    std::vector<uint8_t> a(NV * 16), b;
    std::vector<float> V(NV * 8), B(N * K);
    for (int i = 0; i < NV * 8; i++) { __nv_bfloat16 v = __float2bfloat16(rnd()); V[i] = tofloat(v); reinterpret_cast<__nv_bfloat16 *>(a.data())[i] = v; }
    const uint32_t b_lbo = N * 16, sbo = 128;
    for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) { __nv_bfloat16 v = __float2bfloat16(rnd()); B[n * K + k] = tofloat(v); put(b, n, k, b_lbo, sbo, v); }
    std::vector<double> ref(128 * N, 0.0);
    for (int r = 0; r < 128; r++) for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) ref[r * N + n] += (double)V[(r + k / 8) * 8 + k % 8] * B[n * K + k];
    std::vector<MmaOp> ops;
    for (int i = 0; i < K / 16; i++) ops.push_back({(uint32_t)(2 * i) * 16, 16, (uint32_t)(2 * i) * b_lbo, b_lbo});
    fails += run<__nv_bfloat16>("T2 Hankel A (LBO=16B, SBO=128B) bf16", a, b, ops, N, umma::BF16, sbo, sbo, ref);
  }
  {  // T3: fp16, N = 96, K = 32
    const int N = 96, K = 32;
    const uint32_t a_lbo = 128 * 16, b_lbo = N * 16, sbo = 128;
    std::vector<uint8_t> a, b;
    std::vector<float> A(128 * K), B(N * K);
    for (int r = 0; r < 128; r++) for (int k = 0; k < K; k++) { __half v = __float2half(rnd() * 8); A[r * K + k] = tofloat(v); put(a, r, k, a_lbo, sbo, v); }
    for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) { __half v = __float2half(rnd()); B[n * K + k] = tofloat(v); put(b, n, k, b_lbo, sbo, v); }
    std::vector<double> ref(128 * N, 0.0);
    for (int r = 0; r < 128; r++) for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) ref[r * N + n] += (double)A[r * K + k] * B[n * K + k];
    std::vector<MmaOp> ops;
    for (int i = 0; i < K / 16; i++) ops.push_back({(uint32_t)(2 * i) * a_lbo, a_lbo, (uint32_t)(2 * i) * b_lbo, b_lbo});
    fails += run<__half>("T3 plain fp16 128x96x32", a, b, ops, N, umma::F16, sbo, sbo, ref);
  }
  {  // T4: Hankel A with a per-instruction LBO: chunk0 at group g0, chunk1 at group g0 + 37 ; 2 MMAs
    const int N = 32, NV = 256;
    std::vector<uint8_t> a(NV * 16), b;
    std::vector<float> V(NV * 8), B(N * 32);
    for (int i = 0; i < NV * 8; i++) { __nv_bfloat16 v = __float2bfloat16(rnd()); V[i] = tofloat(v); reinterpret_cast<__nv_bfloat16 *>(a.data())[i] = v; }
    const uint32_t b_lbo = N * 16, sbo = 128;
    for (int n = 0; n < N; n++) for (int k = 0; k < 32; k++) { __nv_bfloat16 v = __float2bfloat16(rnd()); B[n * 32 + k] = tofloat(v); put(b, n, k, b_lbo, sbo, v); }
    const int g[4] = {3, 40, 41, 100};  // start groups of the 4 K-chunks
    std::vector<double> ref(128 * N, 0.0);
    for (int r = 0; r < 128; r++) for (int n = 0; n < N; n++) for (int k = 0; k < 32; k++) ref[r * N + n] += (double)V[(r + g[k / 8]) * 8 + k % 8] * B[n * 32 + k];
    std::vector<MmaOp> ops = {{(uint32_t)g[0] * 16, (uint32_t)(g[1] - g[0]) * 16, 0, b_lbo},
                              {(uint32_t)g[2] * 16, (uint32_t)(g[3] - g[2]) * 16, 2 * b_lbo, b_lbo}};
    fails += run<__nv_bfloat16>("T4 Hankel A, per-instruction LBO", a, b, ops, N, umma::BF16, sbo, sbo, ref);
  }
  printf("%s\n", fails ? "PROBE FAILED" : "PROBE PASSED");
  return fails;
}
