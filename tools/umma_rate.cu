// umma_rate.cu — issue-rate probe of tcgen05.mma on B200 (development aid): cycles per instruction for the operand shapes the
// implicit-GEMM convolutions of lenet_tc.cu use (M = 128, K-major no-swizzle "Hankel" A with 16-byte row stride), as a function
// of kind (i8 K = 32 / f16 K = 16), N, the alignment of the A start address, the K-chunk distance (LBO) and whether consecutive
// instructions accumulate into the same TMEM tile.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_rate tools/umma_rate.cu ; run on one GPU.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../gpd_b200/csrc/umma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {  // D = s32, A = u8, B = s8
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Cfg {
  int kind;      // 0 f16, 1 i8
  int N;
  int a_off;     // byte offset of the A start inside the plane (multiple of 16)
  int a_lbo;     // bytes between the two K-chunks of A
  int a_sbo;     // bytes between 8-row groups of A (128 = Hankel / dense 16-byte rows)
  int ndbuf;     // 1: every instruction accumulates into the same tile; 2: alternate between two tiles
  int nctas;     // CTAs per launch (one per SM): 1 or 148
  int reps;
  int group;     // > 0: tcgen05.commit after every `group` instructions (as the convolution kernels do per tile)
  int fresh;     // 1: the first instruction of a group overwrites the accumulator (accumulate = false)
};

__global__ void __launch_bounds__(128, 1) k_rate(Cfg c, unsigned long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  // A plane: 64 KB, B: 256 rows x 32 B = 8 KB per chunk pair, zero data
  for (int i = threadIdx.x; i < (96 * 1024) / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    umma::mbar_init(&mbar, c.group > 0 ? (uint32_t)(c.reps / c.group) : 1u);
    umma::fence_mbar_init();
  }
  if (threadIdx.x < 32) umma::tmem_alloc(&tmem_base, 512);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base;
  const uint32_t sA = umma::smem_u32(smem) + c.a_off, sB = umma::smem_u32(smem) + 64 * 1024;
  const uint64_t da = umma::desc_from(sA, c.a_lbo, c.a_sbo);
  const uint64_t db = umma::desc_from(sB, c.N * 16, 128);  // B: [chunk][N rows][16 B]
  const uint32_t idesc = c.kind ? idesc_i8(128, c.N) : umma::instr_desc(128, c.N, umma::F16);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    if (c.group > 0) {
      for (int i = 0; i < c.reps; i += c.group) {
        for (int j = 0; j < c.group; j++) {  // descriptors vary as in the convolutions: A by filter tap, B by K-chunk pair
          const uint64_t daj = da + (uint64_t)((j % 5) + (j / 5 % 5) * 60), dbj = db + (uint64_t)((j % 8) * (c.N * 32 >> 4));
          const uint32_t dcol = tb + ((c.ndbuf == 2 && ((i / c.group) & 1)) ? 256u : 0u);
          if (c.kind) umma::mma_i8(dcol, daj, dbj, idesc, !(c.fresh && j == 0));
          else umma::mma_f16(dcol, daj, dbj, idesc, !(c.fresh && j == 0));
        }
        if (i + c.group < c.reps) umma::commit(&mbar);
      }
    } else if (c.kind) {
      for (int i = 0; i < c.reps; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) umma::mma_i8(tb + ((c.ndbuf == 2 && (j & 1)) ? 256u : 0u), da, db, idesc, true);
      }
    } else {
      for (int i = 0; i < c.reps; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) umma::mma_f16(tb + ((c.ndbuf == 2 && (j & 1)) ? 256u : 0u), da, db, idesc, true);
      }
    }
    umma::commit(&mbar);
  }
  umma::mbar_wait(&mbar, 0);
  umma::fence_after_sync();
  if (threadIdx.x == 0) {
    t1 = clock64();
    out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tb, 512);
}


// ---- second probe: the MMA issue loop of k_conv1_i8 as it is (13 unrolled kind::i8 instructions per tile, descriptors =
// uniform base + compile-time constants, one elected lane, commit per tile), without any consumer. flags: 1 = commit per
// tile, 2 = four more warps read the accumulators with tcgen05.ld all the time (the epilogue's TMEM traffic), 4 = SBO 960
// (16 x 8 tiles) instead of 128, 8 = the other warps execute shared-memory loads all the time (LSU traffic), 16 = pseudo-random
// operand bytes instead of zeros
__host__ __device__ constexpr uint32_t t_off(int c) { return (uint32_t)((((c >= 25 ? 24 : c) / 5) * 60 + (c >= 25 ? 24 : c) % 5) * 16); }
__global__ void __launch_bounds__(192, 1) k_tile(int ntiles, int flags, unsigned long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4], done;
  __shared__ uint32_t tmem_base;
  __shared__ int stop;
  for (int i = threadIdx.x; i < (96 * 1024) / 16; i += 192) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (flags & 16) {  // pseudo-random operand bytes (zero operands flatter the tensor pipe's power draw)
      unsigned x = (unsigned)i * 2654435761u + 12345u;
      x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
      v = make_uint4(x, x * 3266489917u, x * 668265263u + 7u, x ^ 0x9e3779b9u);
    }
    reinterpret_cast<uint4 *>(smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < 4; b++) umma::mbar_init(&bars[b], 1);
    umma::mbar_init(&done, 1);
    umma::fence_mbar_init();
    stop = 0;
  }
  if (threadIdx.x < 32) umma::tmem_alloc(&tmem_base, 256);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t sbo = (flags & 4) ? 960u : 128u;
  if (warp == 4) {
    const uint32_t sPl_u = umma::smem_u32(smem), sB_u = umma::smem_u32(smem) + 64 * 1024;
    constexpr uint32_t idesc = idesc_i8(128, 64);
    long long t0 = clock64();
    for (int t = 0; t < ntiles; t++) {
      const int tt = t % 28;
      const uint32_t arow = sPl_u + (uint32_t)((flags & 4) ? ((tt / 7) * 8 * 60 + (tt % 7) * 8) * 16 : tt * 120 * 16) % 8192u;
      const uint32_t dcol = tb + (uint32_t)(t & 3) * 64;
      if (umma::elect_one()) {
#pragma unroll
        for (int i = 0; i < 13; i++) {
          const uint32_t a0 = t_off(2 * i), a1 = t_off(2 * i + 1);
          const uint32_t lbo = (2 * i + 1 >= 25) ? 16u : (a1 - a0);
          umma::mma_i8(dcol, umma::desc_from(arow + a0, lbo, sbo), umma::desc_from(sB_u + (uint32_t)(2 * i) * 1024, 1024, 128), idesc, i > 0);
        }
        if (flags & 1) umma::commit(&bars[t & 3]);
      }
      __syncwarp();
    }
    if (umma::elect_one()) umma::commit(&done);
    __syncwarp();
    umma::mbar_wait(&done, 0);
    umma::fence_after_sync();
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) {
      out[blockIdx.x] = (unsigned long long)(t1 - t0);
      *(volatile int *)&stop = 1;
    }
  } else if (warp < 4 && (flags & 2)) {
    float acc = 0.f;
    while (!*(volatile int *)&stop) {
      float d[64];
#pragma unroll
      for (int cb = 0; cb < 4; cb++) umma::tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + cb * 16, d + cb * 16);
      umma::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 64; j++) acc += d[j];
    }
    if (acc == 12345.678f) out[200] = 1;
  } else if (warp < 4 && (flags & 8)) {
    float acc = 0.f;
    const float *sp = reinterpret_cast<const float *>(smem) + threadIdx.x * 4;
    while (!*(volatile int *)&stop) {
#pragma unroll
      for (int j = 0; j < 16; j++) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(umma::smem_u32(sp + j * 512)));
        acc += v.x + v.w;
      }
    }
    if (acc == 12345.678f) out[200] = 1;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tb, 256);
}

int main() {
  unsigned long long *d_out;
  CK(cudaMalloc(&d_out, 148 * sizeof(unsigned long long)));
  CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  std::vector<Cfg> cfgs;
  const int R = 4096;
  for (int nct : {148}) {
    cfgs.push_back({1, 64, 16, 16, 128, 1, nct, R, 0, 0});
    cfgs.push_back({0, 128, 16, 16, 128, 1, nct, R, 0, 0});
  }
  printf("%-5s %4s %6s %6s %5s %6s %6s %6s %6s | %10s %10s\n", "kind", "N", "a_off", "a_lbo", "sbo", "ndbuf", "ctas", "group", "fresh", "cyc/mma", "MAC/clk/SM");
  for (const Cfg &c : cfgs) {
    unsigned long long h[148];
    for (int w = 0; w < 2; w++) {  // second run is the measurement
      k_rate<<<c.nctas, 128, 96 * 1024>>>(c, d_out);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d_out, c.nctas * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double mx = 0;
    for (int i = 0; i < c.nctas; i++) mx = h[i] > mx ? (double)h[i] : mx;
    const double cyc = mx / c.reps, macs = 128.0 * c.N * (c.kind ? 32 : 16);
    printf("%-5s %4d %6d %6d %5d %6d %6d %6d %6d | %10.1f %10.0f\n", c.kind ? "i8" : "f16", c.N, c.a_off, c.a_lbo, c.a_sbo, c.ndbuf, c.nctas,
           c.group, c.fresh, cyc, macs / cyc);
  }
  CK(cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  printf("\nconv1 issue loop (13 x i8 128x64x32 per tile): cycles per tile (13 x 48 = 624 when the operand stream is the limit)\n");
  for (int flags : {1, 17, 19, 27, 21}) {
    unsigned long long h[148];
    for (int w = 0; w < 2; w++) {
      k_tile<<<148, 192, 96 * 1024>>>(2800, flags, d_out);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d_out, 148 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double mx = 0;
    for (int i = 0; i < 148; i++) mx = h[i] > mx ? (double)h[i] : mx;
    printf("flags %2d (commit %d, tmem_ld warps %d, sbo %d, lds warps %d, random operands %d): %8.1f cycles / tile\n", flags, flags & 1,
           (flags >> 1) & 1, (flags & 4) ? 960 : 128, (flags >> 3) & 1, (flags >> 4) & 1, mx / 2800);
  }
  return 0;
}
