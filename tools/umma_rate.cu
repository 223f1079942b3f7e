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
};

__global__ void __launch_bounds__(128, 1) k_rate(Cfg c, unsigned long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  // A plane: 64 KB, B: 256 rows x 32 B = 8 KB per chunk pair, zero data
  for (int i = threadIdx.x; i < (96 * 1024) / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    umma::mbar_init(&mbar, 1);
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
    if (c.kind) {
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

int main() {
  unsigned long long *d_out;
  CK(cudaMalloc(&d_out, 148 * sizeof(unsigned long long)));
  CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  std::vector<Cfg> cfgs;
  const int R = 4096;
  for (int nct : {1, 148}) {
    // conv1 shapes: i8, N = 64, Hankel A (sbo 128), chunk distance 16 B, aligned / unaligned start
    cfgs.push_back({1, 64, 0, 16, 128, 1, nct, R});
    cfgs.push_back({1, 64, 16, 16, 128, 1, nct, R});
    cfgs.push_back({1, 64, 64, 16, 128, 1, nct, R});
    cfgs.push_back({1, 64, 0, 896, 128, 1, nct, R});   // distant K-chunks
    cfgs.push_back({1, 64, 0, 2048, 256, 1, nct, R});  // non-overlapping 8-row groups (dense core matrices, canonical layout)
    cfgs.push_back({1, 64, 0, 16, 128, 2, nct, R});    // alternate accumulators
    cfgs.push_back({1, 128, 0, 16, 128, 1, nct, R});
    cfgs.push_back({1, 128, 16, 16, 128, 1, nct, R});
    cfgs.push_back({1, 256, 0, 16, 128, 1, nct, R});
    cfgs.push_back({1, 32, 0, 16, 128, 1, nct, R});
    // conv2 shapes: f16, N = 128 / 64 / 112
    cfgs.push_back({0, 128, 0, 16, 128, 1, nct, R});
    cfgs.push_back({0, 128, 16, 16, 128, 1, nct, R});
    cfgs.push_back({0, 128, 0, 2048, 256, 1, nct, R});
    cfgs.push_back({0, 64, 0, 16, 128, 1, nct, R});
    cfgs.push_back({0, 64, 16, 16, 128, 1, nct, R});
    cfgs.push_back({0, 112, 16, 16, 128, 1, nct, R});
    cfgs.push_back({0, 256, 16, 16, 128, 1, nct, R});
    cfgs.push_back({0, 128, 16, 16, 128, 2, nct, R});
  }
  printf("%-5s %4s %6s %6s %5s %6s %6s | %10s %10s\n", "kind", "N", "a_off", "a_lbo", "sbo", "ndbuf", "ctas", "cyc/mma", "MAC/clk/SM");
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
    printf("%-5s %4d %6d %6d %5d %6d %6d | %10.1f %10.0f\n", c.kind ? "i8" : "f16", c.N, c.a_off, c.a_lbo, c.a_sbo, c.ndbuf, c.nctas, cyc,
           macs / cyc);
  }
  return 0;
}
