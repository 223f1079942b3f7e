// umma.cuh — thin inline-PTX layer for the 5th-generation tensor cores of sm_100a (tcgen05 + TMEM),
// used by lenet_tc.cu. Encodings follow the PTX ISA "tcgen05" matrix/instruction descriptors:
//   shared-memory matrix descriptor (64 bit): start address >> 4 [0,14), leading-dimension byte offset >> 4
//   [16,30), stride-dimension byte offset >> 4 [32,46), version = 1 [46,48), swizzle mode [61,64) (0 = none);
//   instruction descriptor (32 bit): D format [4,6) (1 = f32), A format [7,10), B format [10,13)
//   (0 = f16, 1 = bf16, 2 = tf32), A/B major [15],[16] (0 = K-major), N >> 3 [17,23), M >> 4 [24,29).
// K-major operands without swizzle are stored as 8-row x 16-byte "core matrices": element (r, k) of a
// 16-bit operand lives at  start + (r % 8) * 16 + (r / 8) * SBO + (k / 8) * LBO + (k % 8) * 2  bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle. lbo / sbo in bytes (multiples of 16).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

enum Fmt { F16 = 0, BF16 = 1, TF32 = 2 };
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int fmt_ab) {
  return (1u << 4) | ((uint32_t)fmt_ab << 7) | ((uint32_t)fmt_ab << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues on behalf of the CTA.
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// integer variant: A uint8 / B int8 (formats in the instruction descriptor), int32 accumulators, K = 32 per instruction
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// elect one lane of a fully converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// descriptor from its two 32-bit halves: lo = (addr >> 4) | (lbo >> 4) << 16 ; hi = (sbo >> 4) | 1 << 14 (version)
__device__ __forceinline__ uint64_t desc_from(uint32_t addr_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint32_t lo = (addr_bytes >> 4) | ((lbo_bytes >> 4) << 16);
  uint32_t hi = (sbo_bytes >> 4) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}

// all previously issued MMAs of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void commit(uint64_t *mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(mbar)),
      "r"(parity)
      : "memory");
}

// consumer-side wait of the epilogue warps for a tile. GPDB_WAIT_MODE: 0 = tight try_wait loop (mbar_wait, the default),
// 1 = __nanosleep(100) back-off, 2 = try_wait with a 100 us suspend-time hint. Measured on B200 (conv1 / conv2 per 50.6 k
// images): 4.73 / 5.11 ms, 4.80 / 5.19 ms, 4.99 / 5.20 ms — the polling does not disturb the operand stream, the wake-up
// latency is on the critical path, so the tight loop stays.
#ifndef GPDB_WAIT_MODE
#define GPDB_WAIT_MODE 0
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *mbar, uint32_t parity) {
#if GPDB_WAIT_MODE == 1
  uint32_t ok = 0;
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(mbar)), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(100);
  }
#elif GPDB_WAIT_MODE == 2
  uint32_t ok = 0;
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(mbar)), "r"(parity), "r"(100000u)
        : "memory");
    if (ok) break;
  }
#else
  mbar_wait(mbar, parity);
#endif
}

__device__ __forceinline__ void mbar_arrive(uint64_t *mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar)) : "memory");
}
// named barrier among a subset of the CTA's warps (id 1..15)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 1-D bulk async copy global -> shared (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(uint64_t *mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMEM allocation by ONE warp; ncols power of two >= 32; the base address is written to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32-bit, 16 consecutive columns: thread t of the warp reads lane (lane_base + t), v[j] = column j
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace umma
