// tcgen05 / TMEM / mbarrier helpers shared by the tensor-core gridder and degridder (sm_100a).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace idgb200 {

// K-major, no-swizzle core-matrix operand layout: a 16-byte chunk = 8 fp16 K-elements of one row;
// 8 rows x 16 B = one 128-byte core matrix; chunks of the same K-slice are row-contiguous.
constexpr int A_CHUNK_BYTES = 128 * 16;          // one 16-byte K-chunk (4 vis) of 128 rows
constexpr int B_CHUNK_BYTES = 16 * 16;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Ablation builds (tools/ablate.py, never the shipped library): bit 0 = do not issue the stage's MMAs,
// bit 1 = the phasor operand's stores are predicated off at run time (ablate_never() is false, but only
// the hardware knows: the instructions still issue, nothing reaches the shared-memory pipe).
#ifndef IDGB200_ABLATE
#define IDGB200_ABLATE 0
#endif
__device__ __forceinline__ bool ablate_never() {
  unsigned n;
  asm volatile("mov.u32 %0, %%nctaid.z;" : "=r"(n));
  return n == 12345u;
}

constexpr unsigned MBAR_SUSPEND_HINT = 0x989680u;   // upper bound of one try_wait's sleep (what CUTLASS passes)

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned done = 0;
  for (int spin = 0; !done; spin++) {
    // suspend-time hint: a waiting warp sleeps in the barrier unit instead of coming back to spin
    // through the issue port it shares with the producers (ncu: 3-6 retries per stage without it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(MBAR_SUSPEND_HINT) : "memory");
    if (spin > (1 << 22)) __trap();   // a lost arrival must fail loudly, not hang the GPU
  }
}
// The same on a 32-bit shared address computed once outside a hot loop: one try_wait and one branch
// when the phase is already complete, the bounded spin (trap on a lost arrival) only otherwise.
__device__ __forceinline__ void mbar_wait_u(unsigned bar, unsigned parity) {
  unsigned done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  for (int spin = 0; !done; spin++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_HINT) : "memory");
    if (spin > (1 << 22)) __trap();
  }
}
// The same with a wall-clock bound instead of a spin count (a suspended try_wait may last up to the hint, so a spin
// count bounds nothing): used by the persistent pipelines, whose roles wait on each other for whole subgrids.
__device__ __forceinline__ void mbar_wait_t(unsigned bar, unsigned parity) {
  unsigned done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  if (done) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_HINT) : "memory");
    if (!done) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) __trap();   // 4 s: a lost arrival must fail loudly, not hang the GPU
    }
  }
}
// Latency-critical waits of the persistent pipelines: nothing but the try_wait and its branch in the loop (the wall-clock
// check above is six more instructions per poll, and a pipeline polls a few times per tile).  2^26 polls of >= ~50 clocks
// bound a lost arrival to seconds.
__device__ __forceinline__ void mbar_wait_spin(unsigned bar, unsigned parity) {
  unsigned done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  for (unsigned spin = 0; !done; spin++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
// For waits that are known to be long (a role that runs ahead of the others): sleep between polls, so that the waiting
// warp leaves the issue slots to the warps it is waiting for (ncu on the first pipelined degridder: a third of all
// executed instructions were try_wait loops of warps with nothing to do).  Bounded like the others: 2^24 sleeps >= 4 s.
__device__ __forceinline__ void mbar_wait_backoff(unsigned bar, unsigned parity, unsigned ns) {
  unsigned done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  for (unsigned spin = 0; !done; spin++) {
    __nanosleep(ns);
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void umma_commit_u(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start[0,14) LBO[16,30) SBO[32,46),
// all >> 4; version[46,48) = 1; layout[61,64) = 0 (no swizzle).  K-major canonical layout:
// 8 rows x 16 B core matrices; SBO = distance of 8-row groups, LBO = distance of 16-byte K chunks.
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr, unsigned lbo, unsigned sbo) {
  return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo >> 4) & 0x3FFF) << 16) |
         ((unsigned long long)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(unsigned tmem_d, unsigned long long da, unsigned long long db,
                                         unsigned idesc, unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(unsigned long long *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// sin / cos of 2 pi r, |r| <= 0.5, on the FP32 pipe: the tensor-core kernel leaves that pipe idle
// while the XU (2 MUFU per item) is the roof, so a compile-time subset of the channels of every
// 16-channel block gets its phasor from these polynomials instead (near-minimax fits, max error
// 6.4e-6 / 4.1e-5 in fp32 - below the fp16 rounding of the operand), after an exact
// round-to-nearest range reduction in revolutions.  Cost: 14 FP32-pipe instructions per item
// against FFMA + FMUL + 2 MUFU.
__device__ __forceinline__ float2 phasor_poly(float t /* revolutions */) {
  const float k = __fadd_rn(__fadd_rn(t, 12582912.0f), -12582912.0f);   // rint(t), |t| < 2^22
  const float r = __fsub_rn(t, k);
  const float x = __fmul_rn(r, r);
  // both Horner chains in one packed register pair: 4 FFMA2 instead of 8 FFMA (issue slots)
  const float2 xx = make_float2(x, x);
  float2 cs = __ffma2_rn(make_float2(45.64655685424805f, 32.7813835144043f), xx,
                         make_float2(-82.40354919433594f, -74.47799682617188f));
  cs = __ffma2_rn(cs, xx, make_float2(64.67343139648438f, 81.36681365966797f));
  cs = __ffma2_rn(cs, xx, make_float2(-19.731040954589844f, -41.331214904785156f));
  cs = __ffma2_rn(cs, xx, make_float2(0.9999597668647766f, 6.283055782318115f));
  return make_float2(cs.x, __fmul_rn(cs.y, r));
}

// (cos, sin) -> packed half2; with SPLIT also the rounding residual as a second half2, which goes
// to a second A buffer and a second MMA against the same B slot: the operand then keeps ~22 bits
template <bool SPLIT>
__device__ __forceinline__ void pack_phasor(const float2 ph, unsigned &hi, unsigned &lo) {
  const __half2 hh = __floats2half2_rn(ph.x, ph.y);
  hi = *reinterpret_cast<const unsigned *>(&hh);
  if (SPLIT) {
    // residual = ph - float(hh), exact (Sterbenz).  sm_100a's mixed-precision FMA reads the fp16 halves
    // of the packed word directly (SASS: FHFMA Rd, Rh.H0|.H1, -1.0h, Rph): one instruction per
    // component instead of two conversions and a packed FADD2 (= 4 dispatch cycles)
    const unsigned short h0 = (unsigned short)(hi & 0xffffu), h1 = (unsigned short)(hi >> 16);
    const unsigned short minus_one = 0xbc00u;
    float r0, r1;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r0) : "h"(h0), "h"(minus_one), "f"(ph.x));
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r1) : "h"(h1), "h"(minus_one), "f"(ph.y));
    const __half2 ll = __floats2half2_rn(r0, r1);
    lo = *reinterpret_cast<const unsigned *>(&ll);
  }
}

}  // namespace idgb200
