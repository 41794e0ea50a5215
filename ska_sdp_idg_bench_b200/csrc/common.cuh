// Device-side helpers shared by the gridder and degridder kernels (sm_100a only).
//
// Math follows app/common/math.hpp:9-92 of the reference and, where it is cheap,
// the exact operation order of the reference's CPU binary (see oracle/idg_oracle.c
// for how that order was established), so that phases are bit-identical to the CPU
// result and the remaining difference is sincos + summation order only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "idg_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "idg_b200 kernels are written for sm_100a (B200) only"
#endif

namespace idgb200 {

constexpr int NR_POL = IDGB200_NR_CORRELATIONS;

// Kernel arguments: the reference's 13 kernel parameters
// (app/CUDA/kernels/gridder_v8.cu:286-291) as one POD.
struct KernelArgs {
  int grid_size;
  int subgrid_size;
  float image_size;
  float w_step_in_lambda;
  int nr_channels;
  int nr_stations;
  const idgb200_uvw *uvw;
  const float *wavenumbers;
  const float2 *visibilities;  // gridder: in, degridder: out
  const float *spheroidal;
  const float2 *aterms;
  const idgb200_metadata *metadata;
  const float2 *subgrids;      // gridder: out, degridder: in
  int nr_subgrids;     // number of subgrids this launch processes ...
  int subgrid_offset;  // ... starting at this index of metadata / subgrids
};

// ---------------------------------------------------------------- packed FP32
// fma.rn.f32x2 -> one FFMA2 issue slot for two FMAs (sm_100+).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  return __ffma2_rn(a, b, c);
}

// ------------------------------------------------------------------- phasors
// returns (cos(phase), sin(phase))
template <int MODE>
__device__ __forceinline__ float2 phasor(float phase) {
  float s, c;
  if (MODE == IDGB200_SINCOS_FAST) {
    __sincosf(phase, &s, &c);
  } else if (MODE == IDGB200_SINCOS_REDUCED) {
    // k = round(phase / 2pi) by the magic-number trick (|phase| < 2^22 * 2pi),
    // r = phase - k*2pi with 2pi split in two so that k*hi is exact for |k| < 2^12
    const float kf = __fadd_rn(__fmaf_rn(phase, 0.15915494309189535f, 12582912.0f), -12582912.0f);
    float r = __fmaf_rn(kf, -6.28125f, phase);                   // hi: 8 significant bits
    r = __fmaf_rn(kf, -1.9353071795864769e-03f, r);              // lo = 2pi - hi
    __sincosf(r, &s, &c);
  } else {
    sincosf(phase, &s, &c);
  }
  return make_float2(c, s);
}

// ------------------------------------------------------- image-plane geometry
// app/common/math.hpp:9-17 (double intermediate, one rounding)
__device__ __forceinline__ float compute_l(int x, int subgrid_size, float image_size) {
  return (float)((x + 0.5 - (subgrid_size / 2)) * (double)image_size / (double)subgrid_size);
}

// app/common/math.hpp:19-24; the CPU binary evaluates l*l + m*m as fma(l, l, m*m)
__device__ __forceinline__ float compute_n(float l, float m) {
  const float tmp = __fmaf_rn(l, l, __fmul_rn(m, m));
  return tmp > 1.0f ? 1.0f : __fdiv_rn(tmp, __fadd_rn(1.0f, __fsqrt_rn(__fsub_rn(1.0f, tmp))));
}

struct SubgridCtx {
  long long time_offset;
  int nr_timesteps;
  int aterm_index;
  int station1, station2;
  float u_offset, v_offset, w_offset;
};

// gridder_reference.cpp:23-39 / degridder_reference.cpp:24-32,77-79
__device__ __forceinline__ SubgridCtx load_ctx(const KernelArgs &a, int s) {
  const idgb200_metadata m = a.metadata[s];
  const int base0 = a.metadata[0].baseline_offset;
  SubgridCtx c;
  c.time_offset = (long long)(m.baseline_offset - base0) + m.time_offset;
  c.nr_timesteps = m.nr_timesteps;
  c.aterm_index = m.aterm_index;
  c.station1 = (int)m.station1;
  c.station2 = (int)m.station2;
  const double two_pi = 6.283185307179586476925286766559;
  const float w_offset_in_lambda = (float)((double)a.w_step_in_lambda * (m.z + 0.5));
  const double scale = two_pi / (double)a.image_size;
  c.u_offset = (float)((m.x + a.subgrid_size / 2 - a.grid_size / 2) * scale);
  c.v_offset = (float)((m.y + a.subgrid_size / 2 - a.grid_size / 2) * scale);
  c.w_offset = (float)(two_pi * (double)w_offset_in_lambda);
  return c;
}

// ------------------------------------------------------------ 2x2 Jones algebra
// std::complex<float> product with the FMA contraction the reference's CPU
// binary uses (oracle/idg_oracle.c: cmul / cmul_x).
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)),
                     __fmaf_rn(a.y, b.x, __fmul_rn(a.x, b.y)));
}
__device__ __forceinline__ float2 cmul_x(float2 a, float2 b) {
  return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)),
                     __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}
__device__ __forceinline__ float2 conjf(float2 a) { return make_float2(a.x, -a.y); }

// app/common/math.hpp:26-36; X01 / X23: rows {0,1} / {2,3} use cmul_x
template <bool X01, bool X23>
__device__ __forceinline__ void matmul2x2(const float2 *a, const float2 *b, float2 *c) {
  if (X01) {
    c[0] = cadd(cmul_x(a[0], b[0]), cmul_x(a[1], b[2]));
    c[1] = cadd(cmul_x(a[0], b[1]), cmul_x(a[1], b[3]));
  } else {
    c[0] = cadd(cmul(a[0], b[0]), cmul(a[1], b[2]));
    c[1] = cadd(cmul(a[0], b[1]), cmul(a[1], b[3]));
  }
  if (X23) {
    c[2] = cadd(cmul_x(a[2], b[0]), cmul_x(a[3], b[2]));
    c[3] = cadd(cmul_x(a[2], b[1]), cmul_x(a[3], b[3]));
  } else {
    c[2] = cadd(cmul(a[2], b[0]), cmul(a[3], b[2]));
    c[3] = cadd(cmul(a[2], b[1]), cmul(a[3], b[3]));
  }
}

// app/common/math.hpp:64-77: P <- A1^H * P * A2
__device__ __forceinline__ void apply_aterm_gridder(float2 *p, const float2 *a1, const float2 *a2) {
  const float2 a1h[4] = {conjf(a1[0]), conjf(a1[2]), conjf(a1[1]), conjf(a1[3])};
  float2 t[4];
  matmul2x2<true, false>(a1h, p, t);
  matmul2x2<false, false>(t, a2, p);
}

// app/common/math.hpp:79-92: P <- A1 * P * A2^H
__device__ __forceinline__ void apply_aterm_degridder(float2 *p, const float2 *a1, const float2 *a2) {
  const float2 a2h[4] = {conjf(a2[0]), conjf(a2[2]), conjf(a2[1]), conjf(a2[3])};
  float2 t[4];
  matmul2x2<true, true>(a1, p, t);
  matmul2x2<false, false>(t, a2h, p);
}

// 4 complex = one Jones matrix = 2 x 16-byte loads (aterms are 32-byte aligned records)
__device__ __forceinline__ void load_jones(const float2 *aterms, size_t index, float2 *a) {
  const float4 *q = reinterpret_cast<const float4 *>(aterms + index);
  const float4 lo = __ldg(q), hi = __ldg(q + 1);
  a[0] = make_float2(lo.x, lo.y);
  a[1] = make_float2(lo.z, lo.w);
  a[2] = make_float2(hi.x, hi.y);
  a[3] = make_float2(hi.z, hi.w);
}

// 16-byte shared-memory load the compiler may not hoist out of a loop (used for values
// that are wanted once per outer iteration and must not occupy registers in between)
__device__ __forceinline__ float4 lds128_pinned(const float4 *p) {
  float4 r;
  const unsigned s = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(s));
  return r;
}

// ----------------------------------------------------------------- async copy
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::); }

}  // namespace idgb200
