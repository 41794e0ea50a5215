// Dispatch cost of generating the phasors of a block of equally spaced channels by rotation
// (gridder_tc.cu: tc_produce_linear), per (pixel, channel) item and SM sub-partition:
//   fp32      ph *= d as FMUL2 + FFMA2, then F2FP to the packed fp16 operand        (what ships)
//   fp16      the rotation itself in half2: HMUL2 + HFMA2, result already packed
//   mixed     fp32 rotation by d^2 for the even channels (+ F2FP), one half2 step for the odd ones
// plus the raw issue rates of HFMA2 and of the mixed-precision FHFMA (fma.rn.f32.f16).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/hrot tools/hrot.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned pack(float2 p) {
  __half2 h = __floats2half2_rn(p.x, p.y);
  return *reinterpret_cast<unsigned *>(&h);
}
__device__ __forceinline__ float2 rot32(float2 ph, float2 dxx, float2 dny) {
  return __ffma2_rn(make_float2(ph.y, ph.x), dny, __fmul2_rn(ph, dxx));
}
__device__ __forceinline__ __half2 rot16(__half2 ph, __half2 dxx, __half2 dny) {
  return __hfma2(__lowhigh2highlow(ph), dny, __hmul2(ph, dxx));
}

// P pixels per thread, 8 channels each per iteration; results xor-ed so that nothing is dead
template <int MODE, int P>
__global__ void __launch_bounds__(256) k_rot(unsigned *out, int iters, float a0) {
  float2 ph0[P], d[P];
  for (int j = 0; j < P; j++) {
    __sincosf(a0 + threadIdx.x * 0.01f + j, &ph0[j].y, &ph0[j].x);
    __sincosf(0.001f * (threadIdx.x + j + 1), &d[j].y, &d[j].x);
  }
  unsigned acc = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < P; j++) {
      float2 ph = ph0[j];
      const float2 dxx = make_float2(d[j].x, d[j].x), dny = make_float2(-d[j].y, d[j].y);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          acc ^= pack(ph);
          if (i < 7) ph = rot32(ph, dxx, dny);
        }
      } else if (MODE == 1) {
        __half2 h = __floats2half2_rn(ph.x, ph.y);
        const __half2 hxx = __floats2half2_rn(d[j].x, d[j].x), hny = __floats2half2_rn(-d[j].y, d[j].y);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          acc ^= *reinterpret_cast<unsigned *>(&h);
          if (i < 7) h = rot16(h, hxx, hny);
        }
      } else {
        // d^2 in fp32 (once per pixel), even channels in fp32, odd channels one half2 step
        const float2 d2 = make_float2(d[j].x * d[j].x - d[j].y * d[j].y, 2.f * d[j].x * d[j].y);
        const float2 exx = make_float2(d2.x, d2.x), eny = make_float2(-d2.y, d2.y);
        const __half2 hxx = __floats2half2_rn(d[j].x, d[j].x), hny = __floats2half2_rn(-d[j].y, d[j].y);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const unsigned e = pack(ph);
          acc ^= e;
          const __half2 o = rot16(*reinterpret_cast<const __half2 *>(&e), hxx, hny);
          acc ^= *reinterpret_cast<const unsigned *>(&o) * 3u;
          if (i < 3) ph = rot32(ph, exx, eny);
        }
      }
      ph0[j].x += 1e-7f;   // keeps the loop from being hoisted
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int KIND>
__global__ void __launch_bounds__(256) k_raw(float *out, int iters, float a0) {
  __half2 h[16];
  float f[16];
  for (int i = 0; i < 16; i++) { h[i] = __floats2half2_rn(a0 + i, a0 - i); f[i] = a0 * i; }
  const __half2 m = __floats2half2_rn(0.999f, 1.001f), c = __floats2half2_rn(1e-3f, -1e-3f);
  const unsigned short hm = 0x3bff;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (KIND == 0) h[i] = __hfma2(h[i], m, c);
      else asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(f[i]) : "h"(hm), "h"((unsigned short)(0x3c00 + i)));
    }
  }
  float s = 0;
  for (int i = 0; i < 16; i++) s += f[i] + __low2float(h[i]) + __high2float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; i++) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int sms = prop.multiProcessorCount;
  const double fmax = khz * 1e3;
  printf("device %s, %d SMs, max clock %.0f MHz; 6 warps / SMSP\n", prop.name, sms, khz * 1e-3);
  void *out; CK(cudaMalloc(&out, 4 * sms * 3 * 256));
  const int iters = 20000, bps = 3, wps = 2 * bps;
  auto rep = [&](const char *name, double ms, double per_smsp, const char *unit) {
    printf("%-44s %8.3f ms  %6.2f cycles / %s / SMSP\n", name, ms, ms * 1e-3 * fmax / per_smsp, unit);
  };
  double ms;
  ms = time_ms([&] { k_raw<0><<<sms * bps, 256>>>((float *)out, iters, 1.f); });
  rep("HFMA2 x16 independent", ms, (double)iters * 16 * wps, "instr");
  ms = time_ms([&] { k_raw<1><<<sms * bps, 256>>>((float *)out, iters, 1.f); });
  rep("FHFMA x16 independent", ms, (double)iters * 16 * wps, "instr");
  ms = time_ms([&] { k_rot<0, 4><<<sms * bps, 256>>>((unsigned *)out, iters, 0.3f); });
  rep("fp32 rotation + F2FP (ships)", ms, (double)iters * 4 * 8 * wps, "item");
  ms = time_ms([&] { k_rot<1, 4><<<sms * bps, 256>>>((unsigned *)out, iters, 0.3f); });
  rep("half2 rotation (7-step chain)", ms, (double)iters * 4 * 8 * wps, "item");
  ms = time_ms([&] { k_rot<2, 4><<<sms * bps, 256>>>((unsigned *)out, iters, 0.3f); });
  rep("mixed: fp32 by d^2 + one half2 step", ms, (double)iters * 4 * 8 * wps, "item");
  return 0;
}
