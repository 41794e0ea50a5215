// Microbenchmark for the tensor-core gridder's producer loop: what does one phasor item
// (FFMA phase, FMUL.RZ range scaling, MUFU.SIN, MUFU.COS, F2FP pack, 1/4 STS.128) cost a
// sub-partition, and which of its instructions share the XU with the MUFUs?
// 6 warps per SMSP (768 threads/SM) like the kernel; cycles are per warp-level item per SMSP.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

enum { M_SINCOS = 0, M_SINCOS_PACK, M_PACK_ONLY, M_PHASE_SINCOS, M_FULL, M_FULL_PRMT, M_SIN_ONLY, M_FULL_NOSTS, M_PACK_FMA, M_FULL_RND };

__device__ __forceinline__ unsigned pack_h2(float a, float b) {
  unsigned r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
  unsigned r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
__device__ __forceinline__ float msin(float x) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mcos(float x) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int MODE>
__global__ void __launch_bounds__(768, 1) k_mix(unsigned *out, int iters, float idx0, float off0, const float *wn_g, long long *cyc) {
  __shared__ __align__(16) unsigned s_out[768 * 4];
  __shared__ float s_wn[16];
  if (threadIdx.x < 16) s_wn[threadIdx.x] = wn_g[threadIdx.x];
  __syncthreads();
  const float idx = idx0 + threadIdx.x * 1e-3f, off = off0 + threadIdx.x * 1e-4f;
  unsigned accu = 0;
  float accf = 0.f;
  long long t0, t1;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) : "f"(idx), "f"(off) : "memory");
  for (int it = 0; it < iters; it++) {
    const float bump = it * 1e-6f;
#pragma unroll
    for (int g = 0; g < 4; g++) {          // 4 groups of 4 items -> 16 items per iteration
      unsigned pk[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float wn = s_wn[g * 4 + i];
        float s, c;
        if (MODE == M_SINCOS || MODE == M_SINCOS_PACK || MODE == M_SIN_ONLY) {
          const float ph = wn + bump + idx;            // 2 FADD instead of the FFMA + FMUL
          s = msin(ph);
          c = MODE == M_SIN_ONLY ? ph : mcos(ph);
        } else if (MODE == M_PACK_ONLY || MODE == M_PACK_FMA) {
          s = wn + bump; c = idx + bump;
          if (MODE == M_PACK_FMA) { s = fmaf(s, idx, off); c = fmaf(c, idx, off); s = fmaf(s, c, off); c = fmaf(c, s, off); }
        } else {
          const float ph = fmaf(-(idx + bump), wn, off);
          s = __sinf(ph);                              // FMUL.RZ + MUFU.SIN
          c = __cosf(ph);
        }
        if (MODE == M_SINCOS || MODE == M_PHASE_SINCOS || MODE == M_SIN_ONLY) { accf += s; accf += c; pk[i] = 0; }
        else if (MODE == M_FULL_PRMT) pk[i] = prmt(__float_as_uint(c), __float_as_uint(s), 0x7632);
        else if (MODE == M_FULL_RND) pk[i] = prmt(__float_as_uint(c) + 0x1000u, __float_as_uint(s) + 0x1000u, 0x7632);
        else pk[i] = pack_h2(c, s);
      }
      if (MODE == M_FULL || MODE == M_FULL_PRMT || MODE == M_FULL_RND)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"((unsigned)__cvta_generic_to_shared(&s_out[threadIdx.x * 4])),
                     "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
      else accu ^= pk[0] ^ pk[1] ^ pk[2] ^ pk[3];
    }
  }
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "r"(accu), "f"(accf) : "memory");
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = accu ^ __float_as_uint(accf) ^ s_out[(threadIdx.x * 5) % (768 * 4)];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static unsigned *out; static float *wn; static long long *cyc; static int sms;

template <int MODE>
void run(const char *name, double xu_ops, const char *what) {
  const int iters = 4000;
  k_mix<MODE><<<sms, 768>>>(out, iters, 3.0f, 0.5f, wn, cyc);
  CK(cudaDeviceSynchronize());
  k_mix<MODE><<<sms, 768>>>(out, iters, 3.0f, 0.5f, wn, cyc);
  CK(cudaDeviceSynchronize());
  long long h[256];
  CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < sms; i++) avg += h[i];
  avg /= sms;
  const double per_item = avg / (iters * 16.0 * 6.0);   // 6 warps per SMSP, 16 items per iteration
  printf("%-14s %7.2f cycles / item / SMSP   (%s; %.0f MUFU)\n", name, per_item, what, xu_ops);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  sms = p.multiProcessorCount;
  CK(cudaMalloc(&out, sizeof(unsigned) * sms * 768));
  CK(cudaMalloc(&cyc, sizeof(long long) * sms));
  CK(cudaMalloc(&wn, 64));
  float h[16]; for (int i = 0; i < 16; i++) h[i] = 0.01f * (i + 1);
  CK(cudaMemcpy(wn, h, 64, cudaMemcpyHostToDevice));
  printf("device %s, %d SMs; 6 warps / SMSP\n", p.name, sms);
  run<M_SIN_ONLY>("sin", 1, "2 FADD + MUFU.SIN");
  run<M_SINCOS>("sincos", 2, "2 FADD + MUFU.SIN + MUFU.COS + 2 FADD");
  run<M_SINCOS_PACK>("sincos+pack", 2, "2 FADD + 2 MUFU + F2FP");
  run<M_PACK_ONLY>("pack", 0, "2 FADD + F2FP");
  run<M_PACK_FMA>("pack+6fma", 0, "2 FADD + 4 FFMA + F2FP");
  run<M_PHASE_SINCOS>("phase+sincos", 2, "FFMA + 2 FMUL.RZ? + 2 MUFU + 2 FADD");
  run<M_FULL>("full", 2, "FFMA + FMUL.RZ + 2 MUFU + F2FP + STS.128/4");
  run<M_FULL_PRMT>("full prmt", 2, "same with PRMT instead of F2FP (bf16-style pack)");
  run<M_FULL_RND>("full rnd+prmt", 2, "same with 2 IADD + PRMT");
  return 0;
}
