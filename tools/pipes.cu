// Second microbenchmark: how do MUFU, FFMA/FFMA2 and LDS share an SM sub-partition?
// Each kernel runs, per loop iteration and per thread, 8 MUFU ops (8 independent
// dependency chains) plus NF FMA-pipe instructions per MUFU (independent
// accumulators) and optionally NL broadcast LDS.128 per 8 MUFU.
// Reports cycles per MUFU per SMSP; if the pipes were independent this would be
// max(8.2, NF * cost_fma).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

enum { OP_SIN = 0, OP_EX2 = 1, OP_NONE = 2 };

template <int OP>
__device__ __forceinline__ float sfu(float x) {
  float r;
  if (OP == OP_SIN) r = __sinf(x);
  else if (OP == OP_EX2) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  else r = x;
  return r;
}

// NF: FMA-pipe instructions per MUFU; PACKED: FFMA2 or FFMA; NL: LDS.128 per 8 MUFU
template <int OP, int NF, bool PACKED, int NL>
__global__ void __launch_bounds__(256, 2) k_ratio(float *out, int iters, float a0, const float4 *g) {
  __shared__ float4 s[256];
  s[threadIdx.x] = g[threadIdx.x];
  __syncthreads();
  float x[8];
  for (int i = 0; i < 8; i++) x[i] = 0.1f * i + threadIdx.x * 1e-3f;
  constexpr int NA = (NF > 0 ? NF : 1) * 8 > 48 ? 48 : (NF > 0 ? NF : 1) * 8;  // accumulators
  float2 acc[NA];
  for (int i = 0; i < NA; i++) acc[i] = make_float2(i, -i);
  float2 a = make_float2(a0, a0 * 0.5f), b = make_float2(1e-7f, 2e-7f);
  float4 ld = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const float4 v = s[(it + l * 16) & 255];
      ld.x += v.x; ld.y += v.y;   // 2 FADD per LDS keep the loads alive (FMA pipe, counted below)
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      x[i] = sfu<OP>(x[i]);
#pragma unroll
      for (int f = 0; f < NF; f++) {
        const int k = (i * NF + f) % NA;
        if (PACKED) acc[k] = __ffma2_rn(a, b, acc[k]);
        else acc[k].x = fmaf(a.x, b.x, acc[k].x);
      }
    }
    a.x += 1e-9f;
  }
  float r = ld.x + ld.y;
  for (int i = 0; i < 8; i++) r += x[i];
  for (int i = 0; i < NA; i++) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
double time_ms(F launch, int reps = 3) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; i++) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

static float *out; static float4 *g; static int sms; static double f_max_hz;
static const int iters = 8000;

template <int OP, int NF, bool PACKED, int NL>
void run(const char *name) {
  const int grid = sms * 2;   // 2 blocks/SM = 4 warps/SMSP
  const double ms = time_ms([&] { k_ratio<OP, NF, PACKED, NL><<<grid, 256>>>(out, iters, 1.0f, g); });
  const double cyc = ms * 1e-3 * f_max_hz;
  const double mufu_per_smsp = (double)iters * 8 * 4;   // warp-level MUFU per SMSP
  printf("%-8s NF=%2d %-6s NL=%d : %7.2f cyc per MUFU-slot/SMSP\n", name, NF, PACKED ? "FFMA2" : "FFMA", NL,
         cyc / mufu_per_smsp);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  sms = prop.multiProcessorCount; f_max_hz = khz * 1e3;
  CK(cudaMalloc(&out, sizeof(float) * sms * 4 * 256));
  CK(cudaMalloc(&g, sizeof(float4) * 256)); CK(cudaMemset(g, 0, sizeof(float4) * 256));
  printf("# cycles per MUFU slot per SMSP (4 warps/SMSP); independent pipes => max(8.2, NF*cost)\n");
  run<OP_NONE, 4, true, 0>("none");  run<OP_NONE, 8, false, 0>("none");
  run<OP_SIN, 0, true, 0>("sin");   run<OP_EX2, 0, true, 0>("ex2");
  run<OP_SIN, 1, true, 0>("sin");   run<OP_SIN, 2, true, 0>("sin");   run<OP_SIN, 3, true, 0>("sin");
  run<OP_SIN, 4, true, 0>("sin");   run<OP_SIN, 5, true, 0>("sin");   run<OP_SIN, 6, true, 0>("sin");
  run<OP_SIN, 8, true, 0>("sin");
  run<OP_EX2, 2, true, 0>("ex2");   run<OP_EX2, 4, true, 0>("ex2");   run<OP_EX2, 6, true, 0>("ex2");
  run<OP_SIN, 2, false, 0>("sin");  run<OP_SIN, 4, false, 0>("sin");  run<OP_SIN, 8, false, 0>("sin");
  run<OP_SIN, 12, false, 0>("sin"); run<OP_EX2, 8, false, 0>("ex2");  run<OP_EX2, 12, false, 0>("ex2");
  run<OP_NONE, 4, true, 2>("none"); run<OP_NONE, 4, true, 4>("none"); run<OP_NONE, 4, true, 8>("none");
  run<OP_SIN, 4, true, 2>("sin");   run<OP_SIN, 4, true, 4>("sin");   run<OP_SIN, 4, true, 8>("sin");
  return 0;
}
