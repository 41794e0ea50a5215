// Third microbenchmark: what does a register operand cost an FFMA2 / FFMA on sm_100a?
// acc[i] = fma(a[i % KA], b[i % KB], acc[i]) over 32 (FFMA2) or 48 (FFMA) accumulators; KA, KB
// control how many distinct multiplicand registers rotate, i.e. how often ptxas can serve an
// operand from the reuse cache.  tools/sass_model.py on the same binary gives the per-instruction
// read counts; together they calibrate the operand-bandwidth model.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int KA, int KB, bool PACKED>
__global__ void __launch_bounds__(256, 2) k_ops(float *out, int iters, const float2 *src) {
  constexpr int NA = PACKED ? 32 : 48;
  float2 a[KA], b[KB], acc[NA];
  for (int i = 0; i < KA; i++) a[i] = src[i + threadIdx.x % 3];
  for (int i = 0; i < KB; i++) b[i] = src[16 + i + threadIdx.x % 5];
  for (int i = 0; i < NA; i++) acc[i] = make_float2(i, -i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NA; i++) {
      if (PACKED) acc[i] = __ffma2_rn(a[i % KA], b[i % KB], acc[i]);
      else acc[i].x = fmaf(a[i % KA].x, b[i % KB].x, acc[i].x);
    }
    if (KA > 0) a[it % KA].x += 1e-9f;   // keep the loop honest
  }
  float r = 0;
  for (int i = 0; i < NA; i++) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static float *out; static float2 *src; static int sms; static double f_hz;
static const int iters = 20000;

template <int KA, int KB, bool PACKED>
void run() {
  const int grid = sms * 2;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_ops<KA, KB, PACKED><<<grid, 256>>>(out, iters, src);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 3; i++) k_ops<KA, KB, PACKED><<<grid, 256>>>(out, iters, src);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
  const double n = (double)iters * (PACKED ? 32 : 48) * 4;   // warp-instructions per SMSP
  printf("k_ops<KA=%d,KB=%d,%s>: %.3f cyc/instr/SMSP\n", KA, KB, PACKED ? "FFMA2" : "FFMA ", ms * 1e-3 * f_hz / n);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  sms = prop.multiProcessorCount; f_hz = khz * 1e3;
  CK(cudaMalloc(&out, sizeof(float) * sms * 2 * 256));
  CK(cudaMalloc(&src, sizeof(float2) * 64)); CK(cudaMemset(src, 0, sizeof(float2) * 64));
  run<1, 1, true>(); run<2, 1, true>(); run<4, 1, true>(); run<8, 1, true>();
  run<3, 2, true>(); run<5, 3, true>(); run<7, 5, true>(); run<8, 7, true>(); run<11, 9, true>();
  run<1, 1, false>(); run<2, 1, false>(); run<4, 1, false>(); run<8, 1, false>();
  run<3, 2, false>(); run<5, 3, false>(); run<7, 5, false>(); run<8, 7, false>(); run<11, 9, false>();
  return 0;
}
