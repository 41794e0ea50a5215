// Pipe-throughput microbenchmarks for the gridder inner loop on sm_100a.
// Measures, per SM sub-partition (SMSP), the cycles one warp-wide "item" costs:
//   item = 1 FFMA (phase) + 1 FMUL.RZ + MUFU.SIN + MUFU.COS + 16 FMA (as 16 FFMA or 8 FFMA2)
// FP32 pipe floor: 18 cycles/item/SMSP (32 lanes per cycle); XU floor: 16 (4 lanes per cycle).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// ---- pure FFMA, 32 independent accumulators (3 distinct register operands)
__global__ void k_ffma(float *out, int iters, float a0, float b0) {
  float acc[32];
  for (int i = 0; i < 32; i++) acc[i] = threadIdx.x * 1e-3f + i;
  float a = a0 + threadIdx.x, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = fmaf(a, b, acc[i]);
    a += 1e-9f;
  }
  float s = 0;
  for (int i = 0; i < 32; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- pure FFMA2, 32 independent float2 accumulators (= 64 FMAs per 32 instructions)
__global__ void k_ffma2(float *out, int iters, float a0, float b0) {
  float2 acc[32];
  for (int i = 0; i < 32; i++) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  float2 a = make_float2(a0 + threadIdx.x, a0), b = make_float2(b0, b0 * 0.5f);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = ffma2(a, b, acc[i]);
    a.x += 1e-9f;
  }
  float s = 0;
  for (int i = 0; i < 32; i++) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- pure MUFU: sin + cos of 8 independent angles per iteration
__global__ void k_mufu(float *out, int iters, float a0) {
  float x[8], s = 0;
  for (int i = 0; i < 8; i++) x[i] = a0 + threadIdx.x * 0.01f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      float sn, cs;
      __sincosf(x[i], &sn, &cs);
      x[i] = sn + cs;   // 1 FADD keeps the chain alive
    }
  }
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- the gridder item mix; P pixels of ILP; PACKED: FFMA2 vs scalar FFMA; LDS: operands from smem
template <int P, bool PACKED, bool USE_LDS>
__global__ void __launch_bounds__(256, 2) k_mix(float *out, int iters, float k0, const float4 *vis_g) {
  __shared__ float4 s_vis[64 * 4];
  for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) s_vis[i] = vis_g[i];
  __syncthreads();
  float2 accA[P][4], accB[P][4];
  float idx[P], off[P];
  for (int j = 0; j < P; j++) {
    idx[j] = 1.0f + threadIdx.x * 1e-3f + j;
    off[j] = 0.5f * j;
    for (int p = 0; p < 4; p++) accA[j][p] = accB[j][p] = make_float2(0, 0);
  }
  float4 v0 = vis_g[0], v1 = vis_g[1], v2 = vis_g[2], v3 = vis_g[3];
  for (int it = 0; it < iters; it++) {
    const float wn = k0 + it * 1e-6f;
    if (USE_LDS) {
      const float4 *vt = s_vis + (it & 63) * 4;
      v0 = vt[0]; v1 = vt[1]; v2 = vt[2]; v3 = vt[3];
    }
#pragma unroll
    for (int j = 0; j < P; j++) {
      float sn, cs;
      __sincosf(fmaf(-idx[j], wn, off[j]), &sn, &cs);
      if (PACKED) {
        const float2 ph = make_float2(cs, sn);
        accA[j][0] = ffma2(make_float2(v0.x, v0.y), ph, accA[j][0]);
        accB[j][0] = ffma2(make_float2(v0.z, v0.w), ph, accB[j][0]);
        accA[j][1] = ffma2(make_float2(v1.x, v1.y), ph, accA[j][1]);
        accB[j][1] = ffma2(make_float2(v1.z, v1.w), ph, accB[j][1]);
        accA[j][2] = ffma2(make_float2(v2.x, v2.y), ph, accA[j][2]);
        accB[j][2] = ffma2(make_float2(v2.z, v2.w), ph, accB[j][2]);
        accA[j][3] = ffma2(make_float2(v3.x, v3.y), ph, accA[j][3]);
        accB[j][3] = ffma2(make_float2(v3.z, v3.w), ph, accB[j][3]);
      } else {
        const float vr[4] = {v0.x, v1.x, v2.x, v3.x}, vi[4] = {v0.z, v1.z, v2.z, v3.z};
#pragma unroll
        for (int p = 0; p < 4; p++) {
          accA[j][p].x = fmaf(vr[p], cs, accA[j][p].x);
          accA[j][p].x = fmaf(-vi[p], sn, accA[j][p].x);
          accA[j][p].y = fmaf(vr[p], sn, accA[j][p].y);
          accA[j][p].y = fmaf(vi[p], cs, accA[j][p].y);
        }
      }
    }
  }
  float s = 0;
  for (int j = 0; j < P; j++)
    for (int p = 0; p < 4; p++) s += accA[j][p].x + accA[j][p].y + accB[j][p].x + accB[j][p].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- same item mix, software pipelined: the phasors of iteration i+1 are produced while
// iteration i is accumulated, so MUFU and FFMA2 interleave inside one warp
template <int P, int MINB>
__global__ void __launch_bounds__(256, MINB) k_mix_sp(float *out, int iters, float k0, const float4 *vis_g) {
  __shared__ float4 s_vis[64 * 4];
  for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) s_vis[i] = vis_g[i];
  __syncthreads();
  float2 accA[P][4], accB[P][4];
  float idx[P], off[P];
  for (int j = 0; j < P; j++) {
    idx[j] = 1.0f + threadIdx.x * 1e-3f + j;
    off[j] = 0.5f * j;
    for (int p = 0; p < 4; p++) accA[j][p] = accB[j][p] = make_float2(0, 0);
  }
  float2 ph[P];
  for (int j = 0; j < P; j++) { float sn, cs; __sincosf(fmaf(-idx[j], k0, off[j]), &sn, &cs); ph[j] = make_float2(cs, sn); }
#pragma unroll 2
  for (int it = 0; it < iters; it++) {
    const float wn = k0 + (it + 1) * 1e-6f;
    const float4 *vt = s_vis + (it & 63) * 4;
    const float4 v0 = vt[0], v1 = vt[1], v2 = vt[2], v3 = vt[3];
    float2 nxt[P];
#pragma unroll
    for (int j = 0; j < P; j++) {
      float sn, cs;
      __sincosf(fmaf(-idx[j], wn, off[j]), &sn, &cs);
      nxt[j] = make_float2(cs, sn);
      const float2 p_ = ph[j];
      accA[j][0] = ffma2(make_float2(v0.x, v0.y), p_, accA[j][0]);
      accB[j][0] = ffma2(make_float2(v0.z, v0.w), p_, accB[j][0]);
      accA[j][1] = ffma2(make_float2(v1.x, v1.y), p_, accA[j][1]);
      accB[j][1] = ffma2(make_float2(v1.z, v1.w), p_, accB[j][1]);
      accA[j][2] = ffma2(make_float2(v2.x, v2.y), p_, accA[j][2]);
      accB[j][2] = ffma2(make_float2(v2.z, v2.w), p_, accB[j][2]);
      accA[j][3] = ffma2(make_float2(v3.x, v3.y), p_, accA[j][3]);
      accB[j][3] = ffma2(make_float2(v3.z, v3.w), p_, accB[j][3]);
    }
#pragma unroll
    for (int j = 0; j < P; j++) ph[j] = nxt[j];
  }
  float s = 0;
  for (int j = 0; j < P; j++)
    for (int p = 0; p < 4; p++) s += accA[j][p].x + accA[j][p].y + accB[j][p].x + accB[j][p].y;
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
  printf("device %s, %d SMs, max clock %.0f MHz\n", prop.name, sms, khz * 1e-3);
  float *out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256));
  float4 *vis; CK(cudaMalloc(&vis, sizeof(float4) * 64 * 4)); CK(cudaMemset(vis, 0x3c, sizeof(float4) * 64 * 4));
  const int iters = 20000;
  // cycles per warp-instruction per SMSP, assuming the max clock (lower bound on the real figure)
  auto report = [&](const char *name, double ms, double warp_instr_per_smsp, double items_per_smsp) {
    const double cyc = ms * 1e-3 * fmax;
    printf("%-34s %8.3f ms  %7.3f cyc/warp-instr/SMSP", name, ms, cyc / warp_instr_per_smsp);
    if (items_per_smsp > 0) printf("  %7.2f cyc/item/SMSP (FP32 floor 18, XU floor 16)", cyc / items_per_smsp);
    printf("\n");
  };
  for (int bps = 1; bps <= 4; bps++) {
    const int grid = sms * bps, warps_per_smsp = 2 * bps;
    printf("--- software pipelined, %d block(s)/SM x 256 threads (%d warps/SMSP)\n", bps, warps_per_smsp);
    double ms;
    if (bps <= 2) {
      ms = time_ms([&] { k_mix_sp<4, 2><<<grid, 256>>>(out, iters, 3.1f, vis); });
      report("SP item mix FFMA2 P=4 + LDS", ms, (double)iters * 4 * 12 * warps_per_smsp, (double)iters * 4 * warps_per_smsp);
    }
    ms = time_ms([&] { k_mix_sp<2, 4><<<grid, 256>>>(out, iters, 3.1f, vis); });
    report("SP item mix FFMA2 P=2 + LDS", ms, (double)iters * 2 * 12 * warps_per_smsp, (double)iters * 2 * warps_per_smsp);
    if (bps <= 3) {
      ms = time_ms([&] { k_mix_sp<3, 3><<<grid, 256>>>(out, iters, 3.1f, vis); });
      report("SP item mix FFMA2 P=3 + LDS", ms, (double)iters * 3 * 12 * warps_per_smsp, (double)iters * 3 * warps_per_smsp);
    }
  }
  for (int bps = 1; bps <= 2; bps++) {   // blocks per SM (256 threads each = 2 warps per SMSP per block)
    const int grid = sms * bps, warps_per_smsp = 2 * bps;
    printf("--- %d block(s)/SM x 256 threads (%d warps/SMSP)\n", bps, warps_per_smsp);
    double ms;
    ms = time_ms([&] { k_ffma<<<grid, 256>>>(out, iters, 1.0f, 1e-7f); });
    report("FFMA x32", ms, (double)iters * 32 * warps_per_smsp, 0);
    ms = time_ms([&] { k_ffma2<<<grid, 256>>>(out, iters, 1.0f, 1e-7f); });
    report("FFMA2 x32 (2 FMA each)", ms, (double)iters * 32 * warps_per_smsp, 0);
    ms = time_ms([&] { k_mufu<<<grid, 256>>>(out, iters, 0.3f); });
    report("MUFU sin+cos x8 (16 MUFU)", ms, (double)iters * 16 * warps_per_smsp, 0);
    ms = time_ms([&] { k_mix<4, true, false><<<grid, 256>>>(out, iters, 3.1f, vis); });
    report("item mix FFMA2 P=4", ms, (double)iters * 4 * 12 * warps_per_smsp, (double)iters * 4 * warps_per_smsp);
    ms = time_ms([&] { k_mix<4, true, true><<<grid, 256>>>(out, iters, 3.1f, vis); });
    report("item mix FFMA2 P=4 + LDS", ms, (double)iters * 4 * 12 * warps_per_smsp, (double)iters * 4 * warps_per_smsp);
    ms = time_ms([&] { k_mix<4, false, false><<<grid, 256>>>(out, iters, 3.1f, vis); });
    report("item mix FFMA  P=4", ms, (double)iters * 4 * 20 * warps_per_smsp, (double)iters * 4 * warps_per_smsp);
    ms = time_ms([&] { k_mix<4, false, true><<<grid, 256>>>(out, iters, 3.1f, vis); });
    report("item mix FFMA  P=4 + LDS", ms, (double)iters * 4 * 20 * warps_per_smsp, (double)iters * 4 * warps_per_smsp);
    ms = time_ms([&] { k_mix<2, true, true><<<grid, 256>>>(out, iters, 3.1f, vis); });
    report("item mix FFMA2 P=2 + LDS", ms, (double)iters * 2 * 12 * warps_per_smsp, (double)iters * 2 * warps_per_smsp);
  }
  return 0;
}
