// Inner-loop formulations of the gridder item mix, timed against each other (sm_100a).
// All variants: P pixels per thread, software-pipelined phasors, visibilities broadcast from smem.
//   V_ROT   : FFMA2, duplicated vis (re,re)(im,im), rotated accumulators A/B      (64 B / vis)
//   V_SWZ   : FFMA2, vis as vr scalars + (-vi,vi) pairs, phasor swizzle LO_HI     (48 B / vis)
//   V_GEMM  : scalar FFMA written as a GEMM micro-kernel (outer product), raw vis (32 B / vis)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/mixv tools/mixv.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

enum { V_ROT = 0, V_SWZ = 1, V_GEMM = 2, V_SWZ_BYV = 3, V_SWZ_NSP = 4, V_SWZ_BOU = 5 };
__device__ __forceinline__ float2 f2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int V, int P, int MINB, int NT = 256, int UNR = 2>
__global__ void __launch_bounds__(NT, MINB) k_mix(float *out, int iters, float k0, const float4 *vis_g) {
  __shared__ float4 s_vis[64 * 4];
  for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) s_vis[i] = vis_g[i];
  __syncthreads();
  float2 acc[P][8];   // V_ROT uses all 8 per pixel, the others 4
  float idx[P], off[P];
#pragma unroll
  for (int j = 0; j < P; j++) {
    idx[j] = 1.0f + threadIdx.x * 1e-3f + j;
    off[j] = 0.5f * j;
#pragma unroll
    for (int p = 0; p < 8; p++) acc[j][p] = make_float2(0, 0);
  }
  float2 ph[P];
#pragma unroll
  for (int j = 0; j < P; j++) { float sn, cs; __sincosf(fmaf(-idx[j], k0, off[j]), &sn, &cs); ph[j] = make_float2(cs, sn); }
#pragma unroll UNR
  for (int it = 0; it < iters; it++) {
    const float wn = k0 + (it + 1) * 1e-6f;
    float2 nxt[P];
    if (V == V_ROT) {
      const float4 *vt = s_vis + (it & 63) * 4;
      const float4 q0 = vt[0], q1 = vt[1], q2 = vt[2], q3 = vt[3];
      const float2 v[8] = {{q0.x, q0.y}, {q0.z, q0.w}, {q1.x, q1.y}, {q1.z, q1.w}, {q2.x, q2.y}, {q2.z, q2.w}, {q3.x, q3.y}, {q3.z, q3.w}};
#pragma unroll
      for (int j = 0; j < P; j++) {
        float sn, cs; __sincosf(fmaf(-idx[j], wn, off[j]), &sn, &cs); nxt[j] = make_float2(cs, sn);
#pragma unroll
        for (int p = 0; p < 8; p++) acc[j][p] = f2(v[p], ph[j], acc[j][p]);
      }
    } else if (V == V_SWZ || V == V_SWZ_BYV) {
      const float4 *vt = s_vis + (it & 63) * 3;
      const float4 q0 = vt[0], q1 = vt[1], q2 = vt[2];
      const float vr[4] = {q0.x, q0.y, q0.z, q0.w};
      const float2 vi[4] = {{q1.x, q1.y}, {q1.z, q1.w}, {q2.x, q2.y}, {q2.z, q2.w}};
      if (V == V_SWZ) {
#pragma unroll
        for (int j = 0; j < P; j++) {
          float sn, cs; __sincosf(fmaf(-idx[j], wn, off[j]), &sn, &cs); nxt[j] = make_float2(cs, sn);
          const float2 p1 = ph[j], p2 = make_float2(ph[j].y, ph[j].x);
#pragma unroll
          for (int p = 0; p < 4; p++) acc[j][p] = f2(make_float2(vr[p], vr[p]), p1, acc[j][p]);
#pragma unroll
          for (int p = 0; p < 4; p++) acc[j][p] = f2(vi[p], p2, acc[j][p]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < P; j++) { float sn, cs; __sincosf(fmaf(-idx[j], wn, off[j]), &sn, &cs); nxt[j] = make_float2(cs, sn); }
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
          for (int j = 0; j < P; j++) acc[j][p] = f2(make_float2(vr[p], vr[p]), ph[j], acc[j][p]);
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
          for (int j = 0; j < P; j++) acc[j][p] = f2(vi[p], make_float2(ph[j].y, ph[j].x), acc[j][p]);
      }
    } else if (V == V_SWZ_NSP) {  // no software pipeline: phasors become ready pixel by pixel
      const float4 *vt = s_vis + (it & 63) * 3;
      const float4 q0 = vt[0], q1 = vt[1], q2 = vt[2];
      const float vr[4] = {q0.x, q0.y, q0.z, q0.w};
      const float2 vi[4] = {{q1.x, q1.y}, {q1.z, q1.w}, {q2.x, q2.y}, {q2.z, q2.w}};
      const float wn0 = k0 + it * 1e-6f;
#pragma unroll
      for (int j = 0; j < P; j++) {
        float sn, cs; __sincosf(fmaf(-idx[j], wn0, off[j]), &sn, &cs);
        const float2 p1 = make_float2(cs, sn), p2 = make_float2(sn, cs);
        nxt[j] = p1;
#pragma unroll
        for (int p = 0; p < 4; p++) acc[j][p] = f2(make_float2(vr[p], vr[p]), p1, acc[j][p]);
#pragma unroll
        for (int p = 0; p < 4; p++) acc[j][p] = f2(vi[p], p2, acc[j][p]);
      }
    } else if (V == V_SWZ_BOU) {  // non-pipelined, operands of consecutive FFMA2 chained (boustrophedon)
      const float4 *vt = s_vis + (it & 63) * 3;
      const float4 q0 = vt[0], q1 = vt[1], q2 = vt[2];
      const float vr[4] = {q0.x, q0.y, q0.z, q0.w};
      const float2 vi[4] = {{q1.x, q1.y}, {q1.z, q1.w}, {q2.x, q2.y}, {q2.z, q2.w}};
      const float wn0 = k0 + it * 1e-6f;
#pragma unroll
      for (int j = 0; j < P; j++) {
        float sn, cs; __sincosf(fmaf(-idx[j], wn0, off[j]), &sn, &cs);
        const float2 p1 = make_float2(cs, sn), p2 = make_float2(sn, cs);
        nxt[j] = p1;
        if ((j & 1) == 0) {
#pragma unroll
          for (int p = 0; p < 4; p++) acc[j][p] = f2(make_float2(vr[p], vr[p]), p1, acc[j][p]);
#pragma unroll
          for (int p = 0; p < 4; p++) acc[j][p] = f2(vi[p], p2, acc[j][p]);
        } else {
#pragma unroll
          for (int p = 3; p >= 0; p--) acc[j][p] = f2(vi[p], p2, acc[j][p]);
#pragma unroll
          for (int p = 3; p >= 0; p--) acc[j][p] = f2(make_float2(vr[p], vr[p]), p1, acc[j][p]);
        }
      }
    } else {  // V_GEMM: C[2P][4] += A[2P] (x) B[4], twice per visibility
      const float4 *vt = s_vis + (it & 63) * 2;
      const float4 q0 = vt[0], q1 = vt[1];
      const float vr[4] = {q0.x, q0.z, q1.x, q1.z}, vi[4] = {q0.y, q0.w, q1.y, q1.w};
#pragma unroll
      for (int j = 0; j < P; j++) { float sn, cs; __sincosf(fmaf(-idx[j], wn, off[j]), &sn, &cs); nxt[j] = make_float2(cs, sn); }
#pragma unroll
      for (int j = 0; j < P; j++)
#pragma unroll
        for (int p = 0; p < 4; p++) {
          acc[j][p].x = fmaf(ph[j].x, vr[p], acc[j][p].x);
          acc[j][p].y = fmaf(ph[j].y, vr[p], acc[j][p].y);
        }
#pragma unroll
      for (int j = 0; j < P; j++)
#pragma unroll
        for (int p = 0; p < 4; p++) {
          acc[j][p].x = fmaf(-ph[j].y, vi[p], acc[j][p].x);
          acc[j][p].y = fmaf(ph[j].x, vi[p], acc[j][p].y);
        }
    }
#pragma unroll
    for (int j = 0; j < P; j++) ph[j] = nxt[j];
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < P; j++)
#pragma unroll
    for (int p = 0; p < 8; p++) s += acc[j][p].x + acc[j][p].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float *out; static float4 *vis; static int sms; static double f_hz;
static const int iters = 12000;

template <int V, int P, int MINB, int NT = 256, int UNR = 2>
void run(const char *name, int bps) {
  const int grid = sms * bps;
  int occ = 0;
  auto kern = k_mix<V, P, MINB, NT, UNR>;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<grid, NT>>>(out, iters, 3.1f, vis);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 3; i++) kern<<<grid, NT>>>(out, iters, 3.1f, vis);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
  const double items = (double)iters * P * (NT / 128) * bps;   // warp-items per SMSP
  printf("%-10s P=%d NT=%d unroll=%d blocks/SM=%d (occ limit %d): %6.2f cyc/item/SMSP  -> %5.1f%% of FP32 peak\n", name, P, NT, UNR, bps, occ,
         ms * 1e-3 * f_hz / items, 100.0 * 18.0 / (ms * 1e-3 * f_hz / items));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  sms = prop.multiProcessorCount; f_hz = khz * 1e3;
  CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256));
  CK(cudaMalloc(&vis, sizeof(float4) * 64 * 4)); CK(cudaMemset(vis, 0x3c, sizeof(float4) * 64 * 4));
  run<V_SWZ_NSP, 8, 1, 256, 1>("swz-nsp", 1);
  run<V_SWZ_BOU, 8, 1, 256, 1>("swz-bou", 1); run<V_SWZ_BOU, 8, 1, 256, 1>("swz-bou", 2); run<V_SWZ_BOU, 8, 2, 128, 1>("swz-bou", 2);
  run<V_SWZ_BOU, 8, 2, 128, 1>("swz-bou", 3); run<V_SWZ_BOU, 8, 4, 128, 1>("swz-bou", 4); run<V_SWZ_BOU, 8, 1, 256, 2>("swz-bou", 1);
  run<V_SWZ_BOU, 4, 4, 256, 1>("swz-bou", 4); run<V_SWZ_BOU, 4, 2, 256, 1>("swz-bou", 2); run<V_SWZ_BOU, 6, 2, 256, 1>("swz-bou", 2);
  return 0;
}
