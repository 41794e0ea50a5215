// What does the tensor pipe sustain for the MMA pattern of the kernels?  tools/smem_mix.cu measured one
// issuing warp: 40 clocks per tcgen05.mma M=128 N=16 K=16 (operands in shared memory) with 256 MMAs per
// commit, 59 with a commit + wait every 16.  The kernels look different: W producer warps per SM, each
// with its own A buffers, accumulator and barrier, issue b MMAs (gridder_tc.cu: 2, degridder_tc8.cu: 4),
// one tcgen05.commit, and wait for that commit before their next stage (single-buffered A).  This
// program runs exactly that, with nothing else in the loop, and prints clocks per MMA per SM:
//   wait = 1: as in the kernels;  wait = 0: commits only (what a commit costs the pipe);
//   spin = n: n dependent FFMA between the wait and the issue (the producer's work, so that the warps
//             do not all sit in the barrier at once).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -I ska_sdp_idg_bench_b200/csrc \
//        -o tools/bin/mma_commit tools/mma_commit.cu
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

using namespace idgb200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(768, 1) k_commit(int W, int b, int wait, int spin, int iters, float *sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;                                   // [W][2][4 KB]
  unsigned char *sB = smem + W * 8192;                        // 512 B
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(sB + 512);   // [W] stage, [W] final
  unsigned *s_tmem = reinterpret_cast<unsigned *>(bar + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (W * 8192 + 512) / 4; i += blockDim.x) reinterpret_cast<unsigned *>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    for (int i = 0; i < 64; i++) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *s_tmem;
  if (warp < W) {
    const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned long long da = smem_desc(smem_u32(sA + warp * 8192), A_CHUNK_BYTES, 128);
    const unsigned long long db = smem_desc(smem_u32(sB), B_CHUNK_BYTES, 128);
    const unsigned my = smem_u32(&bar[warp]), fin = smem_u32(&bar[32 + warp]);
    const unsigned d = tmem + warp * 16;
    float x = (float)tid;
    for (int i = 0; i < iters; i++) {
      if (wait && i >= 1) mbar_wait_u(my, (i - 1) & 1);
      for (int s = 0; s < spin; s++) x = __fmaf_rn(x, 1.0001f, 0.5f);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      __syncwarp();
      if (elect_one()) {
        for (int j = 0; j < b; j++)
          umma_f16(d, da + (unsigned long long)((j & 1) * (4096 >> 4)), db, idesc, (i | j) ? 1u : 0u);
        umma_commit_u(my);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit_u(fin);
    __syncwarp();
    mbar_wait_u(fin, 0);
    if (x == 12345.678f) sink[0] = x;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  float *sink; CK(cudaMalloc(&sink, 4));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, %d MHz\n", prop.name, sms, khz / 1000);
  printf("%3s %2s %4s %5s  %9s  %s\n", "W", "b", "wait", "spin", "ms", "clocks per MMA per SM");
  auto run = [&](int W, int b, int wait, int spin) {
    const size_t smem = (size_t)W * 8192 + 512 + 64 * 8 + 16;
    CK(cudaFuncSetAttribute(k_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 48000 / (W * b) * 8;      // ~384 k MMAs per SM
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_commit<<<sms, W * 32, smem>>>(W, b, wait, spin, iters, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_commit<<<sms, W * 32, smem>>>(W, b, wait, spin, iters, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cyc = ms * 1e-3 * khz * 1e3;
    printf("%3d %2d %4d %5d  %9.3f  %.1f\n", W, b, wait, spin, ms, cyc / ((double)iters * W * b));
  };
  const int Ws[] = {1, 8, 12, 24};
  for (int W : Ws)
    for (int b : {1, 2, 4, 16})
      for (int wait = 0; wait < 2; wait++) run(W, b, wait, 0);
  // the kernels' operating points: 24 warps x 2 MMAs (gridder), 12 warps x 4 MMAs (degridder), with the
  // producer's work between the wait and the issue
  for (int spin : {256, 1024, 4096}) {
    run(24, 2, 1, spin);
    run(12, 4, 1, spin);
  }
  return 0;
}
