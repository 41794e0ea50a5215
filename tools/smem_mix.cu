// Do the LSU (STS.128) and the tensor core's operand reads share one shared-memory crossbar?
// One CTA per SM: warps 1..8 stream STS.128 into a scratch area (4 wavefronts of 128 B per warp
// instruction), an elected lane of warp 0 issues tcgen05.mma M=128 N=16 K=16 with both operands in
// shared memory (A 4 KB + B 512 B = 36 wavefronts per MMA).  Three runs: stores only, MMAs only, both.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -I ska_sdp_idg_bench_b200/csrc \
//        -o tools/bin/smem_mix tools/smem_mix.cu
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

using namespace idgb200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// mode bit 2: every MMA accumulates into the same 16 TMEM columns (a dependent chain); without it the
// MMAs rotate over 8 accumulators, as the 8 tiles of a CTA do in the kernels
__global__ void __launch_bounds__(288, 1) k_mix(int mode, int sts_iters, int mma_count, long long *cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;                       // 2 x 4 KB
  unsigned char *sB = smem + 8192;                // 512 B
  unsigned char *scratch = smem + 16384;          // 8 warps x 4 KB
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + 16384 + 32768);
  unsigned *s_tmem = reinterpret_cast<unsigned *>(bar + 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<unsigned *>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = *s_tmem;
  const long long t0 = clock64();
  if (warp == 0) {
    if (mode & 2) {
      const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      // mode bit 8: K chunks of an 8-row group side by side (LBO 128 B, SBO 256 B) instead of the kernels'
      // chunk-major tile (LBO 2048 B, SBO 128 B)
      const unsigned long long da = (mode & 8) ? smem_desc(smem_u32(sA), 128, 256) : smem_desc(smem_u32(sA), A_CHUNK_BYTES, 128);
      const unsigned long long db = smem_desc(smem_u32(sB), B_CHUNK_BYTES, 128);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      unsigned phase = 0;
      const int batch = (mode & 16) ? 256 : 16;   // MMAs between two commits (the wait drains the pipe)
      for (int i = 0; i < mma_count; i += batch) {
        if (elect_one()) {
          for (int j = 0; j < batch; j++)
            umma_f16(tmem + ((mode & 4) ? 0 : (j & 7) * 16), da + (unsigned long long)((j & 1) * (4096 >> 4)), db, idesc,
                     (i + j) >= 8 ? 1u : 0u);
          umma_commit(&bar[0]);
        }
        __syncwarp();
        mbar_wait(&bar[0], phase);
        phase ^= 1u;
      }
    }
  } else if (mode & 1) {
    unsigned char *mine = scratch + (warp - 1) * 4096 + lane * 16;
    uint4 v = make_uint4(tid, tid + 1, tid + 2, tid + 3);
    for (int i = 0; i < sts_iters; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(mine + j * 512)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      v.x += i;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  long long *d; CK(cudaMalloc(&d, sizeof(long long) * sms));
  const size_t smem = 16384 + 32768 + 64;
  CK(cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sts_iters = 20000, mma_count = 32000;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  auto run = [&](int mode, const char *name) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_mix<<<sms, 288, smem>>>(mode, sts_iters, mma_count, d);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_mix<<<sms, 288, smem>>>(mode, sts_iters, mma_count, d);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long c; CK(cudaMemcpy(&c, d, sizeof c, cudaMemcpyDeviceToHost));
    const double cyc = ms * 1e-3 * khz * 1e3;    // at the maximum clock
    const double sts_wf = (mode & 1) ? 8.0 * sts_iters * 8 * 4 : 0, mma_wf = (mode & 2) ? 36.0 * mma_count : 0;   // 32 wavefronts of A + 4 of B
    printf("%-26s %8.3f ms = %9.0f cycles (clock64: %lld)  STS wavefronts/clk %.3f  MMA operand wavefronts/clk %.3f (%.1f clk/MMA)  sum %.3f\n",
           name, ms, cyc, c, sts_wf / cyc, mma_wf / cyc, (mode & 2) ? cyc / mma_count : 0.0, (sts_wf + mma_wf) / cyc);
  };
  printf("device %s, %d SMs\n", prop.name, sms);
  run(1, "STS.128 only (8 warps)");
  run(2, "MMA SS, 8 accumulators");
  run(6, "MMA SS, 1 accumulator");
  run(10, "MMA SS, interleaved A layout");
  run(18, "MMA SS, 256 per commit");
  run(3, "STS + MMA (8 accumulators)");
  return 0;
}
