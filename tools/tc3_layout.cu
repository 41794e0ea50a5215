// Experiment for the third-generation tensor-core kernels: the A operand of tcgen05.mma taken
// from TMEM (".ts" form), written there straight from registers with tcgen05.st - no shared-memory
// round trip for the phasor tile.  This file checks the operand layout numerically:
//   D[128][16] = A[128][K] * B[K][16],  A[m][k] written by thread m as packed half2 columns,
//   B in shared memory (K-major, no swizzle), K = 16 * NK.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../ska_sdp_idg_bench_b200/csrc/tc_common.cuh"
using namespace idgb200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void umma_f16_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long db, unsigned idesc,
                                            unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int NK = 2;   // K = 32

__global__ void __launch_bounds__(128, 1) k_layout(const float *Ag /*[128][K]*/, const float *Bg /*[K][16]*/, float *Dg) {
  __shared__ __align__(1024) unsigned char sB[NK * 512];
  __shared__ unsigned long long bar;
  __shared__ unsigned s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = 16 * NK;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // B: K-major core matrices: chunk (kc, n) = 8 consecutive k of column n at kc*256 + n*16
  for (int i = tid; i < K * 16; i += 128) {
    const int k = i / 16, n = i % 16;
    const int blk = k / 16, kk = k % 16, kc = kk / 8, ke = kk % 8;
    reinterpret_cast<__half *>(sB + blk * 512 + kc * 256 + n * 16)[ke] = __float2half_rn(Bg[k * 16 + n]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = s_tmem;
  const unsigned tmem_d = tmem_base, tmem_a = tmem_base + 16;   // D: columns 0..15, A: columns 16..16+K/2
  // thread m = tid writes row m: K/2 packed columns
  unsigned r[8 * NK];
  for (int j = 0; j < 8 * NK; j++) {
    const __half2 h = __floats2half2_rn(Ag[tid * K + 2 * j], Ag[tid * K + 2 * j + 1]);
    r[j] = *reinterpret_cast<const unsigned *>(&h);
  }
  const unsigned lane_base = (unsigned)(warp * 32) << 16;
#pragma unroll
  for (int blk = 0; blk < NK; blk++)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(tmem_a + lane_base + blk * 8), "r"(r[blk * 8 + 0]), "r"(r[blk * 8 + 1]), "r"(r[blk * 8 + 2]),
                   "r"(r[blk * 8 + 3]), "r"(r[blk * 8 + 4]), "r"(r[blk * 8 + 5]), "r"(r[blk * 8 + 6]), "r"(r[blk * 8 + 7])
                 : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    for (int blk = 0; blk < NK; blk++)
      umma_f16_ts(tmem_d, tmem_a + blk * 8, smem_desc(smem_u32(sB + blk * 512), B_CHUNK_BYTES, 128), idesc, blk > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  unsigned d[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
        "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
      : "r"(tmem_d + lane_base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 16; n++) Dg[tid * 16 + n] = __uint_as_float(d[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64));
}

int main() {
  const int K = 16 * NK;
  std::vector<float> A(128 * K), B(K * 16), D(128 * 16), R(128 * 16, 0.f);
  for (int m = 0; m < 128; m++) for (int k = 0; k < K; k++) A[m * K + k] = (float)((m * 3 + k * 5) % 11 - 5);
  for (int k = 0; k < K; k++) for (int n = 0; n < 16; n++) B[k * 16 + n] = (float)((k * 7 + n * 3) % 13 - 6);
  for (int m = 0; m < 128; m++) for (int n = 0; n < 16; n++) for (int k = 0; k < K; k++) R[m * 16 + n] += A[m * K + k] * B[k * 16 + n];
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  k_layout<<<1, 128>>>(dA, dB, dD);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 128 * 16; i++) if (D[i] != R[i]) { if (bad < 8) printf("mismatch m=%d n=%d got %g want %g\n", i / 16, i % 16, D[i], R[i]); bad++; }
  printf("tcgen05.mma A-from-TMEM layout check: %d mismatches of %d\n", bad, 128 * 16);
  return bad != 0;
}
