// TMEM -> register bandwidth of tcgen05.ld (what bounds the row-column degridder's sum over the rows: the accumulator of a
// 128-visibility tile is 128 lanes x 256 columns x 4 B = 128 KB, read once per tile).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/tmem_ld tools/tmem_ld.cu && tools/bin/tmem_ld
// One CTA per SM, W warps, each reads its lane quadrant's columns REPS times with shape 32x32b.xN, D loads in flight.
#include <cstdio>
#include <cuda_runtime.h>

template <int N> struct Ld;
#define LD_ASM(N, ...) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #N ".b32 {" __VA_ARGS__ "}, [%" #N "];"
template <> struct Ld<16> {
  __device__ static void go(unsigned addr, unsigned *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
  }
};
template <> struct Ld<32> {
  __device__ static void go(unsigned addr, unsigned *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
  }
};

template <int N, int D>
__global__ void __launch_bounds__(1024, 1) k(int reps, long long *out, unsigned *sink) {
  __shared__ unsigned s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_tmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned base = s_tmem + ((unsigned)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < reps; i++) {
    unsigned r[D][N];
#pragma unroll
    for (int d = 0; d < D; d++) Ld<N>::go(base + ((i * D + d) * N) % 512, r[d]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int d = 0; d < D; d++)
#pragma unroll
      for (int j = 0; j < N; j++) acc ^= r[d][j];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(512));
}

template <int N, int D>
void run(int warps, long long *d_out, unsigned *d_sink) {
  const int reps = 2000;
  k<N, D><<<148, warps * 32>>>(reps, d_out, d_sink);
  k<N, D><<<148, warps * 32>>>(reps, d_out, d_sink);
  long long c = 0;
  cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double)warps * reps * D * N * 128.0;
  printf("32x32b.x%-3d  %d in flight  %2d warps: %8lld clocks, %7.1f B/clock/SM, %6.1f clocks per load\n", N, D, warps, c, bytes / c,
         (double)c / (reps * D));
}

int main() {
  long long *d_out;
  unsigned *d_sink;
  cudaMalloc(&d_out, 8);
  cudaMalloc(&d_sink, 4);
  for (int w : {1, 2, 4, 8, 16, 32}) run<32, 1>(w, d_out, d_sink);
  for (int w : {4, 8, 16}) run<32, 2>(w, d_out, d_sink);
  for (int w : {4, 8, 16, 32}) run<16, 1>(w, d_out, d_sink);
  for (int w : {4, 8, 16}) run<16, 4>(w, d_out, d_sink);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
