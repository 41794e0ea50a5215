// Experiment (north_star: "tensor cores only if ... a shared-memory-generated phasor tile
// contracted against the visibility tile beats the FP32/SFU path"): a tcgen05 gridder core.
//
//   D[pixel][8] += A[pixel][k] * B[k][8]        per subgrid, pixels = 1024 = 8 M-tiles of 128
//     A[pixel][(v,0)] = cos(phase), A[pixel][(v,1)] = sin(phase)            fp16, made by MUFU
//     B[(v,0)][re,pol] =  vr   B[(v,1)][re,pol] = -vi                        fp16 hi + lo halves
//     B[(v,0)][im,pol] =  vi   B[(v,1)][im,pol] =  vr                        (N = 8 x 2 = 16)
//   accumulators: fp32 in TMEM, 8 tiles x 16 columns.
//
// Standalone: builds random inputs, runs this kernel and a plain FP32 kernel, compares the raw
// pixel sums (no A-terms / taper: those are identical epilogue work) and times both.
// Every wait is bounded (trap instead of hang).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/tc_gridder tools/tc_gridder.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int N = 32, NPIX = N * N, C = 16, NT = 256, TILES = 8;
constexpr int A_CHUNK_BYTES = 128 * 16;                       // one 16-byte K-chunk of 128 rows
constexpr int A_TILE_BYTES = 4 * A_CHUNK_BYTES;               // K = 32 halfs = 4 chunks
constexpr int A_STAGE_BYTES = TILES * A_TILE_BYTES;           // 64 KB
constexpr int B_CHUNK_BYTES = 16 * 16;
constexpr int B_STAGE_BYTES = 4 * B_CHUNK_BYTES;              // 1 KB
constexpr int SMEM_BYTES = 2 * A_STAGE_BYTES + 2 * B_STAGE_BYTES + 64 + C * 4;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned done = 0;
  for (long long spin = 0; !done; spin++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spin > (1ll << 24)) { printf("mbar_wait timeout block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr, unsigned lbo, unsigned sbo) {
  // cute::UMMA::SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1 layout[61,64)=0 (no swizzle)
  return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo >> 4) & 0x3FFF) << 16) |
         ((unsigned long long)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                         unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

struct Args {
  const float *uvw;      // [S][T][3]
  const float *wn;       // [C]
  const float2 *vis;     // [S][T][C][4]
  const float *uoff;     // [S][2]  (u_offset, v_offset)
  float2 *out;           // [S][4][NPIX]   raw pixel sums
  int T;
  float image_size;
};

__device__ __forceinline__ float pix_l(int x, float image_size) { return (x + 0.5f - N / 2) * image_size / N; }

// ------------------------------------------------------------------ tensor-core kernel
__global__ void __launch_bounds__(NT, 1) k_tc(const Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;
  unsigned char *sB = smem + 2 * A_STAGE_BYTES;
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(sB + 2 * B_STAGE_BYTES);  // [3]
  unsigned *s_tmem = reinterpret_cast<unsigned *>(mbar + 3);
  float *s_wn = reinterpret_cast<float *>(s_tmem + 2);

  const int s = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T;

  if (tid < C) s_wn[tid] = a.wn[tid];
  if (tid == 0) {
    mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_init(&mbar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = *s_tmem;

  // this thread's 4 pixels: tile = warp, rows lane + 32 j
  float l[4], m[4], off[4];
  const float uo = a.uoff[s * 2 + 0], vo = a.uoff[s * 2 + 1];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int q = warp * 128 + lane + 32 * j;
    l[j] = pix_l(q % N, a.image_size);
    m[j] = pix_l(q / N, a.image_size);
    off[j] = fmaf(uo, l[j], vo * m[j]);
  }
  // instruction descriptor: D=F32 [4,6)=1, A=B=F16, K-major, N>>3 at [17,23), M>>4 at [24,29)
  const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
  const float *uvw = a.uvw + (size_t)s * T * 3;
  const float2 *vis = a.vis + (size_t)s * T * C * 4;

  for (int t = 0; t < T; t++) {
    const int stage = t & 1;
    const int use = t >> 1;
    if (use >= 1) mbar_wait(&mbar[stage], (use - 1) & 1);   // MMAs that read this stage are done
    unsigned char *A = sA + stage * A_STAGE_BYTES;
    unsigned char *B = sB + stage * B_STAGE_BYTES;

    // B operand: 64 threads, one 16-byte chunk each: row n (hi/lo, pol, re/im) x 4 visibilities
    uint4 bchunk = make_uint4(0, 0, 0, 0);
    if (tid < 64) {
      const int n = tid & 15, kc = tid >> 4;
      const int h = n >> 3, p = (n >> 1) & 3, im = n & 1;
      unsigned w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 v = vis[((size_t)t * C + kc * 4 + i) * 4 + p];
        float x0 = im ? v.y : v.x;        // multiplies cos
        float x1 = im ? v.x : -v.y;       // multiplies sin
        __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
        if (h) { h0 = __float2half_rn(x0 - __half2float(h0)); h1 = __float2half_rn(x1 - __half2float(h1)); }
        w[i] = (unsigned)__half_as_ushort(h0) | ((unsigned)__half_as_ushort(h1) << 16);
      }
      bchunk = make_uint4(w[0], w[1], w[2], w[3]);
    }

    // A operand: phasors of this thread's 4 pixels x 16 channels
    const float u = uvw[t * 3 + 0], v = uvw[t * 3 + 1];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float idx = fmaf(u, l[j], v * m[j]);
      const int row = lane + 32 * j;
#pragma unroll
      for (int kc = 0; kc < 4; kc++) {
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float sn, cs;
          __sincosf(fmaf(-idx, s_wn[kc * 4 + i], off[j]), &sn, &cs);
          const __half2 hh = __floats2half2_rn(cs, sn);
          w[i] = *reinterpret_cast<const unsigned *>(&hh);
        }
        *reinterpret_cast<uint4 *>(A + warp * A_TILE_BYTES + kc * A_CHUNK_BYTES + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    if (tid < 64) *reinterpret_cast<uint4 *>(B + (tid >> 4) * B_CHUNK_BYTES + (tid & 15) * 16) = bchunk;

    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned a_base = smem_u32(A), b_base = smem_u32(B);
#pragma unroll
      for (int tile = 0; tile < TILES; tile++)
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
          const unsigned long long da = smem_desc(a_base + tile * A_TILE_BYTES + kk * 2 * A_CHUNK_BYTES, A_CHUNK_BYTES, 128);
          const unsigned long long db = smem_desc(b_base + kk * 2 * B_CHUNK_BYTES, B_CHUNK_BYTES, 128);
          umma_f16(tmem_base + tile * 16, da, db, idesc, (t > 0 || kk > 0) ? 1u : 0u);
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[stage])) : "memory");
      if (t == T - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[2])) : "memory");
    }
  }

  // epilogue: TMEM -> registers -> global; warp w reads lanes 32 (w % 4) .. +31 of tiles w/4, w/4+2, ...
  mbar_wait(&mbar[2], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float2 *out = a.out + (size_t)s * 4 * NPIX;
  const int q4 = warp & 3;
  for (int tile = warp >> 2; tile < TILES; tile += 2) {
    unsigned r[16];
    const unsigned taddr = tmem_base + ((unsigned)(q4 * 32) << 16) + tile * 16;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int pixel = tile * 128 + q4 * 32 + lane;
#pragma unroll
    for (int p = 0; p < 4; p++)
      out[p * NPIX + pixel] = make_float2(__uint_as_float(r[2 * p]) + __uint_as_float(r[8 + 2 * p]),
                                          __uint_as_float(r[2 * p + 1]) + __uint_as_float(r[8 + 2 * p + 1]));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
}


// ------------------------------------------------------------------ v2: warp specialised
// 16 producer warps (2 pixels per thread) + 1 MMA warp that also builds the B operand;
// producers and the MMA warp only meet at mbarriers (full / empty per stage), no block barrier
// in the main loop.
constexpr int NT2 = 17 * 32;
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(NT2, 1) k_tc2(const Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;
  unsigned char *sB = smem + 2 * A_STAGE_BYTES;
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(sB + 2 * B_STAGE_BYTES);  // full[2] empty[2] done
  unsigned *s_tmem = reinterpret_cast<unsigned *>(mbar + 5);
  float *s_wn = reinterpret_cast<float *>(s_tmem + 2);
  unsigned long long *full = mbar, *empty = mbar + 2, *done = mbar + 4;

  const int s = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T;
  if (tid < C) s_wn[tid] = a.wn[tid];
  if (tid == 0) {
    mbar_init(&full[0], 512); mbar_init(&full[1], 512);
    mbar_init(&empty[0], 1); mbar_init(&empty[1], 1); mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = *s_tmem;
  const float *uvw = a.uvw + (size_t)s * T * 3;
  const float2 *vis = a.vis + (size_t)s * T * C * 4;

  if (warp < 16) {
    // ---------------- producers: tile = warp / 2, rows (warp & 1) * 64 + lane + 32 j
    const int tile = warp >> 1;
    float l[2], m[2], off[2];
    int row[2];
    const float uo = a.uoff[s * 2 + 0], vo = a.uoff[s * 2 + 1];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      row[j] = (warp & 1) * 64 + lane + 32 * j;
      const int q = tile * 128 + row[j];
      l[j] = pix_l(q % N, a.image_size);
      m[j] = pix_l(q / N, a.image_size);
      off[j] = fmaf(uo, l[j], vo * m[j]);
    }
    for (int t = 0; t < T; t++) {
      const int stage = t & 1, use = t >> 1;
      if (use >= 1) mbar_wait(&empty[stage], (use - 1) & 1);
      unsigned char *A = sA + stage * A_STAGE_BYTES + tile * A_TILE_BYTES;
      const float u = uvw[t * 3 + 0], v = uvw[t * 3 + 1];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const float idx = fmaf(u, l[j], v * m[j]);
#pragma unroll
        for (int kc = 0; kc < 4; kc++) {
          unsigned w[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            float sn, cs;
            __sincosf(fmaf(-idx, s_wn[kc * 4 + i], off[j]), &sn, &cs);
            const __half2 hh = __floats2half2_rn(cs, sn);
            w[i] = *reinterpret_cast<const unsigned *>(&hh);
          }
          *reinterpret_cast<uint4 *>(A + kc * A_CHUNK_BYTES + row[j] * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&full[stage]);
    }
  } else {
    // ---------------- MMA warp: builds B (2 chunks per lane), issues the MMAs
    const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    auto load_b = [&](int t, float2 (&raw)[2][4]) {
#pragma unroll
      for (int h2 = 0; h2 < 2; h2++) {
        const int chunk = lane + 32 * h2, n = chunk & 15, kc = chunk >> 4, p = (n >> 1) & 3;
#pragma unroll
        for (int i = 0; i < 4; i++) raw[h2][i] = vis[((size_t)t * C + kc * 4 + i) * 4 + p];
      }
    };
    float2 raw[2][4];
    load_b(0, raw);
    for (int t = 0; t < T; t++) {
      const int stage = t & 1, use = t >> 1;
      if (use >= 1) mbar_wait(&empty[stage], (use - 1) & 1);
      unsigned char *B = sB + stage * B_STAGE_BYTES;
#pragma unroll
      for (int h2 = 0; h2 < 2; h2++) {
        const int chunk = lane + 32 * h2, n = chunk & 15, kc = chunk >> 4, hl = n >> 3, im = n & 1;
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float2 vv = raw[h2][i];
          const float x0 = im ? vv.y : vv.x, x1 = im ? vv.x : -vv.y;
          __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
          if (hl) { h0 = __float2half_rn(x0 - __half2float(h0)); h1 = __float2half_rn(x1 - __half2float(h1)); }
          w[i] = (unsigned)__half_as_ushort(h0) | ((unsigned)__half_as_ushort(h1) << 16);
        }
        *reinterpret_cast<uint4 *>(B + kc * B_CHUNK_BYTES + n * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (t + 1 < T) load_b(t + 1, raw);
      if (lane == 0) {
        mbar_wait(&full[stage], use & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned a_base = smem_u32(sA + stage * A_STAGE_BYTES), b_base = smem_u32(B);
#pragma unroll
        for (int tile = 0; tile < TILES; tile++)
#pragma unroll
          for (int kk = 0; kk < 2; kk++) {
            const unsigned long long da = smem_desc(a_base + tile * A_TILE_BYTES + kk * 2 * A_CHUNK_BYTES, A_CHUNK_BYTES, 128);
            const unsigned long long db = smem_desc(b_base + kk * 2 * B_CHUNK_BYTES, B_CHUNK_BYTES, 128);
            umma_f16(tmem_base + tile * 16, da, db, idesc, (t > 0 || kk > 0) ? 1u : 0u);
          }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
        if (t == T - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
      }
      __syncwarp();
    }
  }

  // epilogue (producer warps): warp w reads TMEM lanes 32 (w % 4) .. +31 of tiles w/4 and w/4 + 4
  if (warp < 16) {
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float2 *out = a.out + (size_t)s * 4 * NPIX;
    const int q4 = warp & 3;
    for (int tile = warp >> 2; tile < TILES; tile += 4) {
      unsigned r[16];
      const unsigned taddr = tmem_base + ((unsigned)(q4 * 32) << 16) + tile * 16;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                     "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int pixel = tile * 128 + q4 * 32 + lane;
#pragma unroll
      for (int p = 0; p < 4; p++)
        out[p * NPIX + pixel] = make_float2(__uint_as_float(r[2 * p]) + __uint_as_float(r[8 + 2 * p]),
                                            __uint_as_float(r[2 * p + 1]) + __uint_as_float(r[8 + 2 * p + 1]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
}

// ------------------------------------------------------------------ plain FP32 kernel (checker)
__global__ void k_ref(const Args a) {
  const int s = blockIdx.x, T = a.T;
  const float uo = a.uoff[s * 2 + 0], vo = a.uoff[s * 2 + 1];
  for (int q = threadIdx.x; q < NPIX; q += blockDim.x) {
    const float l = pix_l(q % N, a.image_size), m = pix_l(q / N, a.image_size);
    const float off = fmaf(uo, l, vo * m);
    float2 acc[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (int t = 0; t < T; t++) {
      const float *uvw = a.uvw + ((size_t)s * T + t) * 3;
      const float idx = fmaf(uvw[0], l, uvw[1] * m);
      for (int c = 0; c < C; c++) {
        float sn, cs;
        sincosf(fmaf(-idx, a.wn[c], off), &sn, &cs);
        const float2 *v = a.vis + (((size_t)s * T + t) * C + c) * 4;
        for (int p = 0; p < 4; p++) {
          acc[p].x += v[p].x * cs - v[p].y * sn;
          acc[p].y += v[p].x * sn + v[p].y * cs;
        }
      }
    }
    for (int p = 0; p < 4; p++) a.out[((size_t)s * 4 + p) * NPIX + q] = acc[p];
  }
}

int main(int argc, char **argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 2368, T = 128;
  printf("tcgen05 gridder experiment: %d subgrids, N=%d, T=%d, C=%d, smem %d B\n", S, N, T, C, SMEM_BYTES);
  std::vector<float> uvw((size_t)S * T * 3), wn(C), uoff((size_t)S * 2);
  std::vector<float2> vis((size_t)S * T * C * 4);
  srand(1);
  auto rnd = [] { return rand() / (float)RAND_MAX; };
  for (int i = 0; i < C; i++) wn[i] = 3.14f + 0.0147f * i;
  for (size_t i = 0; i < uvw.size(); i++) uvw[i] = (i % 3 == 2) ? 0.f : (rnd() - 0.5f) * 2000.f;
  for (size_t i = 0; i < uoff.size(); i++) uoff[i] = (rnd() - 0.5f) * 512.f * 628.3f;
  for (size_t i = 0; i < vis.size(); i++) vis[i] = make_float2(rnd() * 2 - 1, rnd() * 2 - 1);
  Args a;
  float *d_uvw, *d_wn, *d_uoff; float2 *d_vis, *d_out, *d_ref;
  CK(cudaMalloc(&d_uvw, uvw.size() * 4)); CK(cudaMalloc(&d_wn, C * 4)); CK(cudaMalloc(&d_uoff, uoff.size() * 4));
  CK(cudaMalloc(&d_vis, vis.size() * 8)); CK(cudaMalloc(&d_out, (size_t)S * 4 * NPIX * 8)); CK(cudaMalloc(&d_ref, (size_t)S * 4 * NPIX * 8));
  CK(cudaMemcpy(d_uvw, uvw.data(), uvw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_wn, wn.data(), C * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_uoff, uoff.data(), uoff.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_vis, vis.data(), vis.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xff, (size_t)S * 4 * NPIX * 8));
  a.uvw = d_uvw; a.wn = d_wn; a.vis = d_vis; a.uoff = d_uoff; a.T = T; a.image_size = 0.01f;
  CK(cudaFuncSetAttribute(k_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CK(cudaFuncSetAttribute(k_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  for (int version = 1; version <= 2; version++) {
  auto launch = [&](int blocks) {
    if (version == 1) k_tc<<<blocks, NT, SMEM_BYTES>>>(a); else k_tc2<<<blocks, NT2, SMEM_BYTES>>>(a);
  };
  printf("--- kernel v%d\n", version);
  CK(cudaMemset(d_out, 0xff, (size_t)S * 4 * NPIX * 8));

  // correctness on the first 8 subgrids
  const int SC = S < 8 ? S : 8;
  a.out = d_out; launch(SC);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  a.out = d_ref; k_ref<<<SC, 256>>>(a);
  CK(cudaDeviceSynchronize());
  std::vector<float2> h_out((size_t)SC * 4 * NPIX), h_ref((size_t)SC * 4 * NPIX);
  CK(cudaMemcpy(h_out.data(), d_out, h_out.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_ref.data(), d_ref, h_ref.size() * 8, cudaMemcpyDeviceToHost));
  double maxd = 0, maxr = 0, sd = 0, sr = 0;
  for (size_t i = 0; i < h_out.size(); i++) {
    const double dx = h_out[i].x - h_ref[i].x, dy = h_out[i].y - h_ref[i].y;
    maxd = fmax(maxd, sqrt(dx * dx + dy * dy));
    maxr = fmax(maxr, hypot(h_ref[i].x, h_ref[i].y));
    sd += dx * dx + dy * dy; sr += (double)h_ref[i].x * h_ref[i].x + (double)h_ref[i].y * h_ref[i].y;
  }
  printf("parity vs plain FP32 kernel: max|d|/max|ref| = %.3e, rel-RMS = %.3e (first = (%.4f,%.4f) vs (%.4f,%.4f))\n",
         maxd / maxr, sqrt(sd / sr), h_out[0].x, h_out[0].y, h_ref[0].x, h_ref[0].y);

  // timing
  a.out = d_out;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(S); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 3; i++) launch(S);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
  const double mvis = 1e-6 * S * T * C, flop = (double)S * T * NPIX * (10 + 34.0 * C);
  printf("tcgen05 core: %.3f ms, %.1f MVis/s, %.2f TFLOP/s-equivalent = %.1f%% of the FP32 peak (74.45)\n", ms, mvis / (ms * 1e-3),
         flop / (ms * 1e-3) * 1e-12, 100 * flop / (ms * 1e-3) * 1e-12 / 74.45);
  }
  return 0;
}
