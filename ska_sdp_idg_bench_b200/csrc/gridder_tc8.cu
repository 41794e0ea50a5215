// IDG gridder on tcgen05 + TMEM, second generation: a producer warp owns TWO M-tiles (variant 26).
//
// Same GEMM, operands, builder warp and epilogue as gridder_tc.cu (D[pixel][n] += A[pixel][k] B[k][n],
// A = fp16 phasors made by the warp, B = the visibilities split into fp16 hi + lo), specialised to
// the regular case that gridder_tc.cu's lean stage loop serves - every block of 8 channels equally
// spaced with ONE spacing and an even number of blocks (what the reference's init.cpp and a real
// telescope's channelisation give) - and reorganised like degridder_tc8.cu:
//   * lane l of producer warp w owns pixels l + 32 j of tile 2 w and of tile 2 w + 1 (8 pixels), so a
//     stage = (timestep, 16 channels) is 128 items per thread: the per-stage bookkeeping (barrier
//     waits, fences, the elected MMA issue, ~100 instructions) is paid half as often per item;
//   * 5 warps per CTA leave 128 registers per thread: l, m, n, offset of all 8 pixels stay in
//     registers (the 72-register kernel reloads 8 spilled values per timestep) and the 8 independent
//     pixel chains interleave.
// Per stage the elected lane issues four MMAs (M=128, N=16, K=16: two tiles x two channel blocks) and
// commits them to the warp's one empty barrier.
//
// Whether the wavenumbers are regular is only known on the device, so the launcher runs a one-warp
// check kernel first (the same per-block test as gridder_tc.cu: common.cuh linear_channels) that
// writes a flag, then THIS kernel, which returns at once unless the flag is set, then gridder_tc.cu's
// variant 24, which returns at once if it is: irregular channels cost two empty launches (~30 us),
// regular ones one.  FAST sincos only.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace idgb200 {

namespace {

constexpr int G8_PAIRS = 4;                      // producer warps per CTA, two M-tiles each
constexpr int G8_TILES = 2 * G8_PAIRS;
constexpr int G8_CB = 8;                         // channels per block -> K = 16 per MMA
constexpr int G8_A_BUF = 2 * A_CHUNK_BYTES;      // 4 KB: one tile, one channel block
constexpr int G8_A_WARP = 4 * G8_A_BUF;          // [tile half][block a | b]
constexpr int G8_B_SLOT = 2 * B_CHUNK_BYTES;     // 512 B
constexpr int G8_NB = 16;                        // B ring slots
constexpr int G8_THREADS = (G8_PAIRS + 1) * 32;

// regular = the recurrence applies to every block with one spacing and the blocks pair up
__global__ void gridder_regular_check_kernel(const float *__restrict__ wavenumbers, const int C, int *flag,
                                             int *zero2) {
  if (zero2 && threadIdx.x < 2) zero2[threadIdx.x] = 0;   // the list counts of gridder_fold.cu's planar check
  const int ncb = (C + G8_CB - 1) / G8_CB;
  bool ok = !(ncb & 1);
  float dw0 = 0.f;
  for (int cb = threadIdx.x; cb < ncb; cb += 32) {
    float dw;
    ok = ok && linear_channels(wavenumbers, cb * G8_CB, min(G8_CB, C - cb * G8_CB), &dw);
    if (cb == 0) dw0 = dw;
  }
  dw0 = __shfl_sync(0xffffffffu, dw0, 0);
  for (int cb = threadIdx.x; cb < ncb; cb += 32) {
    float dw;
    linear_channels(wavenumbers, cb * G8_CB, min(G8_CB, C - cb * G8_CB), &dw);
    ok = ok && dw == dw0;
  }
  ok = __all_sync(0xffffffffu, ok);
  if (threadIdx.x == 0) *flag = ok ? 1 : 0;
}

// one block of 8 equally spaced channels for 4 pixels (rows lane + 32 j of one tile): first channel by
// sincos (the reference's angle, bit for bit), second by rotation, the rest by the three-term
// recurrence (gridder_tc.cu: tc_produce_linear has the error argument)
__device__ __forceinline__ void g8_produce(unsigned char *A, const float wn0, const float2 *rot, const float *idx,
                                           const float *off, const int lane) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float2 prev = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(-idx[j], wn0, off[j]));   // gridder_reference.cpp:69
    const float2 d = rot[j];
    const float2 dxx = make_float2(d.x, d.x), dny = make_float2(-d.y, d.y);
    const float c2 = __fadd_rn(d.x, d.x);
    const float2 cc = make_float2(c2, c2);
    unsigned pk[8], unused;
    pack_phasor<false>(prev, pk[0], unused);
    float2 cur = ffma2(make_float2(prev.y, prev.x), dny, __fmul2_rn(prev, dxx));
    pack_phasor<false>(cur, pk[1], unused);
#pragma unroll
    for (int i = 2; i < 8; i++) {
      const float2 nxt = ffma2(cur, cc, make_float2(-prev.x, -prev.y));
      pack_phasor<false>(nxt, pk[i], unused);
      prev = cur;
      cur = nxt;
    }
    *reinterpret_cast<uint4 *>(A + (lane + 32 * j) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4 *>(A + A_CHUNK_BYTES + (lane + 32 * j) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

__global__ void __launch_bounds__(G8_THREADS, 3)
gridder_tc8_kernel(const KernelArgs a, const int slabs, const int tiles_per_cta, const int tmem_cols,
                   const int *__restrict__ regular_flag) {
  if (*regular_flag == 0) return;   // irregular channels: gridder_tc.cu runs instead
  extern __shared__ __align__(1024) unsigned char smem[];
  const int N = a.subgrid_size, C = a.nr_channels, npix = N * N;
  const int s_local = blockIdx.x / slabs;
  const int slab = blockIdx.x - s_local * slabs;
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler too
  constexpr int NW = G8_PAIRS;                                // producer warps; warp NW builds B
  const int pix0 = slab * tiles_per_cta * 128;
  const int ntiles = min(tiles_per_cta, (npix - pix0 + 127) / 128);
  if (ntiles <= 0) return;
  const int npw = (ntiles + 1) >> 1;                          // producer warps that have a tile

  unsigned char *sA = smem;                                              // [warp][half][block][4 KB]
  unsigned char *sB = sA + G8_PAIRS * G8_A_WARP;                         // [G8_NB][512 B]
  unsigned long long *aempty = reinterpret_cast<unsigned long long *>(sB + G8_NB * G8_B_SLOT);  // [warp]
  unsigned long long *bfull = aempty + G8_PAIRS;                         // [G8_NB]
  unsigned long long *bempty = bfull + G8_NB;                            // [2] half rings
  unsigned long long *done = bempty + 2;
  unsigned *s_tmem = reinterpret_cast<unsigned *>(done + 1);
  float *s_red = reinterpret_cast<float *>(s_tmem + 2);     // [12] block reduction scratch + scale
  float *s_wn = s_red + 12;                                 // [ncb * 8], zero padded

  const SubgridCtx ctx = load_ctx(a, s);
  const int nt = ctx.nr_timesteps;
  const int ncb = (C + G8_CB - 1) / G8_CB;                  // even (regular_flag)
  const int nstages = nt * ncb;                             // B slots

  for (int c = tid; c < ncb * G8_CB; c += blockDim.x) s_wn[c] = c < C ? a.wavenumbers[c] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < G8_PAIRS; i++) mbar_init(&aempty[i], 1);
    for (int i = 0; i < G8_NB; i++) mbar_init(&bfull[i], 1);
    mbar_init(&bempty[0], npw);
    mbar_init(&bempty[1], npw);
    mbar_init(done, npw);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = *s_tmem;

  const float *g_uvw = reinterpret_cast<const float *>(a.uvw) + (size_t)ctx.time_offset * 3;
  const float2 *g_vis = a.visibilities + (size_t)ctx.time_offset * C * NR_POL;

  // power-of-two scaling of this subgrid's visibilities into fp16 range (as gridder_tc.cu)
  {
    float amax = 0.f;
    const float4 *v4 = reinterpret_cast<const float4 *>(g_vis);
    for (int i = tid; i < nt * C * 2; i += blockDim.x) {
      const float4 q = __ldg(&v4[i]);
      amax = fmaxf(fmaxf(amax, fmaxf(fabsf(q.x), fabsf(q.y))), fmaxf(fabsf(q.z), fabsf(q.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) s_red[warp] = amax;
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i <= NW; i++) amax = fmaxf(amax, s_red[i]);
      const unsigned eb = (__float_as_uint(amax) >> 23) & 0xffu;          // biased exponent
      const bool ok = eb >= 14u && eb <= 253u;                             // finite, not tiny
      s_red[10] = ok ? __uint_as_float((267u - eb) << 23) : 1.f;           // 2^(13 - E)
      s_red[11] = ok ? __uint_as_float((eb - 13u) << 23) : 1.f;            // 2^(E - 13)
    }
    __syncthreads();
  }
  const float vis_scale = s_red[10], vis_unscale = s_red[11];

  if (warp < NW) {
    // ------------------------------------------------------------------ producers (+ their own MMAs)
    const int tile0 = 2 * warp;
    if (tile0 < ntiles) {
      const bool has2 = tile0 + 1 < ntiles;
      const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      float l[8], m[8], n[8], off[8];
#pragma unroll
      for (int p = 0; p < 8; p++) {   // p = 4 h + j: pixel lane + 32 j of tile 2 w + h
        const int q = min(pix0 + (tile0 + (p >> 2)) * 128 + lane + 32 * (p & 3), npix - 1);
        const int y = q / N, x = q - y * N;
        l[p] = compute_l(x, N, a.image_size);
        m[p] = compute_l(y, N, a.image_size);
        n[p] = compute_n(l[p], m[p]);
        // gridder_reference.cpp:64 as the CPU binary contracts it
        off[p] = __fmaf_rn(ctx.w_offset, n[p], __fmaf_rn(ctx.u_offset, l[p], __fmul_rn(ctx.v_offset, m[p])));
      }
      unsigned char *A_warp = sA + warp * G8_A_WARP;
      unsigned long long da0 = smem_desc(smem_u32(A_warp), A_CHUNK_BYTES, 128);
      unsigned long long db0 = smem_desc(smem_u32(sB), B_CHUNK_BYTES, 128);
      unsigned tmem_d = tmem_base + tile0 * 16;
      unsigned my_empty_u = smem_u32(aempty + warp), bfull_u = smem_u32(bfull), bempty_u = smem_u32(bempty),
               done_u = smem_u32(done), wn_u = smem_u32(s_wn);
      asm volatile("" : "+l"(da0), "+l"(db0), "+r"(tmem_d), "+r"(my_empty_u), "+r"(bfull_u), "+r"(bempty_u), "+r"(done_u),
                   "+r"(wn_u));
      // the spacing (one for all blocks: regular_flag), from the first block as gridder_tc.cu computes it
      float dw0;
      linear_channels(s_wn, 0, min(G8_CB, C), &dw0);
      float un = 0.f, vn = 0.f, wnx = 0.f;   // uvw of the next timestep, fetched one timestep ahead
      if (nt > 0) { un = __ldg(&g_uvw[0]); vn = __ldg(&g_uvw[1]); wnx = __ldg(&g_uvw[2]); }
      unsigned k = 0, sk = 0;                 // B slots and stages done
      unsigned slot2 = 0, ring_phase = 0;     // B slot pair of the stage (k % 16), lap parity of the ring
      const unsigned last_k = (unsigned)nstages - 2u;
      for (int t = 0; t < nt; t++) {
        const float u = un, v = vn, w = wnx;
        if (t + 1 < nt) { un = __ldg(&g_uvw[3 * t + 3]); vn = __ldg(&g_uvw[3 * t + 4]); wnx = __ldg(&g_uvw[3 * t + 5]); }
        float idx[8];
        float2 rot[8];
#pragma unroll
        for (int p = 0; p < 8; p++) {  // gridder_reference.cpp:61 as contracted by the CPU binary
          idx[p] = __fmaf_rn(w, n[p], __fmaf_rn(u, l[p], __fmul_rn(v, m[p])));
          rot[p] = phasor<IDGB200_SINCOS_FAST>(__fmul_rn(-idx[p], dw0));
        }
        for (int cb0 = 0; cb0 < ncb; cb0 += 2, sk++, k += 2) {
          if (sk >= 1) mbar_wait_u(my_empty_u, (sk - 1) & 1);
          float wn0a, wn0b;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wn0a) : "r"(wn_u + (unsigned)cb0 * (G8_CB * 4)));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wn0b) : "r"(wn_u + (unsigned)cb0 * (G8_CB * 4) + G8_CB * 4));
          g8_produce(A_warp, wn0a, rot, idx, off, lane);                        // tile 0, block a
          g8_produce(A_warp + G8_A_BUF, wn0b, rot, idx, off, lane);             // tile 0, block b
          g8_produce(A_warp + 2 * G8_A_BUF, wn0a, rot + 4, idx + 4, off + 4, lane);   // tile 1, block a
          g8_produce(A_warp + 3 * G8_A_BUF, wn0b, rot + 4, idx + 4, off + 4, lane);   // tile 1, block b
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_wait_u(bfull_u + slot2 * 8, ring_phase);
          mbar_wait_u(bfull_u + slot2 * 8 + 8, ring_phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          __syncwarp();
          if (elect_one()) {
            const unsigned long long dba = db0 + (unsigned long long)(slot2 * (G8_B_SLOT >> 4));
            const unsigned long long dbb = dba + (unsigned long long)(G8_B_SLOT >> 4);
            const unsigned acc = k > 0 ? 1u : 0u;
            umma_f16(tmem_d, da0, dba, idesc, acc);
            umma_f16(tmem_d, da0 + (unsigned long long)(G8_A_BUF >> 4), dbb, idesc, 1u);
            if (has2) {
              umma_f16(tmem_d + 16, da0 + (unsigned long long)(2 * G8_A_BUF >> 4), dba, idesc, acc);
              umma_f16(tmem_d + 16, da0 + (unsigned long long)(3 * G8_A_BUF >> 4), dbb, idesc, 1u);
            }
            umma_commit_u(my_empty_u);
            if ((slot2 & 7u) == 6u) umma_commit_u(bempty_u + (slot2 >> 3) * 8);   // half ring consumed
            if (k == last_k) umma_commit_u(done_u);
          }
          __syncwarp();
          slot2 += 2;
          if (slot2 == (unsigned)G8_NB) { slot2 = 0; ring_phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ B builder warp (as gridder_tc.cu)
    // lane = (kc, n): one 16-byte chunk = 4 channels x (cos-row, sin-row) of column n = (hi|lo, pol, re|im)
    const int nrow = lane & 15, kc = lane >> 4, lo = nrow >> 3, p = (nrow >> 1) & 3, im = nrow & 1;
    auto load_b = [&](int kk, float2 (&raw)[4]) {
      const int t = kk / ncb, cb = kk - t * ncb;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int c = cb * G8_CB + kc * 4 + i;
        raw[i] = c < C ? __ldg(&g_vis[((size_t)t * C + c) * NR_POL + p]) : make_float2(0.f, 0.f);
      }
    };
    float2 raw[4];
    if (nstages > 0) load_b(0, raw);
    for (int k = 0; k < nstages; k++) {
      const int slot = k % G8_NB;
      if ((k & 7) == 0 && k >= G8_NB) mbar_wait(&bempty[(k >> 3) & 1], ((k / G8_NB) - 1) & 1);
      unsigned pk[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 vv = raw[i];
        const float x0 = (im ? vv.y : vv.x) * vis_scale;    // multiplies cos
        const float x1 = (im ? vv.x : -vv.y) * vis_scale;   // multiplies sin
        __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
        if (lo) {
          h0 = __float2half_rn(x0 - __half2float(h0));
          h1 = __float2half_rn(x1 - __half2float(h1));
        }
        pk[i] = (unsigned)__half_as_ushort(h0) | ((unsigned)__half_as_ushort(h1) << 16);
      }
      if (k + 1 < nstages) load_b(k + 1, raw);
      *reinterpret_cast<uint4 *>(sB + slot * G8_B_SLOT + kc * B_CHUNK_BYTES + nrow * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bfull[slot]);
    }
  }

  // ---- epilogue (producer warps): accumulators -> A-terms, taper, store (gridder_reference.cpp:84-110)
  if (warp < NW) {
    if (nstages > 0) {
      mbar_wait(done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const size_t plane = (size_t)npix;
    const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
    const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
    float2 *out = const_cast<float2 *>(a.subgrids) + (size_t)s * NR_POL * plane;
    // warp w reads TMEM lanes 32 w .. 32 w + 31 of every tile
    for (int tile = 0; tile < ntiles; tile++) {
      unsigned r[16];
      if (nstages > 0) {
        const unsigned taddr = tmem_base + ((unsigned)(warp * 32) << 16) + tile * 16;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) r[i] = 0u;
      }
      const int pixel = pix0 + tile * 128 + warp * 32 + lane;
      if (pixel < npix) {
        float2 px[NR_POL];
#pragma unroll
        for (int p = 0; p < NR_POL; p++)
          px[p] = make_float2((__uint_as_float(r[2 * p]) + __uint_as_float(r[8 + 2 * p])) * vis_unscale,
                              (__uint_as_float(r[2 * p + 1]) + __uint_as_float(r[8 + 2 * p + 1])) * vis_unscale);
        float2 a1[4], a2[4];
        load_jones(a.aterms, (at1 + pixel) * NR_POL, a1);
        load_jones(a.aterms, (at2 + pixel) * NR_POL, a2);
        apply_aterm_gridder(px, a1, a2);
        const float sph = __ldg(&a.spheroidal[pixel]);
        const int dst = subgrid_slot(pixel, a.subgrid_size, a.flags);
#pragma unroll
        for (int p = 0; p < NR_POL; p++)
          out[p * plane + dst] = make_float2(__fmul_rn(px[p].x, sph), __fmul_rn(px[p].y, sph));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}

}  // namespace

cudaError_t launch_gridder_regular_check(const KernelArgs &a, int *d_flag, cudaStream_t stream, int *d_zero2) {
  gridder_regular_check_kernel<<<1, 32, 0, stream>>>(a.wavenumbers, a.nr_channels, d_flag, d_zero2);
  return cudaGetLastError();
}

// regular_flag: written by launch_gridder_regular_check on the same stream; the kernel is a no-op when 0
cudaError_t launch_gridder_tc8(const KernelArgs &a, const int *d_regular_flag, cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  const int npix = a.subgrid_size * a.subgrid_size;
  const int tiles_total = (npix + 127) / 128;
  const int slabs = (tiles_total + G8_TILES - 1) / G8_TILES;
  const int tiles_per_cta = tiles_total > 64 ? G8_TILES : (tiles_total + slabs - 1) / slabs;
  const int nslabs = (tiles_total + tiles_per_cta - 1) / tiles_per_cta;
  int tmem_cols = 32;
  while (tmem_cols < tiles_per_cta * 16) tmem_cols *= 2;
  const int ncb = (a.nr_channels + G8_CB - 1) / G8_CB;
  const size_t smem = (size_t)G8_PAIRS * G8_A_WARP + G8_NB * G8_B_SLOT + (G8_PAIRS + G8_NB + 3) * 8 + 8 + 48 +
                      (size_t)ncb * G8_CB * 4;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  auto k = gridder_tc8_kernel;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<dim3((unsigned)a.nr_subgrids * nslabs), dim3(G8_THREADS), smem, stream>>>(a, nslabs, tiles_per_cta, tmem_cols,
                                                                                 d_regular_flag);
  return cudaGetLastError();
}

}  // namespace idgb200
