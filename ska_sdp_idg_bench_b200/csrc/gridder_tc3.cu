// IDG gridder on tcgen05 with the phasor operand written straight from registers to TMEM
// (tcgen05.st + the A-from-TMEM form of tcgen05.mma): the phasor tile never touches shared memory.
//
// Same GEMM as gridder_tc.cu (D[pixel][n] += A[pixel][k] B[k][n], n = (hi|lo, pol, re|im), B = the
// visibilities in fp16 hi + lo through a shared-memory ring), but
//   * TMEM lane = tile row = pixel = thread: a tile is produced by the four warps whose
//     warp % 4 gives them access to its four lane quadrants; a thread has ONE pixel, makes the
//     8 phasors of a stage (timestep, 8 channels), tcgen05.st's them as packed half2 columns into
//     one of NBUF buffers; the four warps meet at a named barrier and the warp of quadrant 0
//     issues the stage's MMAs;
//   * SPLIT: the phasor is stored as fp16 hi + fp16 lo (two K=16 blocks per stage against the
//     same B slot), so the operand keeps ~22 bits and the result is FP32-class; without SPLIT it
//     is the fp16 phasor of gridder_tc.cu;
//   * the epilogue needs no redistribution: every thread reads its own pixel's accumulators.
// TMEM per tile: D 16 columns + 4 buffers x 8 columns (fp16) or 3 buffers x 16 columns (hi + lo) of A.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace idgb200 {

namespace {

constexpr int G3_TILES = 4;                       // tiles (128 pixels) per CTA
constexpr int G3_WARPS = 4 * G3_TILES;            // producer warps
constexpr int G3_THREADS = (G3_WARPS + 1) * 32;   // + B builder warp
constexpr int G3_CB = 8;                          // channels per stage
constexpr int G3_B_SLOT = 2 * B_CHUNK_BYTES;      // 512 B
constexpr int G3_NB = 16;                         // B ring slots

__device__ __forceinline__ void umma_f16_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long db, unsigned idesc,
                                            unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const unsigned (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// the 8 phasors of one stage of one pixel as packed half2 (cos, sin): hi, and with SPLIT the
// rounding residual lo.  MASK8 bit i set -> channel i from phasor_poly instead of MUFU.
template <unsigned MASK8, bool SPLIT>
__device__ __forceinline__ void tc3_produce(const float (&wn)[8], const float idx, const float idxr, const float off,
                                            const float offr, unsigned (&hi)[8], unsigned (&lo)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float2 ph = ((MASK8 >> i) & 1u) ? phasor_poly(__fmaf_rn(-idxr, wn[i], offr))
                                          : phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(-idx, wn[i], off));  // :69
    const __half2 hh = __floats2half2_rn(ph.x, ph.y);
    hi[i] = *reinterpret_cast<const unsigned *>(&hh);
    if (SPLIT) {
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(__fsub_rn(ph.x, hf.x), __fsub_rn(ph.y, hf.y));
      lo[i] = *reinterpret_cast<const unsigned *>(&ll);
    } else {
      lo[i] = 0u;
    }
  }
}

// MASK16: bit c set -> channel c of every 16-channel group uses phasor_poly instead of MUFU
template <unsigned MASK16, bool SPLIT>
__global__ void __launch_bounds__(G3_THREADS, 2)
gridder_tc3_kernel(const KernelArgs a, const int slabs, const int tiles_per_cta) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int ACOLS = SPLIT ? 16 : 8;                 // TMEM columns of one A buffer
  constexpr int NBUF = SPLIT ? 3 : 4;                   // A buffers per tile
  constexpr int TCOLS = 16 + NBUF * ACOLS;              // per tile: D + the A buffers (48 / 64 columns)
  constexpr int TMEM_COLS = G3_TILES * TCOLS <= 128 ? 128 : 256;
  static_assert(G3_TILES * TCOLS <= 256, "two CTAs per SM share the 512 TMEM columns");
  // shared-memory carve-up (byte offsets, so that every access keeps the shared state space)
  constexpr int OFF_AFULL = G3_NB * G3_B_SLOT;                  // [tile][NBUF] mbarriers, 4 arrivals each
  constexpr int OFF_AEMPTY = OFF_AFULL + G3_TILES * NBUF * 8;   // [tile][NBUF] tcgen05.commit targets
  constexpr int OFF_BFULL = OFF_AEMPTY + G3_TILES * NBUF * 8;   // [G3_NB]
  constexpr int OFF_BEMPTY = OFF_BFULL + G3_NB * 8;             // [2] half rings
  constexpr int OFF_DONE = OFF_BEMPTY + 2 * 8;
  constexpr int OFF_TMEM = OFF_DONE + 8;
  constexpr int OFF_RED = OFF_TMEM + 8;                         // [20] floats
  constexpr int OFF_WN = (OFF_RED + 20 * 4 + 15) & ~15;         // [ncb * 8] floats, zero padded
  const int N = a.subgrid_size, C = a.nr_channels, npix = N * N;
  const int s_local = blockIdx.x / slabs;
  const int slab = blockIdx.x - s_local * slabs;
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler too
  const int pix0 = slab * tiles_per_cta * 128;
  const int ntiles = min(tiles_per_cta, (npix - pix0 + 127) / 128);
  if (ntiles <= 0) return;

  unsigned char *sB = smem;                                                       // [G3_NB][512 B]
  unsigned long long *afull = reinterpret_cast<unsigned long long *>(smem + OFF_AFULL);
  unsigned long long *aempty = reinterpret_cast<unsigned long long *>(smem + OFF_AEMPTY);
  unsigned long long *bfull = reinterpret_cast<unsigned long long *>(smem + OFF_BFULL);
  unsigned long long *bempty = reinterpret_cast<unsigned long long *>(smem + OFF_BEMPTY);
  unsigned long long *done = reinterpret_cast<unsigned long long *>(smem + OFF_DONE);
  unsigned *s_tmem = reinterpret_cast<unsigned *>(smem + OFF_TMEM);
  float *s_red = reinterpret_cast<float *>(smem + OFF_RED);
  float *s_wn = reinterpret_cast<float *>(smem + OFF_WN);

  const SubgridCtx ctx = load_ctx(a, s);
  const int nt = ctx.nr_timesteps;
  const int ncb = (C + G3_CB - 1) / G3_CB;
  const int nstages = nt * ncb;

  for (int c = tid; c < ncb * G3_CB; c += G3_THREADS) s_wn[c] = c < C ? a.wavenumbers[c] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < G3_TILES * NBUF; i++) { mbar_init(&afull[i], 4); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < G3_NB; i++) mbar_init(&bfull[i], 1);
    mbar_init(&bempty[0], ntiles);
    mbar_init(&bempty[1], ntiles);
    mbar_init(done, ntiles);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = *s_tmem;

  const float *g_uvw = reinterpret_cast<const float *>(a.uvw) + (size_t)ctx.time_offset * 3;
  const float2 *g_vis = a.visibilities + (size_t)ctx.time_offset * C * NR_POL;

  // power-of-two scaling of this subgrid's visibilities into fp16 range (gridder_tc.cu)
  {
    float amax = 0.f;
    const float4 *v4 = reinterpret_cast<const float4 *>(g_vis);
    for (int i = tid; i < nt * C * 2; i += G3_THREADS) {
      const float4 q = __ldg(&v4[i]);
      amax = fmaxf(fmaxf(amax, fmaxf(fabsf(q.x), fabsf(q.y))), fmaxf(fabsf(q.z), fabsf(q.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) s_red[warp] = amax;
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i <= G3_WARPS; i++) amax = fmaxf(amax, s_red[i]);
      const unsigned eb = (__float_as_uint(amax) >> 23) & 0xffu;          // biased exponent
      const bool ok = eb >= 14u && eb <= 253u;                             // finite, not tiny
      s_red[18] = ok ? __uint_as_float((267u - eb) << 23) : 1.f;           // 2^(13 - E)
      s_red[19] = ok ? __uint_as_float((eb - 13u) << 23) : 1.f;            // 2^(E - 13)
    }
    __syncthreads();
  }
  const float vis_scale = s_red[18], vis_unscale = s_red[19];

  const int tile = warp >> 2, q4 = warp & 3;                  // producer warps: tile, TMEM lane quadrant
  const unsigned lane_base = (unsigned)(q4 * 32) << 16;
  const unsigned tmem_tile = tmem_base + tile * TCOLS;       // D at +0, A buffer b at +16 + b * ACOLS

  if (warp < G3_WARPS) {
    if (tile < ntiles) {
      // ---------------------------------------------------------------- producers
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6) = 1, A = B = F16 (0),
      // both K-major (0), N >> 3 at [17,23), M >> 4 at [24,29)
      const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      const int pixel = pix0 + tile * 128 + q4 * 32 + lane;
      const int q = min(pixel, npix - 1);
      const int y = q / N, x = q - y * N;
      const float l = compute_l(x, N, a.image_size), m = compute_l(y, N, a.image_size), n = compute_n(l, m);
      // gridder_reference.cpp:64 as the CPU binary contracts it
      const float off = __fmaf_rn(ctx.w_offset, n, __fmaf_rn(ctx.u_offset, l, __fmul_rn(ctx.v_offset, m)));
      const float offr = __fmul_rn(off, 0.15915494309189535f);
      const unsigned long long db0 = smem_desc(smem_u32(sB), B_CHUNK_BYTES, 128);
      unsigned long long *my_empty = aempty + tile * NBUF;

      // The four warps of a tile meet at a named hardware barrier once per stage (a waiting warp
      // costs no issue slots, unlike an mbarrier spin); the warp of quadrant 0 then issues the
      // stage's MMAs, so the sums are accumulated in stage order (deterministic).
      auto issue = [&](int kk) {
        const int pb = kk % NBUF, slot = kk % G3_NB;
        mbar_wait(&bfull[slot], (kk / G3_NB) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const unsigned long long db = db0 + (unsigned long long)(slot * (G3_B_SLOT >> 4));
          const unsigned a_t = tmem_tile + 16 + pb * ACOLS;
          umma_f16_ts(tmem_tile, a_t, db, idesc, kk > 0 ? 1u : 0u);
          if (SPLIT) umma_f16_ts(tmem_tile, a_t + 8, db, idesc, 1u);
          umma_commit(&my_empty[pb]);
          if ((kk & 7) == 7) umma_commit(&bempty[(kk >> 3) & 1]);
          if (kk == nstages - 1) umma_commit(done);
        }
        __syncwarp();
      };

      float un = 0.f, vn = 0.f, wnx = 0.f;   // uvw of the next timestep, fetched one timestep ahead
      if (nt > 0) { un = __ldg(&g_uvw[0]); vn = __ldg(&g_uvw[1]); wnx = __ldg(&g_uvw[2]); }
      int k = 0;
      for (int t = 0; t < nt; t++) {
        const float u = un, v = vn, w = wnx;
        if (t + 1 < nt) { un = __ldg(&g_uvw[3 * t + 3]); vn = __ldg(&g_uvw[3 * t + 4]); wnx = __ldg(&g_uvw[3 * t + 5]); }
        // gridder_reference.cpp:61 as contracted by the CPU binary
        const float idx = __fmaf_rn(w, n, __fmaf_rn(u, l, __fmul_rn(v, m)));
        const float idxr = __fmul_rn(idx, 0.15915494309189535f);   // in revolutions, for phasor_poly
        for (int cb = 0; cb < ncb; cb++, k++) {
          const int buf = k % NBUF, use = k / NBUF;
          const float4 w0 = *reinterpret_cast<const float4 *>(s_wn + cb * G3_CB);
          const float4 w1 = *reinterpret_cast<const float4 *>(s_wn + cb * G3_CB + 4);
          const float wn[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          unsigned hi[8], lo[8];
          if ((MASK16 >> 8) == (MASK16 & 0xffu) || !(cb & 1))
            tc3_produce<(MASK16 & 0xffu), SPLIT>(wn, idx, idxr, off, offr, hi, lo);
          else
            tc3_produce<(MASK16 >> 8), SPLIT>(wn, idx, idxr, off, offr, hi, lo);
          if (use >= 1) mbar_wait(&my_empty[buf], (use - 1) & 1);   // the MMAs of stage k - NBUF have read this buffer
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const unsigned ta = tmem_tile + lane_base + 16 + buf * ACOLS;
          tmem_st8(ta, hi);
          if (SPLIT) tmem_st8(ta + 8, lo);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1 + tile) : "memory");
          if (q4 == 0) issue(k);
        }
      }

      // ---- epilogue: every thread owns its pixel's accumulators (gridder_reference.cpp:84-110)
      unsigned r[16];
      if (nstages > 0) {
        mbar_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem_tile + lane_base));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) r[i] = 0u;
      }
      if (pixel < npix) {
        const size_t plane = (size_t)npix;
        const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
        const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
        float2 *out = const_cast<float2 *>(a.subgrids) + (size_t)s * NR_POL * plane;
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
  } else {
    // ------------------------------------------------------------------ B builder warp (as gridder_tc.cu)
    const int nrow = lane & 15, kc = lane >> 4, lo = nrow >> 3, p = (nrow >> 1) & 3, im = nrow & 1;
    auto load_b = [&](int kk, float2 (&raw)[4]) {
      const int t = kk / ncb, cb = kk - t * ncb;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int c = cb * G3_CB + kc * 4 + i;
        raw[i] = c < C ? __ldg(&g_vis[((size_t)t * C + c) * NR_POL + p]) : make_float2(0.f, 0.f);
      }
    };
    float2 raw[4];
    if (nstages > 0) load_b(0, raw);
    for (int k = 0; k < nstages; k++) {
      const int slot = k % G3_NB;
      if ((k & 7) == 0 && k >= G3_NB) mbar_wait(&bempty[(k >> 3) & 1], ((k / G3_NB) - 1) & 1);
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
      *reinterpret_cast<uint4 *>(sB + slot * G3_B_SLOT + kc * B_CHUNK_BYTES + nrow * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bfull[slot]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

}  // namespace

// mode: 0 = fp16 phasors, MUFU only; 1 = fp16 phasors, 4/16 polynomial;
//       2 = fp16 hi + lo phasors, MUFU only; 3 = hi + lo, 2/16 polynomial
cudaError_t launch_gridder_tc3(const KernelArgs &a, int mode, cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  const int npix = a.subgrid_size * a.subgrid_size;
  const int tiles_total = (npix + 127) / 128;
  const int slabs = (tiles_total + G3_TILES - 1) / G3_TILES;
  const int tiles_per_cta = G3_TILES;
  const int ncb = (a.nr_channels + G3_CB - 1) / G3_CB;
  const size_t smem = (size_t)G3_NB * G3_B_SLOT + (G3_TILES * 4 * 2 + G3_NB + 3) * 8 + 8 + 20 * 4 + 16 +
                      (size_t)ncb * G3_CB * 4 + 64;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  void (*k)(const KernelArgs, int, int) = nullptr;
  switch (mode) {
    case 0: k = gridder_tc3_kernel<0x0000u, false>; break;
    case 1: k = gridder_tc3_kernel<0x4444u, false>; break;
    case 2: k = gridder_tc3_kernel<0x0000u, true>; break;
    case 3: k = gridder_tc3_kernel<0x4040u, true>; break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<dim3((unsigned)a.nr_subgrids * slabs), dim3(G3_THREADS), smem, stream>>>(a, slabs, tiles_per_cta);
  return cudaGetLastError();
}

}  // namespace idgb200
