// IDG degridder on tcgen05 + TMEM, second generation: a producer warp owns TWO M-tiles (variant 24,
// the default for FAST sincos when the channel count is a multiple of 8).
//
// Same GEMM as degridder_tc.cu (D[vis][n] += A[vis][k] B[k][n]; A = fp16 hi + lo phasors made by the
// warp, B = P' = A1 (sph . subgrid) A2^H split into fp16 hi + lo by the builder warp, the phase in the
// CPU binary's operation order), but the rows are enumerated in OCTS = (timestep, 8 channels): lane l
// of producer warp w owns oct 32 (4 r + w) + l of round r, its channels 0-3 are rows l + 32 j of the
// warp's first tile, channels 4-7 the same rows of its second tile.  What that buys over the quads of
// degridder_tc.cu (ncu source page of both, profiles/):
//   * per pixel one phase index, one sincos for the first channel and one for the spacing serve 8
//     items instead of 4, and 6 of the 8 phasors come from the three-term recurrence (one FFMA2);
//   * a stage is 8 pixels x 8 channels = 64 items per thread: the per-stage bookkeeping (ring and
//     barrier state, fences, the elected MMA issue, ~100 instructions) is paid half as often per item;
//   * 5 warps per CTA leave 136 registers per thread, so the 64 packed words of a 4-pixel chunk
//     (hi + lo for 8 rows) stay in registers and the eight independent pixel chains interleave.
// Per stage the elected lane issues four MMAs (M=128, N=16, K=16: two tiles x hi, lo) against the same
// B slot and commits them to the warp's one empty barrier.  Ring, builder warp, pixel scaling and the
// epilogue are those of degridder_tc.cu.  FAST sincos only; channel counts that are not a multiple of
// 8 stay on degridder_tc.cu (variant 22).
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace idgb200 {

namespace {

constexpr int E_PAIRS = 4;                       // producer warps per CTA, two M-tiles each
constexpr int E_TILES = 2 * E_PAIRS;
constexpr int E_PB = 8;                          // pixels per stage -> K = 16
constexpr int E_A_BUF = 2 * A_CHUNK_BYTES;       // 4 KB: one tile, one stage, hi or lo
constexpr int E_A_WARP = 4 * E_A_BUF;            // [tile half][hi | lo]
constexpr int E_GROUP = 4;                       // stages per ring group (32 pixels = one builder pass)
constexpr int E_NG = 3;                          // ring groups
constexpr int E_B_SLOT = 2 * B_CHUNK_BYTES;      // 512 B of B per stage
constexpr int E_G_SLOT = E_PB * 16;              // 128 B of (l, m, n, offset) per stage
constexpr int E_THREADS = (E_PAIRS + 1) * 32;
constexpr int E_TMEM_COLS = E_TILES * 16;        // 128 columns

// P' of pixel q (degridder_reference.cpp:38-74): taper, A1 . P . A2^H
__device__ __forceinline__ void pixel_after_aterms8(const KernelArgs &a, const float2 *sub, size_t plane, size_t at1,
                                                    size_t at2, int q, float2 (&px)[NR_POL]) {
  const float sph = __ldg(&a.spheroidal[q]);
  const int src = subgrid_slot(q, a.subgrid_size, a.flags);
#pragma unroll
  for (int p = 0; p < NR_POL; p++) {
    const float2 v = __ldg(&sub[p * plane + src]);
    px[p] = make_float2(__fmul_rn(sph, v.x), __fmul_rn(sph, v.y));
  }
  float2 a1[4], a2[4];
  load_jones(a.aterms, (at1 + q) * NR_POL, a1);
  load_jones(a.aterms, (at2 + q) * NR_POL, a2);
  apply_aterm_degridder(px, a1, a2);
}

// the packed words of one 4-pixel chunk: channel c of the oct -> tile half c >> 2, row lane + 32 (c & 3)
__device__ __forceinline__ void store_chunk(unsigned char *A, int kc, int lane, const unsigned (&pk)[8][4],
                                            const unsigned (&pl)[8][4]) {
#pragma unroll
  for (int c = 0; c < 8; c++) {
    unsigned char *row = A + (c >> 2) * (2 * E_A_BUF) + kc * A_CHUNK_BYTES + (lane + 32 * (c & 3)) * 16;
    *reinterpret_cast<uint4 *>(row) = make_uint4(pk[c][0], pk[c][1], pk[c][2], pk[c][3]);
    *reinterpret_cast<uint4 *>(row + E_A_BUF) = make_uint4(pl[c][0], pl[c][1], pl[c][2], pl[c][3]);
  }
}

// one stage of one thread, 8 pixels x 8 equally spaced channels (common.cuh: linear_channels): per pixel
// the first channel's sincos (the reference's angle, bit for bit), the spacing's e^{i idx dw}, one
// complex multiplication and six steps of ph[c+1] = 2 cos(delta) ph[c] - ph[c-1] (gridder_tc.cu:
// tc_produce_linear has the error argument: <= 21 roundings after 7 channels, ~1e-6, against the
// ~1e-4 rad the FAST sincos itself is off at these phases)
__device__ __forceinline__ void etc_produce_linear(unsigned char *A, const float4 *geo, const float u, const float v,
                                                   const float w, const float wn0, const float dw, const int lane) {
  // phase A, all 8 pixels of the stage: geometry, phase index, the two sincos, the first rotation - 8
  // independent chains cover the MUFU and LDS latencies with only ~4 warps per sub-partition resident
  float2 prev[8], cur[8], cc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float4 g = geo[i];   // broadcast: all lanes read the same 16 bytes
    // degridder_reference.cpp:106 as the CPU binary evaluates it (w term unfused)
    const float idx = __fadd_rn(__fmaf_rn(u, g.x, __fmul_rn(v, g.y)), __fmul_rn(w, g.z));
    prev[i] = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(idx, wn0, -g.w));   // :112, (cos, sin)
    const float2 d = phasor<IDGB200_SINCOS_FAST>(__fmul_rn(idx, dw));
    const float2 dxx = make_float2(d.x, d.x), dny = make_float2(-d.y, d.y);
    const float c2 = __fadd_rn(d.x, d.x);
    cc[i] = make_float2(c2, c2);
    cur[i] = ffma2(make_float2(prev[i].y, prev[i].x), dny, __fmul2_rn(prev[i], dxx));
  }
  // phase B, channel-major over each chunk of 4 pixels: every recurrence step is 4 independent FFMA2 and
  // the 4 packed words of a row (one 16-byte store) are produced back to back
#pragma unroll
  for (int kc = 0; kc < 2; kc++) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
      unsigned hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int q = kc * 4 + i;
        if (c == 0) {
          pack_phasor<true>(prev[q], hi[i], lo[i]);
        } else if (c == 1) {
          pack_phasor<true>(cur[q], hi[i], lo[i]);
        } else {
          const float2 nxt = ffma2(cur[q], cc[q], make_float2(-prev[q].x, -prev[q].y));
          pack_phasor<true>(nxt, hi[i], lo[i]);
          prev[q] = cur[q];
          cur[q] = nxt;
        }
      }
      if ((IDGB200_ABLATE & 2) && !ablate_never()) continue;
      unsigned char *row = A + (c >> 2) * (2 * E_A_BUF) + kc * A_CHUNK_BYTES + (lane + 32 * (c & 3)) * 16;
      *reinterpret_cast<uint4 *>(row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4 *>(row + E_A_BUF) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// arbitrary wavenumbers: one sincos per channel (wn8: the oct's 8 wavenumbers in global memory)
__device__ __forceinline__ void etc_produce(unsigned char *A, const float4 *geo, const float u, const float v,
                                            const float w, const float *__restrict__ wn8, const int lane) {
#pragma unroll
  for (int kc = 0; kc < 2; kc++) {
    unsigned pk[8][4], pl[8][4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float4 g = geo[kc * 4 + i];
      const float idx = __fadd_rn(__fmaf_rn(u, g.x, __fmul_rn(v, g.y)), __fmul_rn(w, g.z));
#pragma unroll
      for (int c = 0; c < 8; c++)
        pack_phasor<true>(phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(idx, __ldg(&wn8[c]), -g.w)), pk[c][i], pl[c][i]);
    }
    store_chunk(A, kc, lane, pk, pl);
  }
}

template <bool LIST>
__device__ __forceinline__ void degridder_tc8_body(const KernelArgs &a, const int recur, const int fold_ok, const int s_local) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int N = a.subgrid_size, C = a.nr_channels, npix = N * N;
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler too
  constexpr int NW = E_PAIRS;                                 // producer warps; warp NW builds B

  const SubgridCtx ctx = load_ctx(a, s);
  const int nt = ctx.nr_timesteps;
  const int nco = C >> 3;                                     // channel groups of 8 (C % 8 == 0)
  const int octs = nt * nco;
  if (octs == 0) return;
  // the subgrid's tile pairs are processed in rounds of <= 4 (one per producer warp), balanced
  const int pairs_sub = (octs + 31) >> 5;
  const int rounds = (pairs_sub + E_PAIRS - 1) / E_PAIRS;
  const int ppr = (pairs_sub + rounds - 1) / rounds;

  unsigned char *sA = smem;                                                      // [warp][half][hi|lo][4 KB]
  unsigned char *sB = sA + E_PAIRS * E_A_WARP;                                   // [NG * GROUP][512 B]
  float4 *sG = reinterpret_cast<float4 *>(sB + E_NG * E_GROUP * E_B_SLOT);       // [NG * GROUP][8] (l, m, n, off)
  float *scratch = reinterpret_cast<float *>(sG + E_NG * E_GROUP * E_PB);        // [2][32][8] builder scratch
  unsigned long long *aempty = reinterpret_cast<unsigned long long *>(scratch + 2 * 32 * 8);  // [warp]
  unsigned long long *bfull = aempty + E_PAIRS;                                  // [NG]
  unsigned long long *bempty = bfull + E_NG;                                     // [NG]
  unsigned long long *done = bempty + E_NG;
  unsigned *s_tmem = reinterpret_cast<unsigned *>(done + 1);
  float *s_red = reinterpret_cast<float *>(s_tmem + 2);     // [12] block reduction scratch + scale

  if (tid == 0) {
    for (int i = 0; i < E_PAIRS; i++) mbar_init(&aempty[i], 1);
    for (int i = 0; i < E_NG; i++) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], ppr);
    }
    mbar_init(done, ppr);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(E_TMEM_COLS));
    if (!LIST) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);   // a list-mode CTA allocates again
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = *s_tmem;

  const size_t plane = (size_t)npix;
  const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
  const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
  const float2 *sub = a.subgrids + (size_t)s * NR_POL * plane;
  const float *g_uvw = reinterpret_cast<const float *>(a.uvw) + (size_t)ctx.time_offset * 3;

  // Planar subgrids (w = 0 for every timestep and no w offset - the reference's own synthetic
  // observation, init.cpp:4-25, and any snapshot of a coplanar array): pixel q and its mirror image
  // q' = npix - 1 - q have l, m negated, so their phases are exact negatives of each other in the
  // reference's own fp32 arithmetic (every operation of degridder_reference.cpp:96-112 is odd in
  // (l, m) once the w terms are +-0) and the phasor of q' is the conjugate of q's, bit for bit.  Then
  //   P[q] ph + P[q'] conj(ph) = (P[q] + P[q']) cos + i (P[q] - P[q']) sin,
  // i.e. the same GEMM over HALF the pixels with B built from the sum and the difference of the
  // pair: half the phasors, half the operand stores, half the MMAs.  Checked per subgrid; anything
  // else (w != 0, odd subgrid size) takes the full sum.
  int planar = fold_ok && ctx.w_offset == 0.f && !(N & 1);
  for (int t = tid; t < nt; t += E_THREADS) planar &= __ldg(&g_uvw[3 * t + 2]) == 0.f;
  const bool fold = __syncthreads_and(planar) != 0;
  const int npix_k = fold ? npix / 2 : npix;                  // pixels (pairs) along K
  const int nstages = (npix_k + E_PB - 1) / E_PB;
  const int ngroups = (nstages + E_GROUP - 1) / E_GROUP;

  // fp16 has 5 exponent bits: scale P' (or the pair's sum and difference) by a power of two so that
  // the largest component lands in [2^13, 2^14) (exact; undone in the epilogue), as in degridder_tc.cu
  {
    float amax = 0.f;
    for (int q = tid; q < npix_k; q += E_THREADS) {
      float2 px[NR_POL];
      pixel_after_aterms8(a, sub, plane, at1, at2, q, px);
      if (fold) {
        float2 py[NR_POL];
        pixel_after_aterms8(a, sub, plane, at1, at2, npix - 1 - q, py);
#pragma unroll
        for (int p = 0; p < NR_POL; p++) {
          amax = fmaxf(amax, fmaxf(fabsf(px[p].x + py[p].x), fabsf(px[p].y + py[p].y)));
          amax = fmaxf(amax, fmaxf(fabsf(px[p].x - py[p].x), fabsf(px[p].y - py[p].y)));
        }
      } else {
#pragma unroll
        for (int p = 0; p < NR_POL; p++) amax = fmaxf(amax, fmaxf(fabsf(px[p].x), fabsf(px[p].y)));
      }
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
  const float pix_scale = s_red[10], pix_unscale = s_red[11];

  if (warp < NW) {
    // ------------------------------------------------------------------ producers (+ their own MMAs)
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6) = 1, A = B = F16 (0),
    // both K-major (0), N >> 3 at [17,23), M >> 4 at [24,29)
    const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    unsigned char *A_warp = sA + warp * E_A_WARP;
    unsigned long long da0 = smem_desc(smem_u32(A_warp), A_CHUNK_BYTES, 128);
    unsigned long long db0 = smem_desc(smem_u32(sB), B_CHUNK_BYTES, 128);
    unsigned tmem_d = tmem_base + warp * 32;     // the warp's two accumulators: 16 columns each
    // barrier addresses as 32-bit shared addresses, computed once and made opaque so that the
    // compiler keeps them instead of rematerialising the address arithmetic in every issue path
    unsigned my_empty_u = smem_u32(aempty + warp), bfull_u = smem_u32(bfull), bempty_u = smem_u32(bempty),
             done_u = smem_u32(done), geo_u = smem_u32(sG);
    asm volatile("" : "+l"(da0), "+l"(db0), "+r"(tmem_d), "+r"(my_empty_u), "+r"(bfull_u), "+r"(bempty_u), "+r"(done_u),
                 "+r"(geo_u));
    float4 *g_vis = reinterpret_cast<float4 *>(const_cast<float2 *>(a.visibilities)) +
                    (size_t)ctx.time_offset * C * 2;
    // per-warp pipeline state, continuing across rounds: ring group + its lap parity, stages issued
    unsigned grp = 0, gphase = 0, use = 0;
    for (int r = 0; r < rounds; r++) {
      const int npairs = min(ppr, pairs_sub - r * ppr);   // tile pairs of this round
      if (warp < npairs) {
        const int oct = min((r * ppr + warp) * 32 + lane, octs - 1);
        const int t = oct / nco, co = oct - t * nco;
        const float u = __ldg(&g_uvw[3 * t]), v = __ldg(&g_uvw[3 * t + 1]), w = __ldg(&g_uvw[3 * t + 2]);
        const float *wn8 = a.wavenumbers + 8 * co;
        float wn[8];
#pragma unroll
        for (int j = 0; j < 8; j++) wn[j] = __ldg(&wn8[j]);
        float dw;
        const bool lin = linear_channels(wn, 0, 8, &dw);
        const bool warp_lin = recur && __all_sync(0xffffffffu, lin);   // warp-uniform choice of the path
        const float wn0 = wn[0];
        for (int g = 0; g < ngroups; g++) {
          mbar_wait_u(bfull_u + grp * 8, gphase);
          const int nst = min(E_GROUP, nstages - g * E_GROUP);
          for (int st = 0; st < nst; st++, use++) {
            const unsigned slot = grp * E_GROUP + st;
            if (use >= 1) mbar_wait_u(my_empty_u, (use - 1) & 1);
            const float4 *geo = sG + slot * E_PB;
            if (warp_lin) etc_produce_linear(A_warp, geo, u, v, w, wn0, dw, lane);
            else etc_produce(A_warp, geo, u, v, w, wn8, lane);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            __syncwarp();
            if (elect_one()) {
              const unsigned long long db = db0 + (unsigned long long)(slot * (E_B_SLOT >> 4));
              const unsigned acc = (g > 0 || st > 0) ? 1u : 0u;
              if (!(IDGB200_ABLATE & 1)) {
              umma_f16(tmem_d, da0, db, idesc, acc);                                              // tile 0, hi
              umma_f16(tmem_d, da0 + (unsigned long long)(E_A_BUF >> 4), db, idesc, 1u);          // tile 0, lo
              umma_f16(tmem_d + 16, da0 + (unsigned long long)(2 * E_A_BUF >> 4), db, idesc, acc);   // tile 1, hi
              umma_f16(tmem_d + 16, da0 + (unsigned long long)(3 * E_A_BUF >> 4), db, idesc, 1u);    // tile 1, lo
              }
              umma_commit_u(my_empty_u);
              if (st == nst - 1) umma_commit_u(bempty_u + grp * 8);
              if (st == nst - 1 && g == ngroups - 1) umma_commit_u(done_u);
            }
            __syncwarp();
          }
          if (++grp == (unsigned)E_NG) { grp = 0; gphase ^= 1u; }
        }
      } else if (warp < ppr) {
        // no tile pair in this (last) round: keep the ring's and the round's arrival counts complete,
        // paced by the builder so that an arrival can never fall into an earlier phase
        for (int g = 0; g < ngroups; g++) {
          mbar_wait_u(bfull_u + grp * 8, gphase);
          if (lane == 0) mbar_arrive(&bempty[grp]);
          if (++grp == (unsigned)E_NG) { grp = 0; gphase ^= 1u; }
        }
        if (lane == 0) mbar_arrive(done);
      } else {
        for (int g = 0; g < ngroups; g++)
          if (++grp == (unsigned)E_NG) { grp = 0; gphase ^= 1u; }
      }

      // ---- epilogue of the round: accumulators -> visibilities (degridder_reference.cpp:118-127)
      mbar_wait(done, r & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // warp w reads TMEM lanes 32 w .. 32 w + 31 of every tile = channel 4 h + w of the octs
      for (int tl = 0; tl < 2 * npairs; tl++) {
        unsigned rr[16];
        const unsigned taddr = tmem_base + ((unsigned)(warp * 32) << 16) + tl * 16;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
              "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int oct = (r * ppr + (tl >> 1)) * 32 + lane;
        const int t = oct / nco, c = 8 * (oct - t * nco) + 4 * (tl & 1) + warp;
        if (oct < octs) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; i++) o[i] = (__uint_as_float(rr[i]) + __uint_as_float(rr[8 + i])) * pix_unscale;
          float4 *dst = g_vis + ((size_t)t * C + c) * 2;
          dst[0] = make_float4(o[0], o[1], o[2], o[3]);
          dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
      if (r + 1 < rounds) {   // the next round's first MMA overwrites the accumulators
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
    }
  } else {
    // ------------------------------------------------------------------ builder warp (as degridder_tc.cu)
    // pass 1: lane = pixel of the group: P' (scaled) -> scratch, geometry -> ring
    // pass 2: lane = (kc, n): one 16-byte chunk of B per stage of the group
    const int nrow = lane & 15, kc = lane >> 4, lo = nrow >> 3, p = (nrow >> 1) & 3, im = nrow & 1;
    int GG = 0;
    for (int r = 0; r < rounds; r++) {
      for (int g = 0; g < ngroups; g++, GG++) {
        const int grp = GG % E_NG;
        if (GG >= E_NG) mbar_wait(&bempty[grp], ((GG / E_NG) - 1) & 1);
        const int q = g * (E_GROUP * E_PB) + lane;
        float2 px[NR_POL];
        float4 geo = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 pd[NR_POL] = {};   // the pair's difference (multiplies sin) when the subgrid is folded
        if (q < npix_k) {
          pixel_after_aterms8(a, sub, plane, at1, at2, q, px);
          if (fold) {
            float2 py[NR_POL];
            pixel_after_aterms8(a, sub, plane, at1, at2, npix - 1 - q, py);
#pragma unroll
            for (int pp = 0; pp < NR_POL; pp++) {
              pd[pp] = make_float2(__fsub_rn(px[pp].x, py[pp].x), __fsub_rn(px[pp].y, py[pp].y));
              px[pp] = make_float2(__fadd_rn(px[pp].x, py[pp].x), __fadd_rn(px[pp].y, py[pp].y));
            }
          }
          const int y = q / N, x = q - y * N;
          const float l = compute_l(x, N, a.image_size);
          const float m = compute_l(y, N, a.image_size);
          const float n = compute_n(l, m);
          // the CPU binary leaves the w term unfused here (oracle/idg_oracle.c)
          geo = make_float4(l, m, n, __fadd_rn(__fmaf_rn(ctx.u_offset, l, __fmul_rn(ctx.v_offset, m)),
                                              __fmul_rn(ctx.w_offset, n)));
        } else {
#pragma unroll
          for (int pp = 0; pp < NR_POL; pp++) px[pp] = make_float2(0.f, 0.f);
        }
        sG[grp * (E_GROUP * E_PB) + lane] = geo;
#pragma unroll
        for (int pp = 0; pp < NR_POL; pp++) {
          scratch[lane * 8 + 2 * pp] = px[pp].x * pix_scale;
          scratch[lane * 8 + 2 * pp + 1] = px[pp].y * pix_scale;
          if (fold) {
            scratch[256 + lane * 8 + 2 * pp] = pd[pp].x * pix_scale;
            scratch[256 + lane * 8 + 2 * pp + 1] = pd[pp].y * pix_scale;
          }
        }
        __syncwarp();
#pragma unroll
        for (int st = 0; st < E_GROUP; st++) {
          unsigned pk[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float2 vv = *reinterpret_cast<const float2 *>(&scratch[(st * E_PB + kc * 4 + i) * 8 + 2 * p]);
            const float2 vd = fold ? *reinterpret_cast<const float2 *>(&scratch[256 + (st * E_PB + kc * 4 + i) * 8 + 2 * p]) : vv;
            const float x0 = im ? vv.y : vv.x;    // multiplies cos
            const float x1 = im ? vd.x : -vd.y;   // multiplies sin
            __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
            if (lo) {
              h0 = __float2half_rn(x0 - __half2float(h0));
              h1 = __float2half_rn(x1 - __half2float(h1));
            }
            pk[i] = (unsigned)__half_as_ushort(h0) | ((unsigned)__half_as_ushort(h1) << 16);
          }
          *reinterpret_cast<uint4 *>(sB + (grp * E_GROUP + st) * E_B_SLOT + kc * B_CHUNK_BYTES + nrow * 16) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&bfull[grp]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(E_TMEM_COLS));
}

// LIST = false: CTA = subgrid blockIdx.x; LIST = true: a fixed number of CTAs loop over the subgrids of a.list
// (what degridder_sep.cu left: the work is rare, the launch must be cheap when the list is empty)
template <bool LIST>
__global__ void __launch_bounds__(E_THREADS, 3)
degridder_tc8_kernel(const KernelArgs a, const int recur, const int fold_ok) {
  if (!LIST) {
    degridder_tc8_body<false>(a, recur, fold_ok, blockIdx.x);
  } else {
    const int total = a.list[0];
    for (int i = blockIdx.x; i < total; i += gridDim.x) {
      degridder_tc8_body<true>(a, recur, fold_ok, a.list[1 + i]);
      __syncthreads();
    }
  }
}

}  // namespace

// nr_channels must be a multiple of 8; recur: octs of equally spaced channels get their phasors by recurrence
// fold: planar subgrids (checked per subgrid on the device) sum over half the pixels, see the kernel
cudaError_t launch_degridder_tc8(const KernelArgs &a, bool recur, bool fold, cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  if (a.nr_channels & 7) return cudaErrorInvalidValue;
  const size_t smem = (size_t)E_PAIRS * E_A_WARP + E_NG * E_GROUP * (E_B_SLOT + E_G_SLOT) + 2 * 32 * 8 * 4 +
                      (E_PAIRS + 2 * E_NG + 1) * 8 + 8 + 48;
  auto k = a.list ? degridder_tc8_kernel<true> : degridder_tc8_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<dim3((unsigned)(a.list && a.nr_subgrids > LIST_MODE_CTAS ? LIST_MODE_CTAS : a.nr_subgrids)), dim3(E_THREADS), smem, stream>>>(a, recur ? 1 : 0, fold ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace idgb200
