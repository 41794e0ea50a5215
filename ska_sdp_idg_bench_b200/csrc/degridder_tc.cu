// IDG degridder on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a: the transpose of
// gridder_tc.cu.
//
//   D[vis][n] += A[vis][k] * B[k][n]               one CTA = one subgrid, rounds of <= 8 tiles x 128 visibilities
//     k = (pixel, {cos, sin})                         A: fp16 phasors, produced by MUFU / FP32 polynomial
//     n = (hi|lo, pol, re|im)  -> N = 16              B: P' = A1 (sph . subgrid) A2^H split into fp16 hi + lo
//       B[(pix,cos)][re,pol] =  Pr    B[(pix,sin)][re,pol] = -Pi
//       B[(pix,cos)][im,pol] =  Pi    B[(pix,sin)][im,pol] =  Pr
//   (degridder_reference.cpp:38-129; the phase is evaluated in the CPU binary's operation order,
//   so the angle fed to sincos is bit-identical to the FP32 kernel's and the reference's.)
//
// Rows: the visibilities of the subgrid are enumerated in quads = (timestep, group of 4 channels);
// a tile holds 32 quads, lane l of producer warp w owns quad 32 w + l and its four rows
// l + 32 j = channel 4 cg + j.  A thread therefore has ONE timestep: u, v, w and 4 wavenumbers
// stay in registers for the whole kernel, and the phase index u l + v m + w n is computed once
// per pixel for 4 items.
//
// As in the gridder every producer warp is its own pipeline: per stage (8 pixels, K = 16) it
// makes 32 phasors per thread, stores them in the K-major core-matrix layout, and an elected
// lane issues the stage's tcgen05.mma (M=128, N=16, K=16) and commits it to the warp's private
// empty barrier.  The B operand and the pixel geometry (l, m, n, phase offset) stream through a
// ring of 3 groups x 32 pixels filled by one builder warp (taper, A-terms, hi/lo split), so
// nothing pixel-sized is resident and any subgrid size runs with the same shared memory.
// FAST sincos only: the fp16 phasor operand is a FAST-class approximation.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace idgb200 {

namespace {

constexpr int D_MAX_TILES = 8;                   // M-tiles (128 visibilities) = producer warps per CTA
constexpr int D_PB = 8;                          // pixels per stage -> K = 16
constexpr int D_A_STAGE = 2 * A_CHUNK_BYTES;     // 4 KB per tile and stage
constexpr int D_STAGES = 2;
constexpr int D_GROUP = 4;                       // stages per ring group (32 pixels = one builder pass)
constexpr int D_NG = 3;                          // ring groups
constexpr int D_B_SLOT = 2 * B_CHUNK_BYTES;      // 512 B of B per stage
constexpr int D_G_SLOT = D_PB * 16;              // 128 B of (l, m, n, offset) per stage
constexpr int D_THREADS = (D_MAX_TILES + 1) * 32;
constexpr int D_TMEM_COLS = D_MAX_TILES * 16;   // 128 columns (power of two >= 32)

// P' of pixel q (degridder_reference.cpp:38-74): taper, A1 . P . A2^H
__device__ __forceinline__ void pixel_after_aterms(const KernelArgs &a, const float2 *sub, size_t plane, size_t at1,
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

// one stage of one thread: 8 pixels x 4 channels.  MASK8 bit i set -> pixel i of the stage gets
// its phasors from phasor_poly (FP32 pipe) instead of MUFU.  SPLIT: the rounding residual of the
// fp16 phasor goes to the tile's second A buffer (see gridder_tc.cu: pack_phasor).
template <unsigned MASK8, bool SPLIT>
__device__ __forceinline__ void dtc_produce(unsigned char *A, const float4 *geo, const float u, const float v,
                                            const float w, const float (&wn)[4], const int lane) {
#pragma unroll
  for (int kc = 0; kc < 2; kc++) {
    unsigned pk[4][4], pl[4][4];   // [row j][pixel i]
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float4 g = geo[kc * 4 + i];   // broadcast: all lanes read the same 16 bytes
      // degridder_reference.cpp:106 as the CPU binary evaluates it (w term unfused)
      const float idx = __fadd_rn(__fmaf_rn(u, g.x, __fmul_rn(v, g.y)), __fmul_rn(w, g.z));
      if ((MASK8 >> (kc * 4 + i)) & 1u) {
        const float idxr = __fmul_rn(idx, 0.15915494309189535f), offr = __fmul_rn(g.w, 0.15915494309189535f);
#pragma unroll
        for (int j = 0; j < 4; j++) pack_phasor<SPLIT>(phasor_poly(__fmaf_rn(idxr, wn[j], -offr)), pk[j][i], pl[j][i]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)   // :112, (cos, sin)
          pack_phasor<SPLIT>(phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(idx, wn[j], -g.w)), pk[j][i], pl[j][i]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      *reinterpret_cast<uint4 *>(A + kc * A_CHUNK_BYTES + (lane + 32 * j) * 16) =
          make_uint4(pk[j][0], pk[j][1], pk[j][2], pk[j][3]);
      if (SPLIT)
        *reinterpret_cast<uint4 *>(A + D_A_STAGE + kc * A_CHUNK_BYTES + (lane + 32 * j) * 16) =
            make_uint4(pl[j][0], pl[j][1], pl[j][2], pl[j][3]);
    }
  }
}

// The same stage when the thread's four channels are equally spaced (common.cuh: linear_channels):
// per pixel one sincos for the first channel (the reference's angle, bit for bit), one for the
// rotation e^{i idx dw}, one complex multiplication and two steps of the three-term recurrence.
template <bool SPLIT>
__device__ __forceinline__ void dtc_produce_linear(unsigned char *A, const float4 *geo, const float u, const float v,
                                                   const float w, const float wn0, const float dw, const int lane) {
#pragma unroll
  for (int kc = 0; kc < 2; kc++) {
    unsigned pk[4][4], pl[4][4];   // [row j][pixel i]
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float4 g = geo[kc * 4 + i];
      const float idx = __fadd_rn(__fmaf_rn(u, g.x, __fmul_rn(v, g.y)), __fmul_rn(w, g.z));
      const float2 p0 = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(idx, wn0, -g.w));
      const float2 d = phasor<IDGB200_SINCOS_FAST>(__fmul_rn(idx, dw));
      const float2 dxx = make_float2(d.x, d.x), dny = make_float2(-d.y, d.y);
      const float c2 = __fadd_rn(d.x, d.x);
      const float2 cc = make_float2(c2, c2);
      // channel 1 by rotation: (x, y) * d = (x, y) * (dx, dx) + (y, x) * (-dy, dy): FMUL2 + FFMA2;
      // channels 2, 3 by the three-term recurrence ph[c+1] = 2 cos(delta) ph[c] - ph[c-1]: one FFMA2
      // each (gridder_tc.cu: tc_produce_linear has the error argument)
      const float2 p1 = ffma2(make_float2(p0.y, p0.x), dny, __fmul2_rn(p0, dxx));
      const float2 p2 = ffma2(p1, cc, make_float2(-p0.x, -p0.y));
      const float2 p3 = ffma2(p2, cc, make_float2(-p1.x, -p1.y));
      pack_phasor<SPLIT>(p0, pk[0][i], pl[0][i]);
      pack_phasor<SPLIT>(p1, pk[1][i], pl[1][i]);
      pack_phasor<SPLIT>(p2, pk[2][i], pl[2][i]);
      pack_phasor<SPLIT>(p3, pk[3][i], pl[3][i]);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      *reinterpret_cast<uint4 *>(A + kc * A_CHUNK_BYTES + (lane + 32 * j) * 16) =
          make_uint4(pk[j][0], pk[j][1], pk[j][2], pk[j][3]);
      if (SPLIT)
        *reinterpret_cast<uint4 *>(A + D_A_STAGE + kc * A_CHUNK_BYTES + (lane + 32 * j) * 16) =
            make_uint4(pl[j][0], pl[j][1], pl[j][2], pl[j][3]);
    }
  }
}

// SPLIT: fp16 hi + lo phasors: the tile's two A buffers hold the hi and the lo block of ONE stage (two
//        MMAs against the same B slot; a warp then waits for its own MMAs every stage)
// recur: quads of equally spaced channels get their phasors by rotation from the quad's first channel
template <unsigned MASK8, bool SPLIT, bool LIST>
__device__ __forceinline__ void degridder_tc_body(const KernelArgs &a, const int recur, const int s_local) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int N = a.subgrid_size, C = a.nr_channels, npix = N * N;
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler too
  constexpr int NW = D_MAX_TILES;                             // producer warps; warp NW builds B

  const SubgridCtx ctx = load_ctx(a, s);
  const int nt = ctx.nr_timesteps;
  const int ncg = (C + 3) >> 2;                               // channel groups of 4
  const int quads = nt * ncg;
  if (quads == 0) return;
  // the subgrid's tiles are processed in rounds of <= 8 (one per producer warp), balanced
  const int tiles_sub = (quads + 31) >> 5;
  const int rounds = (tiles_sub + D_MAX_TILES - 1) / D_MAX_TILES;
  const int tpr = (tiles_sub + rounds - 1) / rounds;
  const int nstages = (npix + D_PB - 1) / D_PB;
  const int ngroups = (nstages + D_GROUP - 1) / D_GROUP;

  unsigned char *sA = smem;                                                      // [tile][stage][4 KB]
  unsigned char *sB = sA + D_MAX_TILES * D_STAGES * D_A_STAGE;                   // [NG * GROUP][512 B]
  float4 *sG = reinterpret_cast<float4 *>(sB + D_NG * D_GROUP * D_B_SLOT);       // [NG * GROUP][8] (l, m, n, off)
  float *scratch = reinterpret_cast<float *>(sG + D_NG * D_GROUP * D_PB);        // [32][8] builder scratch
  unsigned long long *aempty = reinterpret_cast<unsigned long long *>(scratch + 32 * 8);  // [tile][stage]
  unsigned long long *bfull = aempty + D_MAX_TILES * D_STAGES;                   // [NG]
  unsigned long long *bempty = bfull + D_NG;                                     // [NG]
  unsigned long long *done = bempty + D_NG;
  unsigned *s_tmem = reinterpret_cast<unsigned *>(done + 1);
  float *s_red = reinterpret_cast<float *>(s_tmem + 2);     // [12] block reduction scratch + scale

  if (tid == 0) {
    for (int i = 0; i < D_MAX_TILES * D_STAGES; i++) mbar_init(&aempty[i], 1);
    for (int i = 0; i < D_NG; i++) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], tpr);
    }
    mbar_init(done, tpr);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(D_TMEM_COLS));
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

  // fp16 has 5 exponent bits: scale P' by a power of two so that its largest component lands in
  // [2^13, 2^14) (exact; undone in the epilogue).  One extra evaluation of the A-term product per
  // pixel and CTA (< 1 % of the CTA's work), and it warms L1/L2 for the builder.
  {
    float amax = 0.f;
    for (int q = tid; q < npix; q += D_THREADS) {
      float2 px[NR_POL];
      pixel_after_aterms(a, sub, plane, at1, at2, q, px);
#pragma unroll
      for (int p = 0; p < NR_POL; p++) amax = fmaxf(amax, fmaxf(fabsf(px[p].x), fabsf(px[p].y)));
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
    // ------------------------------------------------------------------ producers (+ their own MMA)
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6) = 1, A = B = F16 (0),
    // both K-major (0), N >> 3 at [17,23), M >> 4 at [24,29)
    const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    unsigned char *A_tile = sA + warp * D_STAGES * D_A_STAGE;
    unsigned long long *my_empty = aempty + warp * D_STAGES;
    unsigned long long da0 = smem_desc(smem_u32(A_tile), A_CHUNK_BYTES, 128);
    unsigned long long db0 = smem_desc(smem_u32(sB), B_CHUNK_BYTES, 128);
    unsigned tmem_d = tmem_base + warp * 16;
    // barrier addresses as 32-bit shared addresses, computed once and made opaque so that the
    // compiler keeps them instead of rematerialising the address arithmetic in every issue path
    unsigned my_empty_u = smem_u32(my_empty), bfull_u = smem_u32(bfull), bempty_u = smem_u32(bempty),
             done_u = smem_u32(done);
    asm volatile("" : "+l"(da0), "+l"(db0), "+r"(tmem_d), "+r"(my_empty_u), "+r"(bfull_u), "+r"(bempty_u), "+r"(done_u));
    float4 *g_vis = reinterpret_cast<float4 *>(const_cast<float2 *>(a.visibilities)) +
                    (size_t)ctx.time_offset * C * 2;
    int kk = 0, GG = 0;   // running stage / ring-group counters: barrier phases continue across rounds
    for (int r = 0; r < rounds; r++) {
      const int tile = r * tpr + warp;
      const int ntiles = min(tpr, tiles_sub - r * tpr);   // tiles of this round
      if (warp < ntiles) {
        const int quad = min(tile * 32 + lane, quads - 1);
        const int t = quad / ncg, cg = quad - t * ncg;
        const float u = __ldg(&g_uvw[3 * t]), v = __ldg(&g_uvw[3 * t + 1]), w = __ldg(&g_uvw[3 * t + 2]);
        float wn[4];
#pragma unroll
        for (int j = 0; j < 4; j++) wn[j] = (4 * cg + j < C) ? __ldg(&a.wavenumbers[4 * cg + j]) : 0.f;
        float dw;
        const bool lin = linear_channels(wn, 0, min(4, C - 4 * cg), &dw);
        const bool warp_lin = recur && __all_sync(0xffffffffu, lin);   // warp-uniform choice of the path
        for (int k = 0; k < nstages; k++, kk++) {
          const int stage = SPLIT ? 0 : (kk & 1), use = SPLIT ? kk : (kk >> 1);
          const int G = GG + (k >> 2), grp = G % D_NG, slot = grp * D_GROUP + (k & (D_GROUP - 1));
          if ((k & (D_GROUP - 1)) == 0) mbar_wait_u(bfull_u + grp * 8, (G / D_NG) & 1);
          if (use >= 1) mbar_wait_u(my_empty_u + stage * 8, (use - 1) & 1);
          if (warp_lin)
            dtc_produce_linear<SPLIT>(A_tile + stage * D_A_STAGE, sG + slot * D_PB, u, v, w, wn[0], dw, lane);
          else
            dtc_produce<MASK8, SPLIT>(A_tile + stage * D_A_STAGE, sG + slot * D_PB, u, v, w, wn, lane);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          __syncwarp();
          if (elect_one()) {
            umma_f16(tmem_d, da0 + (unsigned long long)(stage * (D_A_STAGE >> 4)),
                     db0 + (unsigned long long)(slot * (D_B_SLOT >> 4)), idesc, k > 0 ? 1u : 0u);
            if (SPLIT)
              umma_f16(tmem_d, da0 + (unsigned long long)(D_A_STAGE >> 4),
                       db0 + (unsigned long long)(slot * (D_B_SLOT >> 4)), idesc, 1u);
            umma_commit_u(my_empty_u + stage * 8);
            if ((k & (D_GROUP - 1)) == D_GROUP - 1 || k == nstages - 1) umma_commit_u(bempty_u + grp * 8);
            if (k == nstages - 1) umma_commit_u(done_u);
          }
        }
      } else if (warp < tpr) {
        // no tile in this (last) round: keep the ring's and the round's arrival counts complete,
        // paced by the builder so that an arrival can never fall into an earlier phase
        for (int g = 0; g < ngroups; g++) {
          const int G = GG + g, grp = G % D_NG;
          mbar_wait(&bfull[grp], (G / D_NG) & 1);
          if (lane == 0) mbar_arrive(&bempty[grp]);
        }
        if (lane == 0) mbar_arrive(done);
      }
      GG += ngroups;

      // ---- epilogue of the round: accumulators -> visibilities (degridder_reference.cpp:118-127)
      mbar_wait(done, r & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q4 = warp & 3;   // a warp reads TMEM lanes 32 (warp % 4) .. +31 = rows of channel 4 cg + q4
      for (int tl = warp >> 2; tl < ntiles; tl += NW / 4) {
        unsigned rr[16];
        const unsigned taddr = tmem_base + ((unsigned)(q4 * 32) << 16) + tl * 16;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
              "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int quad = (r * tpr + tl) * 32 + lane;
        const int t = quad / ncg, c = 4 * (quad - t * ncg) + q4;
        if (quad < quads && c < C) {
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
    // ------------------------------------------------------------------ builder warp
    // pass 1: lane = pixel of the group: P' (scaled) -> scratch, geometry -> ring
    // pass 2: lane = (kc, n): one 16-byte chunk of B per stage of the group
    const int nrow = lane & 15, kc = lane >> 4, lo = nrow >> 3, p = (nrow >> 1) & 3, im = nrow & 1;
    int GG = 0;
    for (int r = 0; r < rounds; r++) {
      for (int g = 0; g < ngroups; g++, GG++) {
        const int grp = GG % D_NG;
        if (GG >= D_NG) mbar_wait(&bempty[grp], ((GG / D_NG) - 1) & 1);
        const int q = g * (D_GROUP * D_PB) + lane;
        float2 px[NR_POL];
        float4 geo = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < npix) {
          pixel_after_aterms(a, sub, plane, at1, at2, q, px);
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
        sG[grp * (D_GROUP * D_PB) + lane] = geo;
#pragma unroll
        for (int pp = 0; pp < NR_POL; pp++) {
          scratch[lane * 8 + 2 * pp] = px[pp].x * pix_scale;
          scratch[lane * 8 + 2 * pp + 1] = px[pp].y * pix_scale;
        }
        __syncwarp();
#pragma unroll
        for (int st = 0; st < D_GROUP; st++) {
          unsigned pk[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float2 vv = *reinterpret_cast<const float2 *>(&scratch[(st * D_PB + kc * 4 + i) * 8 + 2 * p]);
            const float x0 = im ? vv.y : vv.x;    // multiplies cos
            const float x1 = im ? vv.x : -vv.y;   // multiplies sin
            __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
            if (lo) {
              h0 = __float2half_rn(x0 - __half2float(h0));
              h1 = __float2half_rn(x1 - __half2float(h1));
            }
            pk[i] = (unsigned)__half_as_ushort(h0) | ((unsigned)__half_as_ushort(h1) << 16);
          }
          *reinterpret_cast<uint4 *>(sB + (grp * D_GROUP + st) * D_B_SLOT + kc * B_CHUNK_BYTES + nrow * 16) =
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(D_TMEM_COLS));
}

// LIST = false: CTA = subgrid blockIdx.x; LIST = true: a fixed number of CTAs loop over the subgrids of a.list
// (what degridder_sep.cu left: the work is rare, the launch must be cheap when the list is empty)
template <unsigned MASK8, bool SPLIT, bool LIST>
__global__ void __launch_bounds__(D_THREADS, 3)
degridder_tc_kernel(const KernelArgs a, const int recur) {
  if (!LIST) {
    degridder_tc_body<MASK8, SPLIT, false>(a, recur, blockIdx.x);
  } else {
    const int total = a.list[0];
    for (int i = blockIdx.x; i < total; i += gridDim.x) {
      degridder_tc_body<MASK8, SPLIT, true>(a, recur, a.list[1 + i]);
      __syncthreads();
    }
  }
}

}  // namespace

// poly: (0..3 = fp16 phasors, by MUFU or partly by FP32 polynomial: outside the stated tolerance on the
//       reference's degridder input, DESIGN.md 4.6 - no longer built);
//       10 = fp16 hi + lo phasors (FP32-class accuracy)
// recur: quads of equally spaced channels get their phasors by rotation
cudaError_t launch_degridder_tc(const KernelArgs &a, int poly, bool recur, cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  const size_t smem = (size_t)D_MAX_TILES * D_STAGES * D_A_STAGE + D_NG * D_GROUP * (D_B_SLOT + D_G_SLOT) + 32 * 8 * 4 +
                      (D_MAX_TILES * D_STAGES + 2 * D_NG + 1) * 8 + 8 + 48;
  void (*k)(const KernelArgs, int) = nullptr;
  switch (poly) {
    case 10:   // hi + lo phasors
      k = a.list ? degridder_tc_kernel<0x00u, true, true> : degridder_tc_kernel<0x00u, true, false>;
      break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<dim3((unsigned)(a.list && a.nr_subgrids > LIST_MODE_CTAS ? LIST_MODE_CTAS : a.nr_subgrids)), dim3(D_THREADS), smem, stream>>>(a, recur ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace idgb200
