// IDG gridder on tcgen05 + TMEM for PLANAR subgrids: half the phasors (variant 29; behind a device-side
// gate the default for FAST sincos when gridder_tc.cu's regular-case loop would run).
//
// When every timestep of a subgrid has w = 0 and the subgrid has no w offset - the reference's own
// synthetic observation (init.cpp:4-25: w = 0, W_STEP = 0) and any snapshot of a coplanar array - pixel q
// and its mirror image q' = npix - 1 - q have l and m negated (math.hpp:9-12, even subgrid sizes), so
// every operation of gridder_reference.cpp:61-69 gives exactly the negated phase in the reference's own
// fp32 arithmetic, and the phasor of q' is the conjugate of q's, bit for bit.  With
//   E = sum_v cos_v vis_v,   F = sum_v sin_v vis_v        (complex, per polarisation)
// the two pixels are  D[q] = E + i F  and  D[q'] = E - i F : ONE phasor row serves both, and the
// phasors are what this kernel's time goes into (tools/ablate.py: 10.5 of 12.9 ms is the instruction
// stream that makes them).  As a GEMM:
//   D[pair][n] += A[pair][k] * B[k][n]       M = 128 pixel pairs per tile, K = 16 per MMA, N = 32
//     k = (visibility v, {cos, sin})            A: fp16 phasors of pixel q, as in gridder_tc.cu
//     n = (E|F, hi|lo, pol, re|im)              B[(v,cos)][E..] = vis, B[(v,sin)][F..] = vis, zeros elsewhere
// so a subgrid of 32 x 32 pixels is 4 tiles instead of 8, each MMA does twice the columns for the
// same 4 KB operand fetch (the fetch, not the math, is what a small-N MMA costs: tools/mma_commit.cu),
// and the epilogue recombines E +- i F before the A-terms and the taper, which are applied per pixel
// exactly as in gridder_tc.cu.
//
// CTA = up to 4 tiles (one producer warp each, rows lane + 32 j) + the B builder warp = 160 threads,
// 49 KB of shared memory, 128 TMEM columns: 4 CTAs per SM.  The stage loop is gridder_tc.cu's regular
// case (every 8-channel block equally spaced with one spacing, an even number of blocks: one stage =
// one timestep x 16 channels, single-buffered, three-term recurrence).  The gate is per subgrid: two small
// check kernels on the same stream sort the launch's subgrids into a fold list (regular channel layout
// AND planar subgrid AND even subgrid size) and a general list; this kernel serves the first,
// gridder_tc.cu (variant 24), launched behind it, the second, and the CTAs either kernel has no
// subgrid for return after one read.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace idgb200 {

namespace {

constexpr int F_TILES = 4;                       // M-tiles (128 pixel pairs) = producer warps per CTA
constexpr int F_CB = 8;                          // channels per block -> K = 16
constexpr int F_A_BLOCK = 2 * A_CHUNK_BYTES;     // 4 KB per tile and channel block
constexpr int F_B_CHUNK = 32 * 16;               // one 16-byte K-chunk of the 32 columns
constexpr int F_B_SLOT = 2 * F_B_CHUNK;          // 1 KB per channel block
constexpr int F_NB = 16;                         // B ring slots
constexpr int F_THREADS = (F_TILES + 1) * 32;

// which subgrids of the launch fold: regular channel layout (*regular_flag, written by
// gridder_regular_check_kernel on the same stream) and a planar subgrid.  One warp per subgrid, its lanes
// stride over the timesteps, so the check is a handful of independent loads deep.  The verdicts are
// compacted into two lists, lists = { n_fold, n_general, fold[nr_subgrids], general[nr_subgrids] } (counts
// zeroed on the stream before), so that in the two kernels behind it CTA i either finds its subgrid in
// its list or returns after one read of a count every CTA shares.
__global__ void gridder_planar_check_kernel(const KernelArgs a, const int *__restrict__ regular_flag,
                                            int *__restrict__ lists) {
  const int s_local = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s_local >= a.nr_subgrids) return;
  const SubgridCtx ctx = load_ctx(a, a.subgrid_offset + s_local);
  bool planar = *regular_flag != 0 && ctx.w_offset == 0.f && !(a.subgrid_size & 1);
  const float *g_uvw = reinterpret_cast<const float *>(a.uvw) + (size_t)ctx.time_offset * 3;
  for (int t = lane; t < ctx.nr_timesteps; t += 32) planar = planar && __ldg(&g_uvw[3 * t + 2]) == 0.f;
  planar = __all_sync(0xffffffffu, planar);
  if (lane == 0) {
    const int i = atomicAdd(&lists[planar ? 0 : 1], 1);
    lists[2 + (planar ? 0 : a.nr_subgrids) + i] = s_local;
  }
}

// one block of 8 equally spaced channels for the 4 pixel pairs of a thread: gridder_tc.cu's
// tc_produce_linear (first channel by sincos - the reference's angle, bit for bit -, second by one
// rotation, the rest by ph[c+1] = 2 cos(delta) ph[c] - ph[c-1])
__device__ __forceinline__ void fold_produce_linear(unsigned char *A, const float wn0, const float2 (&rot)[4],
                                                    const float (&idx)[4], const float (&off)[4], const int lane) {
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

__global__ void __launch_bounds__(F_THREADS, 4)
gridder_fold_kernel(const KernelArgs a, const int slabs, const int *__restrict__ lists) {
  const int i_cta = blockIdx.x / slabs;
  if (i_cta >= lists[0]) return;          // the other subgrids (irregular channels, off the plane): gridder_tc.cu
  const int s_local = lists[2 + i_cta];
  extern __shared__ __align__(1024) unsigned char smem[];
  const int N = a.subgrid_size, C = a.nr_channels, npix = N * N, npairs = npix >> 1;
  const int slab = blockIdx.x - i_cta * slabs;
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler too
  const int pair0 = slab * F_TILES * 128;
  const int ntiles = min(F_TILES, (npairs - pair0 + 127) / 128);
  if (ntiles <= 0) return;

  unsigned char *sA = smem;                                              // [tile][block a | b][4 KB]
  unsigned char *sB = sA + F_TILES * 2 * F_A_BLOCK;                      // [F_NB][1 KB]
  unsigned long long *aempty = reinterpret_cast<unsigned long long *>(sB + F_NB * F_B_SLOT);  // [tile]
  unsigned long long *bfull = aempty + F_TILES;                          // [F_NB]
  unsigned long long *bempty = bfull + F_NB;                             // [2] half rings
  unsigned long long *done = bempty + 2;
  unsigned *s_tmem = reinterpret_cast<unsigned *>(done + 1);
  float *s_red = reinterpret_cast<float *>(s_tmem + 2);     // [12] block reduction scratch + scale
  float *s_wn = s_red + 12;                                 // [ncb * 8]

  const SubgridCtx ctx = load_ctx(a, s);
  const int nt = ctx.nr_timesteps;
  const int ncb = C / F_CB;                                 // even, every block full (the gate)
  const int nstages = nt * ncb;                             // channel blocks = B slots

  for (int c = tid; c < C; c += F_THREADS) s_wn[c] = a.wavenumbers[c];
  if (tid == 0) {
    for (int i = 0; i < F_TILES; i++) mbar_init(&aempty[i], 1);
    for (int i = 0; i < F_NB; i++) mbar_init(&bfull[i], 1);
    mbar_init(&bempty[0], ntiles);
    mbar_init(&bempty[1], ntiles);
    mbar_init(done, ntiles);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(F_TILES * 32));
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
    for (int i = tid; i < nt * C * 2; i += F_THREADS) {
      const float4 q = __ldg(&v4[i]);
      amax = fmaxf(fmaxf(amax, fmaxf(fabsf(q.x), fabsf(q.y))), fmaxf(fabsf(q.z), fabsf(q.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) s_red[warp] = amax;
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i <= F_TILES; i++) amax = fmaxf(amax, s_red[i]);
      const unsigned eb = (__float_as_uint(amax) >> 23) & 0xffu;          // biased exponent
      const bool ok = eb >= 14u && eb <= 253u;                             // finite, not tiny
      s_red[10] = ok ? __uint_as_float((267u - eb) << 23) : 1.f;           // 2^(13 - E)
      s_red[11] = ok ? __uint_as_float((eb - 13u) << 23) : 1.f;            // 2^(E - 13)
    }
    __syncthreads();
  }
  const float vis_scale = s_red[10], vis_unscale = s_red[11];

  if (warp < F_TILES) {
    // ------------------------------------------------------------------ producers (+ their own MMAs)
    const int tile = warp;
    if (tile < ntiles) {
      // instruction descriptor: D = F32, A = B = F16, both K-major, N = 32, M = 128
      const unsigned idesc = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      float l[4], m[4], off[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int q = min(pair0 + tile * 128 + lane + 32 * j, npairs - 1);
        const int y = q / N, x = q - y * N;
        l[j] = compute_l(x, N, a.image_size);
        m[j] = compute_l(y, N, a.image_size);
        // gridder_reference.cpp:64 as the CPU binary contracts it; the w term is +-0 (planar)
        off[j] = __fmaf_rn(ctx.u_offset, l[j], __fmul_rn(ctx.v_offset, m[j]));
      }
      unsigned char *A_tile = sA + tile * 2 * F_A_BLOCK;
      unsigned long long da0 = smem_desc(smem_u32(A_tile), A_CHUNK_BYTES, 128);
      unsigned long long db0 = smem_desc(smem_u32(sB), F_B_CHUNK, 128);
      unsigned tmem_d = tmem_base + tile * 32;
      unsigned my_empty_u = smem_u32(aempty + tile), bfull_u = smem_u32(bfull), bempty_u = smem_u32(bempty),
               done_u = smem_u32(done), wn_u = smem_u32(s_wn);
      asm volatile("" : "+l"(da0), "+l"(db0), "+r"(tmem_d), "+r"(my_empty_u), "+r"(bfull_u), "+r"(bempty_u), "+r"(done_u),
                   "+r"(wn_u));
      float dw0;
      linear_channels(s_wn, 0, F_CB, &dw0);                // the one spacing of every block (the gate)
      float un = 0.f, vn = 0.f;                            // uv of the next timestep, fetched one ahead
      if (nt > 0) { un = __ldg(&g_uvw[0]); vn = __ldg(&g_uvw[1]); }
      unsigned k = 0, sk = 0, slot2 = 0, ring_phase = 0;
      const unsigned last_k = (unsigned)nstages - 2u;
      for (int t = 0; t < nt; t++) {
        const float u = un, v = vn;
        if (t + 1 < nt) { un = __ldg(&g_uvw[3 * t + 3]); vn = __ldg(&g_uvw[3 * t + 4]); }
        float idx[4];
        float2 rot[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {  // gridder_reference.cpp:61 with w = +-0
          idx[j] = __fmaf_rn(u, l[j], __fmul_rn(v, m[j]));
          rot[j] = phasor<IDGB200_SINCOS_FAST>(__fmul_rn(-idx[j], dw0));
        }
        for (int cb0 = 0; cb0 < ncb; cb0 += 2, sk++, k += 2) {
          if (sk >= 1) mbar_wait_u(my_empty_u, (sk - 1) & 1);
          float wn0a, wn0b;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wn0a) : "r"(wn_u + (unsigned)cb0 * (F_CB * 4)));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wn0b) : "r"(wn_u + (unsigned)cb0 * (F_CB * 4) + F_CB * 4));
          fold_produce_linear(A_tile, wn0a, rot, idx, off, lane);
          fold_produce_linear(A_tile + F_A_BLOCK, wn0b, rot, idx, off, lane);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_wait_u(bfull_u + slot2 * 8, ring_phase);
          mbar_wait_u(bfull_u + slot2 * 8 + 8, ring_phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          __syncwarp();
          if (elect_one()) {
            const unsigned long long db = db0 + (unsigned long long)(slot2 * (F_B_SLOT >> 4));
            umma_f16(tmem_d, da0, db, idesc, k > 0 ? 1u : 0u);
            umma_f16(tmem_d, da0 + (unsigned long long)(F_A_BLOCK >> 4), db + (unsigned long long)(F_B_SLOT >> 4), idesc, 1u);
            umma_commit_u(my_empty_u);
            if ((slot2 & 7u) == 6u) umma_commit_u(bempty_u + (slot2 >> 3) * 8);   // half ring consumed
            if (k == last_k) umma_commit_u(done_u);
          }
          __syncwarp();
          slot2 += 2;
          if (slot2 == (unsigned)F_NB) { slot2 = 0; ring_phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ B builder warp
    // lane = (kc, n16): 4 channels x column n16 = (hi|lo, pol, re|im).  The value goes to the cos row of
    // column n16 (E) and to the sin row of column 16 + n16 (F); the other halves are zero.
    const int nrow = lane & 15, kc = lane >> 4, lo = nrow >> 3, p = (nrow >> 1) & 3, im = nrow & 1;
    auto load_b = [&](int kk, float2 (&raw)[4]) {
      const int t = kk / ncb, cb = kk - t * ncb;
#pragma unroll
      for (int i = 0; i < 4; i++)
        raw[i] = __ldg(&g_vis[((size_t)t * C + cb * F_CB + kc * 4 + i) * NR_POL + p]);
    };
    float2 raw[4];
    if (nstages > 0) load_b(0, raw);
    for (int k = 0; k < nstages; k++) {
      const int slot = k % F_NB;
      if ((k & 7) == 0 && k >= F_NB) mbar_wait(&bempty[(k >> 3) & 1], ((k / F_NB) - 1) & 1);
      unsigned pe[4], pf[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float e = (im ? raw[i].y : raw[i].x) * vis_scale;
        __half h = __float2half_rn(e);
        if (lo) h = __float2half_rn(e - __half2float(h));
        pe[i] = (unsigned)__half_as_ushort(h);            // (cos row, sin row) = (e, 0)
        pf[i] = (unsigned)__half_as_ushort(h) << 16;      //                      (0, e)
      }
      if (k + 1 < nstages) load_b(k + 1, raw);
      unsigned char *chunk = sB + slot * F_B_SLOT + kc * F_B_CHUNK;
      *reinterpret_cast<uint4 *>(chunk + nrow * 16) = make_uint4(pe[0], pe[1], pe[2], pe[3]);
      *reinterpret_cast<uint4 *>(chunk + (16 + nrow) * 16) = make_uint4(pf[0], pf[1], pf[2], pf[3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bfull[slot]);
    }
  }

  // ---- epilogue (the 4 producer warps, one per TMEM lane quadrant): E +- i F -> A-terms, taper, store
  // (gridder_reference.cpp:84-110) for the pair's two pixels
  if (warp < F_TILES) {
    if (nstages > 0) {
      mbar_wait(done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const size_t plane = (size_t)npix;
    const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
    const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
    float2 *out = const_cast<float2 *>(a.subgrids) + (size_t)s * NR_POL * plane;
    for (int tile = 0; tile < ntiles; tile++) {
      unsigned r[32];
      if (nstages > 0) {
        const unsigned taddr = tmem_base + ((unsigned)(warp * 32) << 16) + tile * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr + 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 32; i++) r[i] = 0u;
      }
      const int pair = pair0 + tile * 128 + warp * 32 + lane;
      if (pair < npairs) {
        float2 E[NR_POL], F[NR_POL];
#pragma unroll
        for (int p = 0; p < NR_POL; p++) {
          E[p] = make_float2(__uint_as_float(r[2 * p]) + __uint_as_float(r[8 + 2 * p]),
                             __uint_as_float(r[2 * p + 1]) + __uint_as_float(r[8 + 2 * p + 1]));
          F[p] = make_float2(__uint_as_float(r[16 + 2 * p]) + __uint_as_float(r[24 + 2 * p]),
                             __uint_as_float(r[16 + 2 * p + 1]) + __uint_as_float(r[24 + 2 * p + 1]));
        }
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const int pixel = side ? npix - 1 - pair : pair;
          float2 px[NR_POL];
#pragma unroll
          for (int p = 0; p < NR_POL; p++)   // E + i F for q, E - i F for its mirror image
            px[p] = side ? make_float2((E[p].x + F[p].y) * vis_unscale, (E[p].y - F[p].x) * vis_unscale)
                         : make_float2((E[p].x - F[p].y) * vis_unscale, (E[p].y + F[p].x) * vis_unscale);
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
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(F_TILES * 32));
}

}  // namespace

// d_lists[2 + 2 nr_subgrids] = { n_fold, n_general, fold[], general[] }: the subgrids that fold (regular channel
// layout, *d_regular_flag from launch_gridder_regular_check on the same stream, and a planar subgrid of even
// size) and the others; the two counts must be zero when this runs
cudaError_t launch_gridder_planar_check(const KernelArgs &a, const int *d_regular_flag, int *d_lists,
                                        cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  gridder_planar_check_kernel<<<(a.nr_subgrids + 3) / 4, 128, 0, stream>>>(a, d_regular_flag, d_lists);
  return cudaGetLastError();
}

// serves the subgrids of the fold list; the other CTAs return at once
cudaError_t launch_gridder_fold(const KernelArgs &a, const int *d_lists, cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  if (a.nr_channels > 1024 || (a.nr_channels & 15) || (a.subgrid_size & 1)) return cudaErrorInvalidValue;
  const int npairs = a.subgrid_size * a.subgrid_size / 2;
  const int tiles_total = (npairs + 127) / 128;
  const int slabs = (tiles_total + F_TILES - 1) / F_TILES;
  const size_t smem = (size_t)F_TILES * 2 * F_A_BLOCK + F_NB * F_B_SLOT + (F_TILES + F_NB + 3) * 8 + 8 + 48 +
                      (size_t)a.nr_channels * 4;
  cudaError_t e = cudaFuncSetAttribute(gridder_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  gridder_fold_kernel<<<dim3((unsigned)a.nr_subgrids * slabs), dim3(F_THREADS), smem, stream>>>(a, slabs, d_lists);
  return cudaGetLastError();
}

}  // namespace idgb200
