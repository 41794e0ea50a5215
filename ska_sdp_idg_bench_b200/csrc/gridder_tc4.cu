// IDG gridder on tcgen05 with the phasor operand in TMEM, coarse-grained (variant 27, experimental).
//
// gridder_tc3.cu showed that the A operand can be written from registers straight to TMEM (tcgen05.st)
// and consumed by the A-from-TMEM form of tcgen05.mma - no STS.128, and an MMA that costs its 8 clocks
// of math instead of the ~40 it takes to fetch a 4 KB A tile from shared memory (tools/smem_mix.cu) -
// but met its four warps (a TMEM lane can only be written by the warp whose warp % 4 owns its
// quadrant) at a barrier every 8 items.  Here the work is cut the other way round:
//   * a GROUP of four warps (quadrants q = 0..3) owns four tiles; lane l of warp q holds row 32 q + l
//     of each of the group's four tiles = 4 pixels, exactly the per-thread work of gridder_tc.cu;
//   * a stage = (timestep, 16 channels) = 64 items per thread: per pixel 16 packed phasors (two blocks
//     of 8 by sincos + rotation + three-term recurrence) go out as ONE tcgen05.st.32x32b.x16 into the
//     tile's 16 A columns; then tcgen05.wait::st, one named barrier of the group's 128 threads, and
//     the warp of quadrant 0 issues the stage's 8 MMAs (4 tiles x 2 channel blocks, A from TMEM) and
//     commits them to the group's empty barrier, which all four warps wait on before the next stage;
//   * TMEM per tile: 16 columns of D + 16 of A (single-buffered) -> 8 tiles = 256 columns = 2 CTAs/SM,
//     shared memory only for the B ring: 9 KB.
// Regular channel layouts only, behind the same device-side check and gate as gridder_tc8.cu.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace idgb200 {

namespace {

constexpr int G4_TILES = 8;                      // tiles (128 pixels) per CTA = 2 groups of 4
constexpr int G4_WARPS = 8;                      // producer warps: group = warp >> 2, quadrant = warp & 3
constexpr int G4_CB = 8;                         // channels per block -> K = 16 per MMA
constexpr int G4_TCOLS = 32;                     // TMEM columns per tile: D 16 + A 16
constexpr int G4_B_SLOT = 2 * B_CHUNK_BYTES;     // 512 B
constexpr int G4_NB = 16;                        // B ring slots
constexpr int G4_THREADS = (G4_WARPS + 1) * 32;

__device__ __forceinline__ void umma_f16_ts4(unsigned tmem_d, unsigned tmem_a, unsigned long long db, unsigned idesc,
                                             unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const unsigned (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}

// 8 equally spaced channels of one pixel: first by sincos (the reference's angle, bit for bit), second by
// rotation, the rest by the three-term recurrence (gridder_tc.cu: tc_produce_linear)
__device__ __forceinline__ void g4_block(const float wn0, const float2 d, const float idx, const float off,
                                         unsigned *pk) {
  float2 prev = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(-idx, wn0, off));   // gridder_reference.cpp:69
  const float2 dxx = make_float2(d.x, d.x), dny = make_float2(-d.y, d.y);
  const float c2 = __fadd_rn(d.x, d.x);
  const float2 cc = make_float2(c2, c2);
  unsigned unused;
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
}

__global__ void __launch_bounds__(G4_THREADS, 2)
gridder_tc4_kernel(const KernelArgs a, const int slabs, const int tiles_per_cta, const int *__restrict__ regular_flag,
                   const int arrive_split) {
  if (*regular_flag == 0) return;   // irregular channels: gridder_tc.cu runs instead
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int TMEM_COLS = G4_TILES * G4_TCOLS;              // 256
  const int N = a.subgrid_size, C = a.nr_channels, npix = N * N;
  const int s_local = blockIdx.x / slabs;
  const int slab = blockIdx.x - s_local * slabs;
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler too
  const int pix0 = slab * tiles_per_cta * 128;
  const int ntiles = min(tiles_per_cta, (npix - pix0 + 127) / 128);
  if (ntiles <= 0) return;
  const int ngroups = (ntiles + 3) >> 2;                      // groups that have a tile

  unsigned char *sB = smem;                                              // [G4_NB][512 B]
  unsigned long long *aempty = reinterpret_cast<unsigned long long *>(sB + G4_NB * G4_B_SLOT);  // [group]
  unsigned long long *bfull = aempty + 2;                                // [G4_NB]
  unsigned long long *bempty = bfull + G4_NB;                            // [2] half rings
  unsigned long long *done = bempty + 2;
  unsigned *s_tmem = reinterpret_cast<unsigned *>(done + 1);
  float *s_red = reinterpret_cast<float *>(s_tmem + 2);     // [12] block reduction scratch + scale
  float *s_wn = s_red + 12;                                 // [ncb * 8], zero padded

  const SubgridCtx ctx = load_ctx(a, s);
  const int nt = ctx.nr_timesteps;
  const int ncb = (C + G4_CB - 1) / G4_CB;                  // even (regular_flag)
  const int nstages = nt * ncb;                             // B slots

  for (int c = tid; c < ncb * G4_CB; c += blockDim.x) s_wn[c] = c < C ? a.wavenumbers[c] : 0.f;
  if (tid == 0) {
    mbar_init(&aempty[0], 1);
    mbar_init(&aempty[1], 1);
    for (int i = 0; i < G4_NB; i++) mbar_init(&bfull[i], 1);
    mbar_init(&bempty[0], ngroups);
    mbar_init(&bempty[1], ngroups);
    mbar_init(done, ngroups);
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
      for (int i = 1; i <= G4_WARPS; i++) amax = fmaxf(amax, s_red[i]);
      const unsigned eb = (__float_as_uint(amax) >> 23) & 0xffu;          // biased exponent
      const bool ok = eb >= 14u && eb <= 253u;                             // finite, not tiny
      s_red[10] = ok ? __uint_as_float((267u - eb) << 23) : 1.f;           // 2^(13 - E)
      s_red[11] = ok ? __uint_as_float((eb - 13u) << 23) : 1.f;            // 2^(E - 13)
    }
    __syncthreads();
  }
  const float vis_scale = s_red[10], vis_unscale = s_red[11];

  const int grp = warp >> 2, q4 = warp & 3;                   // producer warps: group, TMEM lane quadrant
  const unsigned lane_base = (unsigned)(q4 * 32) << 16;
  const unsigned tmem_grp = tmem_base + grp * 4 * G4_TCOLS;   // tile j of the group at + j * 32: D +0, A +16

  if (warp < G4_WARPS) {
    if (grp < ngroups) {
      // ---------------------------------------------------------------- producers
      const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      const int gtiles = min(4, ntiles - 4 * grp);             // tiles of this group that exist
      float l[4], m[4], n[4], off[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {   // pixel 32 q + lane of tile 4 grp + j
        const int q = min(pix0 + (4 * grp + j) * 128 + q4 * 32 + lane, npix - 1);
        const int y = q / N, x = q - y * N;
        l[j] = compute_l(x, N, a.image_size);
        m[j] = compute_l(y, N, a.image_size);
        n[j] = compute_n(l[j], m[j]);
        // gridder_reference.cpp:64 as the CPU binary contracts it
        off[j] = __fmaf_rn(ctx.w_offset, n[j], __fmaf_rn(ctx.u_offset, l[j], __fmul_rn(ctx.v_offset, m[j])));
      }
      unsigned long long db0 = smem_desc(smem_u32(sB), B_CHUNK_BYTES, 128);
      unsigned my_empty_u = smem_u32(aempty + grp), bfull_u = smem_u32(bfull), bempty_u = smem_u32(bempty),
               done_u = smem_u32(done), wn_u = smem_u32(s_wn);
      asm volatile("" : "+l"(db0), "+r"(my_empty_u), "+r"(bfull_u), "+r"(bempty_u), "+r"(done_u), "+r"(wn_u));
      float dw0;
      linear_channels(s_wn, 0, min(G4_CB, C), &dw0);
      float un = 0.f, vn = 0.f, wnx = 0.f;   // uvw of the next timestep, fetched one timestep ahead
      if (nt > 0) { un = __ldg(&g_uvw[0]); vn = __ldg(&g_uvw[1]); wnx = __ldg(&g_uvw[2]); }
      unsigned k = 0, sk = 0;                 // B slots and stages done
      unsigned slot2 = 0, ring_phase = 0;     // B slot pair of the stage (k % 16), lap parity of the ring
      const unsigned last_k = (unsigned)nstages - 2u;
      for (int t = 0; t < nt; t++) {
        const float u = un, v = vn, w = wnx;
        if (t + 1 < nt) { un = __ldg(&g_uvw[3 * t + 3]); vn = __ldg(&g_uvw[3 * t + 4]); wnx = __ldg(&g_uvw[3 * t + 5]); }
        float idx[4];
        float2 rot[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {  // gridder_reference.cpp:61 as contracted by the CPU binary
          idx[j] = __fmaf_rn(w, n[j], __fmaf_rn(u, l[j], __fmul_rn(v, m[j])));
          rot[j] = phasor<IDGB200_SINCOS_FAST>(__fmul_rn(-idx[j], dw0));
        }
        for (int cb0 = 0; cb0 < ncb; cb0 += 2, sk++, k += 2) {
          float wn0a, wn0b;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wn0a) : "r"(wn_u + (unsigned)cb0 * (G4_CB * 4)));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wn0b) : "r"(wn_u + (unsigned)cb0 * (G4_CB * 4) + G4_CB * 4));
          // all 64 phasors of the stage first, in registers: the wait for the previous stage's MMAs (which
          // still read the single-buffered A columns) then overlaps with this stage's work
          unsigned pk[4][16];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            g4_block(wn0a, rot[j], idx[j], off[j], pk[j]);
            g4_block(wn0b, rot[j], idx[j], off[j], pk[j] + 8);
          }
          if (sk >= 1) mbar_wait_u(my_empty_u, (sk - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 4; j++) tmem_st16(tmem_grp + j * G4_TCOLS + 16 + lane_base, pk[j]);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          // the group's named barrier.  arrive_split (IDGB200_TC4_SPLIT=1, experimental): quadrants 1-3 only
          // arrive and go on to the next stage's phasors - they meet the single-buffered A columns again at
          // the empty barrier - and only the issuing warp waits for all four
          if (arrive_split && q4 != 0)
            asm volatile("bar.arrive %0, 128;" ::"r"(1 + grp) : "memory");
          else
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          if (q4 == 0) {
            mbar_wait_u(bfull_u + slot2 * 8, ring_phase);
            mbar_wait_u(bfull_u + slot2 * 8 + 8, ring_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
              const unsigned long long dba = db0 + (unsigned long long)(slot2 * (G4_B_SLOT >> 4));
              const unsigned long long dbb = dba + (unsigned long long)(G4_B_SLOT >> 4);
              const unsigned acc = k > 0 ? 1u : 0u;
              for (int j = 0; j < gtiles; j++) {
                const unsigned td = tmem_grp + j * G4_TCOLS;
                umma_f16_ts4(td, td + 16, dba, idesc, acc);
                umma_f16_ts4(td, td + 24, dbb, idesc, 1u);
              }
              umma_commit_u(my_empty_u);
              if ((slot2 & 7u) == 6u) umma_commit_u(bempty_u + (slot2 >> 3) * 8);   // half ring consumed
              if (k == last_k) umma_commit_u(done_u);
            }
            __syncwarp();
          }
          slot2 += 2;
          if (slot2 == (unsigned)G4_NB) { slot2 = 0; ring_phase ^= 1u; }
        }
      }

      // ---- epilogue: every thread owns its 4 pixels' accumulators (gridder_reference.cpp:84-110)
      if (nstages > 0) {
        mbar_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const size_t plane = (size_t)npix;
      const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
      const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
      float2 *out = const_cast<float2 *>(a.subgrids) + (size_t)s * NR_POL * plane;
      for (int j = 0; j < gtiles; j++) {
        unsigned r[16];
        if (nstages > 0) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(tmem_grp + j * G4_TCOLS + lane_base));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int i = 0; i < 16; i++) r[i] = 0u;
        }
        const int pixel = pix0 + (4 * grp + j) * 128 + q4 * 32 + lane;
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
  } else {
    // ------------------------------------------------------------------ B builder warp (as gridder_tc.cu)
    // lane = (kc, n): one 16-byte chunk = 4 channels x (cos-row, sin-row) of column n = (hi|lo, pol, re|im)
    const int nrow = lane & 15, kc = lane >> 4, lo = nrow >> 3, p = (nrow >> 1) & 3, im = nrow & 1;
    auto load_b = [&](int kk, float2 (&raw)[4]) {
      const int t = kk / ncb, cb = kk - t * ncb;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int c = cb * G4_CB + kc * 4 + i;
        raw[i] = c < C ? __ldg(&g_vis[((size_t)t * C + c) * NR_POL + p]) : make_float2(0.f, 0.f);
      }
    };
    float2 raw[4];
    if (nstages > 0) load_b(0, raw);
    for (int k = 0; k < nstages; k++) {
      const int slot = k % G4_NB;
      if ((k & 7) == 0 && k >= G4_NB) mbar_wait(&bempty[(k >> 3) & 1], ((k / G4_NB) - 1) & 1);
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
      *reinterpret_cast<uint4 *>(sB + slot * G4_B_SLOT + kc * B_CHUNK_BYTES + nrow * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
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

// regular_flag: written by launch_gridder_regular_check on the same stream; the kernel is a no-op when 0
cudaError_t launch_gridder_tc4(const KernelArgs &a, const int *d_regular_flag, cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  const int npix = a.subgrid_size * a.subgrid_size;
  const int tiles_total = (npix + 127) / 128;
  const int slabs = (tiles_total + G4_TILES - 1) / G4_TILES;
  const int tiles_per_cta = tiles_total > 64 ? G4_TILES : (tiles_total + slabs - 1) / slabs;
  const int nslabs = (tiles_total + tiles_per_cta - 1) / tiles_per_cta;
  const int ncb = (a.nr_channels + G4_CB - 1) / G4_CB;
  const size_t smem = (size_t)G4_NB * G4_B_SLOT + (2 + G4_NB + 3) * 8 + 8 + 48 + (size_t)ncb * G4_CB * 4;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  auto k = gridder_tc4_kernel;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const char *split = std::getenv("IDGB200_TC4_SPLIT");
  k<<<dim3((unsigned)a.nr_subgrids * nslabs), dim3(G4_THREADS), smem, stream>>>(a, nslabs, tiles_per_cta, d_regular_flag,
                                                                                 split && split[0] == '1' ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace idgb200
