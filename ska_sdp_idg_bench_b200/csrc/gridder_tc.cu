// IDG gridder on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// north_star: "Tensor cores: used only if ncu shows that a shared-memory-generated phasor
// tile contracted against the visibility tile beats the FP32/SFU path."  It does
// (profiles/r01_tc_*.log): the FP32 kernel (gridder.cu) is bound by the
// dispatch port at ~26 cycles per (pixel, timestep, channel) item, of which 16 are the
// complex multiply-adds; moving those onto the tensor pipe leaves 2 MUFU + 3 FP32-pipe
// instructions per item, i.e. the XU (17.3 cycles/item, tools/xu_mix.cu) becomes the roof.
//
//   D[pixel][n] += A[pixel][k] * B[k][n]           one CTA = one slab of <= 8 tiles x 128 pixels
//     k = (visibility v, {cos, sin})                  A: fp16 phasors, produced by MUFU / FP32 polynomial
//     n = (hi|lo, pol, re|im)  -> N = 16              B: visibilities split into fp16 hi + lo
//       B[(v,cos)][re,pol] =  vr    B[(v,sin)][re,pol] = -vi
//       B[(v,cos)][im,pol] =  vi    B[(v,sin)][im,pol] =  vr
//     D: fp32 accumulators in TMEM, 16 columns per tile; hi and lo columns are added in
//        the epilogue, so the visibilities keep ~22 significant bits; the phasors are rounded
//        to fp16 (11 bits: 2.4e-4 relative, better than TF32's 10), which adds ~6e-5 rel-RMS
//        to the result - the same class as the fast-sincos error (tests: FAST tolerance only).
//
// Every producer warp is its own pipeline (no block-wide or cross-warp barrier in the main loop):
//   producer warp w owns the whole 128-pixel tile w (rows lane + 32 j, j = 0..3).  Per stage =
//     (timestep, block of 8 channels) each thread makes 32 phasors: phase exactly as in the FP32
//     kernel (bit-identical angle), __sincosf or the polynomial, pack to half2, one 16-byte
//     st.shared per 4 visibilities into the K-major, no-swizzle core-matrix layout; then
//     fence.proxy.async, and one elected lane of the same warp issues the tcgen05.mma of the stage
//     (M=128, N=16, K=16, fp16 in, fp32 out) and commits it to the warp's private empty barrier.
//     A producer only ever waits for its own MMA of two stages ago, which has long retired.
//     (The first version of this kernel - one MMA warp, one full[] barrier per stage for all
//     producers - spent 29 % of its issue slots in try_wait spins that share the MIO queue with
//     the MUFUs: profiles/r01_tc_gen1_ncu.txt; this layout runs the XU at 85 %.)
//   B builder warp: the only shared state is the B operand, a 16-slot ring (8 KB) released half a
//     ring at a time by tcgen05.commit from every tile;
//   epilogue: tcgen05.ld of the accumulators (a warp can only read its own 32-lane quadrant),
//     A-terms, taper, coalesced stores - identical math to gridder.cu.
// Every mbarrier wait is bounded (trap instead of hang) and carries a suspend-time hint, so that a
// waiting warp sleeps in the barrier unit instead of polling through the shared issue port.
//
// On top of that, selected by template / launch parameters:
//   recur - blocks of 8 equally spaced channels (checked per block, common.cuh: linear_channels) get
//           their phasors from the block's first channel: one complex rotation (FMUL2 + FFMA2) for
//           the second, then the three-term recurrence ph[c+1] = 2 cos(delta) ph[c] - ph[c-1], one
//           FFMA2 per phasor (the reference's gridder_v8 rotates every channel and assumes the
//           spacing unconditionally): the XU stops being the roof;
//   WIDE  - both channel blocks of a timestep in the tile's two A buffers as one K = 32 stage; when
//           every block is linear with one spacing and their number is even (the benchmark, SKA-Low)
//           the stage loop is a second, lean one: three counters of per-stage state instead of
//           ~170 instructions of flags, slot / phase arithmetic and rematerialised addresses;
//   SPLIT - fp16 hi + lo phasors (second A buffer, second MMA): FP32-class accuracy;
//   MASK16 - where the recurrence does not apply, part of the channels' phasors from an FP32
//           polynomial instead of MUFU.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace idgb200 {

namespace {

constexpr int T2_MAX_TILES = 8;                  // M-tiles (128 pixels) = producer warps per CTA
constexpr int T2_CB = 8;                         // channels per stage -> K = 16
constexpr int T2_A_STAGE = 2 * A_CHUNK_BYTES;    // 4 KB per tile and stage
constexpr int T2_STAGES = 2;
constexpr int T2_B_SLOT = 2 * B_CHUNK_BYTES;     // 512 B
constexpr int T2_NB = 16;                        // B ring slots
constexpr int T2_THREADS = (T2_MAX_TILES + 1) * 32;

template <unsigned MASK8, bool SPLIT>
__device__ __forceinline__ void tc_produce(unsigned char *A, const float *wn8, const float (&idx)[4],
                                           const float (&idxr)[4], const float (&off)[4],
                                           const float (&offr)[4], const int lane) {
#pragma unroll
  for (int kc = 0; kc < 2; kc++) {
    const float4 wq = *reinterpret_cast<const float4 *>(wn8 + 4 * kc);
    const float wn[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      unsigned pk[4], pl[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 ph = ((MASK8 >> (kc * 4 + i)) & 1u)
                              ? phasor_poly(__fmaf_rn(-idxr[j], wn[i], offr[j]))
                              : phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(-idx[j], wn[i], off[j]));  // :69
        pack_phasor<SPLIT>(ph, pk[i], pl[i]);
      }
      *reinterpret_cast<uint4 *>(A + kc * A_CHUNK_BYTES + (lane + 32 * j) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if (SPLIT)
        *reinterpret_cast<uint4 *>(A + T2_A_STAGE + kc * A_CHUNK_BYTES + (lane + 32 * j) * 16) =
            make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
  }
}

// The same stage for a block of equally spaced channels (common.cuh: linear_channels): one sincos
// for the first channel (the same angle as above, bit for bit), the per-channel rotation
// d = e^{-i idx dw} (rot) for the second, and from there the three-term recurrence of equally spaced
// angles, ph[c+1] = 2 cos(delta) ph[c] - ph[c-1]: ONE packed FFMA2 per phasor (2 dispatch cycles)
// where the complex multiplication ph[c] * d costs FMUL2 + FFMA2 (4).  A rounding error e made at one
// step comes back as e sin(k delta) / sin(delta) <= k e after k steps, so over the 6 steps of a block
// the phasor stays within ~2e-6 of the rotated one - 1 % of the fp16 rounding of the operand it
// becomes (tools/hrot.cu has the dispatch costs, tests/test_gpu_parity.py the parity cases).
template <bool SPLIT>
__device__ __forceinline__ void tc_produce_linear(unsigned char *A, const float wn0, const float2 (&rot)[4],
                                                  const float (&idx)[4], const float (&off)[4], const int lane) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float2 prev = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(-idx[j], wn0, off[j]));   // :69
    const float2 d = rot[j];
    const float2 dxx = make_float2(d.x, d.x), dny = make_float2(-d.y, d.y);
    const float c2 = __fadd_rn(d.x, d.x);
    const float2 cc = make_float2(c2, c2);
    unsigned pk[8], pl[8];
    pack_phasor<SPLIT>(prev, pk[0], pl[0]);
    // (x, y) * d = (x, y) * (dx, dx) + (y, x) * (-dy, dy): FMUL2 + FFMA2 with free operand modes
    // (broadcast scalar, LO_HI swizzle)
    float2 cur = ffma2(make_float2(prev.y, prev.x), dny, __fmul2_rn(prev, dxx));
    pack_phasor<SPLIT>(cur, pk[1], pl[1]);
#pragma unroll
    for (int i = 2; i < 8; i++) {
      const float2 nxt = ffma2(cur, cc, make_float2(-prev.x, -prev.y));
      pack_phasor<SPLIT>(nxt, pk[i], pl[i]);
      prev = cur;
      cur = nxt;
    }
    if ((IDGB200_ABLATE & 2) && !ablate_never()) continue;
    *reinterpret_cast<uint4 *>(A + (lane + 32 * j) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4 *>(A + A_CHUNK_BYTES + (lane + 32 * j) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    if (SPLIT) {
      *reinterpret_cast<uint4 *>(A + T2_A_STAGE + (lane + 32 * j) * 16) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      *reinterpret_cast<uint4 *>(A + T2_A_STAGE + A_CHUNK_BYTES + (lane + 32 * j) * 16) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
    }
  }
}

// MASK16: bit c set -> channel c of every 16-channel group uses phasor_poly instead of MUFU
// recur: blocks of equally spaced channels use the rotation recurrence instead
// SPLIT: fp16 hi + lo phasors: the tile's two A buffers hold the hi and the lo block of ONE stage (two
//        MMAs against the same B slot), so a warp waits for its own MMAs every stage - the other
//        warps of the sub-partition cover that latency
// WIDE:  the two A buffers hold the two channel blocks of one K = 32 stage (16 channels): the
//        per-stage bookkeeping is amortised over 64 items instead of 32, single-buffered like SPLIT
template <unsigned MASK16, bool SPLIT, bool WIDE, bool LIST>
__device__ __forceinline__ void gridder_tc_body(const KernelArgs &a, const int tiles_per_cta, const int tmem_cols, const int recur,
                                                const int s_local, const int slab) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int N = a.subgrid_size, C = a.nr_channels, npix = N * N;
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler too
  const int nwarps = blockDim.x >> 5, NW = nwarps - 1;        // producer warps; warp NW builds B
  const int pix0 = slab * tiles_per_cta * 128;
  const int ntiles = min(tiles_per_cta, (npix - pix0 + 127) / 128);
  if (ntiles <= 0) return;

  unsigned char *sA = smem;                                              // [tile][stage][4 KB]
  unsigned char *sB = sA + tiles_per_cta * T2_STAGES * T2_A_STAGE;       // [T2_NB][512 B]
  unsigned long long *aempty = reinterpret_cast<unsigned long long *>(sB + T2_NB * T2_B_SLOT);  // [tile][stage]
  unsigned long long *bfull = aempty + T2_MAX_TILES * T2_STAGES;         // [T2_NB]
  unsigned long long *bempty = bfull + T2_NB;                            // [2] half rings
  unsigned long long *done = bempty + 2;
  unsigned *s_tmem = reinterpret_cast<unsigned *>(done + 1);
  float *s_red = reinterpret_cast<float *>(s_tmem + 2);     // [12] block reduction scratch + scale
  float *s_wn = s_red + 12;                                 // [ncb * 8], zero padded
  float *s_dw = s_wn + ((a.nr_channels + T2_CB - 1) / T2_CB) * T2_CB;   // [ncb] channel spacing of a block, if linear
  int *s_lin = reinterpret_cast<int *>(s_dw + (a.nr_channels + T2_CB - 1) / T2_CB);   // [ncb]

  const SubgridCtx ctx = load_ctx(a, s);
  const int nt = ctx.nr_timesteps;
  const int ncb = (C + T2_CB - 1) / T2_CB;
  const int nstages = nt * ncb;

  for (int c = tid; c < ncb * T2_CB; c += blockDim.x) s_wn[c] = c < C ? a.wavenumbers[c] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < T2_MAX_TILES * T2_STAGES; i++) mbar_init(&aempty[i], 1);
    for (int i = 0; i < T2_NB; i++) mbar_init(&bfull[i], 1);
    mbar_init(&bempty[0], ntiles);
    mbar_init(&bempty[1], ntiles);
    mbar_init(done, ntiles);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(tmem_cols));
    if (!LIST) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);   // a list-mode CTA allocates again
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = *s_tmem;
  // per 8-channel block: equally spaced? (read by the producers after the next __syncthreads)
  for (int cb = tid; cb < ncb; cb += blockDim.x) {
    float dw;
    const bool lin = linear_channels(s_wn, cb * T2_CB, min(T2_CB, C - cb * T2_CB), &dw);
    s_dw[cb] = dw;
    s_lin[cb] = (recur && lin) ? 1 : 0;
  }

  const float *g_uvw = reinterpret_cast<const float *>(a.uvw) + (size_t)ctx.time_offset * 3;
  const float2 *g_vis = a.visibilities + (size_t)ctx.time_offset * C * NR_POL;

  // power-of-two scaling of this subgrid's visibilities into fp16 range (see above)
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
      for (int i = 1; i < nwarps; i++) amax = fmaxf(amax, s_red[i]);
      const unsigned eb = (__float_as_uint(amax) >> 23) & 0xffu;          // biased exponent
      const bool ok = eb >= 14u && eb <= 253u;                             // finite, not tiny
      s_red[10] = ok ? __uint_as_float((267u - eb) << 23) : 1.f;           // 2^(13 - E)
      s_red[11] = ok ? __uint_as_float((eb - 13u) << 23) : 1.f;            // 2^(E - 13)
    }
    __syncthreads();
  }
  const float vis_scale = s_red[10], vis_unscale = s_red[11];
  // every block linear with the same spacing (bitwise)?  (s_lin / s_dw are complete: two barriers ago)
  bool same_dw = true;
  for (int cb = 0; cb < ncb; cb++) same_dw = same_dw && s_lin[cb] && s_dw[cb] == s_dw[0];

  if (warp < NW) {
    // ------------------------------------------------------------------ producers (+ their own MMA)
    const int tile = warp;
    if (tile < ntiles) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6) = 1, A = B = F16 (0),
      // both K-major (0), N >> 3 at [17,23), M >> 4 at [24,29)
      const unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      float l[4], m[4], n[4], off[4], offr[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int q = min(pix0 + tile * 128 + lane + 32 * j, npix - 1);
        const int y = q / N, x = q - y * N;
        l[j] = compute_l(x, N, a.image_size);
        m[j] = compute_l(y, N, a.image_size);
        n[j] = compute_n(l[j], m[j]);
        // gridder_reference.cpp:64 as the CPU binary contracts it
        off[j] = __fmaf_rn(ctx.w_offset, n[j], __fmaf_rn(ctx.u_offset, l[j], __fmul_rn(ctx.v_offset, m[j])));
        offr[j] = __fmul_rn(off[j], 0.15915494309189535f);
      }
      unsigned char *A_tile = sA + tile * T2_STAGES * T2_A_STAGE;
      unsigned long long da0 = smem_desc(smem_u32(A_tile), A_CHUNK_BYTES, 128);
      unsigned long long db0 = smem_desc(smem_u32(sB), B_CHUNK_BYTES, 128);
      unsigned tmem_d = tmem_base + tile * 16;
      // barrier addresses as 32-bit shared addresses, computed once; the empty asm statements make
      // the values opaque, so that the compiler keeps them instead of recomputing ~30 uniform
      // instructions per stage in the issue path (which is what bounds this kernel)
      unsigned my_empty_u = smem_u32(aempty + tile * T2_STAGES), bfull_u = smem_u32(bfull),
               bempty_u = smem_u32(bempty), done_u = smem_u32(done);
      asm volatile("" : "+l"(da0), "+l"(db0), "+r"(tmem_d), "+r"(my_empty_u), "+r"(bfull_u), "+r"(bempty_u), "+r"(done_u));
      float un = 0.f, vn = 0.f, wnx = 0.f;   // uvw of the next timestep, fetched one timestep ahead
      if (nt > 0) { un = __ldg(&g_uvw[0]); vn = __ldg(&g_uvw[1]); wnx = __ldg(&g_uvw[2]); }
      unsigned k = 0, sk = 0;   // channel blocks (= B slots) and stages done; unsigned: masks and shifts
      if (WIDE && !SPLIT && same_dw && !(ncb & 1)) {
        // ---- the regular case (an even number of linear channel blocks, one spacing): every stage is
        // two blocks, nothing about it depends on data.  The generic loop below spends as many issue
        // slots per stage on block flags, slot / phase arithmetic and rematerialised shared-memory
        // addresses (~170 instructions, ncu source page) as on 20 of the 64 phasors it makes; here the
        // per-stage state is three counters and the addresses are opaque registers.
        unsigned wn_u = smem_u32(s_wn);
        const float dw0 = s_dw[0];
        asm volatile("" : "+r"(wn_u));
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
            if (sk >= 1) mbar_wait_u(my_empty_u, (sk - 1) & 1);
            float wn0a, wn0b;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wn0a) : "r"(wn_u + (unsigned)cb0 * (T2_CB * 4)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wn0b) : "r"(wn_u + (unsigned)cb0 * (T2_CB * 4) + T2_CB * 4));
            tc_produce_linear<false>(A_tile, wn0a, rot, idx, off, lane);
            tc_produce_linear<false>(A_tile + T2_A_STAGE, wn0b, rot, idx, off, lane);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_wait_u(bfull_u + slot2 * 8, ring_phase);
            mbar_wait_u(bfull_u + slot2 * 8 + 8, ring_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            __syncwarp();
            if (elect_one()) {
              const unsigned long long db = db0 + (unsigned long long)(slot2 * (T2_B_SLOT >> 4));
              if (!(IDGB200_ABLATE & 1)) {
              umma_f16(tmem_d, da0, db, idesc, k > 0 ? 1u : 0u);
              umma_f16(tmem_d, da0 + (unsigned long long)(T2_A_STAGE >> 4), db + (unsigned long long)(T2_B_SLOT >> 4), idesc, 1u);
              }
              umma_commit_u(my_empty_u);
              if ((slot2 & 7u) == 6u) umma_commit_u(bempty_u + (slot2 >> 3) * 8);   // half ring consumed
              if (k == last_k) umma_commit_u(done_u);
            }
            __syncwarp();
            slot2 += 2;
            if (slot2 == (unsigned)T2_NB) { slot2 = 0; ring_phase ^= 1u; }
          }
        }
      } else
      for (int t = 0; t < nt; t++) {
        const float u = un, v = vn, w = wnx;
        if (t + 1 < nt) { un = __ldg(&g_uvw[3 * t + 3]); vn = __ldg(&g_uvw[3 * t + 4]); wnx = __ldg(&g_uvw[3 * t + 5]); }
        float idx[4], idxr[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {  // gridder_reference.cpp:61 as contracted by the CPU binary
          idx[j] = __fmaf_rn(w, n[j], __fmaf_rn(u, l[j], __fmul_rn(v, m[j])));
          idxr[j] = __fmul_rn(idx[j], 0.15915494309189535f);   // in revolutions, for phasor_poly
        }
        float2 rot[4];
#pragma unroll
        for (int j = 0; j < 4; j++) rot[j] = make_float2(1.f, 0.f);
        for (int cb0 = 0; cb0 < ncb; cb0 += WIDE ? 2 : 1, sk++) {
          const int nb = WIDE ? min(2, ncb - cb0) : 1;           // channel blocks of this stage
          const unsigned stage = (SPLIT || WIDE) ? 0u : (sk & 1u), use = (SPLIT || WIDE) ? sk : (sk >> 1);
          if (use >= 1) mbar_wait_u(my_empty_u + stage * 8, (use - 1) & 1);
#pragma unroll
          for (int b = 0; b < (WIDE ? 2 : 1); b++) {
            if (b < nb) {
              const int cb = cb0 + b;
              unsigned char *A = A_tile + (stage + b) * T2_A_STAGE;
              if (s_lin[cb]) {
                // the rotation step -idx * dw: once per timestep when every block has the same spacing
                if (!same_dw || cb == 0) {
#pragma unroll
                  for (int j = 0; j < 4; j++) rot[j] = phasor<IDGB200_SINCOS_FAST>(__fmul_rn(-idx[j], s_dw[cb]));
                }
                tc_produce_linear<SPLIT>(A, s_wn[cb * T2_CB], rot, idx, off, lane);
              }
              else if ((MASK16 >> 8) == (MASK16 & 0xffu) || !(cb & 1))
                tc_produce<(MASK16 & 0xffu), SPLIT>(A, s_wn + cb * T2_CB, idx, idxr, off, offr, lane);
              else
                tc_produce<(MASK16 >> 8), SPLIT>(A, s_wn + cb * T2_CB, idx, idxr, off, offr, lane);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          // the whole warp waits for the B slot(s) (one SYNCS either way) and stays converged, so the
          // descriptors live in uniform registers and one elected lane issues MMA + commits
#pragma unroll
          for (int b = 0; b < (WIDE ? 2 : 1); b++)
            if (b < nb) mbar_wait_u(bfull_u + ((k + b) % (unsigned)T2_NB) * 8, ((k + b) / (unsigned)T2_NB) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          __syncwarp();
          if (elect_one()) {
#pragma unroll
            for (int b = 0; b < (WIDE ? 2 : 1); b++) {
              if (b < nb) {
                const unsigned kb = k + b, slot = kb % (unsigned)T2_NB;
                umma_f16(tmem_d, da0 + (unsigned long long)((stage + b) * (T2_A_STAGE >> 4)),
                         db0 + (unsigned long long)(slot * (T2_B_SLOT >> 4)), idesc, kb > 0 ? 1u : 0u);
                if (SPLIT)
                  umma_f16(tmem_d, da0 + (unsigned long long)(T2_A_STAGE >> 4),
                           db0 + (unsigned long long)(slot * (T2_B_SLOT >> 4)), idesc, 1u);
              }
            }
            umma_commit_u(my_empty_u + stage * 8);
#pragma unroll
            for (int b = 0; b < (WIDE ? 2 : 1); b++)
              if (b < nb && ((k + b) & 7) == 7) umma_commit_u(bempty_u + (((k + b) >> 3) & 1) * 8);
            if (k + nb == (unsigned)nstages) umma_commit_u(done_u);
          }
          __syncwarp();
          k += nb;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ B builder warp
    // lane = (kc, n): one 16-byte chunk = 4 channels x (cos-row, sin-row) of column n = (hi|lo, pol, re|im)
    const int nrow = lane & 15, kc = lane >> 4, lo = nrow >> 3, p = (nrow >> 1) & 3, im = nrow & 1;
    auto load_b = [&](int kk, float2 (&raw)[4]) {
      const int t = kk / ncb, cb = kk - t * ncb;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int c = cb * T2_CB + kc * 4 + i;
        raw[i] = c < C ? __ldg(&g_vis[((size_t)t * C + c) * NR_POL + p]) : make_float2(0.f, 0.f);
      }
    };
    float2 raw[4];
    if (nstages > 0) load_b(0, raw);
    for (int k = 0; k < nstages; k++) {
      const int slot = k % T2_NB;
      if ((k & 7) == 0 && k >= T2_NB) mbar_wait(&bempty[(k >> 3) & 1], ((k / T2_NB) - 1) & 1);
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
      *reinterpret_cast<uint4 *>(sB + slot * T2_B_SLOT + kc * B_CHUNK_BYTES + nrow * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bfull[slot]);
    }
  }

  // ---- epilogue (producer warps): accumulators -> A-terms, taper, store (gridder_reference.cpp:84-110)
  const int groups = NW >> 2;   // sets of 4 warps, one per TMEM lane quadrant
  if (warp < 4 * groups) {
    if (nstages > 0) {
      mbar_wait(done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const size_t plane = (size_t)npix;
    const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
    const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
    float2 *out = const_cast<float2 *>(a.subgrids) + (size_t)s * NR_POL * plane;
    const int q4 = warp & 3;   // a warp reads TMEM lanes 32 (warp % 4) .. +31
    for (int tile = warp >> 2; tile < ntiles; tile += groups) {
      unsigned r[16];
      if (nstages > 0) {
        const unsigned taddr = tmem_base + ((unsigned)(q4 * 32) << 16) + tile * 16;
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
      const int pixel = pix0 + tile * 128 + q4 * 32 + lane;
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

// LIST = false: CTA (s, slab) = blockIdx.x; LIST = true: a fixed number of CTAs loop over the subgrids of a.list
// (what gridder_sep.cu left: the work is rare, the launch must be cheap when the list is empty)
template <unsigned MASK16, bool SPLIT, bool WIDE, bool LIST>
__global__ void __launch_bounds__(T2_THREADS, 3)
gridder_tc_kernel(const KernelArgs a, const int slabs, const int tiles_per_cta, const int tmem_cols, const int recur) {
  if (!LIST) {
    const int s_local = blockIdx.x / slabs;
    gridder_tc_body<MASK16, SPLIT, WIDE, false>(a, tiles_per_cta, tmem_cols, recur, s_local, blockIdx.x - s_local * slabs);
  } else {
    const int total = a.list[0] * slabs;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int i = item / slabs;
      gridder_tc_body<MASK16, SPLIT, WIDE, true>(a, tiles_per_cta, tmem_cols, recur, a.list[1 + i], item - i * slabs);
      __syncthreads();
    }
  }
}

}  // namespace

// FAST sincos only: the fp16 phasor operand is a FAST-class approximation.
// poly: 3 = where the recurrence does not apply, 6 of every 16 channels' phasors by FP32 polynomial, the rest by
//       MUFU (the measured optimum: DESIGN.md 4.5); 10 = fp16 hi + lo phasors (FP32-class accuracy), all by
//       MUFU / rotation; 11 = as 3 with 16 channels (K = 32) per stage
// recur: blocks of 8 equally spaced channels get their phasors by rotation from the first one
cudaError_t launch_gridder_tc(const KernelArgs &a, int poly, bool recur, cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  const int npix = a.subgrid_size * a.subgrid_size;
  const int tiles_total = (npix + 127) / 128;
  const int slabs = (tiles_total + T2_MAX_TILES - 1) / T2_MAX_TILES;
  const int tiles_per_cta = tiles_total > 64 ? T2_MAX_TILES : (tiles_total + slabs - 1) / slabs;
  const int nslabs = (tiles_total + tiles_per_cta - 1) / tiles_per_cta;
  const int producer_warps = tiles_per_cta < 4 ? 4 : tiles_per_cta;
  int tmem_cols = 32;
  while (tmem_cols < tiles_per_cta * 16) tmem_cols *= 2;
  const int ncb = (a.nr_channels + T2_CB - 1) / T2_CB;
  const size_t smem = (size_t)tiles_per_cta * T2_STAGES * T2_A_STAGE + T2_NB * T2_B_SLOT +
                      (T2_MAX_TILES * T2_STAGES + T2_NB + 3) * 8 + 8 + 48 + (size_t)ncb * (T2_CB + 2) * 4;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  void (*k)(const KernelArgs, int, int, int, int) = nullptr;
  const bool list = a.list != nullptr;
  switch (poly) {
    case 3:    // 6 of 16 by polynomial: channels 1,4,6 | 9,12,14
      k = list ? gridder_tc_kernel<0x5252u, false, false, true> : gridder_tc_kernel<0x5252u, false, false, false>;
      break;
    case 10:   // fp16 hi + lo phasors
      if (list) return cudaErrorInvalidValue;
      k = gridder_tc_kernel<0x0000u, true, false, false>;
      break;
    case 11:   // as 3 with K = 32 stages
      k = list ? gridder_tc_kernel<0x5252u, false, true, true> : gridder_tc_kernel<0x5252u, false, true, false>;
      break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long ctas = (long long)a.nr_subgrids * nslabs;
  k<<<dim3((unsigned)(list && ctas > LIST_MODE_CTAS ? LIST_MODE_CTAS : ctas)), dim3((producer_warps + 1) * 32), smem, stream>>>(
      a, nslabs, tiles_per_cta, tmem_cols, recur ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace idgb200
