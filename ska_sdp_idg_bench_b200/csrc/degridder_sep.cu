// IDG degridder in row-column form on tcgen05 + TMEM (variant 30, the default for FAST sincos and
// subgrids of up to 64 x 64 pixels).
//
// The transpose of gridder_sep.cu.  With the phase split into a column and a row part (see there; the
// degridder's sign: degridder_reference.cpp:96-112),
//   vis_v[p] = sum_y Y_v(y) [ sum_x X_v(x) P'[y][x][p] ],      P' = A1 (sph . subgrid) A2^H   (:38-74)
//   X_v(x) = exp i[(u l_x + w f(l_x^2)) k_c - (u_off l_x + w_off f(l_x^2))],   Y_v(y) likewise with v, m_y
// (subgrids whose dropped phase term |gamma| r exceeds SEP_PHASE_TOL are left to the per-pixel kernel
// launched behind this one).  The inner sum is a GEMM per tile of 128 visibilities,
//   Q[v][(y, p, re|im)] = A[v][(x, cos|sin)] * B[(x, cos|sin)][(y, p, re|im)]      M = 128, K = 2 N, N = 8 N
//     A = X in fp16 hi + lo (32 phasors per visibility instead of 1024: three-term recurrence over equally
//         spaced channels, else one sincos per channel),
//     B = P' in fp16 hi + lo, built once per subgrid (power-of-two scaling into the range of the parts),
//     an error-compensated product per K step: hi hi as an fp16 MMA, the two cross products lo hi + hi lo as ONE e4m3
//     MMA (DS_FP8_LO below): the reference's smooth degridder input needs more than fp16's 11 bits on both
//     operands (its visibilities are small sums of large terms: DESIGN.md 4.6),
// and the outer sum runs on the CUDA cores straight out of TMEM: thread = visibility (TMEM lane = row), one
// sincos per (visibility, row y), four complex multiply-adds, then one coalesced 32-byte store per visibility:
// every visibility of the subgrid's time range is written exactly once.
//
// Two kernels with this arithmetic (bit-identical results):
//   degridder_sep_kernel       CTA = one subgrid, a tile's phases one after the other (A rows, MMAs, sum over the rows);
//                              up to 32 x 32 pixels 8 warps and 2 CTAs per SM; above, slabs of 32 rows and 16 warps
//   degridder_sep_pipe_kernel  one persistent CTA per SM, warp-specialised (producers / issuer / consumers / setup of the
//                              next subgrid), two A buffers, two accumulators, two B buffers: the default up to 32 x 32
// The rows of the GEMM are enumerated in blocks of 8 channels (row = (timestep * blocks + block) * 8 + channel, channels
// padded to blocks), 16 blocks per tile of 128 visibilities; a warp makes the A rows of whole blocks (lane = column x; the
// 8 phasors of a block unrolled), one thread issues the tile's MMAs, and the two (four) warps that may read TMEM lane
// quadrant q each sum their share of the rows y for its 32 visibilities; the shares meet in shared memory.
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace idgb200 {

namespace {

constexpr int DS_A_CH = 128 * 16 + 16;           // one 16-byte K chunk of 128 rows, padded: the 8 chunks a warp's
                                                 // 4-byte stores touch fall into different banks
constexpr float SEP_PHASE_TOL = 1e-4f;           // largest dropped phase |gamma| r (rad), as in gridder_sep.cu
constexpr int DS_UVW_STAGED = 384;               // timesteps whose uvw are staged in shared memory (longer subgrids read global)

__device__ __forceinline__ unsigned pack_h2(const float lo, const float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const unsigned *>(&h);
}
__device__ __forceinline__ float residual_h(const float x, const unsigned short h) {
  const unsigned short minus_one = 0xbc00u;
  float r;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(h), "h"(minus_one), "f"(x));
  return r;
}

// The second word of an operand element.  DS_FP8_LO = 0: the fp16 rounding residuals (lo), and three fp16 products per K
// step (hi hi, lo hi, hi lo).  DS_FP8_LO = 1 (default): the two cross products only have to be good to a few bits - they
// are 2^-11 of the result - so they run as ONE kind::f8f6f4 MMA over four e4m3 values per element,
//   A word = (lo8_c, lo8_s, hi8_c, hi8_s)   against   B word = (hi8_c, hi8_s, lo8_c, lo8_s)
// with hi8 = e4m3(fp16 hi), lo8 = e4m3(x - fp16 hi): K = 32 e4m3 take the bytes and the clocks of K = 16 fp16, so a tile
// is 8 MMAs instead of 12 (the kernel sits at the board's power cap and the tensor pipe is its largest consumer).  The
// operands are scaled so that both parts sit in e4m3's range: A by 2^8 (|X| <= 1 -> hi8 <= 256 < 448, lo8 <= 2^-4 against
// a subnormal step of 2^-9), B to [2^7, 2^8); an operand then keeps ~2^-17 of its largest value instead of fp16's 2^-12
// (tools/sep_prototype.py --fp8: config 1 rel-RMS 1.5e-4 against 1.3e-4 for three fp16 products and 7.1e-4 for one).
#ifndef DS_FP8_LO
#define DS_FP8_LO 1
#endif
constexpr float DS_A_SCALE = DS_FP8_LO ? 256.f : 1.f;
constexpr unsigned DS_B_EXP = DS_FP8_LO ? 7u : 13u;               // B is scaled to [2^DS_B_EXP, 2^(DS_B_EXP + 1))
constexpr unsigned DS_UNSCALE_EXP = DS_FP8_LO ? 15u : 13u;        // log2 of the product of the two scalings
constexpr float DS_UNSCALE_RAW = DS_FP8_LO ? 1.f / 256.f : 1.f;   // where B is not scaled (all-zero or non-finite pixels)

__device__ __forceinline__ unsigned e4m3x2(const float lo, const float hi) {
  return (unsigned)__nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ unsigned e4m3x2_h2(const unsigned h2) {
  __half2_raw r;
  r.x = (unsigned short)(h2 & 0xffffu);
  r.y = (unsigned short)(h2 >> 16);
  return (unsigned)__nv_cvt_halfraw2_to_fp8x2(r, __NV_SATFINITE, __NV_E4M3);
}
// A: second word of (first, second) = the K pair (cos, sin) of a column, given its fp16 word
__device__ __forceinline__ unsigned second_word_a(const float first, const float second, const unsigned hi) {
  const float r0 = residual_h(first, (unsigned short)(hi & 0xffffu)), r1 = residual_h(second, (unsigned short)(hi >> 16));
#if DS_FP8_LO
  return e4m3x2(r0, r1) | (e4m3x2_h2(hi) << 16);
#else
  return pack_h2(r0, r1);
#endif
}
// B: second word given the fp16 word and the residuals of its two halves
__device__ __forceinline__ unsigned second_word_b(const float r0, const float r1, const unsigned hi) {
#if DS_FP8_LO
  return e4m3x2_h2(hi) | (e4m3x2(r0, r1) << 16);
#else
  return pack_h2(r0, r1);
#endif
}
// the cross products of one K step (2 chunks): lo hi + hi lo
__device__ __forceinline__ void umma_cross(const unsigned d, const unsigned long long a_hi, const unsigned long long a_lo,
                                           const unsigned long long b_hi, const unsigned long long b_lo, const unsigned idesc) {
#if DS_FP8_LO
  // the same instruction descriptor: format code 0 is F16 for kind::f16 and E4M3 for kind::f8f6f4
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a_lo), "l"(b_lo), "r"(idesc) : "memory");
#else
  umma_f16(d, a_lo, b_hi, idesc, 1u);
  umma_f16(d, a_hi, b_lo, idesc, 1u);
#endif
}

// 16 consecutive accumulator columns of the thread's TMEM lane.  Two of these, not one .x32: tools/tmem_ld.cu measures
// 143 clocks per 32x32b.x32 load and at most 120 B/clock/SM however many warps issue them, against 39 clocks and 410
// B/clock/SM for .x16 - with .x32 the 128 KB accumulator of a tile took longer to read (~1600 clocks) than to compute.
__device__ __forceinline__ void tmem_ld16(const unsigned taddr, unsigned *r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}

// All 8 warps wait for the tile's MMAs (~1 us); a try_wait's own suspension is short (ncu: ~10 polls per warp and
// tile, 13 % of the kernel's instructions), so they sleep between polls and leave the issue slots to the other CTA
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long *bar_ptr, const unsigned parity) {
  const unsigned bar = smem_u32(bar_ptr);
  unsigned done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  for (int spin = 0; !done; spin++) {
    __nanosleep(150);
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1 << 22)) __trap();   // a lost arrival must fail loudly, not hang the GPU
  }
}

// P' of pixel q (degridder_reference.cpp:38-74): sph . subgrid, A1 . A2^H
__device__ __forceinline__ void pprime(const KernelArgs &a, const float2 *sub, const size_t at1, const size_t at2,
                                       const int q, const int N, const size_t plane, float2 *px) {
  const float sph = __ldg(&a.spheroidal[q]);
  const int src = subgrid_slot(q, N, a.flags);
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

// XPL: columns x per lane of the A rows (1: subgrids of up to 32 x 32 pixels, 2 CTAs per SM; 2: up to 64 x 64, one CTA
// per SM).  Subgrids of more than 32 rows are done in SLABS of 32 rows (8 x 32 = 256 accumulator columns each), one after
// the other in the same CTA: B is rebuilt for the slab, the tiles run again (the A rows are made again: they are the
// cheaper operand), and from the second slab on the partial visibilities are read back, added to and stored again by the
// thread that wrote them - no atomics, the same bits every run.  With one CTA per SM there is room for 16 warps: one
// block of A rows per warp, four warps per TMEM lane quadrant in the sum over the rows.
template <int XPL>
__global__ void __launch_bounds__(XPL == 1 ? 256 : 512, XPL == 1 ? 2 : 1)
degridder_sep_kernel(const KernelArgs a, const int tmem_cols, int *__restrict__ todo) {
  constexpr int NW = XPL == 1 ? 8 : 16, DS_THREADS = NW * 32;   // warps
  constexpr int NP = NW / 4;                                    // warps per TMEM lane quadrant = parts of the sum over the rows
  constexpr int BPW = 16 / NW;                                  // blocks of A rows per warp and tile
  extern __shared__ __align__(1024) unsigned char smem[];
  const int N = a.subgrid_size, C = a.nr_channels;
  const int s_local = blockIdx.x, s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int KC = N >> 2, KCp = (KC + 1) & ~1;      // 16-byte K chunks (4 columns x (cos, sin)); K steps of 2 chunks
  const int nslab = (N + 31) >> 5, ny_max = min(N, 32);
  const int b_ch = 8 * ny_max * 16 + 16;           // one K chunk of B (n = y * 8 + p * 2 + (re|im) of a slab), padded like DS_A_CH
  const int ncb = (C + 7) >> 3;

  unsigned char *sB = smem;                                             // [hi|lo][KCp][b_ch]
  unsigned char *sA = sB + 2 * KCp * b_ch;                              // [hi|lo][KCp][DS_A_CH]
  float4 *sGeo = reinterpret_cast<float4 *>(sA + 2 * KCp * DS_A_CH);    // [N] (m_y, f(m_y^2), offset_y, 0)
  float4 *sPart = sGeo + N;                                             // [NP - 1][128][2] partial sums of a quadrant's warps 1 .. NP - 1
  float *s_uvw = reinterpret_cast<float *>(sPart + (NP - 1) * 256);                // [DS_UVW_STAGED][3]
  unsigned long long *mma_done = reinterpret_cast<unsigned long long *>(s_uvw + DS_UVW_STAGED * 3);
  unsigned *s_tmem = reinterpret_cast<unsigned *>(mma_done + 1);
  float *s_red = reinterpret_cast<float *>(s_tmem + 2);                 // [24]: [0, 16) per warp, [16, 19) scale, unscale, verdict
  float *s_wn = s_red + 24;                                             // [ncb * 8]
  float *s_dw = s_wn + ncb * 8;                                         // [ncb]
  int *s_lin = reinterpret_cast<int *>(s_dw + ncb);                     // [ncb]

  const SubgridCtx ctx = load_ctx(a, s);
  const int nt = ctx.nr_timesteps;
  const int nblk = nt * ncb;                       // blocks of 8 rows
  const int ntiles = (nblk + 15) >> 4;
  const float *g_uvw = reinterpret_cast<const float *>(a.uvw) + (size_t)ctx.time_offset * 3;

  for (int c = tid; c < ncb * 8; c += DS_THREADS) s_wn[c] = c < C ? a.wavenumbers[c] : 0.f;
  const bool staged = nt <= DS_UVW_STAGED;   // uvw of the whole subgrid in shared memory: no global load in the tile loop
  if (staged)
    for (int i = tid; i < nt * 3; i += DS_THREADS) s_uvw[i] = __ldg(&g_uvw[i]);
  auto uvw_at = [&](int i) { return staged ? s_uvw[i] : __ldg(&g_uvw[i]); };
  if (tid < N) {
    const float m = compute_l(tid, N, a.image_size);
    const float n_y = compute_n(m, 0.f);
    sGeo[tid] = make_float4(m, n_y, __fmaf_rn(ctx.w_offset, n_y, __fmul_rn(ctx.v_offset, m)), 0.f);
  }
  if (tid == 0) {
    mbar_init(mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- separability check (gridder_sep.cu) and channel layout
  {
    float wmax = 0.f;
    for (int t = tid; t < nt; t += DS_THREADS) wmax = fmaxf(wmax, fabsf(__ldg(&g_uvw[3 * t + 2])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_red[warp] = wmax;
    for (int cb = tid; cb < ncb; cb += DS_THREADS) {
      float dw;
      s_lin[cb] = linear_channels(s_wn, cb * 8, min(8, C - cb * 8), &dw) ? 1 : 0;
      s_dw[cb] = dw;
    }
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < NW; i++) wmax = fmaxf(wmax, s_red[i]);
      float kmax = 0.f;
      for (int c = 0; c < C; c++) kmax = fmaxf(kmax, fabsf(s_wn[c]));
      const double l0 = (0.5 - (N / 2)) * (double)a.image_size / (double)N;
      const double s1 = l0 * l0;
      const double fn = s1 / (1.0 + sqrt(1.0 - s1)), s2 = 2.0 * s1;
      const double r = s2 > 1.0 ? 1.0 : fabs(s2 / (1.0 + sqrt(1.0 - s2)) - 2.0 * fn);
      const double gmax = (double)fabsf(ctx.w_offset) + (double)wmax * (double)kmax;
      const bool sep = gmax * r <= (double)SEP_PHASE_TOL && isfinite(gmax);
      s_red[18] = sep ? 1.f : 0.f;
      if (!sep) todo[1 + atomicAdd(&todo[0], 1)] = s_local;   // work list of the per-pixel kernel
    }
    __syncthreads();
  }
  if (s_red[18] == 0.f) return;          // the per-pixel kernel behind this launch takes the subgrid
  if (nblk == 0) return;

  if (KCp != KC) {                       // the K padding: zero in A and B (never written again)
    for (int i = tid; i < 128 * 4; i += DS_THREADS) {
      *reinterpret_cast<unsigned *>(sA + KC * DS_A_CH + i * 4) = 0u;
      *reinterpret_cast<unsigned *>(sA + (KCp + KC) * DS_A_CH + i * 4) = 0u;
    }
    for (int i = tid; i < 8 * ny_max * 4; i += DS_THREADS) {
      *reinterpret_cast<unsigned *>(sB + KC * b_ch + i * 4) = 0u;
      *reinterpret_cast<unsigned *>(sB + (KCp + KC) * b_ch + i * 4) = 0u;
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = *s_tmem;

  // lane = column x (and x + 32) of the A operand
  float l[XPL], n_x[XPL], off_x[XPL];
#pragma unroll
  for (int xi = 0; xi < XPL; xi++) {
    l[xi] = compute_l(min(lane + 32 * xi, N - 1), N, a.image_size);
    n_x[xi] = compute_n(l[xi], 0.f);
    off_x[xi] = __fmaf_rn(ctx.w_offset, n_x[xi], __fmul_rn(ctx.u_offset, l[xi]));
  }
  float2 *g_out = const_cast<float2 *>(a.visibilities) + (size_t)ctx.time_offset * C * NR_POL;
  const int q4 = warp & 3, part = warp >> 2;
  const int step_t = 16 / ncb, step_cb = 16 - step_t * ncb;
  const size_t plane = (size_t)N * N;
  const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
  const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
  const float2 *sub = a.subgrids + (size_t)s * NR_POL * plane;
  int it = 0;                                      // tiles so far: the parity of mma_done

  for (int slab = 0; slab < nslab; slab++) {
    const int y0 = slab * 32, ny = min(32, N - y0), npix = ny * N, ncols = 8 * ny;
    // ---- B = P' (degridder_reference.cpp:38-74) of the slab's rows in fp16 hi + lo; the fp32 pixels wait in the A buffer
    {
      float4 *sT = reinterpret_cast<float4 *>(sA);     // [npix][2]: 32 ny N <= 2 KCp DS_A_CH
      float amax = 0.f;
      for (int q = tid; q < npix; q += DS_THREADS) {
        const int qq = y0 * N + q;
        float2 px[NR_POL];
        pprime(a, sub, at1, at2, qq, N, plane, px);
        sT[2 * q] = make_float4(px[0].x, px[0].y, px[1].x, px[1].y);
        sT[2 * q + 1] = make_float4(px[2].x, px[2].y, px[3].x, px[3].y);
#pragma unroll
        for (int p = 0; p < NR_POL; p++) amax = fmaxf(amax, fmaxf(fabsf(px[p].x), fabsf(px[p].y)));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0) s_red[warp] = amax;
      __syncthreads();
      if (tid == 0) {
        for (int i = 1; i < NW; i++) amax = fmaxf(amax, s_red[i]);
        const unsigned eb = (__float_as_uint(amax) >> 23) & 0xffu;
        const bool ok = eb >= 16u && eb <= 253u;
        s_red[16] = ok ? __uint_as_float((254u + DS_B_EXP - eb) << 23) : 1.f;           // 2^(DS_B_EXP - E)
        s_red[17] = ok ? __uint_as_float((eb - DS_UNSCALE_EXP) << 23) : DS_UNSCALE_RAW;  // 2^(E - DS_UNSCALE_EXP)
      }
      __syncthreads();
      const float scale = s_red[16];
      for (int q = tid; q < npix; q += DS_THREADS) {
        const int y = q / N, x = q - y * N;
        const float4 v01 = sT[2 * q], v23 = sT[2 * q + 1];
        const float pv[8] = {v01.x, v01.y, v01.z, v01.w, v23.x, v23.y, v23.z, v23.w};
        unsigned char *col = sB + (x >> 2) * b_ch + (x & 3) * 4 + (y * 8) * 16;
#pragma unroll
        for (int p = 0; p < NR_POL; p++) {
          const float re = pv[2 * p] * scale, im = pv[2 * p + 1] * scale;
          // row (y, p, re): (re, -im) against (cos, sin); row (y, p, im): (im, re)
          const unsigned h_im = pack_h2(im, re);
          const unsigned h_re = pack_h2(re, -im);
          const float r_im = residual_h(im, (unsigned short)(h_im & 0xffffu));
          const float r_re = residual_h(re, (unsigned short)(h_im >> 16));
          unsigned char *row = col + (2 * p) * 16;
          *reinterpret_cast<unsigned *>(row) = h_re;
          *reinterpret_cast<unsigned *>(row + 16) = h_im;
          *reinterpret_cast<unsigned *>(row + KCp * b_ch) = second_word_b(r_re, -r_im, h_re);
          *reinterpret_cast<unsigned *>(row + KCp * b_ch + 16) = second_word_b(r_im, r_re, h_im);
        }
      }
      __syncthreads();                    // the fp32 pixels have been read: the A buffer is free
      if (KCp != KC && ny * N * 32 > KC * DS_A_CH) {   // the staging reached into A's K padding: zero it again
        for (int i = tid; i < 128 * 4; i += DS_THREADS) {
          *reinterpret_cast<unsigned *>(sA + KC * DS_A_CH + i * 4) = 0u;
          *reinterpret_cast<unsigned *>(sA + (KCp + KC) * DS_A_CH + i * 4) = 0u;
        }
      }
    }
    const float unscale = s_red[17];
    // instruction descriptor: D = F32, A = B = F16, K-major, N = 8 ny, M = 128
    const unsigned idesc = (1u << 4) | (((unsigned)ncols >> 3) << 17) | ((128u >> 4) << 24);
    // the rows y of this warp in the sum over the rows: a quadrant's NP warps each take a share of the groups of 4
    const int ng = ny >> 2;
    const int g_lo = NP == 2 ? (part ? (ng + 1) >> 1 : 0) : (part * ng) / NP, g_hi = NP == 2 ? (part ? ng : (ng + 1) >> 1) : ((part + 1) * ng) / NP;

    // (timestep, block) of the warp's first block of the tile, stepped by 16 blocks per tile
    int pt = (warp * BPW) / ncb, pcb = warp * BPW - pt * ncb;
    for (int tile = 0; tile < ntiles; tile++, it++) {
      // ---- A rows of this warp's BPW blocks of the tile: X_v(x) for the block's 8 channels, hi + lo
#pragma unroll
      for (int bi = 0; bi < BPW; bi++) {   // unrolled: independent chains in flight
        const int blk = tile * 16 + warp * BPW + bi;
        int t = pt, cb = pcb + bi;
        if (cb >= ncb) { cb -= ncb; t++; }
        const bool live = blk < nblk;
        const float ut = live ? uvw_at(3 * t) : 0.f, wt = live ? uvw_at(3 * t + 2) : 0.f;
        const float *wn8 = s_wn + (live ? cb : 0) * 8;
        const bool lin = live && s_lin[cb];
        const float dwc = live ? s_dw[cb] : 0.f;
#pragma unroll
        for (int xi = 0; xi < XPL; xi++) {
          unsigned hi[8], lo[8];
          if (live) {
            const float idx = __fmaf_rn(wt, n_x[xi], __fmul_rn(ut, l[xi]));
            float2 ph[8];
            if (lin) {   // first channel by sincos, second by one rotation, then the three-term recurrence
              ph[0] = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(idx, wn8[0], -off_x[xi]));
              if (DS_FP8_LO) ph[0] = __fmul2_rn(ph[0], make_float2(DS_A_SCALE, DS_A_SCALE));
              const float2 d = phasor<IDGB200_SINCOS_FAST>(__fmul_rn(idx, dwc));
              ph[1] = ffma2(make_float2(ph[0].y, ph[0].x), make_float2(-d.y, d.y), __fmul2_rn(ph[0], make_float2(d.x, d.x)));
              const float c2 = __fadd_rn(d.x, d.x);
#pragma unroll
              for (int i = 2; i < 8; i++) ph[i] = ffma2(ph[i - 1], make_float2(c2, c2), make_float2(-ph[i - 2].x, -ph[i - 2].y));
            } else {
#pragma unroll
              for (int i = 0; i < 8; i++) {
                ph[i] = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(idx, wn8[i], -off_x[xi]));
                if (DS_FP8_LO) ph[i] = __fmul2_rn(ph[i], make_float2(DS_A_SCALE, DS_A_SCALE));
              }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
              hi[i] = pack_h2(ph[i].x, ph[i].y);
              lo[i] = second_word_a(ph[i].x, ph[i].y, hi[i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; i++) hi[i] = lo[i] = 0u;
          }
          const int x = lane + 32 * xi;
          if (x < N) {
            unsigned char *row = sA + (x >> 2) * DS_A_CH + (x & 3) * 4 + ((warp * BPW + bi) * 8) * 16;
#pragma unroll
            for (int i = 0; i < 8; i++) {
              *reinterpret_cast<unsigned *>(row + i * 16) = hi[i];
              *reinterpret_cast<unsigned *>(row + i * 16 + KCp * DS_A_CH) = lo[i];
            }
          }
        }
      }
      pt += step_t; pcb += step_cb;
      if (pcb >= ncb) { pcb -= ncb; pt++; }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const unsigned a_u = smem_u32(sA), b_u = smem_u32(sB);
          for (int ks = 0; ks < KCp / 2; ks++) {
            const unsigned long long a_hi = smem_desc(a_u + 2 * ks * DS_A_CH, DS_A_CH, 128);
            const unsigned long long a_lo = smem_desc(a_u + (KCp + 2 * ks) * DS_A_CH, DS_A_CH, 128);
            const unsigned long long b_hi = smem_desc(b_u + 2 * ks * b_ch, b_ch, 128);
            const unsigned long long b_lo = smem_desc(b_u + (KCp + 2 * ks) * b_ch, b_ch, 128);
            umma_f16(tmem_base, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
            umma_cross(tmem_base, a_hi, a_lo, b_hi, b_lo, idesc);
          }
          umma_commit(mma_done);
        }
        __syncwarp();
      }
      mbar_wait_sleep(mma_done, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      // ---- the sum over the rows y: thread = row 32 q4 + lane of the tile, warps q4 and q4 + 4 half of the rows each
      {
        const int r_tile = q4 * 32 + lane;
        const int blk = tile * 16 + (r_tile >> 3);
        const bool in_range = blk < nblk;
        const int t = in_range ? blk / ncb : 0, cb = in_range ? blk - t * ncb : 0;
        const int c = cb * 8 + (r_tile & 7);
        const bool valid = in_range && c < C;
        const float k = s_wn[cb * 8 + (r_tile & 7)];
        const float vt = uvw_at(3 * t + 1), wt = uvw_at(3 * t + 2);
        // acc = sum Q ph with Q = qr + i qi, ph = (c, s): two packed accumulators per polarisation, ar += qr (c, s) and
        // ai += qi (c, s), combined once per tile (acc = (ar.x - ai.y, ar.y + ai.x)): no (-s, c) operand to assemble
        float2 ar[NR_POL], ai[NR_POL];
#pragma unroll
        for (int p = 0; p < NR_POL; p++) ar[p] = ai[p] = make_float2(0.f, 0.f);
        for (int g = g_lo; g < g_hi; g++) {
          unsigned r[32];
          const unsigned taddr = tmem_base + ((unsigned)(q4 * 32) << 16) + g * 32;
          tmem_ld16(taddr, r);
          tmem_ld16(taddr + 16, r + 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int yy = 0; yy < 4; yy++) {
            const float4 geo = sGeo[y0 + g * 4 + yy];        // broadcast
            const float idx = __fmaf_rn(wt, geo.y, __fmul_rn(vt, geo.x));
            const float2 ph = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(idx, k, -geo.z));
#pragma unroll
            for (int p = 0; p < NR_POL; p++) {
              const float qr = __uint_as_float(r[yy * 8 + 2 * p]), qi = __uint_as_float(r[yy * 8 + 2 * p + 1]);
              ar[p] = ffma2(make_float2(qr, qr), ph, ar[p]);
              ai[p] = ffma2(make_float2(qi, qi), ph, ai[p]);
            }
          }
        }
        float2 acc[NR_POL];
#pragma unroll
        for (int p = 0; p < NR_POL; p++) acc[p] = make_float2(ar[p].x - ai[p].y, ar[p].y + ai[p].x);
        if (part) {
          sPart[(part - 1) * 256 + 2 * r_tile] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
          sPart[(part - 1) * 256 + 2 * r_tile + 1] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                  // accumulator and A buffer are free for the next tile; the upper halves are in
        if (!part && valid) {
#pragma unroll
          for (int hh = 0; hh < NP - 1; hh++) {     // the other parts' sums, in order
            const float4 p0 = sPart[hh * 256 + 2 * r_tile], p1 = sPart[hh * 256 + 2 * r_tile + 1];
            acc[0] = make_float2(acc[0].x + p0.x, acc[0].y + p0.y);
            acc[1] = make_float2(acc[1].x + p0.z, acc[1].y + p0.w);
            acc[2] = make_float2(acc[2].x + p1.x, acc[2].y + p1.y);
            acc[3] = make_float2(acc[3].x + p1.z, acc[3].y + p1.w);
          }
          float4 *o = reinterpret_cast<float4 *>(g_out + ((size_t)t * C + c) * NR_POL);
          float4 o0 = make_float4(acc[0].x * unscale, acc[0].y * unscale, acc[1].x * unscale, acc[1].y * unscale);
          float4 o1 = make_float4(acc[2].x * unscale, acc[2].y * unscale, acc[3].x * unscale, acc[3].y * unscale);
          if (slab > 0) {                 // the earlier slabs' partial visibility: written by this very thread
            const float4 e0 = o[0], e1 = o[1];
            o0 = make_float4(e0.x + o0.x, e0.y + o0.y, e0.z + o0.z, e0.w + o0.w);
            o1 = make_float4(e1.x + o1.x, e1.y + o1.y, e1.z + o1.z, e1.w + o1.w);
          }
          o[0] = o0;
          o[1] = o1;
        }
      }
    }
  }
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}

// ---------------------------------------------------------------------------------------------------------------------
// The same arithmetic as a warp-specialised, persistent pipeline: ONE CTA per SM that loops over subgrids, 20 warps.
//   warps  0- 7  producers: the A rows (column phasors X, fp16 + e4m3 parts) of tile i + 1 / i + 2 into one of two A buffers
//   warps  8-15  consumers: the sum over the rows y of tile i out of one of two accumulators (all 512 TMEM columns);
//                two warps per TMEM lane quadrant, half of the rows each
//   warp   16    issuer: the tile's 8 MMAs as soon as its A buffer is full and its accumulator has been drained
//   warps 17-19  setup: the NEXT subgrid's separability check, B operand (P' in fp16 + e4m3 parts, one of two B buffers),
//                staged uvw and row geometry, while the others are in the tile loop of the current subgrid
// so that the tensor pipe, the producers' dependency chains and the consumers' TMEM reads overlap inside one CTA
// instead of across two (the kernel above: a tile's phases one after the other, 2 CTAs per SM), and the per-subgrid
// prologue (32 KB of pixels + A-terms) is hidden behind the previous subgrid.  Hand-offs are mbarriers only:
//   b_full[2]  (3 setup warps)   -> everyone     B buffer, meta, geometry, uvw of subgrid j are ready
//   b_empty[2] (17 warps)        -> setup        producers, consumers and the issuer are done with subgrid j's buffers
//   a_full[2]  (8 producers)     -> issuer
//   mma_done[2] (tcgen05.commit) -> consumers (accumulator full) and producers (A buffer free)
//   d_empty[2] (8 consumers)     -> issuer       accumulator drained
//   part_full[2][4] (upper consumer warp of a quadrant) -> its lower warp (the two halves of the sum over the rows meet)
//   part_empty[2][4] (lower warp of a quadrant) -> its upper warp (the partial sums of tile it - 2 have been read)
// Tiles are counted over the whole run of a CTA (`it`), buffer = it & 1, phase = (it >> 1) & 1.  Results are bit-identical
// to the kernel above (same MMAs in the same order, same split of the sum over the rows).
// 8 consumer warps: 2 per TMEM lane quadrant, half of the rows y each.  Tried and dropped: 16 consumer warps,
// either four per quadrant on one tile or two groups of 8 alternating accumulators (4.74 / 4.45 against 4.41 ms: the
// dispatch ports are what is short, not warps), and making a group's row phasors ahead of its TMEM load's wait (4.61
// against 4.43 ms: the extra live registers cost more than the overlap buys).
constexpr int DP_CONSUMERS = 8;
constexpr int DP_WARPS = 8 + DP_CONSUMERS + 1 + 3, DP_THREADS = DP_WARPS * 32;   // 20 warps of 96 registers
constexpr int DP_ISSUER = 8 + DP_CONSUMERS, DP_SETUP0 = DP_ISSUER + 1, DP_SETUP_WARPS = DP_WARPS - DP_SETUP0, DP_SETUP_THREADS = DP_SETUP_WARPS * 32;
constexpr int DP_UVW_STAGED = 256;

// wait flavours of the pipeline's per-tile hand-offs (A/B knobs, tools/build_ab.sh): 0 = try_wait with suspend hint and
// wall-clock bound, 1 = bare try_wait loop, 2 / 3 / 4 = test_wait + nanosleep(32 / 128 / 256)
#ifndef DP_WAIT_CONS
#define DP_WAIT_CONS 1
#endif
#ifndef DP_WAIT_PROD
#define DP_WAIT_PROD 3
#endif
#ifndef DP_WAIT_ISSUE
#define DP_WAIT_ISSUE 1
#endif
// DP_TRACE builds (tools/build_ab.sh + tools/pipe_trace.py, never shipped): CTA 0 records clock64() at the hand-offs of
// the tiles of its subgrids DP_TRACE_J0 .. DP_TRACE_J0 + 3 (steady state), one row per event
#ifdef DP_TRACE
constexpr int DP_TRACE_J0 = 20, DP_TRACE_TILES = 64, DP_TRACE_EVENTS = 56;
__device__ long long dp_trace_buf[DP_TRACE_EVENTS * DP_TRACE_TILES];
#define DP_MARK(ev, j, tile, cond)                                                                             \
  do {                                                                                                         \
    if (blockIdx.x == 0 && (cond) && (j) >= DP_TRACE_J0 && (j) < DP_TRACE_J0 + 4 && (tile) < 16)               \
      dp_trace_buf[(ev) * DP_TRACE_TILES + ((j) - DP_TRACE_J0) * 16 + (tile)] = clock64();                     \
  } while (0)
#else
#define DP_MARK(ev, j, tile, cond) do { } while (0)
#endif
template <int FLAVOUR>
__device__ __forceinline__ void dp_wait(const unsigned bar, const unsigned parity) {
  if (FLAVOUR == 0) mbar_wait_t(bar, parity);
  else if (FLAVOUR == 1) mbar_wait_spin(bar, parity);
  else mbar_wait_backoff(bar, parity, FLAVOUR == 2 ? 32 : FLAVOUR == 3 ? 128 : 256);
}

struct PipeMeta {        // per-subgrid scalars, written by the setup warps
  int run;               // 1: the row-column tiles of this subgrid run here; 0: nothing to do (declined or no timesteps)
  int nt;
  long long time_offset;
  float u_offset, w_offset, unscale;
  int pad;
};

__device__ __forceinline__ void setup_bar() { asm volatile("bar.sync 1, %0;" ::"n"(DP_SETUP_THREADS) : "memory"); }

__global__ void __launch_bounds__(DP_THREADS, 1)
degridder_sep_pipe_kernel(const KernelArgs a, int *__restrict__ todo) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int N = a.subgrid_size, C = a.nr_channels, npix = N * N;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int KC = N >> 2, KCp = (KC + 1) & ~1;
  const int ncols = 8 * N;
  const int b_ch = ncols * 16 + 16;
  const int ncb = (C + 7) >> 3;
  const int b_buf = 2 * KCp * b_ch, a_buf = 2 * KCp * DS_A_CH;

  unsigned char *sB = smem;                                             // [2][hi|lo][KCp][b_ch]
  unsigned char *sA = sB + 2 * b_buf;                                   // [2][hi|lo][KCp][DS_A_CH]
  float4 *sGeo = reinterpret_cast<float4 *>(sA + 2 * a_buf);            // [2][N] (m_y, f(m_y^2), offset_y, 0)
  float4 *sPart = sGeo + 2 * N;                                         // [2][128][2] partial sums of the upper consumer warps, per accumulator
  float *s_uvw = reinterpret_cast<float *>(sPart + 512);                // [2][DP_UVW_STAGED][3]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(s_uvw + 2 * DP_UVW_STAGED * 3);   // [26]
  PipeMeta *meta = reinterpret_cast<PipeMeta *>(bars + 26);             // [2]
  unsigned *s_tmem = reinterpret_cast<unsigned *>(meta + 2);            // [2]
  float *s_red = reinterpret_cast<float *>(s_tmem + 2);                 // [48]: [0, 28) kmax partials, [32, 48) setup partials
  float *s_wn = s_red + 48;                                             // [ncb * 8]
  float *s_dw = s_wn + ncb * 8;                                         // [ncb]
  int *s_lin = reinterpret_cast<int *>(s_dw + ncb);                     // [ncb]
  const unsigned bar_u = smem_u32(bars);
  // barrier b of buffer i: bar_u + (2 * b + i) * 8
  enum { A_FULL = 0, MMA_DONE = 1, D_EMPTY = 2, B_FULL = 3, B_EMPTY = 4, PART_FULL = 5, PART_EMPTY = 9 };   // the last two: per accumulator and quadrant
  auto bar_at = [&](int b, int i) { return bar_u + (unsigned)(2 * b + i) * 8u; };
  auto arrive_u = [](unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); };

  // ---- once per CTA: channel layout, barriers, K padding, TMEM
  {
    float kmax = 0.f;
    for (int c = tid; c < ncb * 8; c += DP_THREADS) {
      const float k = c < C ? a.wavenumbers[c] : 0.f;
      s_wn[c] = k;
      kmax = fmaxf(kmax, fabsf(k));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kmax = fmaxf(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    if (lane == 0) s_red[warp] = kmax;
  }
  if (tid == 0) {
    for (int i = 0; i < 2; i++) {
      mbar_init(&bars[2 * A_FULL + i], 8);
      mbar_init(&bars[2 * MMA_DONE + i], 1);
      mbar_init(&bars[2 * D_EMPTY + i], 8);
      mbar_init(&bars[2 * B_FULL + i], DP_SETUP_WARPS);
      mbar_init(&bars[2 * B_EMPTY + i], 8 + DP_CONSUMERS + 1);
    }
    for (int q = 0; q < 8; q++) {
      mbar_init(&bars[2 * PART_FULL + q], 1);
      mbar_init(&bars[2 * PART_EMPTY + q], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (KCp != KC) {                      // the K padding: zero in A and B (never written again)
    for (int i = tid; i < 128 * 4; i += DP_THREADS)
      for (int b = 0; b < 2; b++) {
        *reinterpret_cast<unsigned *>(sA + b * a_buf + KC * DS_A_CH + i * 4) = 0u;
        *reinterpret_cast<unsigned *>(sA + b * a_buf + (KCp + KC) * DS_A_CH + i * 4) = 0u;
      }
    for (int i = tid; i < ncols * 4; i += DP_THREADS)
      for (int b = 0; b < 2; b++) {
        *reinterpret_cast<unsigned *>(sB + b * b_buf + KC * b_ch + i * 4) = 0u;
        *reinterpret_cast<unsigned *>(sB + b * b_buf + (KCp + KC) * b_ch + i * 4) = 0u;
      }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == DP_ISSUER) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cb = tid; cb < ncb; cb += DP_THREADS) {     // s_wn is complete
    float dw;
    s_lin[cb] = linear_channels(s_wn, cb * 8, min(8, C - cb * 8), &dw) ? 1 : 0;
    s_dw[cb] = dw;
  }
  __syncthreads();
  const unsigned tmem_base = *s_tmem;
  const int S = a.nr_subgrids;

  if (warp < 8) {
    // ================================================================================= producers: lane = column x
    const float l = compute_l(min(lane, N - 1), N, a.image_size);
    const float n_x = compute_n(l, 0.f);
    const int a_off = (lane >> 2) * DS_A_CH + (lane & 3) * 4;
    const int step_t = 16 / ncb, step_cb = 16 - step_t * ncb;
    int it = 0;
    for (int s_local = blockIdx.x, j = 0; s_local < S; s_local += gridDim.x, j++) {
      const int bb = j & 1;
      mbar_wait_t(bar_at(B_FULL, bb), (j >> 1) & 1);
      const PipeMeta m = meta[bb];
      if (m.run) {
        const int nt = m.nt, nblk = nt * ncb, ntiles = (nblk + 15) >> 4;
        const float off_x = __fmaf_rn(m.w_offset, n_x, __fmul_rn(m.u_offset, l));
        const bool staged = nt <= DP_UVW_STAGED;
        const float *uvw_s = s_uvw + bb * DP_UVW_STAGED * 3, *uvw_g = reinterpret_cast<const float *>(a.uvw) + (size_t)m.time_offset * 3;
        auto uvw_at = [&](int i) { return staged ? uvw_s[i] : __ldg(&uvw_g[i]); };
        int pt = (warp * 2) / ncb, pcb = warp * 2 - pt * ncb;
        for (int tile = 0; tile < ntiles; tile++, it++) {
          const int buf = it & 1;
          if (it >= 2) dp_wait<DP_WAIT_PROD>(bar_at(MMA_DONE, buf), ((it >> 1) - 1) & 1);   // the MMAs of tile it - 2 have read the buffer
          DP_MARK(2 * warp, j, tile, lane == 0);
          unsigned char *a_col = sA + buf * a_buf + a_off;
#pragma unroll
          for (int bi = 0; bi < 2; bi++) {
            if ((IDGB200_ABLATE & 8) && !ablate_never()) continue;      // ablation: no A rows
            const int blk = tile * 16 + warp * 2 + bi;
            unsigned char *row = a_col + ((warp * 2 + bi) * 8) * 16;
            if (blk < nblk) {
              int t = pt, cb = pcb + bi;
              if (cb >= ncb) { cb -= ncb; t++; }
              const float idx = __fmaf_rn(uvw_at(3 * t + 2), n_x, __fmul_rn(uvw_at(3 * t), l));
              const float *wn8 = s_wn + cb * 8;
              float2 ph[8];
              if (s_lin[cb]) {
                ph[0] = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(idx, wn8[0], -off_x));
                if (DS_FP8_LO) ph[0] = __fmul2_rn(ph[0], make_float2(DS_A_SCALE, DS_A_SCALE));
                const float2 d = phasor<IDGB200_SINCOS_FAST>(__fmul_rn(idx, s_dw[cb]));
                ph[1] = ffma2(make_float2(ph[0].y, ph[0].x), make_float2(-d.y, d.y), __fmul2_rn(ph[0], make_float2(d.x, d.x)));
                const float c2 = __fadd_rn(d.x, d.x);
#pragma unroll
                for (int i = 2; i < 8; i++) ph[i] = ffma2(ph[i - 1], make_float2(c2, c2), make_float2(-ph[i - 2].x, -ph[i - 2].y));
              } else {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                  ph[i] = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(idx, wn8[i], -off_x));
                  if (DS_FP8_LO) ph[i] = __fmul2_rn(ph[i], make_float2(DS_A_SCALE, DS_A_SCALE));
                }
              }
              if (lane < N) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                  const unsigned hi = pack_h2(ph[i].x, ph[i].y);
                  *reinterpret_cast<unsigned *>(row + i * 16) = hi;
                  *reinterpret_cast<unsigned *>(row + i * 16 + KCp * DS_A_CH) = second_word_a(ph[i].x, ph[i].y, hi);
                }
              }
            } else if (lane < N) {     // rows beyond the subgrid's last block (the last tile only): zeros
#pragma unroll
              for (int i = 0; i < 8; i++) {
                *reinterpret_cast<unsigned *>(row + i * 16) = 0u;
                *reinterpret_cast<unsigned *>(row + i * 16 + KCp * DS_A_CH) = 0u;
              }
            }
          }
          pt += step_t; pcb += step_cb;
          if (pcb >= ncb) { pcb -= ncb; pt++; }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) arrive_u(bar_at(A_FULL, buf));
          DP_MARK(2 * warp + 1, j, tile, lane == 0);
        }
      }
      __syncwarp();
      if (lane == 0) arrive_u(bar_at(B_EMPTY, bb));
    }
  } else if (warp < DP_ISSUER) {
    // ========================================================= consumers: thread = visibility (TMEM lane = row of the tile)
    // quadrant q4's two warps (h = 0, 1) each sum half of the row groups, and warp 1 hands its sums to warp 0 through sPart
    const int q4 = warp & 3, h = (warp - 8) >> 2;
    const int ng = N >> 2, g_lo = h ? (ng + 1) >> 1 : 0, g_hi = h ? ng : (ng + 1) >> 1;
    const int r_tile = q4 * 32 + lane;
    int it = 0;
    for (int s_local = blockIdx.x, j = 0; s_local < S; s_local += gridDim.x, j++) {
      const int bb = j & 1;
      mbar_wait_t(bar_at(B_FULL, bb), (j >> 1) & 1);
      const PipeMeta m = meta[bb];
      if (m.run) {
        const int nt = m.nt, nblk = nt * ncb, ntiles = (nblk + 15) >> 4;
        const float unscale = m.unscale;
        const bool staged = nt <= DP_UVW_STAGED;
        const float *uvw_s = s_uvw + bb * DP_UVW_STAGED * 3, *uvw_g = reinterpret_cast<const float *>(a.uvw) + (size_t)m.time_offset * 3;
        auto uvw_at = [&](int i) { return staged ? uvw_s[i] : __ldg(&uvw_g[i]); };
        const float4 *geo_b = sGeo + bb * N;
        float2 *g_out = const_cast<float2 *>(a.visibilities) + (size_t)m.time_offset * C * NR_POL;
        // (timestep, block) of this thread's row of the tile, stepped by 16 blocks per tile like the producers'
        const int c_step_t = 16 / ncb, c_step_cb = 16 - c_step_t * ncb;
        int ct = (r_tile >> 3) / ncb, ccb = (r_tile >> 3) - ct * ncb;
        for (int tile = 0; tile < ntiles; tile++, it++) {
          const int buf = it & 1;
          const int blk = tile * 16 + (r_tile >> 3);
          const bool in_range = blk < nblk;
          const int t = in_range ? ct : 0, cb = in_range ? ccb : 0;
          ct += c_step_t; ccb += c_step_cb;
          if (ccb >= ncb) { ccb -= ncb; ct++; }
          const unsigned part_full = bar_at(PART_FULL, 0) + 8u * (buf * 4 + q4), part_empty = bar_at(PART_EMPTY, 0) + 8u * (buf * 4 + q4);
          const int c = cb * 8 + (r_tile & 7);
          const bool valid = in_range && c < C;
          const float k = s_wn[c];
          const float vt = uvw_at(3 * t + 1), wt = uvw_at(3 * t + 2);
          // acc = sum Q ph with Q = qr + i qi, ph = (c, s): two packed accumulators per polarisation, ar += qr (c, s) and
          // ai += qi (c, s), combined once per tile (acc = (ar.x - ai.y, ar.y + ai.x)): no (-s, c) operand to assemble
          float2 ar[NR_POL], ai[NR_POL];
#pragma unroll
          for (int p = 0; p < NR_POL; p++) ar[p] = ai[p] = make_float2(0.f, 0.f);
          // the row phasors Y_v(y) of a group of 4 rows
          auto make_ph = [&](const int g, float2 *ph) {
#pragma unroll
            for (int yy = 0; yy < 4; yy++) {
              const float4 geo = geo_b[g * 4 + yy];        // broadcast
              const float idx = __fmaf_rn(wt, geo.y, __fmul_rn(vt, geo.x));
              ph[yy] = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(idx, k, -geo.z));
            }
          };
          dp_wait<DP_WAIT_CONS>(bar_at(MMA_DONE, buf), (it >> 1) & 1);
          DP_MARK(16 + 2 * (warp - 8), j, tile, lane == 0);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int g = g_lo; g < g_hi; g++) {
            if ((IDGB200_ABLATE & 4) && !ablate_never()) continue;      // ablation: no sum over the rows
            unsigned r[32];
            const unsigned taddr = tmem_base + (unsigned)(buf * 256) + ((unsigned)(q4 * 32) << 16) + g * 32;
            tmem_ld16(taddr, r);
            tmem_ld16(taddr + 16, r + 16);
            float2 ph[4];
            make_ph(g, ph);              // independent of the loads: overlaps their latency
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int yy = 0; yy < 4; yy++) {
#pragma unroll
              for (int p = 0; p < NR_POL; p++) {
                const float qr = __uint_as_float(r[yy * 8 + 2 * p]), qi = __uint_as_float(r[yy * 8 + 2 * p + 1]);
                ar[p] = ffma2(make_float2(qr, qr), ph[yy], ar[p]);
                ai[p] = ffma2(make_float2(qi, qi), ph[yy], ai[p]);
              }
            }
          }
          // the accumulator has been read (its values are in registers): the issuer may overwrite it
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) arrive_u(bar_at(D_EMPTY, buf));
          DP_MARK(17 + 2 * (warp - 8), j, tile, lane == 0);
          float2 acc[NR_POL];
#pragma unroll
          for (int p = 0; p < NR_POL; p++) acc[p] = make_float2(ar[p].x - ai[p].y, ar[p].y + ai[p].x);
          if (h > 0) {
            if (it >= 2) dp_wait<DP_WAIT_CONS>(part_empty, ((it >> 1) - 1) & 1);    // warp 0 has read tile it - 2's
            float4 *part = sPart + buf * 256 + 2 * r_tile;
            part[0] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
            part[1] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
            __syncwarp();
            if (lane == 0) arrive_u(part_full);
          } else {
            dp_wait<DP_WAIT_CONS>(part_full, (it >> 1) & 1);
            const float4 *part = sPart + buf * 256 + 2 * r_tile;
            const float4 p0 = part[0], p1 = part[1];
            acc[0] = make_float2(acc[0].x + p0.x, acc[0].y + p0.y);
            acc[1] = make_float2(acc[1].x + p0.z, acc[1].y + p0.w);
            acc[2] = make_float2(acc[2].x + p1.x, acc[2].y + p1.y);
            acc[3] = make_float2(acc[3].x + p1.z, acc[3].y + p1.w);
            __syncwarp();
            if (lane == 0) arrive_u(part_empty);
            if (valid) {
              float4 *o = reinterpret_cast<float4 *>(g_out + ((size_t)t * C + c) * NR_POL);
              o[0] = make_float4(acc[0].x * unscale, acc[0].y * unscale, acc[1].x * unscale, acc[1].y * unscale);
              o[1] = make_float4(acc[2].x * unscale, acc[2].y * unscale, acc[3].x * unscale, acc[3].y * unscale);
            }
          }
          DP_MARK(h ? 53 : 52, j, tile, q4 == 0 && h < 2 && lane == 0);
        }
      }
      __syncwarp();
      if (lane == 0) arrive_u(bar_at(B_EMPTY, bb));
    }
  } else if (warp == DP_ISSUER) {
    // ================================================================================================== issuer
    const unsigned idesc = (1u << 4) | (((unsigned)ncols >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned a_u = smem_u32(sA), b_u = smem_u32(sB);
    int it = 0;
    for (int s_local = blockIdx.x, j = 0; s_local < S; s_local += gridDim.x, j++) {
      const int bb = j & 1;
      mbar_wait_t(bar_at(B_FULL, bb), (j >> 1) & 1);
      const PipeMeta m = meta[bb];
      if (m.run) {
        const int ntiles = (m.nt * ncb + 15) >> 4;
        for (int tile = 0; tile < ntiles; tile++, it++) {
          const int buf = it & 1;
          dp_wait<DP_WAIT_ISSUE>(bar_at(A_FULL, buf), (it >> 1) & 1);
          if (it >= 2) dp_wait<DP_WAIT_ISSUE>(bar_at(D_EMPTY, buf), ((it >> 1) - 1) & 1);
          DP_MARK(48, j, tile, lane == 0);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const unsigned d_t = tmem_base + (unsigned)(buf * 256);
            const unsigned a_b = a_u + buf * a_buf, b_b = b_u + bb * b_buf;
            for (int ks = 0; ks < KCp / 2; ks++) {
              if ((IDGB200_ABLATE & 1) && !ablate_never()) continue;    // ablation: no MMAs
              const unsigned long long a_hi = smem_desc(a_b + 2 * ks * DS_A_CH, DS_A_CH, 128);
              const unsigned long long a_lo = smem_desc(a_b + (KCp + 2 * ks) * DS_A_CH, DS_A_CH, 128);
              const unsigned long long b_hi = smem_desc(b_b + 2 * ks * b_ch, b_ch, 128);
              const unsigned long long b_lo = smem_desc(b_b + (KCp + 2 * ks) * b_ch, b_ch, 128);
              umma_f16(d_t, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
              umma_cross(d_t, a_hi, a_lo, b_hi, b_lo, idesc);
            }
            umma_commit_u(bar_at(MMA_DONE, buf));
          }
          __syncwarp();
          DP_MARK(49, j, tile, lane == 0);
        }
      }
      __syncwarp();
      if (lane == 0) arrive_u(bar_at(B_EMPTY, bb));
    }
  } else {
    // =================================================================== setup: the next subgrid's check and B operand
    const int ts = tid - DP_SETUP0 * 32, sw = warp - DP_SETUP0;
    float kmax = 0.f;
    for (int i = 0; i < DP_WARPS; i++) kmax = fmaxf(kmax, s_red[i]);
    // the dropped phase term r at the corner pixel (gridder_sep.cu)
    const double l0 = (0.5 - (N / 2)) * (double)a.image_size / (double)N;
    const double s1 = l0 * l0;
    const double fn = s1 / (1.0 + sqrt(1.0 - s1)), s2 = 2.0 * s1;
    const double r_corner = s2 > 1.0 ? 1.0 : fabs(s2 / (1.0 + sqrt(1.0 - s2)) - 2.0 * fn);
    const size_t plane = (size_t)npix;
    for (int s_local = blockIdx.x, j = 0; s_local < S; s_local += gridDim.x, j++) {
      const int bb = j & 1;
      if (j >= 2) mbar_wait_backoff(bar_at(B_EMPTY, bb), ((j >> 1) - 1) & 1, 1000);
      DP_MARK(50, j, 0, ts == 0);
      const int s = a.subgrid_offset + s_local;
      const SubgridCtx ctx = load_ctx(a, s);
      const int nt = ctx.nr_timesteps;
      const float *g_uvw = reinterpret_cast<const float *>(a.uvw) + (size_t)ctx.time_offset * 3;
      float wmax = 0.f;
      if (nt <= DP_UVW_STAGED) {
        float *dst = s_uvw + bb * DP_UVW_STAGED * 3;
        for (int i = ts; i < nt * 3; i += DP_SETUP_THREADS) {
          const float x = __ldg(&g_uvw[i]);
          dst[i] = x;
          if (i % 3 == 2) wmax = fmaxf(wmax, fabsf(x));
        }
      } else {
        for (int t = ts; t < nt; t += DP_SETUP_THREADS) wmax = fmaxf(wmax, fabsf(__ldg(&g_uvw[3 * t + 2])));
      }
      const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
      const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
      const float2 *sub = a.subgrids + (size_t)s * NR_POL * plane;
      // pass 1: P' in fp32, parked in the words of the B buffer's lo half that the pixel's own lo parts will take
      // (row n = y * 8 + 2 p + re|im of K chunk x / 4: eight words 16 bytes apart, only ever touched by the pixel's
      // own thread).  The subgrid's pixels (DRAM) are brought there by cp.async, all of a thread's pixels in flight at
      // once and none of them in registers; the A-terms (L2: the array is small) come through registers, one pixel ahead.
      float amax = 0.f;
      unsigned char *sBlo = sB + bb * b_buf + KCp * b_ch;
      if (nt > 0 && !((IDGB200_ABLATE & 32) && !ablate_never())) {     // ablation: no P'
        auto slot_of = [&](const int q) {
          const int y = q / N, x = q - y * N;
          return reinterpret_cast<float *>(sBlo + (x >> 2) * b_ch + (x & 3) * 4 + (y * 8) * 16);
        };
        for (int q = ts; q < npix; q += DP_SETUP_THREADS) {
          const float *src = reinterpret_cast<const float *>(sub + subgrid_slot(q, N, a.flags));
          float *slot = slot_of(q);
#pragma unroll
          for (int p = 0; p < NR_POL; p++) {
            cp_async4(slot + 8 * p, src + 2 * p * plane);
            cp_async4(slot + 8 * p + 4, src + 2 * p * plane + 1);
          }
        }
        cp_async_commit();
        float2 a1[4], a2[4];
        load_jones(a.aterms, (at1 + ts) * NR_POL, a1);
        load_jones(a.aterms, (at2 + ts) * NR_POL, a2);
        float sph = __ldg(&a.spheroidal[ts]);
        asm volatile("cp.async.wait_all;" ::: "memory");     // this thread's pixels have landed (and the loads below stay below)
        for (int q = ts; q < npix; q += DP_SETUP_THREADS) {
          float2 b1[4], b2[4];
          float sph_n = 0.f;
          const int qn = q + DP_SETUP_THREADS;
          if (qn < npix) {
            load_jones(a.aterms, (at1 + qn) * NR_POL, b1);
            load_jones(a.aterms, (at2 + qn) * NR_POL, b2);
            sph_n = __ldg(&a.spheroidal[qn]);
          }
          float *slot = slot_of(q);
          float2 px[NR_POL];
#pragma unroll
          for (int p = 0; p < NR_POL; p++) px[p] = make_float2(__fmul_rn(sph, slot[8 * p]), __fmul_rn(sph, slot[8 * p + 4]));
          apply_aterm_degridder(px, a1, a2);
#pragma unroll
          for (int p = 0; p < NR_POL; p++) {
            slot[8 * p] = px[p].x;
            slot[8 * p + 4] = px[p].y;
            amax = fmaxf(amax, fmaxf(fabsf(px[p].x), fabsf(px[p].y)));
          }
#pragma unroll
          for (int i = 0; i < 4; i++) {
            a1[i] = b1[i];
            a2[i] = b2[i];
          }
          sph = sph_n;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      }
      float *red = s_red + 32 + bb * 8;          // [wmax of the setup warps | amax of the setup warps], one set per buffer
      if (lane == 0) {
        red[sw] = wmax;
        red[4 + sw] = amax;
      }
      setup_bar();                               // also: every thread's fp32 P' is in the B buffer
#pragma unroll
      for (int i = 0; i < DP_SETUP_WARPS; i++) {
        wmax = fmaxf(wmax, red[i]);
        amax = fmaxf(amax, red[4 + i]);
      }
      const double gmax = (double)fabsf(ctx.w_offset) + (double)wmax * (double)kmax;
      const bool sep = gmax * r_corner <= (double)SEP_PHASE_TOL && isfinite(gmax);
      if (!sep && ts == 0) todo[1 + atomicAdd(&todo[0], 1)] = s_local;   // work list of the per-pixel kernel
      const bool run = sep && nt > 0;
      const unsigned eb = (__float_as_uint(amax) >> 23) & 0xffu;
      const bool ok = eb >= 16u && eb <= 253u;
      const float scale = ok ? __uint_as_float((254u + DS_B_EXP - eb) << 23) : 1.f;   // 2^(DS_B_EXP - E)
      if (run && !((IDGB200_ABLATE & 16) && !ablate_never())) {     // ablation: no B operand
        unsigned char *sBb = sB + bb * b_buf;
#pragma unroll 2
        for (int q = ts; q < npix; q += DP_SETUP_THREADS) {
          const int y = q / N, x = q - y * N;
          unsigned char *col = sBb + (x >> 2) * b_ch + (x & 3) * 4 + (y * 8) * 16;
          const float *slot = reinterpret_cast<const float *>(col + KCp * b_ch);
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; i++) pv[i] = slot[4 * i];
#pragma unroll
          for (int p = 0; p < NR_POL; p++) {
            const float re = pv[2 * p] * scale, im = pv[2 * p + 1] * scale;
            const unsigned h_im = pack_h2(im, re);
            const unsigned h_re = pack_h2(re, -im);
            const float r_im = residual_h(im, (unsigned short)(h_im & 0xffffu));
            const float r_re = residual_h(re, (unsigned short)(h_im >> 16));
            unsigned char *row = col + (2 * p) * 16;
            *reinterpret_cast<unsigned *>(row) = h_re;
            *reinterpret_cast<unsigned *>(row + 16) = h_im;
            *reinterpret_cast<unsigned *>(row + KCp * b_ch) = second_word_b(r_re, -r_im, h_re);
            *reinterpret_cast<unsigned *>(row + KCp * b_ch + 16) = second_word_b(r_im, r_re, h_im);
          }
        }
        if (ts < N) {
          const float mm = compute_l(ts, N, a.image_size);
          const float n_y = compute_n(mm, 0.f);
          sGeo[bb * N + ts] = make_float4(mm, n_y, __fmaf_rn(ctx.w_offset, n_y, __fmul_rn(ctx.v_offset, mm)), 0.f);
        }
      }
      if (ts == 0) {
        PipeMeta m;
        m.run = run ? 1 : 0;
        m.nt = nt;
        m.time_offset = ctx.time_offset;
        m.u_offset = ctx.u_offset;
        m.w_offset = ctx.w_offset;
        m.unscale = ok ? __uint_as_float((eb - DS_UNSCALE_EXP) << 23) : DS_UNSCALE_RAW;   // 2^(E - DS_UNSCALE_EXP)
        m.pad = 0;
        meta[bb] = m;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) arrive_u(bar_at(B_FULL, bb));
      DP_MARK(51, j, 0, ts == 0);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == DP_ISSUER)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

}  // namespace

bool degridder_sep_supports(int subgrid_size, int nr_channels) {
  return subgrid_size >= 4 && subgrid_size % 4 == 0 && subgrid_size <= 64 && nr_channels >= 1 && nr_channels <= 4096;
}

static size_t pipe_smem_bytes(int N, int C) {
  const int KC = N / 4, KCp = (KC + 1) & ~1, ncols = 8 * N, ncb = (C + 7) / 8;
  return (size_t)2 * 2 * KCp * (ncols * 16 + 16) + (size_t)2 * 2 * KCp * DS_A_CH + (size_t)2 * N * 16 + 512 * 16 +
         (size_t)2 * DP_UVW_STAGED * 12 + 26 * 8 + 2 * sizeof(PipeMeta) + 8 + 48 * 4 + (size_t)ncb * 10 * 4;
}

// d_todo = { n, subgrid[n] } (device; n zeroed by the caller on the same stream): the subgrids left to the per-pixel
// kernel launched behind this one.  mode: 0 = the pipelined persistent kernel where its buffers fit (they do up to ~2000
// channels), 1 = the one-subgrid-per-CTA kernel, 2 = the pipelined kernel or an error
cudaError_t launch_degridder_sep(const KernelArgs &a, int *d_todo, cudaStream_t stream, int mode) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  if (!degridder_sep_supports(a.subgrid_size, a.nr_channels) || !d_todo) return cudaErrorInvalidValue;
  const int N = a.subgrid_size, KC = N / 4, KCp = (KC + 1) & ~1, ncols = 8 * (N < 32 ? N : 32), ncb = (a.nr_channels + 7) / 8;
  const size_t smem_pipe = pipe_smem_bytes(N, a.nr_channels);
  const bool pipe_fits = N <= 32 && smem_pipe <= 227 * 1024;
  if (mode == 2 && !pipe_fits) return cudaErrorInvalidValue;
  if (mode != 1 && pipe_fits) {
    static int sms[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!sms[dev]) {
      int n = 0;
      e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      if (e != cudaSuccess) return e;
      sms[dev] = n;
    }
    e = cudaFuncSetAttribute(degridder_sep_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pipe);
    if (e != cudaSuccess) return e;
    const int ctas = a.nr_subgrids < sms[dev] ? a.nr_subgrids : sms[dev];
    degridder_sep_pipe_kernel<<<dim3((unsigned)ctas), dim3(DP_THREADS), smem_pipe, stream>>>(a, d_todo);
    return cudaGetLastError();
  }
  int tmem_cols = 32;
  while (tmem_cols < ncols) tmem_cols *= 2;
  const int threads = N > 32 ? 512 : 256;
  const size_t smem = (size_t)2 * KCp * (ncols * 16 + 16) + (size_t)2 * KCp * DS_A_CH + (size_t)N * 16 + (size_t)(threads / 128 - 1) * 256 * 16 +
                      DS_UVW_STAGED * 12 + 8 + 8 + 96 + (size_t)ncb * 10 * 4;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  auto kernel = N > 32 ? degridder_sep_kernel<2> : degridder_sep_kernel<1>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kernel<<<dim3((unsigned)a.nr_subgrids), dim3(threads), smem, stream>>>(a, tmem_cols, d_todo);
  return cudaGetLastError();
}

}  // namespace idgb200

#ifdef DP_TRACE
extern "C" int idgb200_dp_trace_read(long long *out) {
  return (int)cudaMemcpyFromSymbol(out, idgb200::dp_trace_buf, sizeof(idgb200::dp_trace_buf));
}
#endif
