// IDG gridder in row-column form on tcgen05 + TMEM (variant 30, the default for FAST sincos).
//
// The reference's phase (gridder_reference.cpp:61-69) is linear in l and m, and its n term is a function
// of l^2 + m^2 that separates to leading order:  n(l, m) = f(l^2 + m^2),  f(s) = s / (1 + sqrt(1 - s)),
//   n(l, m) = f(l^2) + f(m^2) + r(l, m),   r = l^2 m^2 / 4 + O(6)
// so with  gamma_v = w_offset - w_v k_c  the phasor of pixel (y, x) and visibility v = (t, c) is
//   e^{i phase} = Y_v(y) X_v(x) e^{i gamma_v r(y, x)},
//   X_v(x) = exp i[(u_off l_x + w_off f(l_x^2)) - (u l_x + w f(l_x^2)) k_c],   Y_v(y) likewise with v, m_y.
// Where |gamma| r stays below SEP_PHASE_TOL for every visibility of a subgrid (checked per subgrid on the
// device: any w at the bench's image size; planar data at any image size) the last factor is dropped and
//   subgrid[p][y][x] = sum_v  (Y_v(y) vis_v[p]) X_v(x)
// is ONE GEMM per subgrid with the visibilities as the K dimension:
//   D[(p, y)][(x, re|im)] += A[(p, y)][(v, re|im)] * B[(v, re|im)][(x, re|im)]      M = 128 = 4 pols x 32 rows
//     A = fp16(Y vis): 32 phasors + 128 complex products per visibility for the whole subgrid,
//     B = X in fp16 hi + lo: 32 phasors per visibility,                                 N = 4 x 32 columns
// instead of one phasor per (pixel, visibility) - 1024 of them - in gridder_tc.cu: ~35 instead of ~220 dispatch
// cycles per visibility and subgrid, and one M=128 N=128 K=16 MMA per 8 visibilities instead of eight N=16 ones.
// Subgrids that fail the check are left to the per-pixel kernel launched behind this one (a work list).
// Error model: one fp16 rounding per term (of Y vis; X keeps ~22 bits), the same class as gridder_tc.cu's fp16
// phasor; the two half phases are evaluated in fp32 like the reference's single one (tools/sep_prototype.py
// measures the formulation against the oracle in float64).
//
// CTA = one tile of one subgrid (32 rows x <= 64 columns of pixels; a 32 x 32 subgrid is one tile):
//   4 producer warps, each its own pipeline: stage = (timestep, block of 8 channels) -> K = 16; lane = row y
//     makes the A rows of its 4 polarisations (phasor by the three-term recurrence over equally spaced
//     channels, else one sincos per channel; visibilities broadcast from a cp.async ring), lane = column x the
//     B rows; fence.proxy.async, arrive on the stage's full barrier.  Stages go round-robin over the warps,
//     two A buffers and one B buffer per warp, so a warp waits only for its own MMAs of one / two rounds ago.
//   1 issuer warp: waits for the stages in order and issues their tcgen05.mma (one thread, so the MMAs
//     that share the accumulator are ordered), committing each to the producer's empty barrier.
//   epilogue: TMEM -> shared memory (a warp reads its lane quadrant = one polarisation), then per pixel
//     hi + lo, A-terms, taper and the coalesced store exactly as in gridder.cu (gridder_reference.cpp:84-110).
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace idgb200 {

namespace {

constexpr int GS_W = 4;                          // producer warps
constexpr int GS_THREADS = (GS_W + 1) * 32;
constexpr int GS_CB = 8;                         // channels per stage -> K = 16
constexpr int GS_A_BYTES = 128 * 32;             // 128 rows x 16 fp16
constexpr int GS_VSLOTS = 4;                     // visibility ring slots per warp (256 B each)
constexpr float SEP_PHASE_TOL = 1e-4f;           // largest dropped phase |gamma| r (rad)
// Cancellation guard.  The A operand is rounded to fp16 once per term, so a pixel's error is ~2e-4 sqrt(sum |vis|^2)
// whatever the sum comes to: relative to the RESULT it only stays inside the tolerance while the sums do not cancel.
// Incoherent (noise-like) visibilities give max_pixels |D_p| ~ 2.6 E_p, E_p = sqrt(sum_v |vis_v[p]|^2); a tile whose
// pixel sums ALL stay below GS_CANCEL E_p in every polarisation is counted in cancel[subgrid], and a subgrid all of
// whose tiles are is redone by the FP32 kernel launched behind this one.  Unflagged: max error <~ 7.5e-4 / GS_CANCEL.
constexpr float GS_CANCEL = 1.0f;
#ifndef GS_ISSUER_SLEEP_NS
#define GS_ISSUER_SLEEP_NS 300
#endif

__device__ __forceinline__ unsigned pack_h2(const float lo, const float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const unsigned *>(&h);
}
// x - float(fp16 half of a packed word): one FHFMA (tc_common.cuh: pack_phasor)
__device__ __forceinline__ float residual_h(const float x, const unsigned short h) {
  const unsigned short minus_one = 0xbc00u;
  float r;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(h), "h"(minus_one), "f"(x));
  return r;
}

// The issuer's wait: a try_wait's own suspension is short (ncu: 6 polls of 8 instructions per stage, 19 % of the
// kernel's instructions), so the warp sleeps ~100 ns between polls - a stage takes a producer several times that
__device__ __forceinline__ void mbar_wait_sleep(const unsigned bar, const unsigned parity) {
  unsigned done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  for (int spin = 0; !done; spin++) {
    __nanosleep(100);
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1 << 22)) __trap();   // a lost arrival must fail loudly, not hang the GPU
  }
}

// XPL: columns per lane (1: tiles of <= 32 columns, 2: of 33..64)
template <int XPL>
__global__ void __launch_bounds__(GS_THREADS, XPL == 1 ? 4 : 2)
gridder_sep_kernel(const KernelArgs a, const int ytiles, const int xtiles, const int tmem_cols, int *__restrict__ todo,
                   int *__restrict__ cancel, int *__restrict__ cancel_tiles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int N = a.subgrid_size, C = a.nr_channels;
  const int tiles = ytiles * xtiles;
  const int s_local = blockIdx.x / tiles, tile = blockIdx.x - s_local * tiles;
  const int ytile = tile / xtiles, xtile = tile - ytile * xtiles;
  const int y0 = ytile * 32, ny = min(32, N - y0);
  const int x0 = xtile * 64, XT = min(64, N - x0);       // XT % 4 == 0 (the launcher's condition)
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int ncb = (C + GS_CB - 1) / GS_CB;
  const int b_bytes = 4 * XT * 32;                        // B: 4 XT rows x 16 fp16
  // per producer warp: two A buffers and ONE B buffer (of the kernel's largest tile).  A stage's A rows are most of its
  // work, so by the time a warp turns to the B rows the MMA of its previous stage - the last reader of B - has long
  // completed: B needs no second buffer, and 12 KB per producer let four CTAs share an SM
  const int warp_bytes = 2 * GS_A_BYTES + 4 * 64 * 32 / (XPL == 1 ? 2 : 1);

  unsigned char *sStage = smem;                                                        // [GS_W][A0 | A1 | B]
  unsigned char *sVis = sStage + GS_W * warp_bytes;                                    // [GS_W][GS_VSLOTS][256]
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sVis + GS_W * GS_VSLOTS * 256);   // [GS_W][2]
  unsigned long long *empty = full + GS_W * 2;                                         // [GS_W][2]
  unsigned long long *done = empty + GS_W * 2;
  unsigned *s_tmem = reinterpret_cast<unsigned *>(done + 1);
  float *s_red = reinterpret_cast<float *>(s_tmem + 2);   // [48] reductions, scale, verdict, sum |vis|^2 per polarisation
  float *s_wn = s_red + 48;                               // [ncb * 8], zero padded
  float *s_dw = s_wn + ncb * GS_CB;                       // [ncb]
  int *s_lin = reinterpret_cast<int *>(s_dw + ncb);       // [ncb]

  const SubgridCtx ctx = load_ctx(a, s);
  const int nt = ctx.nr_timesteps;
  const int nstages = nt * ncb;
  const float *g_uvw = reinterpret_cast<const float *>(a.uvw) + (size_t)ctx.time_offset * 3;
  const float2 *g_vis = a.visibilities + (size_t)ctx.time_offset * C * NR_POL;

  for (int c = tid; c < ncb * GS_CB; c += GS_THREADS) s_wn[c] = c < C ? a.wavenumbers[c] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < GS_W * 2; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- per subgrid: fp16 scale of the visibilities (gridder_tc.cu) and the separability check
  {
    float amax = 0.f, wmax = 0.f, ss0 = 0.f, ss1 = 0.f;   // ss: sum |vis|^2 of polarisations (0, 1) (even tid) / (2, 3)
    const float4 *v4 = reinterpret_cast<const float4 *>(g_vis);
    for (int i = tid; i < nt * C * 2; i += GS_THREADS) {    // GS_THREADS is even: a thread keeps its parity
      const float4 q = __ldg(&v4[i]);
      amax = fmaxf(fmaxf(amax, fmaxf(fabsf(q.x), fabsf(q.y))), fmaxf(fabsf(q.z), fabsf(q.w)));
      ss0 = __fmaf_rn(q.x, q.x, __fmaf_rn(q.y, q.y, ss0));
      ss1 = __fmaf_rn(q.z, q.z, __fmaf_rn(q.w, q.w, ss1));
    }
    for (int t = tid; t < nt; t += GS_THREADS) wmax = fmaxf(wmax, fabsf(__ldg(&g_uvw[3 * t + 2])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
      if (o > 1) {   // lanes of one parity
        ss0 += __shfl_xor_sync(0xffffffffu, ss0, o);
        ss1 += __shfl_xor_sync(0xffffffffu, ss1, o);
      }
    }
    if (lane == 0) { s_red[warp] = amax; s_red[5 + warp] = wmax; }
    if (lane < 2) { s_red[20 + warp * 4 + 2 * lane] = ss0; s_red[20 + warp * 4 + 2 * lane + 1] = ss1; }
    for (int cb = tid; cb < ncb; cb += GS_THREADS) {
      float dw;
      s_lin[cb] = linear_channels(s_wn, cb * GS_CB, min(GS_CB, C - cb * GS_CB), &dw) ? 1 : 0;
      s_dw[cb] = dw;
    }
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i <= GS_W; i++) { amax = fmaxf(amax, s_red[i]); wmax = fmaxf(wmax, s_red[5 + i]); }
      const unsigned eb = (__float_as_uint(amax) >> 23) & 0xffu;
      const bool ok = eb >= 14u && eb <= 253u;
      s_red[10] = ok ? __uint_as_float((267u - eb) << 23) : 1.f;           // 2^(13 - E)
      s_red[11] = ok ? __uint_as_float((eb - 13u) << 23) : 1.f;            // 2^(E - 13)
      // largest dropped phase: |gamma| <= |w_offset| + max|w| max|k|, r largest at the corner pixel
      float kmax = 0.f;
      for (int c = 0; c < C; c++) kmax = fmaxf(kmax, fabsf(s_wn[c]));
      const double l0 = (0.5 - (N / 2)) * (double)a.image_size / (double)N;   // math.hpp:9-12, x = 0
      const double s1 = l0 * l0;
      const double fn = s1 / (1.0 + sqrt(1.0 - s1)), s2 = 2.0 * s1;
      const double r = s2 > 1.0 ? 1.0 : fabs(s2 / (1.0 + sqrt(1.0 - s2)) - 2.0 * fn);
      const double gmax = (double)fabsf(ctx.w_offset) + (double)wmax * (double)kmax;
      const bool sep = gmax * r <= (double)SEP_PHASE_TOL && isfinite(gmax);
      s_red[12] = sep ? 1.f : 0.f;
      if (tile == 0 && !sep) todo[1 + atomicAdd(&todo[0], 1)] = s_local;   // work list of the per-pixel kernel
      for (int p = 0; p < NR_POL; p++) {          // E_p^2 = sum_v |vis_v[p]|^2
        float e2 = 0.f;
        for (int i = 0; i <= GS_W; i++) e2 += s_red[20 + i * 4 + p];
        s_red[16 + p] = e2;
      }
    }
    __syncthreads();
  }
  if (s_red[12] == 0.f) return;          // the per-pixel kernel behind this launch takes the subgrid
  const float vis_scale = s_red[10], vis_unscale = s_red[11];

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = *s_tmem;

  if (warp < GS_W) {
    // ------------------------------------------------------------------------------------ producers
    // lane = row y of the tile (A) and column(s) x (B); rows / columns beyond the tile are clamped and
    // either ignored by the epilogue (A rows) or not stored (B rows)
    const float m = compute_l(y0 + min(lane, ny - 1), N, a.image_size);
    const float n_y = compute_n(m, 0.f);
    const float off_y = __fmaf_rn(ctx.w_offset, n_y, __fmul_rn(ctx.v_offset, m));
    float l[XPL], n_x[XPL], off_x[XPL];
#pragma unroll
    for (int i = 0; i < XPL; i++) {
      l[i] = compute_l(x0 + min(lane + 32 * i, XT - 1), N, a.image_size);
      n_x[i] = compute_n(l[i], 0.f);
      off_x[i] = __fmaf_rn(ctx.w_offset, n_x[i], __fmul_rn(ctx.u_offset, l[i]));
    }
    unsigned char *my_stage = sStage + warp * warp_bytes;
    unsigned char *my_vis = sVis + warp * GS_VSLOTS * 256;
    const unsigned full_u = smem_u32(full + warp * 2), empty_u = smem_u32(empty + warp * 2);

    // visibilities of stage (t, cb): 8 channels x 32 B, 16 lanes x 16 B; channels beyond C are zero-filled
    auto fetch_vis = [&](int t, int cb, int slot) {
      if (lane < 16) {
        const int c = cb * GS_CB + (lane >> 1);
        const bool valid = c < C;
        const float2 *src = g_vis + ((size_t)t * C + (valid ? c : 0)) * NR_POL + (lane & 1) * 2;
        const unsigned dst = smem_u32(my_vis + slot * 256 + lane * 16);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0));
      }
    };
    // my stages: flat index warp + j * GS_W -> (t, cb); one step = dt timesteps and dcb blocks.  Global loads are
    // issued at the END of an iteration, behind the stage's fence (fence.proxy.async is a MEMBAR.ALL.CTA in SASS: it
    // waits for every load in flight, so anything issued ahead of it is paid at the fence): the visibilities of the
    // stage three ahead into the ring (4 slots), uvw of the stage two ahead into registers.
    const int dt = GS_W / ncb, dcb = GS_W - dt * ncb;
    auto advance = [&](int &tt, int &cc) {
      tt += dt; cc += dcb;
      if (cc >= ncb) { cc -= ncb; tt++; }
    };
    int ft = warp / ncb, fcb = warp - ft * ncb;  // stage whose visibilities are fetched next
    for (int pre = 0; pre < 3; pre++) {
      if (ft < nt) fetch_vis(ft, fcb, pre);
      cp_async_commit();
      advance(ft, fcb);
    }
    int t = warp / ncb, cb = warp - t * ncb;
    int ut = t, ucb = cb;                        // stage whose uvw are loaded next
    float u = 0.f, v = 0.f, w = 0.f, u1 = 0.f, v1 = 0.f, w1 = 0.f;
    if (ut < nt) { u = __ldg(&g_uvw[3 * ut]); v = __ldg(&g_uvw[3 * ut + 1]); w = __ldg(&g_uvw[3 * ut + 2]); }
    advance(ut, ucb);
    if (ut < nt) { u1 = __ldg(&g_uvw[3 * ut]); v1 = __ldg(&g_uvw[3 * ut + 1]); w1 = __ldg(&g_uvw[3 * ut + 2]); }
    advance(ut, ucb);
    for (int j = 0; t < nt; j++) {
      const int buf = j & 1;
      int tn = t, cbn = cb;
      advance(tn, cbn);
      // ---- the stage's phasors first, row and column chains together in one straight-line block (they only need
      // registers, so they run before the waits for the visibilities and for the stage buffer)
      const float *wn8 = s_wn + cb * GS_CB;
      const bool lin = s_lin[cb] != 0;
      const float dw = s_dw[cb];
      float2 phy[GS_CB], phx[XPL][GS_CB];
      {
        float idx[1 + XPL], off[1 + XPL];
        idx[0] = __fmaf_rn(w, n_y, __fmul_rn(v, m));
        off[0] = off_y;
#pragma unroll
        for (int xi = 0; xi < XPL; xi++) {
          idx[1 + xi] = __fmaf_rn(w, n_x[xi], __fmul_rn(u, l[xi]));
          off[1 + xi] = off_x[xi];
        }
        if (lin) {   // first channel by sincos, second by one rotation, then ph[c+1] = 2 cos(delta) ph[c] - ph[c-1]
          float2 p0[1 + XPL], p1[1 + XPL];
          float c2[1 + XPL];
#pragma unroll
          for (int q = 0; q < 1 + XPL; q++) {
            p0[q] = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(-idx[q], wn8[0], off[q]));
            const float2 d = phasor<IDGB200_SINCOS_FAST>(__fmul_rn(-idx[q], dw));
            if (q == 0) p0[q] = __fmul2_rn(p0[q], make_float2(vis_scale, vis_scale));
            p1[q] = ffma2(make_float2(p0[q].y, p0[q].x), make_float2(-d.y, d.y), __fmul2_rn(p0[q], make_float2(d.x, d.x)));
            c2[q] = __fadd_rn(d.x, d.x);
          }
          phy[0] = p0[0]; phy[1] = p1[0];
#pragma unroll
          for (int xi = 0; xi < XPL; xi++) { phx[xi][0] = p0[1 + xi]; phx[xi][1] = p1[1 + xi]; }
#pragma unroll
          for (int i = 2; i < GS_CB; i++) {
            phy[i] = ffma2(phy[i - 1], make_float2(c2[0], c2[0]), make_float2(-phy[i - 2].x, -phy[i - 2].y));
#pragma unroll
            for (int xi = 0; xi < XPL; xi++)
              phx[xi][i] = ffma2(phx[xi][i - 1], make_float2(c2[1 + xi], c2[1 + xi]), make_float2(-phx[xi][i - 2].x, -phx[xi][i - 2].y));
          }
        } else {
#pragma unroll
          for (int i = 0; i < GS_CB; i++) {
            phy[i] = __fmul2_rn(phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(-idx[0], wn8[i], off[0])), make_float2(vis_scale, vis_scale));
#pragma unroll
            for (int xi = 0; xi < XPL; xi++) phx[xi][i] = phasor<IDGB200_SINCOS_FAST>(__fmaf_rn(-idx[1 + xi], wn8[i], off[1 + xi]));
          }
        }
      }
      asm volatile("cp.async.wait_group 2;\n" ::: "memory");
      __syncwarp();
      if (j >= 2) mbar_wait_u(empty_u + buf * 8, ((j >> 1) - 1) & 1);

      unsigned char *A = my_stage + buf * GS_A_BYTES, *B = my_stage + 2 * GS_A_BYTES;
      if (!((IDGB200_ABLATE & 4) && !ablate_never())) {       // ablation builds (tools/power_probe.py): no A rows
        // ---- A rows (p, y): fp16 of scale * Y_c(y) * vis[c][p]
        const float4 *vs = reinterpret_cast<const float4 *>(my_vis + (j & (GS_VSLOTS - 1)) * 256);
#pragma unroll
        for (int kc = 0; kc < 2; kc++) {
          unsigned pk[NR_POL][4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float2 y = phy[kc * 4 + i];
            const float2 yxx = make_float2(y.x, y.x), yny = make_float2(-y.y, y.y);
            const float4 q0 = vs[(kc * 4 + i) * 2], q1 = vs[(kc * 4 + i) * 2 + 1];   // broadcast loads
            const float2 vv[NR_POL] = {make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y),
                                       make_float2(q1.z, q1.w)};
#pragma unroll
            for (int p = 0; p < NR_POL; p++) {   // (vr, vi) * (yr, yr) + (vi, vr) * (-yi, yi)
              const float2 prod = ffma2(make_float2(vv[p].y, vv[p].x), yny, __fmul2_rn(vv[p], yxx));
              pk[p][i] = pack_h2(prod.x, prod.y);
            }
          }
#pragma unroll
          for (int p = 0; p < NR_POL; p++)
            *reinterpret_cast<uint4 *>(A + kc * (128 * 16) + (p * 32 + lane) * 16) = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
        }
      }
      if (j >= 1) mbar_wait_u(empty_u + (buf ^ 1) * 8, ((j - 1) >> 1) & 1);   // the previous stage's MMA has read B
#pragma unroll
      for (int xi = 0; xi < XPL; xi++) {
        if ((IDGB200_ABLATE & 8) && !ablate_never()) continue;     // ablation: no B rows
        // ---- B rows of column x: (hi|lo, re|im) x XT; re row = (cos, -sin), im row = (sin, cos) per visibility
        const int xx = lane + 32 * xi;
#pragma unroll
        for (int kc = 0; kc < 2; kc++) {
          unsigned hre[4], him[4], lre[4], lim[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float2 x = phx[xi][kc * 4 + i];
            him[i] = pack_h2(x.y, x.x);
            hre[i] = pack_h2(x.x, -x.y);
            const float rs = residual_h(x.y, (unsigned short)(him[i] & 0xffffu));
            const float rc = residual_h(x.x, (unsigned short)(him[i] >> 16));
            lim[i] = pack_h2(rs, rc);
            lre[i] = pack_h2(rc, -rs);
          }
          if (xx < XT) {
            unsigned char *Bk = B + kc * (4 * XT * 16) + xx * 16;
            *reinterpret_cast<uint4 *>(Bk) = make_uint4(hre[0], hre[1], hre[2], hre[3]);
            *reinterpret_cast<uint4 *>(Bk + XT * 16) = make_uint4(him[0], him[1], him[2], him[3]);
            *reinterpret_cast<uint4 *>(Bk + 2 * XT * 16) = make_uint4(lre[0], lre[1], lre[2], lre[3]);
            *reinterpret_cast<uint4 *>(Bk + 3 * XT * 16) = make_uint4(lim[0], lim[1], lim[2], lim[3]);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_u + buf * 8) : "memory");
      // behind the fence: visibilities of stage j + 3 (its slot held stage j - 1), uvw of stage j + 2
      if (ft < nt) fetch_vis(ft, fcb, (j + 3) & (GS_VSLOTS - 1));
      cp_async_commit();
      advance(ft, fcb);
      float u2 = 0.f, v2 = 0.f, w2 = 0.f;
      if (ut < nt) { u2 = __ldg(&g_uvw[3 * ut]); v2 = __ldg(&g_uvw[3 * ut + 1]); w2 = __ldg(&g_uvw[3 * ut + 2]); }
      advance(ut, ucb);
      t = tn; cb = cbn;
      u = u1; v = v1; w = w1;
      u1 = u2; v1 = v2; w1 = w2;
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  } else {
    // ------------------------------------------------------------------------------------ issuer warp
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, K-major, N = 4 XT, M = 128
    const unsigned idesc = (1u << 4) | (((unsigned)(4 * XT) >> 3) << 17) | ((128u >> 4) << 24);
    const unsigned full_u = smem_u32(full), empty_u = smem_u32(empty), done_u = smem_u32(done);
    const unsigned stage_u = smem_u32(sStage);
    // flat stage sidx = j GS_W + pw (producer warp pw's j-th stage, buffer j & 1).  Unrolled over the 2 GS_W (pw, buffer)
    // combinations so that barrier addresses and operand descriptors are loop invariants (they were ~20 uniform
    // instructions per stage), and the warp sleeps longer between polls: a producer needs ~3000 clocks for a stage and has
    // two stage buffers, so an MMA that starts a few hundred clocks late delays nobody (was: 9 polls of 6 instructions
    // per stage, a tenth of the kernel's instructions)
    for (int base = 0; base < nstages; base += 2 * GS_W) {
      const unsigned parity = (unsigned)(base / (2 * GS_W)) & 1u;
#pragma unroll
      for (int u = 0; u < 2 * GS_W; u++) {
        const int sidx = base + u;
        if (sidx >= nstages) break;
        const int pw = u % GS_W, buf = u / GS_W;
        {
          const unsigned bar = full_u + (pw * 2 + buf) * 8;
          unsigned ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
          for (int spin = 0; !ok; spin++) {
            __nanosleep(GS_ISSUER_SLEEP_NS);
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            if (spin > (1 << 22)) __trap();   // a lost arrival must fail loudly, not hang the GPU
          }
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const unsigned w_addr = stage_u + pw * warp_bytes;
          if (!((IDGB200_ABLATE & 1) && !ablate_never()))       // ablation: no MMAs
            umma_f16(tmem_base, smem_desc(w_addr + buf * GS_A_BYTES, 128 * 16, 128), smem_desc(w_addr + 2 * GS_A_BYTES, 4 * XT * 16, 128), idesc,
                     sidx > 0 ? 1u : 0u);
          umma_commit_u(empty_u + (pw * 2 + buf) * 8);
          if (sidx == nstages - 1) umma_commit_u(done_u);
        }
        __syncwarp();
      }
    }
    (void)b_bytes;
  }

  // ---- epilogue: accumulators -> shared memory [p][row][XT + 1] -> per pixel A-terms, taper, store
  if (nstages > 0) {
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  __syncthreads();                                  // every producer is past its last stage: the stage buffers are free
  float2 *sD = reinterpret_cast<float2 *>(smem);    // [4][32][XT + 1]
  const int pitch = XT + 1;
  if (warp < GS_W) {
    float2 *row = sD + (warp * 32 + lane) * pitch;
    float m2 = 0.f;                                 // max |D|^2 of this row (cancellation guard)
    for (int c0 = 0; c0 < XT; c0 += 4) {
      unsigned r[16];
      if (nstages > 0) {
        const unsigned taddr = tmem_base + ((unsigned)(warp * 32) << 16) + c0;
#pragma unroll
        for (int g = 0; g < 4; g++)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(r[4 * g]), "=r"(r[4 * g + 1]), "=r"(r[4 * g + 2]), "=r"(r[4 * g + 3])
                       : "r"(taddr + g * XT));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) r[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {   // column groups: hi re, hi im, lo re, lo im
        const float2 d = make_float2((__uint_as_float(r[i]) + __uint_as_float(r[8 + i])) * vis_unscale,
                                     (__uint_as_float(r[4 + i]) + __uint_as_float(r[12 + i])) * vis_unscale);
        row[c0 + i] = d;
        m2 = fmaxf(m2, __fmaf_rn(d.x, d.x, __fmul_rn(d.y, d.y)));
      }
    }
    if (lane >= ny) m2 = 0.f;                       // rows beyond the tile
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    // polarisation `warp`: every pixel sum of the tile below GS_CANCEL E_p (a polarisation without signal does not object)
    if (lane == 0) s_red[40 + warp] = (m2 < GS_CANCEL * GS_CANCEL * s_red[16 + warp] || s_red[16 + warp] == 0.f) ? 1.f : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (!IDGB200_ABLATE &&              // (an ablated kernel's sums are garbage: no redo behind it)
      tid == 0 && nstages > 0 && s_red[40] + s_red[41] + s_red[42] + s_red[43] == 4.f &&
      s_red[16] + s_red[17] + s_red[18] + s_red[19] > 0.f &&
      atomicAdd(&cancel_tiles[s_local], 1) == tiles - 1)            // the last of the subgrid's tiles to say so
    cancel[1 + atomicAdd(&cancel[0], 1)] = s_local;                  // work list of the FP32 kernel
  {
    const size_t plane = (size_t)N * N;
    const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
    const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
    float2 *out = const_cast<float2 *>(a.subgrids) + (size_t)s * NR_POL * plane;
    for (int i = tid; i < ny * XT; i += GS_THREADS) {
      const int yy = i / XT, xx = i - yy * XT;
      const int pixel = (y0 + yy) * N + x0 + xx;
      float2 px[NR_POL];
#pragma unroll
      for (int p = 0; p < NR_POL; p++) px[p] = sD[(p * 32 + yy) * pitch + xx];
      float2 a1[4], a2[4];
      load_jones(a.aterms, (at1 + pixel) * NR_POL, a1);
      load_jones(a.aterms, (at2 + pixel) * NR_POL, a2);
      apply_aterm_gridder(px, a1, a2);
      const float sph = __ldg(&a.spheroidal[pixel]);
      const int dst = subgrid_slot(pixel, N, a.flags);
#pragma unroll
      for (int p = 0; p < NR_POL; p++)
        out[p * plane + dst] = make_float2(__fmul_rn(px[p].x, sph), __fmul_rn(px[p].y, sph));
    }
  }
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}

}  // namespace

bool gridder_sep_supports(int subgrid_size, int nr_channels) {
  return subgrid_size >= 4 && subgrid_size % 4 == 0 && subgrid_size <= 2048 && nr_channels >= 1 && nr_channels <= 4096;
}

// Work lists for the kernels launched behind this one (device; counts and tile counters zeroed by the caller on the
// same stream): d_todo = { n, subgrid[n] } whose dropped phase term exceeds SEP_PHASE_TOL (per-pixel kernel), d_cancel =
// { n, subgrid[n] } all of whose tiles' sums cancel (FP32 kernel), d_cancel_tiles[nr_subgrids] = tiles counted so far
cudaError_t launch_gridder_sep(const KernelArgs &a, int *d_todo, int *d_cancel, int *d_cancel_tiles, cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  if (!gridder_sep_supports(a.subgrid_size, a.nr_channels) || !d_todo || !d_cancel || !d_cancel_tiles) return cudaErrorInvalidValue;
  const int N = a.subgrid_size;
  const int ytiles = (N + 31) / 32, xtiles = (N + 63) / 64;
  const int xt_max = N < 64 ? N : 64;
  const bool wide = xt_max > 32;
  int tmem_cols = 32;
  while (tmem_cols < 4 * xt_max) tmem_cols *= 2;
  const int ncb = (a.nr_channels + GS_CB - 1) / GS_CB;
  const size_t warp_bytes = 2 * GS_A_BYTES + (wide ? 4 * 64 * 32 : 4 * 32 * 32);
  const size_t smem = GS_W * warp_bytes + GS_W * GS_VSLOTS * 256 + (4 * GS_W + 1) * 8 + 8 + 192 + (size_t)ncb * (GS_CB + 2) * 4;
  auto k = wide ? gridder_sep_kernel<2> : gridder_sep_kernel<1>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<dim3((unsigned)a.nr_subgrids * ytiles * xtiles), dim3(GS_THREADS), smem, stream>>>(a, ytiles, xtiles, tmem_cols, d_todo, d_cancel, d_cancel_tiles);
  return cudaGetLastError();
}

}  // namespace idgb200
