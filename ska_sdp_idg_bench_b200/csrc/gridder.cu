// IDG gridder for B200 (sm_100a).
//
//   subgrid[s][pol][y][x] = sph[y][x] * ( A1^H * ( sum_{t,c} vis[t][c] * e^{i phase} ) * A2 )[pol]
//   phase = phase_offset(x,y) - phase_index(t,x,y) * wavenumber[c]
//
// Behaviour follows cpu::kernel_gridder_reference
// (app/CPU/kernels/gridder_reference.cpp:6-114); the design is new and is shaped by
// what was measured on B200 (tools/microbench.cu, pipes.cu, operands.cu, mixv.cu):
//
//  * The SM sub-partition's dispatch port is the roof, not a pipe: every instruction
//    holds it for one cycle, a packed FFMA2 for two (three when it needs three register
//    reads from one bank).  MUFU work therefore does not "hide" under the FMAs; what
//    counts is instructions per (pixel, timestep, channel) item and register-file reads
//    per FMA.  The floor is 16 (FMA) + 1 (phase) + 1 (MUFU range scale) + 2 (MUFU) = 20
//    cycles per item against 18 at FP32 peak.
//  * One thread block per (subgrid, slab of NT*P pixels); a thread keeps P pixels x 4
//    polarisations of complex sums in registers for the whole time x channel reduction,
//    so nothing is read-modify-written in global memory (the reference's CUDA kernels RMW
//    the subgrid once per 8-channel group, gridder_v4.cu:167-170).  P = 8 makes the shared
//    operand loads (3 LDS.128 + 1 LDS per visibility) and the per-timestep phase index
//    cost 0.6 instructions per item.
//  * The (timestep, channel) visibility tile is streamed through shared memory in chunks:
//    cp.async (LDGSTS, 16 B per thread, coalesced) lands the raw 32-byte records while the
//    previous chunk is consumed; a short smem->smem pass re-lays each record out as
//        (vr0 vr1 vr2 vr3) (-vi0 vi0 -vi1 vi1) (-vi2 vi2 -vi3 vi3)          [48 bytes]
//    so that the complex multiply-add needs no sign or shuffle instruction:
//        acc[pol] += vr.F32 (broadcast) * (cos,sin)        FFMA2 R, R.F32, R.F32x2.HI_LO, R
//        acc[pol] += (-vi,vi)           * (sin,cos)        FFMA2 R, R.F32x2, R.F32x2.LO_HI, R
//    (scalar-broadcast and LO_HI-swizzled operands are free FFMA2 operand modes).  The phasor
//    pair comes straight out of MUFU.COS / MUFU.SIN into adjacent registers.
//  * Pixels are processed one after the other inside a visibility so that the 8 FFMA2 of a
//    pixel share the phasor operand slot (register reuse cache) and stay at 2 dispatch cycles.
//  * phase_offset is hoisted out of the time loop (it does not depend on t) and phase /
//    phase_index are evaluated in the CPU binary's operation order, so the angle fed to
//    sincos is bit-identical to the reference's.
#include "common.cuh"
#include "kernels.h"

namespace idgb200 {

namespace {

// ---------------------------------------------------------------------------------------
// SCHEME 3 (default): swizzled FFMA2 on the 48-byte records
// SCHEME 0: FFMA2 on duplicated (re,re)(im,im) records with rotated accumulators,
//           software pipelined phasors (first design; kept for A/B)
// SCHEME 1: scalar FFMA on the raw records (A/B baseline without packed math)
template <int SCHEME>
struct Layout;
template <>
struct Layout<3> { static constexpr int F4_PER_VIS = 3; static constexpr int ACC = 4; };
template <>
struct Layout<0> { static constexpr int F4_PER_VIS = 4; static constexpr int ACC = 8; };
template <>
struct Layout<1> { static constexpr int F4_PER_VIS = 2; static constexpr int ACC = 4; };

template <int NT, int P, int SCHEME, int MODE>
__device__ __forceinline__ void gridder_body(const KernelArgs &a, const int vis_per_chunk, const int s_local, const int slab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int F4 = Layout<SCHEME>::F4_PER_VIS;
  constexpr int NACC = Layout<SCHEME>::ACC;
  constexpr bool RELAYOUT = (SCHEME != 1);
  constexpr bool LMN_IN_SMEM = (P % 4 == 0);

  const int N = a.subgrid_size;
  const int C = a.nr_channels;
  const int npix = N * N;
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x;

  const SubgridCtx ctx = load_ctx(a, s);

  // chunk geometry: TB timesteps x C channels <= vis_per_chunk visibilities
  const int TB = max(1, vis_per_chunk / C);
  const int chunk_vis = TB * C;

  // smem carve-up
  //   s_raw  [chunk_vis][2]   float4  raw records (landing zone of cp.async)
  //   s_vis  [chunk_vis][F4]  float4  re-laid-out records (RELAYOUT schemes)
  //   s_lmn  [3][P/4][NT]     float4  l, m, n of this thread's pixels (P % 4 == 0)
  //   s_uvw  [2][TB][3]       float   double buffered (consumed in place)
  //   s_wn   [C + 1]          float   (+1: the software pipeline reads one past the end)
  float4 *s_raw = reinterpret_cast<float4 *>(smem_raw);
  float4 *s_vis = s_raw + (size_t)chunk_vis * 2;
  float4 *s_lmn = s_vis + (RELAYOUT ? (size_t)chunk_vis * F4 : 0);
  float *s_uvw = reinterpret_cast<float *>(s_lmn + (LMN_IN_SMEM ? 3 * (P / 4) * NT : 0));
  float *s_wn = s_uvw + 2 * TB * 3;

  for (int c = tid; c <= C; c += NT) s_wn[c] = a.wavenumbers[c < C ? c : 0];

  // per-pixel constants
  float l[P], m[P], n[P], off[P];
  int pix[P];
#pragma unroll
  for (int j = 0; j < P; j++) {
    pix[j] = slab * (NT * P) + j * NT + tid;
    const int q = min(pix[j], npix - 1);
    const int y = q / N, x = q - y * N;
    l[j] = compute_l(x, N, a.image_size);
    m[j] = compute_l(y, N, a.image_size);
    n[j] = compute_n(l[j], m[j]);
    // gridder_reference.cpp:64 as the CPU binary contracts it
    off[j] = __fmaf_rn(ctx.w_offset, n[j], __fmaf_rn(ctx.u_offset, l[j], __fmul_rn(ctx.v_offset, m[j])));
  }
  if (LMN_IN_SMEM) {  // l, m, n are needed once per timestep only: park them in smem
#pragma unroll
    for (int g = 0; g < P / 4; g++) {
      s_lmn[(0 * (P / 4) + g) * NT + tid] = make_float4(l[4 * g], l[4 * g + 1], l[4 * g + 2], l[4 * g + 3]);
      s_lmn[(1 * (P / 4) + g) * NT + tid] = make_float4(m[4 * g], m[4 * g + 1], m[4 * g + 2], m[4 * g + 3]);
      s_lmn[(2 * (P / 4) + g) * NT + tid] = make_float4(n[4 * g], n[4 * g + 1], n[4 * g + 2], n[4 * g + 3]);
    }
  }

  float2 acc[P][NACC];
#pragma unroll
  for (int j = 0; j < P; j++)
#pragma unroll
    for (int p = 0; p < NACC; p++) acc[j][p] = make_float2(0.f, 0.f);

  const float4 *g_vis = reinterpret_cast<const float4 *>(a.visibilities) + (size_t)ctx.time_offset * C * 2;
  const float *g_uvw = reinterpret_cast<const float *>(a.uvw) + (size_t)ctx.time_offset * 3;
  const int nt = ctx.nr_timesteps;
  const int nchunks = (nt + TB - 1) / TB;

  // issue the async copies of chunk k (raw visibilities + uvw)
  auto prefetch = [&](int k) {
    const int t0 = k * TB;
    const int tb = min(TB, nt - t0);
    const float4 *src = g_vis + (size_t)t0 * C * 2;
    for (int i = tid; i < tb * C * 2; i += NT) cp_async16(&s_raw[i], &src[i]);
    float *dst_uvw = s_uvw + (k & 1) * TB * 3;
    for (int i = tid; i < tb * 3; i += NT) cp_async4(&dst_uvw[i], &g_uvw[(size_t)t0 * 3 + i]);
    cp_async_commit();
  };

  auto relayout = [&](int k) {
    const int tb = min(TB, nt - k * TB);
    if (SCHEME == 3) {
      for (int i = tid; i < tb * C; i += NT) {
        const float4 r0 = s_raw[2 * i], r1 = s_raw[2 * i + 1];
        s_vis[3 * i + 0] = make_float4(r0.x, r0.z, r1.x, r1.z);
        s_vis[3 * i + 1] = make_float4(-r0.y, r0.y, -r0.w, r0.w);
        s_vis[3 * i + 2] = make_float4(-r1.y, r1.y, -r1.w, r1.w);
      }
    } else if (SCHEME == 0) {  // (r0,i0,r1,i1) -> (r0,r0,i0,i0) (r1,r1,i1,i1)
      for (int i = tid; i < tb * C * 2; i += NT) {
        const float4 r = s_raw[i];
        s_vis[2 * i] = make_float4(r.x, r.x, r.y, r.y);
        s_vis[2 * i + 1] = make_float4(r.z, r.z, r.w, r.w);
      }
    }
  };

  // gridder_reference.cpp:61 as contracted by the CPU binary
  auto phase_index = [&](const float *uvw_t, float (&idx)[P]) {
    const float u = uvw_t[0], v = uvw_t[1], w = uvw_t[2];
    if (LMN_IN_SMEM) {
#pragma unroll
      for (int g = 0; g < P / 4; g++) {
        const float4 l4 = lds128_pinned(&s_lmn[(0 * (P / 4) + g) * NT + tid]);
        const float4 m4 = lds128_pinned(&s_lmn[(1 * (P / 4) + g) * NT + tid]);
        const float4 n4 = lds128_pinned(&s_lmn[(2 * (P / 4) + g) * NT + tid]);
        idx[4 * g + 0] = __fmaf_rn(w, n4.x, __fmaf_rn(u, l4.x, __fmul_rn(v, m4.x)));
        idx[4 * g + 1] = __fmaf_rn(w, n4.y, __fmaf_rn(u, l4.y, __fmul_rn(v, m4.y)));
        idx[4 * g + 2] = __fmaf_rn(w, n4.z, __fmaf_rn(u, l4.z, __fmul_rn(v, m4.z)));
        idx[4 * g + 3] = __fmaf_rn(w, n4.w, __fmaf_rn(u, l4.w, __fmul_rn(v, m4.w)));
      }
    } else {
#pragma unroll
      for (int j = 0; j < P; j++) idx[j] = __fmaf_rn(w, n[j], __fmaf_rn(u, l[j], __fmul_rn(v, m[j])));
    }
  };

  if (nchunks > 0) prefetch(0);

  for (int k = 0; k < nchunks; k++) {
    cp_async_wait_all();
    __syncthreads();  // chunk k landed; everybody is done with the previous s_vis
    relayout(k);
    if (RELAYOUT) __syncthreads();          // s_vis ready, s_raw free again
    if (RELAYOUT && k + 1 < nchunks) prefetch(k + 1);

    const int tb = min(TB, nt - k * TB);
    const float *uvw_k = s_uvw + (k & 1) * TB * 3;

    if (SCHEME == 3) {
      for (int t = 0; t < tb; t++) {
        float idx[P];
        phase_index(uvw_k + 3 * t, idx);
        const float4 *vt = s_vis + (size_t)t * C * 3;
        const float4 *vp = vt;
#pragma unroll 1
        for (int c = 0; c < C; c++, vp += 3) {
          const float wn = s_wn[c];
          const float4 q0 = vp[0], q1 = vp[1], q2 = vp[2];
          // the phasor of pixel j+1 is started before the FFMA2 group of pixel j so that
          // the group does not open with a scoreboard wait on the MUFU results
          float2 ph = phasor<MODE>(__fmaf_rn(-idx[0], wn, off[0]));  // :69, (cos, sin)
#pragma unroll
          for (int j = 0; j < P; j++) {
            float2 nph = ph;
            if (j + 1 < P) nph = phasor<MODE>(__fmaf_rn(-idx[j + 1], wn, off[j + 1]));
            const float2 hp = make_float2(ph.y, ph.x);                        // LO_HI swizzle
            acc[j][0] = ffma2(make_float2(q0.x, q0.x), ph, acc[j][0]);
            acc[j][1] = ffma2(make_float2(q0.y, q0.y), ph, acc[j][1]);
            acc[j][2] = ffma2(make_float2(q0.z, q0.z), ph, acc[j][2]);
            acc[j][3] = ffma2(make_float2(q0.w, q0.w), ph, acc[j][3]);
            acc[j][0] = ffma2(make_float2(q1.x, q1.y), hp, acc[j][0]);
            acc[j][1] = ffma2(make_float2(q1.z, q1.w), hp, acc[j][1]);
            acc[j][2] = ffma2(make_float2(q2.x, q2.y), hp, acc[j][2]);
            acc[j][3] = ffma2(make_float2(q2.z, q2.w), hp, acc[j][3]);
            ph = nph;
          }
        }
      }
    } else if (SCHEME == 0) {
      // software pipelined: ph[] holds the phasors of the visibility about to be
      // accumulated; those of the following one are made while it is consumed
      float idx[P], idxn[P];
      float2 ph[P];
      phase_index(uvw_k, idx);
#pragma unroll
      for (int j = 0; j < P; j++) ph[j] = phasor<MODE>(__fmaf_rn(-idx[j], s_wn[0], off[j]));

      auto mac = [&](int j, const float4 &v0, const float4 &v1, const float4 &v2, const float4 &v3) {
        acc[j][0] = ffma2(make_float2(v0.x, v0.y), ph[j], acc[j][0]);
        acc[j][4] = ffma2(make_float2(v0.z, v0.w), ph[j], acc[j][4]);
        acc[j][1] = ffma2(make_float2(v1.x, v1.y), ph[j], acc[j][1]);
        acc[j][5] = ffma2(make_float2(v1.z, v1.w), ph[j], acc[j][5]);
        acc[j][2] = ffma2(make_float2(v2.x, v2.y), ph[j], acc[j][2]);
        acc[j][6] = ffma2(make_float2(v2.z, v2.w), ph[j], acc[j][6]);
        acc[j][3] = ffma2(make_float2(v3.x, v3.y), ph[j], acc[j][3]);
        acc[j][7] = ffma2(make_float2(v3.z, v3.w), ph[j], acc[j][7]);
      };
      for (int t = 0; t < tb; t++) {
        phase_index(uvw_k + 3 * min(t + 1, tb - 1), idxn);  // last one: result discarded
        const float4 *vt = s_vis + (size_t)t * C * 4;
#pragma unroll 2
        for (int c = 0; c < C - 1; c++) {
          const float wn = s_wn[c + 1];
          const float4 v0 = vt[c * 4 + 0], v1 = vt[c * 4 + 1], v2 = vt[c * 4 + 2], v3 = vt[c * 4 + 3];
#pragma unroll
          for (int j = 0; j < P; j++) {
            const float2 nx = phasor<MODE>(__fmaf_rn(-idx[j], wn, off[j]));
            mac(j, v0, v1, v2, v3);
            ph[j] = nx;
          }
        }
        {  // last channel of t: the next phasors belong to (t+1, channel 0)
          const float wn = s_wn[0];
          const int c = C - 1;
          const float4 v0 = vt[c * 4 + 0], v1 = vt[c * 4 + 1], v2 = vt[c * 4 + 2], v3 = vt[c * 4 + 3];
#pragma unroll
          for (int j = 0; j < P; j++) {
            const float2 nx = phasor<MODE>(__fmaf_rn(-idxn[j], wn, off[j]));
            mac(j, v0, v1, v2, v3);
            ph[j] = nx;
            idx[j] = idxn[j];
          }
        }
      }
    } else {
      for (int t = 0; t < tb; t++) {
        float idx[P];
        phase_index(uvw_k + 3 * t, idx);
        const float4 *vt = s_raw + (size_t)t * C * 2;
#pragma unroll 2
        for (int c = 0; c < C; c++) {
          const float wn = s_wn[c];
          const float4 v01 = vt[c * 2 + 0], v23 = vt[c * 2 + 1];
          const float2 vv[4] = {make_float2(v01.x, v01.y), make_float2(v01.z, v01.w),
                                make_float2(v23.x, v23.y), make_float2(v23.z, v23.w)};
#pragma unroll
          for (int j = 0; j < P; j++) {
            const float2 ph = phasor<MODE>(__fmaf_rn(-idx[j], wn, off[j]));
#pragma unroll
            for (int p = 0; p < NR_POL; p++) {
              acc[j][p].x = fmaf(vv[p].x, ph.x, acc[j][p].x);
              acc[j][p].x = fmaf(-vv[p].y, ph.y, acc[j][p].x);
              acc[j][p].y = fmaf(vv[p].x, ph.y, acc[j][p].y);
              acc[j][p].y = fmaf(vv[p].y, ph.x, acc[j][p].y);
            }
          }
        }
      }
    }
    if (!RELAYOUT) {
      __syncthreads();  // everybody done reading s_raw
      if (k + 1 < nchunks) prefetch(k + 1);
    }
  }

  // ---- epilogue: A-terms, taper, store (gridder_reference.cpp:84-110)
  const size_t plane = (size_t)npix;
  const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
  const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
  float2 *out = const_cast<float2 *>(a.subgrids) + (size_t)s * NR_POL * plane;
#pragma unroll
  for (int j = 0; j < P; j++) {
    if (pix[j] < npix) {
      float2 px[NR_POL];
#pragma unroll
      for (int p = 0; p < NR_POL; p++) {
        if (SCHEME == 0)  // rotated sums: A + i B
          px[p] = make_float2(acc[j][p].x - acc[j][p + 4 < NACC ? p + 4 : p].y,
                              acc[j][p].y + acc[j][p + 4 < NACC ? p + 4 : p].x);
        else
          px[p] = acc[j][p];
      }
      float2 a1[4], a2[4];
      load_jones(a.aterms, (at1 + pix[j]) * NR_POL, a1);
      load_jones(a.aterms, (at2 + pix[j]) * NR_POL, a2);
      apply_aterm_gridder(px, a1, a2);
      const float sph = __ldg(&a.spheroidal[pix[j]]);
      const int slot = subgrid_slot(pix[j], a.subgrid_size, a.flags);
#pragma unroll
      for (int p = 0; p < NR_POL; p++)
        out[p * plane + slot] = make_float2(__fmul_rn(px[p].x, sph), __fmul_rn(px[p].y, sph));
    }
  }
}

// LIST = false: CTA (s, slab) = blockIdx.x; LIST = true: a fixed number of CTAs loop over the subgrids of a.list
template <int NT, int P, int SCHEME, int MODE, int MINB, bool LIST>
__global__ void __launch_bounds__(NT, MINB)
gridder_kernel(const KernelArgs a, const int slabs, const int vis_per_chunk) {
  if (!LIST) {
    const int s_local = blockIdx.x / slabs;
    gridder_body<NT, P, SCHEME, MODE>(a, vis_per_chunk, s_local, blockIdx.x - s_local * slabs);
  } else {
    const int total = a.list[0] * slabs;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int i = item / slabs;
      gridder_body<NT, P, SCHEME, MODE>(a, vis_per_chunk, a.list[1 + i], item - i * slabs);
      __syncthreads();
    }
  }
}

template <int NT, int P, int SCHEME, int MINB>
cudaError_t launch_t(const KernelArgs &a, int mode, cudaStream_t stream) {
  const int npix = a.subgrid_size * a.subgrid_size;
  const int slabs = (npix + NT * P - 1) / (NT * P);
  const int C = a.nr_channels;
  const int vis_per_chunk = max(256, C);   // 384 / 512 measured within 0.15 % (34.60 / 34.58 / 34.55 ms), 1024 19 % slower
  const int TB = max(1, vis_per_chunk / C);
  const int chunk_vis = TB * C;
  const size_t smem = (size_t)chunk_vis * 32 + (SCHEME != 1 ? (size_t)chunk_vis * 16 * Layout<SCHEME>::F4_PER_VIS : 0) +
                      (P % 4 == 0 ? (size_t)3 * (P / 4) * NT * 16 : 0) + (size_t)2 * TB * 3 * 4 +
                      (size_t)(C + 1) * 4;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  void (*k)(const KernelArgs, int, int) = nullptr;
  const bool list = a.list != nullptr;
  switch (mode) {
    case IDGB200_SINCOS_FAST:
      k = list ? gridder_kernel<NT, P, SCHEME, IDGB200_SINCOS_FAST, MINB, true> : gridder_kernel<NT, P, SCHEME, IDGB200_SINCOS_FAST, MINB, false>;
      break;
    case IDGB200_SINCOS_REDUCED: k = gridder_kernel<NT, P, SCHEME, IDGB200_SINCOS_REDUCED, MINB, false>; break;
    case IDGB200_SINCOS_ACCURATE: k = gridder_kernel<NT, P, SCHEME, IDGB200_SINCOS_ACCURATE, MINB, false>; break;
    default: return cudaErrorInvalidValue;
  }
  if (list && mode != IDGB200_SINCOS_FAST) return cudaErrorInvalidValue;   // work lists come from the FAST row-column kernels
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long ctas = (long long)a.nr_subgrids * slabs;
  k<<<dim3((unsigned)(list && ctas > LIST_MODE_CTAS ? LIST_MODE_CTAS : ctas)), dim3(NT), smem, stream>>>(a, slabs, vis_per_chunk);
  return cudaGetLastError();
}

}  // namespace

// variant: 0 default: FAST sincos -> 30, the row-column kernel (gridder_sep.cu; subgrid sizes that are a multiple
//            of 4), else the per-pixel tensor-core kernel (24 / 21) where the shape fills >= 3/4 of its
//            8-channel blocks and 128-pixel tiles, else 10; other sincos modes -> 10
//         10 FP32 kernel (this file): swizzled FFMA2; 128 threads x 8 pixels (x 4 blocks/SM) when the
//            subgrid has >= 1024 pixels, 128 x 4 below that
//         21 per-pixel tensor-core kernel (gridder_tc.cu, FAST sincos only); blocks of 8 equally spaced channels
//            get their phasors by rotation + three-term recurrence, other blocks by MUFU / FP32 polynomial
//         24 the same with two channel blocks (K = 32) per stage, single-buffered
//         22 / 23 as 21 with fp16 hi + lo phasors (FP32-class accuracy), with / without the recurrence
//         30 gridder_sep.cu with the per-pixel kernel of the shape (24 / 21 / 10) behind it for the subgrids
//            it declines
// the per-pixel kernel of a FAST launch: 24 / 21 (gridder_tc.cu) where the shape fills its tiles, else the FP32 kernel
static int fallback_gridder_variant(int subgrid_size, int nr_channels) {
  const int npix = subgrid_size * subgrid_size;
  const int ncb = (nr_channels + 7) / 8, tiles = (npix + 127) / 128;
  const bool tc = 4 * nr_channels >= 3 * ncb * 8 && 4 * npix >= 3 * tiles * 128 && nr_channels <= 1024;
  return !tc ? 10 : (ncb % 2 == 0 ? 24 : 21);
}

int resolve_gridder_variant(int subgrid_size, int nr_channels, int sincos_mode, int variant) {
  if (variant != 0) return variant;
  if (sincos_mode != IDGB200_SINCOS_FAST) return 10;
  // 30: the row-column kernel (gridder_sep.cu), with the per-pixel kernel of this shape behind it
  if (gridder_sep_supports(subgrid_size, nr_channels)) return 30;
  return fallback_gridder_variant(subgrid_size, nr_channels);
}

cudaError_t launch_gridder_fp32(const KernelArgs &a, int sincos_mode, cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  return a.subgrid_size * a.subgrid_size >= 1024 ? launch_t<128, 8, 3, 4>(a, sincos_mode, stream)
                                                 : launch_t<128, 4, 3, 4>(a, sincos_mode, stream);
}

cudaError_t launch_gridder(const KernelArgs &a, int sincos_mode, int variant, cudaStream_t stream, int *kernels) {
  int nk_local = 0;
  int &nk = kernels ? *kernels : nk_local;
  nk = 0;
  if (a.nr_subgrids == 0) return cudaSuccess;
  variant = resolve_gridder_variant(a.subgrid_size, a.nr_channels, sincos_mode, variant);
  const bool fast = sincos_mode == IDGB200_SINCOS_FAST;
  cudaError_t e = cudaErrorInvalidValue;
  switch (variant) {
    case 10:
      nk = 1;
      return launch_gridder_fp32(a, sincos_mode, stream);
    case 22:   // as 21 with fp16 hi + lo phasors (FP32-class accuracy); 23: the same without the rotation
      nk = 1;
      return fast ? launch_gridder_tc(a, 10, true, stream) : cudaErrorInvalidValue;
    case 23:
      nk = 1;
      return fast ? launch_gridder_tc(a, 10, false, stream) : cudaErrorInvalidValue;
    case 24:   // as 21 with 16 channels (K = 32) per stage, single-buffered
      nk = 1;
      return fast ? launch_gridder_tc(a, 11, true, stream) : cudaErrorInvalidValue;
    case 21:   // tensor-core kernel, phasors of equally spaced channel blocks by rotation (else as 12)
      nk = 1;
      return fast ? launch_gridder_tc(a, 3, true, stream) : cudaErrorInvalidValue;
    case 30: {   // gridder_sep.cu (row-column form) and, behind it, two list-mode launches over what it finds: the
                 // per-pixel kernel of the shape for the subgrids that are not separable and the FP32 kernel for those
                 // whose sums cancel below the fp16 operand's error model.  Scratch (zeroed on the stream first):
                 // { n_todo, todo[S], n_cancel, cancel[S], cancel tiles per subgrid [S] }
      if (!fast || !gridder_sep_supports(a.subgrid_size, a.nr_channels)) return cudaErrorInvalidValue;
      const size_t S = (size_t)a.nr_subgrids;
      ScratchLease lease;
      e = scratch_acquire(3 * S + 2, stream, &lease);
      if (e != cudaSuccess) return e;
      int *todo = lease.ptr, *cancel = lease.ptr + S + 1, *cancel_tiles = lease.ptr + 2 * S + 2;
      nk = 3;
      e = cudaMemsetAsync(lease.ptr, 0, (3 * S + 2) * sizeof(int), stream);
      if (e == cudaSuccess) e = launch_gridder_sep(a, todo, cancel, cancel_tiles, stream);
      KernelArgs b = a;
      if (e == cudaSuccess) {
        b.list = todo;
        const int fb = fallback_gridder_variant(a.subgrid_size, a.nr_channels);
        e = fb == 10 ? launch_gridder_fp32(b, sincos_mode, stream) : launch_gridder_tc(b, fb == 24 ? 11 : 3, true, stream);
      }
      if (e == cudaSuccess) {
        b.list = cancel;
        e = launch_gridder_fp32(b, sincos_mode, stream);
      }
      const cudaError_t e2 = scratch_release(lease, stream);
      return e != cudaSuccess ? e : e2;
    }
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace idgb200
