// IDG gridder for B200 (sm_100a).
//
//   subgrid[s][pol][y][x] = sph[y][x] * ( A1^H * ( sum_{t,c} vis[t][c] * e^{i phase} ) * A2 )[pol]
//   phase = phase_offset(x,y) - phase_index(t,x,y) * wavenumber[c]
//
// Behaviour follows cpu::kernel_gridder_reference
// (app/CPU/kernels/gridder_reference.cpp:6-114); the design is new:
//
//  * one thread block per (subgrid, slab of NT*P pixels); a thread keeps P pixels
//    x 4 polarisations of complex sums in registers for the whole time x channel
//    reduction, so nothing is read-modify-written in global memory (the reference's
//    CUDA kernels RMW the subgrid once per 8-channel group, gridder_v4.cu:167-170);
//  * the (timestep, channel) visibility tile of the subgrid is streamed through
//    shared memory in chunks: cp.async (LDGSTS, 16 B per thread, coalesced) brings
//    the raw 32-byte records in while the previous chunk is being consumed, then a
//    short smem->smem pass re-lays them out for the arithmetic below;
//  * the inner product uses packed FP32 (fma.rn.f32x2 -> FFMA2): a visibility
//    component is stored duplicated, (re,re) and (im,im), the phasor is the
//    natural pair (cos,sin) straight out of the two MUFU ops, and each pixel keeps
//       A[pol] += (v.re,v.re) * (cos,sin)      B[pol] += (v.im,v.im) * (cos,sin)
//    which is recombined once at the end:  sum = (A.x - B.y) + i (A.y + B.x).
//    Per (pixel, t, c) item that is 8 FFMA2 + 1 FFMA (phase) + 1 FMUL + 2 MUFU:
//    12 issue slots for 18 FP32-pipe cycles, which is what lets the XU (MUFU)
//    work hide under the FMA pipe instead of competing for issue slots;
//  * phase_offset is hoisted out of the time loop (it does not depend on t) and
//    phase / phase_index are evaluated in the CPU binary's operation order, so the
//    angle fed to sincos is bit-identical to the reference's.
#include "common.cuh"
#include "kernels.h"

namespace idgb200 {

namespace {

// Pixel sums in the packed layout described above.
template <int P>
struct Acc {
  float2 a[P][NR_POL];
  float2 b[P][NR_POL];
};

// 8 FFMA2: one visibility (4 polarisations, duplicated layout) times one phasor
__device__ __forceinline__ void mac_packed(float2 (&A)[NR_POL], float2 (&B)[NR_POL], const float4 &v0,
                                           const float4 &v1, const float4 &v2, const float4 &v3,
                                           const float2 ph) {
  A[0] = ffma2(make_float2(v0.x, v0.y), ph, A[0]);
  B[0] = ffma2(make_float2(v0.z, v0.w), ph, B[0]);
  A[1] = ffma2(make_float2(v1.x, v1.y), ph, A[1]);
  B[1] = ffma2(make_float2(v1.z, v1.w), ph, B[1]);
  A[2] = ffma2(make_float2(v2.x, v2.y), ph, A[2]);
  B[2] = ffma2(make_float2(v2.z, v2.w), ph, B[2]);
  A[3] = ffma2(make_float2(v3.x, v3.y), ph, A[3]);
  B[3] = ffma2(make_float2(v3.z, v3.w), ph, B[3]);
}

// SCHEME 0: FFMA2, duplicated visibilities, software pipelined: the phasors of
//           visibility v+1 are produced (FFMA + FMUL + 2 MUFU per pixel) in between the
//           FFMA2 groups of visibility v, so that one warp keeps the XU and the FMA
//           pipe busy at the same time (default)
// SCHEME 1: scalar FFMA on the raw records (A/B baseline without packed math)
// SCHEME 2: FFMA2 without the software pipeline (A/B)
template <int NT, int P, int SCHEME, int MODE>
__global__ void __launch_bounds__(NT, (NT * P >= 2048 || P >= 8) ? 1 : 2)
gridder_kernel(const KernelArgs a, const int slabs, const int vis_per_chunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool PACKED = (SCHEME != 1);

  const int N = a.subgrid_size;
  const int C = a.nr_channels;
  const int npix = N * N;
  const int s_local = blockIdx.x / slabs;
  const int slab = blockIdx.x - s_local * slabs;
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x;

  const SubgridCtx ctx = load_ctx(a, s);

  // chunk geometry: TB timesteps x C channels <= vis_per_chunk visibilities
  const int TB = max(1, vis_per_chunk / C);
  const int chunk_vis = TB * C;

  // smem carve-up
  //   s_raw  [chunk_vis][8]  floats   raw records (landing zone of cp.async)
  //   s_vis  [chunk_vis][16] floats   duplicated layout   (packed schemes only)
  //   s_uvw  [2][TB][3]      floats   double buffered (consumed in place)
  //   s_wn   [C + 1]         floats   (+1: the pipeline reads one past the end)
  float4 *s_raw = reinterpret_cast<float4 *>(smem_raw);
  float4 *s_vis = s_raw + (size_t)chunk_vis * 2;
  float *s_uvw = reinterpret_cast<float *>(s_vis + (PACKED ? (size_t)chunk_vis * 4 : 0));
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

  Acc<P> acc;
#pragma unroll
  for (int j = 0; j < P; j++)
#pragma unroll
    for (int p = 0; p < NR_POL; p++) {
      acc.a[j][p] = make_float2(0.f, 0.f);
      acc.b[j][p] = make_float2(0.f, 0.f);
    }

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

  // raw -> duplicated layout: (r0,i0,r1,i1) -> (r0,r0,i0,i0) (r1,r1,i1,i1)
  auto relayout = [&](int k) {
    if (PACKED) {
      const int tb = min(TB, nt - k * TB);
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
#pragma unroll
    for (int j = 0; j < P; j++) idx[j] = __fmaf_rn(w, n[j], __fmaf_rn(u, l[j], __fmul_rn(v, m[j])));
  };

  if (nchunks > 0) prefetch(0);

  for (int k = 0; k < nchunks; k++) {
    cp_async_wait_all();
    __syncthreads();  // chunk k landed; everybody is done with the previous s_vis
    relayout(k);
    if (PACKED) __syncthreads();          // s_vis ready, s_raw free again
    if (PACKED && k + 1 < nchunks) prefetch(k + 1);

    const int tb = min(TB, nt - k * TB);
    const float *uvw_k = s_uvw + (k & 1) * TB * 3;

    if (SCHEME == 0) {
      // ---- software pipelined: ph[] always holds the phasors of the visibility about
      // to be accumulated; those of the following one are made while it is consumed
      float idx[P], idxn[P];
      float2 ph[P];
      phase_index(uvw_k, idx);
#pragma unroll
      for (int j = 0; j < P; j++) ph[j] = phasor<MODE>(__fmaf_rn(-idx[j], s_wn[0], off[j]));  // :69

      for (int t = 0; t < tb; t++) {
        // timestep t+1 (the last one of the chunk re-uses t: its phasors are discarded)
        phase_index(uvw_k + 3 * min(t + 1, tb - 1), idxn);
        const float4 *vt = s_vis + (size_t)t * C * 4;
#pragma unroll 2
        for (int c = 0; c < C - 1; c++) {
          const float wn = s_wn[c + 1];
          const float4 v0 = vt[c * 4 + 0], v1 = vt[c * 4 + 1], v2 = vt[c * 4 + 2], v3 = vt[c * 4 + 3];
#pragma unroll
          for (int j = 0; j < P; j++) {
            const float2 nx = phasor<MODE>(__fmaf_rn(-idx[j], wn, off[j]));
            mac_packed(acc.a[j], acc.b[j], v0, v1, v2, v3, ph[j]);
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
            mac_packed(acc.a[j], acc.b[j], v0, v1, v2, v3, ph[j]);
            ph[j] = nx;
            idx[j] = idxn[j];
          }
        }
      }
    } else {
      for (int t = 0; t < tb; t++) {
        float idx[P];
        phase_index(uvw_k + 3 * t, idx);
        if (SCHEME == 2) {
          const float4 *vt = s_vis + (size_t)t * C * 4;
#pragma unroll 2
          for (int c = 0; c < C; c++) {
            const float wn = s_wn[c];
            const float4 v0 = vt[c * 4 + 0], v1 = vt[c * 4 + 1], v2 = vt[c * 4 + 2], v3 = vt[c * 4 + 3];
#pragma unroll
            for (int j = 0; j < P; j++)
              mac_packed(acc.a[j], acc.b[j], v0, v1, v2, v3, phasor<MODE>(__fmaf_rn(-idx[j], wn, off[j])));
          }
        } else {
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
                acc.a[j][p].x = fmaf(vv[p].x, ph.x, acc.a[j][p].x);
                acc.a[j][p].x = fmaf(-vv[p].y, ph.y, acc.a[j][p].x);
                acc.a[j][p].y = fmaf(vv[p].x, ph.y, acc.a[j][p].y);
                acc.a[j][p].y = fmaf(vv[p].y, ph.x, acc.a[j][p].y);
              }
            }
          }
        }
      }
    }
    if (!PACKED) {
      __syncthreads();  // everybody done reading s_raw
      if (k + 1 < nchunks) prefetch(k + 1);
    }
  }

  // ---- epilogue: recombine, A-terms, taper, store (gridder_reference.cpp:84-110)
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
        if (PACKED)
          px[p] = make_float2(acc.a[j][p].x - acc.b[j][p].y, acc.a[j][p].y + acc.b[j][p].x);
        else
          px[p] = acc.a[j][p];
      }
      float2 a1[4], a2[4];
      load_jones(a.aterms, (at1 + pix[j]) * NR_POL, a1);
      load_jones(a.aterms, (at2 + pix[j]) * NR_POL, a2);
      apply_aterm_gridder(px, a1, a2);
      const float sph = __ldg(&a.spheroidal[pix[j]]);
#pragma unroll
      for (int p = 0; p < NR_POL; p++)
        out[p * plane + pix[j]] = make_float2(__fmul_rn(px[p].x, sph), __fmul_rn(px[p].y, sph));
    }
  }
}

template <int NT, int P, int SCHEME>
cudaError_t launch_t(const KernelArgs &a, int mode, cudaStream_t stream) {
  const int npix = a.subgrid_size * a.subgrid_size;
  const int slabs = (npix + NT * P - 1) / (NT * P);
  const int C = a.nr_channels;
  const int vis_per_chunk = max(256, C);
  const int TB = max(1, vis_per_chunk / C);
  const int chunk_vis = TB * C;
  const size_t smem = (size_t)chunk_vis * 32 + (SCHEME != 1 ? (size_t)chunk_vis * 64 : 0) +
                      (size_t)2 * TB * 3 * 4 + (size_t)(C + 1) * 4;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  void (*k)(const KernelArgs, int, int) = nullptr;
  switch (mode) {
    case IDGB200_SINCOS_FAST: k = gridder_kernel<NT, P, SCHEME, IDGB200_SINCOS_FAST>; break;
    case IDGB200_SINCOS_REDUCED: k = gridder_kernel<NT, P, SCHEME, IDGB200_SINCOS_REDUCED>; break;
    case IDGB200_SINCOS_ACCURATE: k = gridder_kernel<NT, P, SCHEME, IDGB200_SINCOS_ACCURATE>; break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<dim3((unsigned)a.nr_subgrids * slabs), dim3(NT), smem, stream>>>(a, slabs, vis_per_chunk);
  return cudaGetLastError();
}

}  // namespace

// variant: 0 default (FFMA2, software pipelined, 256 threads x 4 pixels)
//          1 scalar-FFMA baseline (256 x 4)
//          2 FFMA2 pipelined, 128 threads x 8 pixels
//          3 FFMA2 pipelined, 256 threads x 2 pixels
//          4 FFMA2 without the software pipeline (256 x 4)
cudaError_t launch_gridder(const KernelArgs &a, int sincos_mode, int variant, cudaStream_t stream) {
  if (a.nr_subgrids == 0) return cudaSuccess;
  switch (variant) {
    case 0: return launch_t<256, 4, 0>(a, sincos_mode, stream);
    case 1: return launch_t<256, 4, 1>(a, sincos_mode, stream);
    case 2: return launch_t<128, 8, 0>(a, sincos_mode, stream);
    case 3: return launch_t<256, 2, 0>(a, sincos_mode, stream);
    case 4: return launch_t<256, 4, 2>(a, sincos_mode, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace idgb200
