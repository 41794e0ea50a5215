// IDG degridder for B200 (sm_100a): the transpose of the gridder.
//
//   P'[y][x]       = A1 * (sph[y][x] * subgrid[s][.][y][x]) * A2^H
//   vis[t][c][pol] = sum_{y,x} P'[y][x][pol] * e^{i phase},
//   phase          = phase_index(t,x,y) * wavenumber[c] - phase_offset(x,y)
//
// Behaviour follows cpu::kernel_degridder_reference
// (app/CPU/kernels/degridder_reference.cpp:6-129); the design is new:
//
//  * one thread block per subgrid.  A prologue applies taper and A-terms once per pixel
//    and leaves P' in shared memory in the 48-byte record layout of the gridder,
//        (Pr0 Pr1 Pr2 Pr3) (-Pi0 Pi0 -Pi1 Pi1) (-Pi2 Pi2 -Pi3 Pi3)
//    plus (l, m, n, phase_offset), so that the complex multiply-add is two FFMA2 with free
//    operand modes and no sign / shuffle instruction (see gridder.cu):
//        acc[pol] += Pr.F32 * (cos,sin)          acc[pol] += (-Pi,Pi) * (sin,cos)
//  * the unit of work is a task = (timestep, block of V = 8 channels): 64 accumulator
//    registers, the pixel record (4 LDS.128) and the phase index (4 FP32 ops) are shared
//    by 8 (pixel, channel) items.  The lanes of a warp are split into 32/PS task slots x
//    PS pixel groups: a lane sums its task over every PS-th pixel, all lanes of a slot
//    read the same shared-memory words (one wavefront per load), and the PS partial sums
//    are combined with __shfl_xor_sync butterflies.  PS is chosen per subgrid (>= 4, up to
//    32 = one warp per task) so that short subgrids still fill the block.  No reference
//    kernel reduces this way: they all loop one thread over all pixels of a visibility
//    (degridder_v6.cu:88-116);
//  * every visibility of the subgrid's time range is written exactly once, by the lanes
//    that own it after the butterfly (no read-modify-write of global memory as in
//    degridder_v4.cu:155-161);
//  * phase_index / phase_offset / phase are evaluated in the CPU binary's operation order
//    (the degridder leaves the w terms unfused, oracle/idg_oracle.c), so the angle fed to
//    sincos is bit-identical to the reference's.
#include "common.cuh"
#include "kernels.h"

namespace idgb200 {

namespace {

constexpr int TILE = 1024;  // pixels resident in shared memory at a time (64 KB)

// SCHEME 3 (default): swizzled FFMA2 on 48-byte pixel records
// SCHEME 1: scalar FFMA baseline (same records)
template <int NT, int V, int SCHEME, int MODE>
__device__ __forceinline__ void degridder_body(const KernelArgs &a, const int s_local) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *s_pix = reinterpret_cast<float4 *>(smem_raw);   // [3][TILE] records, component-major
  float4 *s_lmno = s_pix + 3 * TILE;                       // [TILE]    (l, m, n, phase_offset)

  const int N = a.subgrid_size;
  const int C = a.nr_channels;
  const int npix = N * N;
  const int s = a.subgrid_offset + s_local;
  const int tid = threadIdx.x;
  const SubgridCtx ctx = load_ctx(a, s);
  const int nt = ctx.nr_timesteps;

  const int ncb = (C + V - 1) / V;   // channel blocks per timestep
  const int ntasks = nt * ncb;
  if (ntasks == 0) return;

  // pixel split: smallest power of two >= NT/ntasks, at least 4, at most 32
  int ps_log2 = 2;
  while ((ntasks << ps_log2) < NT && ps_log2 < 5) ps_log2++;
  const int PS = 1 << ps_log2;
  const int slots = NT >> ps_log2;
  const int slot = tid >> ps_log2;
  const int pg = tid & (PS - 1);
  const int nrounds = (ntasks + slots - 1) / slots;
  const int ntiles = (npix + TILE - 1) / TILE;

  const size_t plane = (size_t)npix;
  const size_t at1 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station1) * plane;
  const size_t at2 = ((size_t)ctx.aterm_index * a.nr_stations + ctx.station2) * plane;
  const float2 *sub = a.subgrids + (size_t)s * NR_POL * plane;
  const float *g_uvw = reinterpret_cast<const float *>(a.uvw) + (size_t)ctx.time_offset * 3;
  float4 *g_vis = reinterpret_cast<float4 *>(const_cast<float2 *>(a.visibilities)) +
                  (size_t)ctx.time_offset * C * 2;

  // degridder_reference.cpp:38-74 for the pixels [tile0, tile0 + tile_n)
  auto prologue = [&](int tile0, int tile_n) {
    for (int i = tid; i < tile_n; i += NT) {
      const int q = tile0 + i;
      const int y = q / N, x = q - y * N;
      const float sph = __ldg(&a.spheroidal[q]);
      const int src = subgrid_slot(q, N, a.flags);
      float2 px[NR_POL];
#pragma unroll
      for (int p = 0; p < NR_POL; p++) {
        const float2 v = __ldg(&sub[p * plane + src]);
        px[p] = make_float2(__fmul_rn(sph, v.x), __fmul_rn(sph, v.y));
      }
      float2 a1[4], a2[4];
      load_jones(a.aterms, (at1 + q) * NR_POL, a1);
      load_jones(a.aterms, (at2 + q) * NR_POL, a2);
      apply_aterm_degridder(px, a1, a2);
      s_pix[0 * TILE + i] = make_float4(px[0].x, px[1].x, px[2].x, px[3].x);
      s_pix[1 * TILE + i] = make_float4(-px[0].y, px[0].y, -px[1].y, px[1].y);
      s_pix[2 * TILE + i] = make_float4(-px[2].y, px[2].y, -px[3].y, px[3].y);
      const float l = compute_l(x, N, a.image_size);
      const float m = compute_l(y, N, a.image_size);
      const float n = compute_n(l, m);
      // the CPU binary leaves the w term unfused here (oracle/idg_oracle.c)
      const float off = __fadd_rn(__fmaf_rn(ctx.u_offset, l, __fmul_rn(ctx.v_offset, m)),
                                  __fmul_rn(ctx.w_offset, n));
      s_lmno[i] = make_float4(l, m, n, off);
    }
  };

  for (int round = 0; round < nrounds; round++) {
    const int task = round * slots + slot;
    const bool live = task < ntasks;
    const int tq = live ? task : 0;
    const int t = tq / ncb;
    const int c0 = (tq - t * ncb) * V;
    const float u = g_uvw[t * 3 + 0], v = g_uvw[t * 3 + 1], w = g_uvw[t * 3 + 2];
    float wn[V];
#pragma unroll
    for (int c = 0; c < V; c++) wn[c] = (c0 + c < C) ? __ldg(&a.wavenumbers[c0 + c]) : 0.f;

    float2 acc[V][NR_POL];
#pragma unroll
    for (int c = 0; c < V; c++)
#pragma unroll
      for (int p = 0; p < NR_POL; p++) acc[c][p] = make_float2(0.f, 0.f);

    for (int tile = 0; tile < ntiles; tile++) {
      const int tile0 = tile * TILE;
      const int tile_n = min(TILE, npix - tile0);
      if (ntiles > 1 || round == 0) {
        __syncthreads();  // previous tile fully consumed
        prologue(tile0, tile_n);
        __syncthreads();
      }

#pragma unroll 1
      for (int i = pg; i < tile_n; i += PS) {
        const float4 g = s_lmno[i];
        const float4 q0 = s_pix[0 * TILE + i], q1 = s_pix[1 * TILE + i], q2 = s_pix[2 * TILE + i];
        // degridder_reference.cpp:106 as the CPU binary evaluates it (w term unfused)
        const float idx = __fadd_rn(__fmaf_rn(u, g.x, __fmul_rn(v, g.y)), __fmul_rn(w, g.z));
#pragma unroll
        for (int c = 0; c < V; c++) {
          const float2 ph = phasor<MODE>(__fmaf_rn(idx, wn[c], -g.w));  // :112, (cos, sin)
          if (SCHEME == 3) {
            const float2 hp = make_float2(ph.y, ph.x);                   // LO_HI swizzle
            acc[c][0] = ffma2(make_float2(q0.x, q0.x), ph, acc[c][0]);
            acc[c][1] = ffma2(make_float2(q0.y, q0.y), ph, acc[c][1]);
            acc[c][2] = ffma2(make_float2(q0.z, q0.z), ph, acc[c][2]);
            acc[c][3] = ffma2(make_float2(q0.w, q0.w), ph, acc[c][3]);
            acc[c][0] = ffma2(make_float2(q1.x, q1.y), hp, acc[c][0]);
            acc[c][1] = ffma2(make_float2(q1.z, q1.w), hp, acc[c][1]);
            acc[c][2] = ffma2(make_float2(q2.x, q2.y), hp, acc[c][2]);
            acc[c][3] = ffma2(make_float2(q2.z, q2.w), hp, acc[c][3]);
          } else {
            const float pr[4] = {q0.x, q0.y, q0.z, q0.w};
            const float pi[4] = {q1.y, q1.w, q2.y, q2.w};
#pragma unroll
            for (int p = 0; p < NR_POL; p++) {
              acc[c][p].x = fmaf(pr[p], ph.x, acc[c][p].x);
              acc[c][p].x = fmaf(-pi[p], ph.y, acc[c][p].x);
              acc[c][p].y = fmaf(pr[p], ph.y, acc[c][p].y);
              acc[c][p].y = fmaf(pi[p], ph.x, acc[c][p].y);
            }
          }
        }
      }
    }

    // butterfly over the PS pixel groups, store
#pragma unroll
    for (int c = 0; c < V; c++) {
      float2 sum[NR_POL];
#pragma unroll
      for (int p = 0; p < NR_POL; p++) {
        sum[p] = acc[c][p];
        for (int o = PS >> 1; o > 0; o >>= 1) {
          sum[p].x += __shfl_xor_sync(0xffffffffu, sum[p].x, o);
          sum[p].y += __shfl_xor_sync(0xffffffffu, sum[p].y, o);
        }
      }
      // after the butterfly every lane of the slot holds the totals; lane pg
      // stores the channels with c == pg (mod PS), so the stores are spread
      // over the slot's lanes
      if (live && pg == (c & (PS - 1)) && c0 + c < C) {
        float4 *dst = g_vis + ((size_t)t * C + (c0 + c)) * 2;
        dst[0] = make_float4(sum[0].x, sum[0].y, sum[1].x, sum[1].y);
        dst[1] = make_float4(sum[2].x, sum[2].y, sum[3].x, sum[3].y);
      }
    }
  }
}

// LIST = false: CTA = subgrid blockIdx.x; LIST = true: a fixed number of CTAs loop over the subgrids of a.list
template <int NT, int V, int SCHEME, int MODE, int MINB, bool LIST>
__global__ void __launch_bounds__(NT, MINB)
degridder_kernel(const KernelArgs a) {
  if (!LIST) {
    degridder_body<NT, V, SCHEME, MODE>(a, blockIdx.x);
  } else {
    const int total = a.list[0];
    for (int i = blockIdx.x; i < total; i += gridDim.x) {
      degridder_body<NT, V, SCHEME, MODE>(a, a.list[1 + i]);
      __syncthreads();
    }
  }
}

template <int NT, int V, int SCHEME, int MINB>
cudaError_t launch_t(const KernelArgs &a, int mode, cudaStream_t stream) {
  const size_t smem = (size_t)4 * TILE * sizeof(float4);
  void (*k)(const KernelArgs) = nullptr;
  const bool list = a.list != nullptr;
  switch (mode) {
    case IDGB200_SINCOS_FAST:
      k = list ? degridder_kernel<NT, V, SCHEME, IDGB200_SINCOS_FAST, MINB, true> : degridder_kernel<NT, V, SCHEME, IDGB200_SINCOS_FAST, MINB, false>;
      break;
    case IDGB200_SINCOS_REDUCED: k = degridder_kernel<NT, V, SCHEME, IDGB200_SINCOS_REDUCED, MINB, false>; break;
    case IDGB200_SINCOS_ACCURATE: k = degridder_kernel<NT, V, SCHEME, IDGB200_SINCOS_ACCURATE, MINB, false>; break;
    default: return cudaErrorInvalidValue;
  }
  if (list && mode != IDGB200_SINCOS_FAST) return cudaErrorInvalidValue;   // work lists come from the FAST row-column kernel
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<dim3((unsigned)(list && a.nr_subgrids > LIST_MODE_CTAS ? LIST_MODE_CTAS : a.nr_subgrids)), dim3(NT), smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace

// variant: 0 default: FAST sincos and a shape that fills its channel quads -> tensor-core kernel with
//            fp16 hi + lo phasors (= variant 22, degridder_tc.cu); otherwise the FP32 kernel (= variant 4)
//          4 FP32 kernel: swizzled FFMA2, 256 threads, 8 channels per task, 2 blocks/SM
//          1 scalar-FFMA baseline (256 threads, 4 channels per task)
//          2 swizzled FFMA2, 4 channels per task (3 blocks/SM)
//          3 swizzled FFMA2, 128 threads, 8 channels per task, 3 blocks/SM
//         11 tensor-core kernel with fp16 phasors (FAST sincos only; opt-in: outside the stated
//            tolerance on smooth images); 12..14 with 2, 3, 4 of every 8 pixels'
//            phasors from the FP32 polynomial; 21 with the channel rotation of variant 22
//         22 tensor-core kernel, fp16 hi + lo phasors (FP32-class accuracy), phasors of equally
//            spaced channel quads by rotation from the quad's first channel; 23 without the rotation
// the per-pixel kernel of a FAST launch: 24 / 22 (degridder_tc8.cu / degridder_tc.cu) where the shape fills
// their tiles, else the FP32 kernel
static int fallback_degridder_variant(int subgrid_size, int nr_channels) {
  // the tensor kernel pads the channels to quads of 4 and the pixels to stages of 8
  const int npix = subgrid_size * subgrid_size, ncg = (nr_channels + 3) / 4;
  const bool tc = 4 * nr_channels >= 3 * ncg * 4 && npix >= 256;
  // channel counts that fill groups of 8: the two-tiles-per-warp kernel (degridder_tc8.cu)
  return tc ? ((nr_channels & 7) ? 22 : 24) : 4;
}

int resolve_degridder_variant(int subgrid_size, int nr_channels, int sincos_mode, int variant) {
  if (variant != 0) return variant;
  if (sincos_mode != IDGB200_SINCOS_FAST) return 4;
  // 30: the row-column kernel (degridder_sep.cu), with the per-pixel kernel of this shape behind it
  if (degridder_sep_supports(subgrid_size, nr_channels)) return 30;
  return fallback_degridder_variant(subgrid_size, nr_channels);
}

cudaError_t launch_degridder(const KernelArgs &a, int sincos_mode, int variant, cudaStream_t stream, int *kernels) {
  int nk_local = 0;
  int &nk = kernels ? *kernels : nk_local;
  nk = 0;
  if (a.nr_subgrids == 0) return cudaSuccess;
  variant = resolve_degridder_variant(a.subgrid_size, a.nr_channels, sincos_mode, variant);
  const bool fast = sincos_mode == IDGB200_SINCOS_FAST;
  nk = 1;
  switch (variant) {
    case 4: return launch_t<256, 8, 3, 2>(a, sincos_mode, stream);
    case 22:   // per-pixel tensor-core kernel (degridder_tc.cu), fp16 hi + lo phasors, rotation + recurrence over
               // each quad of equally spaced channels; 23: without the recurrence
      return fast ? launch_degridder_tc(a, 10, true, stream) : cudaErrorInvalidValue;
    case 23:
      return fast ? launch_degridder_tc(a, 10, false, stream) : cudaErrorInvalidValue;
    case 24:   // two M-tiles per warp, groups of 8 channels, fp16 hi + lo phasors (nr_channels % 8 == 0), planar
    case 25:   // subgrids (w = 0) folded onto half the pixels; 25: no recurrence, 28: no folding (24 before it)
    case 28:
      return fast && !(a.nr_channels & 7) ? launch_degridder_tc8(a, variant != 25, variant == 24, stream)
                                          : cudaErrorInvalidValue;
    case 30:     // degridder_sep.cu (row-column form); subgrids it declines go to the per-pixel kernel behind it
    case 31:     // 31: its one-subgrid-per-CTA kernel, 32: its pipelined persistent kernel (30 = 32 where the buffers fit)
    case 32: {
      if (!fast || !degridder_sep_supports(a.subgrid_size, a.nr_channels)) return cudaErrorInvalidValue;
      ScratchLease lease;   // { n_todo, todo[S] }: the subgrids the row-column kernel leaves to the per-pixel kernel
      cudaError_t e = scratch_acquire((size_t)a.nr_subgrids + 1, stream, &lease);
      if (e != cudaSuccess) return e;
      nk = 2;
      e = cudaMemsetAsync(lease.ptr, 0, sizeof(int), stream);
      if (e == cudaSuccess) e = launch_degridder_sep(a, lease.ptr, stream, variant - 30);
      if (e == cudaSuccess) {
        KernelArgs b = a;
        b.list = lease.ptr;
        const int fb = fallback_degridder_variant(a.subgrid_size, a.nr_channels);
        e = fb == 4 ? launch_t<256, 8, 3, 2>(b, sincos_mode, stream)
            : fb == 24 ? launch_degridder_tc8(b, true, true, stream) : launch_degridder_tc(b, 10, true, stream);
      }
      const cudaError_t e2 = scratch_release(lease, stream);
      return e != cudaSuccess ? e : e2;
    }
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace idgb200
