// Grid adder and splitter ("next" rows around the gridder / degridder).
// Adder: every subgrid is accumulated into the grid at its coordinate,
//     grid[pol][y0 + y][x0 + x] += subgrids[s][pol][y'][x'],     pixels outside the grid dropped,
// splitter (its inverse, feeding the degridder): subgrids[s][pol][y'][x'] = grid[pol][y0 + y][x0 + x],
// 0 where the subgrid overhangs the grid; (y', x') = (y, x), or with IDGB200_FLAG_FFT_SHIFT
// ((y + N/2) mod N, (x + N/2) mod N) - the index shift the full pipeline applies between the
// subgrid FFT and the grid.  Both are HBM / L2-bound byte movers: 8 B per pixel read and one 8-byte
// reduction (adder) or store (splitter), a warp along x so that global accesses are contiguous.
// The grid is cut into `nr_parts` blocks of `rows_per_part` rows, each behind its own base
// pointer ([pol][rows_per_part][grid_size] complex64 per part).  On one GPU that is one part.
// On N GPUs part r is the slice rank r owns after the reduction, and the pointer may be a
// *peer* address (NVLink / NVSwitch): the red.global.add.v2.f32 of this kernel then IS the
// reduce-scatter - every rank adds its own subgrids straight into the owners' slices, no partial
// grid, no second pass.  tools/adder_reduce_scatter.py measures it against local grid +
// ncclReduceScatter: which one wins depends on how often a grid cell is hit - with
// many subgrids per cell the local grid absorbs the overlap in L2 before anything crosses NVLink.
// The reference has no adder (only idg::Grid, app/common/types.hpp:358-370); oracle/idg_next_oracle.c
// states the sum the tests check (parity unpinned).  Summation order across subgrids is that of the
// atomics, so results agree with the oracle to fp32 rounding of the sum, not bit for bit.
#include "common.cuh"
#include "kernels.h"

namespace idgb200 {

namespace {

constexpr int ADDER_MAX_PARTS = 16;
struct AdderParts {
  float2 *base[ADDER_MAX_PARTS];
};

// SYS: the parts may live on other GPUs -> system-scope reduction
__device__ __forceinline__ void red_add_v2(float2 *addr, float2 v, bool sys) {
  if (sys)
    asm volatile("red.relaxed.sys.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(v.x), "f"(v.y) : "memory");
  else
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(v.x), "f"(v.y) : "memory");
}
// two neighbouring pixels in one 16-byte reduction (addr 16-byte aligned)
__device__ __forceinline__ void red_add_v4(float2 *addr, float2 lo, float2 hi, bool sys) {
  if (sys)
    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(lo.x), "f"(lo.y),
                 "f"(hi.x), "f"(hi.y) : "memory");
  else
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(lo.x), "f"(lo.y),
                 "f"(hi.x), "f"(hi.y) : "memory");
}

// row = idx / P without the integer-division sequence (exact while rows * P < 2^22)
__device__ __forceinline__ int div_small(int idx, int P, float inv_p) {
  int r = __float2int_rz(__fmul_rn(__int2float_rn(idx) + 0.5f, inv_p));
  (void)P;
  return r;
}

// A CTA owns a subgrid.  Work item = (row y, pixel pair): the pair starts at an even *grid* column
// (x0 + x even), so that both pixels go out as one red.v4.f32 - half the reductions the L2 has to
// serialise; a subgrid at an odd column gets a leading and a trailing single pixel.  The four
// polarisations of an item are independent loads in flight.  VEC = false (odd grid size or
// unaligned part pointers): every pixel on its own.
template <bool SYS, bool VEC>
__global__ void __launch_bounds__(256)
adder_kernel(const int grid_size, const int subgrid_size, const idgb200_metadata *__restrict__ metadata,
             const float2 *__restrict__ subgrids, const __grid_constant__ AdderParts parts, const int rows_per_part,
             const int subgrid_offset, const int shift) {
  const int s = subgrid_offset + blockIdx.x;
  const int N = subgrid_size, G = grid_size;
  const int x0 = metadata[s].x, y0 = metadata[s].y;
  const float2 *src = subgrids + (size_t)s * NR_POL * N * N;
  const size_t part_plane = (size_t)rows_per_part * G;
  const int plane = N * N;
  const int a = x0 & 1;                 // pairs start at x = a - 2, a, a + 2, ...: x0 + x even
  const int P = (N + a + 1) / 2 + a;    // slots per row (the first one is x = -1, 0 when a = 1)
  const float inv_p = __frcp_rn(__int2float_rn(P));
  for (int idx = threadIdx.x; idx < N * P; idx += blockDim.x) {
    const int y = div_small(idx, P, inv_p), jj = idx - y * P;
    const int Y = y0 + y;
    if (Y < 0 || Y >= G) continue;
    const int xl = 2 * jj - a, xh = xl + 1, Xl = x0 + xl;
    const bool okl = xl >= 0 && xl < N && Xl >= 0 && Xl < G;
    const bool okh = xh < N && Xl + 1 >= 0 && Xl + 1 < G;
    if (!okl && !okh) continue;
    const int part = Y / rows_per_part, row = Y - part * rows_per_part;
    int ys = y + shift, xsl = xl + shift, xsh = xh + shift;    // shift = 0 or N/2
    if (ys >= N) ys -= N;
    if (xsl >= N) xsl -= N;
    if (xsh >= N) xsh -= N;
    const float2 *srow = src + ys * N;
    float2 *g = parts.base[part] + (size_t)row * G + Xl;
    float2 vl[NR_POL], vh[NR_POL];
#pragma unroll
    for (int pol = 0; pol < NR_POL; pol++) {
      vl[pol] = okl ? __ldg(srow + pol * plane + xsl) : make_float2(0.f, 0.f);
      vh[pol] = okh ? __ldg(srow + pol * plane + xsh) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int pol = 0; pol < NR_POL; pol++) {
      float2 *d = g + (size_t)pol * part_plane;
      if (VEC && okl && okh) {
        red_add_v4(d, vl[pol], vh[pol], SYS);
      } else {
        if (okl) red_add_v2(d, vl[pol], SYS);
        if (okh) red_add_v2(d + 1, vh[pol], SYS);
      }
    }
  }
}

__device__ __forceinline__ float2 grid_load(const float2 *g, bool sys) {
  float2 v;
  // a peer's slice may be rewritten between launches: no non-coherent (texture path) load for it
  if (sys)
    asm volatile("ld.relaxed.sys.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(g) : "memory");
  else
    v = __ldg(g);
  return v;
}

// Splitter: pairs aligned to the *subgrid* (its stores are the HBM traffic; the grid reads hit L2):
// one 16-byte store per pair where the slot is aligned, the two grid pixels by 8-byte loads.
template <bool SYS>
__global__ void __launch_bounds__(256)
splitter_kernel(const int grid_size, const int subgrid_size, const idgb200_metadata *__restrict__ metadata,
                float2 *__restrict__ subgrids, const __grid_constant__ AdderParts parts, const int rows_per_part,
                const int subgrid_offset, const int shift, const int vec /* pair slots are 16-byte aligned */) {
  const int s = subgrid_offset + blockIdx.x;
  const int N = subgrid_size, G = grid_size;
  const int x0 = metadata[s].x, y0 = metadata[s].y;
  float2 *dst = subgrids + (size_t)s * NR_POL * N * N;
  const size_t part_plane = (size_t)rows_per_part * G;
  const int plane = N * N;
  const int P = (N + 1) / 2;
  const float inv_p = __frcp_rn(__int2float_rn(P));
  for (int idx = threadIdx.x; idx < N * P; idx += blockDim.x) {
    const int y = div_small(idx, P, inv_p), jj = idx - y * P;
    const int xl = 2 * jj, xh = xl + 1, Xl = x0 + xl, Y = y0 + y;
    const bool inrow = Y >= 0 && Y < G;
    const bool okl = inrow && Xl >= 0 && Xl < G;
    const bool okh = inrow && xh < N && Xl + 1 >= 0 && Xl + 1 < G;
    const float2 *g = nullptr;
    if (inrow) {
      const int part = Y / rows_per_part, row = Y - part * rows_per_part;
      g = parts.base[part] + (size_t)row * G + Xl;
    }
    float2 vl[NR_POL], vh[NR_POL];
#pragma unroll
    for (int pol = 0; pol < NR_POL; pol++) {
      vl[pol] = okl ? grid_load(g + (size_t)pol * part_plane, SYS) : make_float2(0.f, 0.f);
      vh[pol] = okh ? grid_load(g + (size_t)pol * part_plane + 1, SYS) : make_float2(0.f, 0.f);
    }
    int ys = y + shift, xsl = xl + shift, xsh = xh + shift;
    if (ys >= N) ys -= N;
    if (xsl >= N) xsl -= N;
    if (xsh >= N) xsh -= N;
    float2 *drow = dst + ys * N;
#pragma unroll
    for (int pol = 0; pol < NR_POL; pol++) {
      if (vec) {
        __stcs(reinterpret_cast<float4 *>(drow + pol * plane + xsl),
               make_float4(vl[pol].x, vl[pol].y, vh[pol].x, vh[pol].y));
      } else {
        drow[pol * plane + xsl] = vl[pol];
        if (xh < N) drow[pol * plane + xsh] = vh[pol];
      }
    }
  }
}

// out[i] = sources[0][i] + sources[1][i] + ... in that order (deterministic, unlike the atomics above).
// The sources may be peer addresses: rank r runs this over its own slice of every rank's local grid -
// a one-shot "pull" reduce-scatter over NVLink with 16-byte loads, no staging and no second copy.
struct ReduceSources {
  const float4 *src[ADDER_MAX_PARTS];
};
__global__ void __launch_bounds__(256)
reduce_parts_kernel(const __grid_constant__ ReduceSources srcs, const int nr_sources, const long long count4,
                    float4 *__restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += stride) {
    float4 acc;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(acc.x), "=f"(acc.y), "=f"(acc.z), "=f"(acc.w) : "l"(srcs.src[0] + i) : "memory");
    for (int k = 1; k < nr_sources; k++) {
      float4 v;
      asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(srcs.src[k] + i) : "memory");
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    out[i] = acc;
  }
}

}  // namespace

cudaError_t launch_reduce_parts(int nr_sources, const float2 *const *sources, long long count, float2 *out,
                                int sm_count, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  if (nr_sources < 1 || nr_sources > ADDER_MAX_PARTS || count < 0 || (count & 1)) return cudaErrorInvalidValue;
  ReduceSources s{};
  for (int i = 0; i < nr_sources; i++) {
    if ((uintptr_t)sources[i] & 15) return cudaErrorInvalidValue;
    s.src[i] = reinterpret_cast<const float4 *>(sources[i]);
  }
  if ((uintptr_t)out & 15) return cudaErrorInvalidValue;
  const long long count4 = count / 2;
  long long ctas = (count4 + 255) / 256;
  const long long cap = (long long)(sm_count > 0 ? sm_count : 148) * 8;   // a few waves, grid-stride beyond
  if (ctas > cap) ctas = cap;
  reduce_parts_kernel<<<dim3((unsigned)ctas), dim3(256), 0, stream>>>(s, nr_sources, count4,
                                                                       reinterpret_cast<float4 *>(out));
  return cudaGetLastError();
}

cudaError_t launch_splitter(int nr_subgrids, int subgrid_offset, int grid_size, int subgrid_size, int flags,
                            const idgb200_metadata *metadata, float2 *subgrids, const float2 *const *parts,
                            int nr_parts, int rows_per_part, cudaStream_t stream) {
  if (nr_subgrids == 0) return cudaSuccess;
  if (nr_parts < 1 || nr_parts > ADDER_MAX_PARTS || rows_per_part < 1 ||
      (long long)nr_parts * rows_per_part < grid_size)
    return cudaErrorInvalidValue;
  AdderParts p{};
  for (int i = 0; i < nr_parts; i++) p.base[i] = const_cast<float2 *>(parts[i]);
  const int shift = (flags & IDGB200_FLAG_FFT_SHIFT) ? subgrid_size / 2 : 0;
  // 16-byte stores need even pair slots (even N and shift) in a 16-byte aligned subgrid array: a complex64 view
  // at an odd element offset takes the 8-byte stores
  const int vec = !(subgrid_size & 1) && !(shift & 1) && !((uintptr_t)subgrids & 15);
  if (nr_parts > 1)
    splitter_kernel<true><<<dim3((unsigned)nr_subgrids), dim3(256), 0, stream>>>(
        grid_size, subgrid_size, metadata, subgrids, p, rows_per_part, subgrid_offset, shift, vec);
  else
    splitter_kernel<false><<<dim3((unsigned)nr_subgrids), dim3(256), 0, stream>>>(
        grid_size, subgrid_size, metadata, subgrids, p, rows_per_part, subgrid_offset, shift, vec);
  return cudaGetLastError();
}

cudaError_t launch_adder(int nr_subgrids, int subgrid_offset, int grid_size, int subgrid_size, int flags,
                         const idgb200_metadata *metadata, const float2 *subgrids, float2 *const *parts,
                         int nr_parts, int rows_per_part, cudaStream_t stream) {
  if (nr_subgrids == 0) return cudaSuccess;
  if (nr_parts < 1 || nr_parts > ADDER_MAX_PARTS || rows_per_part < 1 ||
      (long long)nr_parts * rows_per_part < grid_size)
    return cudaErrorInvalidValue;
  AdderParts p{};
  for (int i = 0; i < nr_parts; i++) p.base[i] = parts[i];
  const int shift = (flags & IDGB200_FLAG_FFT_SHIFT) ? subgrid_size / 2 : 0;
  // 16-byte reductions need even rows (grid_size even) behind 16-byte aligned part pointers
  bool vec = !(grid_size & 1);
  for (int i = 0; i < nr_parts; i++) vec = vec && !((uintptr_t)parts[i] & 15);
  const dim3 grid((unsigned)nr_subgrids), block(256);
#define IDGB200_ADDER_LAUNCH(SYS, VEC)                                                                   \
  adder_kernel<SYS, VEC><<<grid, block, 0, stream>>>(grid_size, subgrid_size, metadata, subgrids, p,   \
                                                     rows_per_part, subgrid_offset, shift)
  if (nr_parts > 1) {
    if (vec) IDGB200_ADDER_LAUNCH(true, true); else IDGB200_ADDER_LAUNCH(true, false);
  } else {
    if (vec) IDGB200_ADDER_LAUNCH(false, true); else IDGB200_ADDER_LAUNCH(false, false);
  }
#undef IDGB200_ADDER_LAUNCH
  return cudaGetLastError();
}

}  // namespace idgb200
