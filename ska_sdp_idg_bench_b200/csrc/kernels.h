// Internal launcher interface between the kernels (*.cu) and the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace idgb200 {

// sincos_mode: IDGB200_SINCOS_*; variant: see the launcher definitions.
// *kernels (may be null) receives the number of kernels the call put on the stream
cudaError_t launch_gridder(const KernelArgs &a, int sincos_mode, int variant, cudaStream_t stream, int *kernels = nullptr);
cudaError_t launch_degridder(const KernelArgs &a, int sincos_mode, int variant, cudaStream_t stream, int *kernels = nullptr);

// grid adder (adder.cu): parts[r] = base of the r-th block of rows_per_part grid rows (may be a peer address)
// flags: IDGB200_FLAG_FFT_SHIFT
cudaError_t launch_adder(int nr_subgrids, int subgrid_offset, int grid_size, int subgrid_size, int flags,
                         const idgb200_metadata *metadata, const float2 *subgrids, float2 *const *parts,
                         int nr_parts, int rows_per_part, cudaStream_t stream);
// splitter (adder.cu): the adder's inverse, subgrids read out of the grid parts
cudaError_t launch_splitter(int nr_subgrids, int subgrid_offset, int grid_size, int subgrid_size, int flags,
                            const idgb200_metadata *metadata, float2 *subgrids, const float2 *const *parts,
                            int nr_parts, int rows_per_part, cudaStream_t stream);

// out[i] = sum over the sources in order (adder.cu); count complex64 elements (even), pointers 16-byte
// aligned, sources may be peer addresses
cudaError_t launch_reduce_parts(int nr_sources, const float2 *const *sources, long long count, float2 *out,
                                int sm_count, cudaStream_t stream);

// subgrid FFT (subgrid_fft.cu): in-place 2-D DFT of nr_planes N x N planes; direction >= 0 forward
// (exp(-i)), < 0 backward (exp(+i), scaled by 1/N^2)
cudaError_t launch_subgrid_fft(long long nr_planes, int subgrid_size, int direction, float2 *planes,
                               cudaStream_t stream);

// what variant 0 means for this shape / sincos mode (other values are returned unchanged)
int resolve_gridder_variant(int subgrid_size, int nr_channels, int sincos_mode, int variant);
int resolve_degridder_variant(int subgrid_size, int nr_channels, int sincos_mode, int variant);

// Every per-pixel kernel below runs in LIST MODE when KernelArgs::list is set: a fixed number of CTAs loop over the
// listed subgrids (what a row-column kernel left), so that the launch is cheap when the list is empty.
// per-pixel tcgen05 / TMEM gridder (gridder_tc.cu); FAST sincos only
cudaError_t launch_gridder_tc(const KernelArgs &a, int poly, bool recur, cudaStream_t stream);
// FP32 gridder (gridder.cu)
cudaError_t launch_gridder_fp32(const KernelArgs &a, int sincos_mode, cudaStream_t stream);
// row-column gridder (gridder_sep.cu): one GEMM per subgrid with the visibilities as K.  Appends to the work lists
// d_todo = { n, subgrid[n] } (its non-separable phase term is too large: per-pixel kernel) and d_cancel = { n, subgrid[n] }
// (every tile's pixel sums cancel below the fp16 operand's error model: FP32 kernel); d_cancel_tiles[nr_subgrids] counts
// tiles.  Counts and counters zeroed by the caller on the same stream.
bool gridder_sep_supports(int subgrid_size, int nr_channels);
cudaError_t launch_gridder_sep(const KernelArgs &a, int *d_todo, int *d_cancel, int *d_cancel_tiles, cudaStream_t stream);
// row-column degridder (degridder_sep.cu); d_todo as for launch_gridder_sep
bool degridder_sep_supports(int subgrid_size, int nr_channels);
// mode: 0 = pipelined persistent kernel where its buffers fit, 1 = one subgrid per CTA, 2 = pipelined or an error
cudaError_t launch_degridder_sep(const KernelArgs &a, int *d_todo, cudaStream_t stream, int mode = 0);
cudaError_t launch_degridder_tc(const KernelArgs &a, int poly, bool recur, cudaStream_t stream);
// two M-tiles per warp, groups of 8 channels (degridder_tc8.cu); nr_channels % 8 == 0
cudaError_t launch_degridder_tc8(const KernelArgs &a, bool recur, bool fold, cudaStream_t stream);

// per-launch device scratch (scratch.cu): acquire .. release brackets the enqueue of the kernels that use it
struct ScratchLease {
  int *ptr = nullptr;
  void *pool = nullptr;
  int slot = 0;
};
cudaError_t scratch_acquire(size_t ints, cudaStream_t stream, ScratchLease *lease);
cudaError_t scratch_release(const ScratchLease &lease, cudaStream_t stream);

}  // namespace idgb200
