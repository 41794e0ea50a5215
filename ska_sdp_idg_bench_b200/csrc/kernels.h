// Internal launcher interface between the kernels (*.cu) and the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace idgb200 {

// sincos_mode: IDGB200_SINCOS_*; variant: see the launcher definitions.
cudaError_t launch_gridder(const KernelArgs &a, int sincos_mode, int variant, cudaStream_t stream);
cudaError_t launch_degridder(const KernelArgs &a, int sincos_mode, int variant, cudaStream_t stream);

// grid adder (adder.cu): parts[r] = base of the r-th block of rows_per_part grid rows (may be a peer address)
cudaError_t launch_adder(int nr_subgrids, int subgrid_offset, int grid_size, int subgrid_size,
                         const idgb200_metadata *metadata, const float2 *subgrids, float2 *const *parts,
                         int nr_parts, int rows_per_part, cudaStream_t stream);

// what variant 0 means for this shape / sincos mode (other values are returned unchanged)
int resolve_gridder_variant(int subgrid_size, int nr_channels, int sincos_mode, int variant);
int resolve_degridder_variant(int subgrid_size, int nr_channels, int sincos_mode, int variant);

// tcgen05 / TMEM gridder (gridder_tc.cu); FAST sincos only
cudaError_t launch_gridder_tc(const KernelArgs &a, int poly, bool recur, cudaStream_t stream);
cudaError_t launch_degridder_tc(const KernelArgs &a, int poly, bool recur, cudaStream_t stream);
// phasor operand written to TMEM from registers (gridder_tc3.cu)
cudaError_t launch_gridder_tc3(const KernelArgs &a, int mode, cudaStream_t stream);

}  // namespace idgb200
