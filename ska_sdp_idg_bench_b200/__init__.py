"""B200-native IDG gridder / degridder behind the ska-sdp-idg-bench entry points.

The product is ``libidgb200.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/idg_b200.h``).  This package is the thin host-side mirror of the
reference's operator interface (``cuda::c_run_gridder`` & co., app/lib-cuda.hpp,
tests/gridder_common.cpp:19-30) for Python callers: numpy arrays for the
host-pointer API, torch CUDA tensors for the device-pointer API.

There is no CPU fallback anywhere in this package: importing it without the
built library raises, and calling it without a CUDA device raises IdgError.
"""
from .api import (FLAG_FFT_SHIFT, IdgError, SINCOS_ACCURATE, SINCOS_FAST, SINCOS_REDUCED, adder, bytes_gridder,
                  c_run_degridder, c_run_gridder, degridder, device_name, flops_gridder, gridder,
                  init_problem_device, launch_count, p_run_degridder, p_run_gridder,
                  print_device_info, reduce_parts, resolve_variant, sm_count, splitter, subgrid_fft)
from .layout import BASELINE_DTYPE, METADATA_DTYPE, NR_CORRELATIONS, IMAGE_SIZE, W_STEP
from .grid_adder_rs import GridAdderRS, adder_rs_mode
from .shard import partition_subgrids, shard_metadata

__all__ = [
    "IdgError", "SINCOS_FAST", "SINCOS_REDUCED", "SINCOS_ACCURATE", "c_run_gridder",
    "c_run_degridder", "gridder", "degridder", "p_run_gridder", "p_run_degridder",
    "flops_gridder", "bytes_gridder", "print_device_info", "device_name", "sm_count",
    "launch_count", "resolve_variant", "adder", "reduce_parts", "splitter", "subgrid_fft", "FLAG_FFT_SHIFT", "init_problem_device", "METADATA_DTYPE", "BASELINE_DTYPE",
    "NR_CORRELATIONS", "IMAGE_SIZE", "W_STEP", "partition_subgrids", "shard_metadata", "GridAdderRS", "adder_rs_mode",
]
