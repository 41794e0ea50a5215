"""Host-side mirror of the reference's operator interface for the gridder /
degridder path (names and argument order of ``cuda::c_run_gridder`` etc.,
tests/gridder_common.cpp:19-30, app/CUDA/util.cpp:172-444), on top of the C ABI.

* ``c_run_gridder`` / ``c_run_degridder``: host (numpy) arrays, the library does
  the copies and the launches (pipelined), like ``cuda::c_run_*``.
* ``gridder`` / ``degridder``: device-resident torch tensors, one asynchronous
  launch on a stream, like the kernel launch inside ``cuda::p_run_kernel``.
* ``p_run_gridder`` / ``p_run_degridder``: env-var driven performance run.

Errors raise IdgError (the reference exits the process, util.cpp:5-15).
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np

from ._lib import Params, Perf, lib
from .layout import METADATA_DTYPE, NR_CORRELATIONS

SINCOS_FAST, SINCOS_REDUCED, SINCOS_ACCURATE = 0, 1, 2
FLAG_FFT_SHIFT = 1  # IDGB200_FLAG_FFT_SHIFT


class IdgError(RuntimeError):
    def __init__(self, code: int):
        self.code = code
        super().__init__(f"idgb200 error {code}: {lib.idgb200_error_string(code).decode()}")


def _check(rc: int) -> None:
    if rc != 0:
        raise IdgError(rc)


# ------------------------------------------------------------------ metric model
def flops_gridder(nr_channels, nr_timesteps, nr_subgrids, subgrid_size, nr_correlations=4) -> int:
    """app/common/common.cpp:100-120 (also used for the degridder, util.cpp:333-335)."""
    return int(lib.idgb200_flops_gridder(nr_channels, nr_timesteps, nr_subgrids, subgrid_size,
                                         nr_correlations))


def bytes_gridder(nr_channels, nr_timesteps, nr_subgrids, subgrid_size, nr_correlations=4) -> int:
    """app/common/common.cpp:122-159."""
    return int(lib.idgb200_bytes_gridder(nr_channels, nr_timesteps, nr_subgrids, subgrid_size,
                                         nr_correlations))


# ------------------------------------------------------------------------ device
def print_device_info() -> None:
    _check(lib.idgb200_print_device_info())


def device_name() -> str:
    buf = C.create_string_buffer(256)
    _check(lib.idgb200_device_name(buf, 256))
    return buf.value.decode()


def sm_count() -> int:
    n = C.c_int(0)
    _check(lib.idgb200_sm_count(C.byref(n)))
    return n.value


def launch_count() -> int:
    return int(lib.idgb200_launch_count())


def _grid_parts(grid_parts, grid_size, rows_per_part):
    """Host array of part pointers; tensor parts are checked (CUDA, contiguous, complex64, 4 * rows * grid_size
    elements), integer device addresses (peer memory) are the caller's responsibility."""
    parts = grid_parts if isinstance(grid_parts, (list, tuple)) else [grid_parts]
    rpp = int(grid_size if rows_per_part is None else rows_per_part)
    need = 4 * rpp * int(grid_size) * 8
    ptrs = [int(q) if isinstance(q, int) else _dev_ptr(q, f"grid part {i}", need, "complex64").value
            for i, q in enumerate(parts)]
    arr = (C.c_void_p * len(parts))(*ptrs)
    return arr, len(parts), rpp


def adder(nr_subgrids, grid_size, subgrid_size, metadata, subgrids, grid_parts, rows_per_part=None,
          stream=None, flags: int = 0) -> None:
    """Grid adder: accumulate the subgrids (torch CUDA tensors) into the grid.

    ``grid_parts``: one complex64 CUDA tensor [4][grid_size][grid_size] (single GPU), or a list of
    tensors / integer device addresses, part r = complex64 [4][rows_per_part][grid_size] holding grid
    rows [r * rows_per_part, ...) - addresses may be peer memory of other GPUs."""
    arr, n, rpp = _grid_parts(grid_parts, grid_size, rows_per_part)
    p = _params(nr_subgrids, grid_size, subgrid_size, 1.0, 0.0, 1, 1, SINCOS_FAST, 0, flags)
    N = int(subgrid_size)
    _check(lib.idgb200_adder(C.byref(p), _dev_ptr(metadata, "metadata", int(nr_subgrids) * 36, "int32"),
                             _dev_ptr(subgrids, "subgrids", int(nr_subgrids) * N * N * 32, "complex64"),
                             arr, n, rpp, _stream_ptr(stream)))


def splitter(nr_subgrids, grid_size, subgrid_size, metadata, subgrids, grid_parts, rows_per_part=None,
             stream=None, flags: int = 0) -> None:
    """Splitter, the adder's inverse: every pixel of ``subgrids`` is overwritten with
    the grid value under it (0 outside the grid).  ``grid_parts`` as for :func:`adder`."""
    arr, n, rpp = _grid_parts(grid_parts, grid_size, rows_per_part)
    p = _params(nr_subgrids, grid_size, subgrid_size, 1.0, 0.0, 1, 1, SINCOS_FAST, 0, flags)
    N = int(subgrid_size)
    _check(lib.idgb200_splitter(C.byref(p), _dev_ptr(metadata, "metadata", int(nr_subgrids) * 36, "int32"),
                                _dev_ptr(subgrids, "subgrids", int(nr_subgrids) * N * N * 32, "complex64"),
                                arr, n, rpp, _stream_ptr(stream)))


def reduce_parts(sources, out, stream=None) -> None:
    """out = sources[0] + sources[1] + ... (in that order).  ``sources``: complex64 CUDA tensors or
    integer device addresses (peer memory allowed) of out.numel() elements each."""
    nbytes = int(out.numel()) * 8
    arr = (C.c_void_p * len(sources))(*[int(q) if isinstance(q, int) else _dev_ptr(q, f"source {i}", nbytes, "complex64").value
                                        for i, q in enumerate(sources)])
    _check(lib.idgb200_reduce_parts(len(sources), arr, int(out.numel()), _dev_ptr(out, "out", nbytes, "complex64"),
                                    _stream_ptr(stream)))


def subgrid_fft(nr_subgrids, subgrid_size, subgrids, direction: int = 1, stream=None) -> None:
    """Subgrid FFT: in-place 2-D DFT of the nr_subgrids * 4 planes of ``subgrids``
    (complex64 CUDA tensor [S][4][N][N]); direction +1 forward (unscaled), -1 backward (1/N^2)."""
    N = int(subgrid_size)
    ptr = _dev_ptr(subgrids, "subgrids", int(nr_subgrids) * N * N * 32)
    _check(lib.idgb200_subgrid_fft(int(nr_subgrids), N, int(direction), ptr, _stream_ptr(stream)))


def resolve_variant(subgrid_size, nr_channels, sincos=SINCOS_FAST, variant=0, gridder=True) -> int:
    """The kernel variant `variant=0` selects for this shape (idgb200_resolve_variant)."""
    p = _params(1, max(subgrid_size, 1), subgrid_size, 1.0, 0.0, nr_channels, 1, sincos, variant)
    rc = lib.idgb200_resolve_variant(C.byref(p), 1 if gridder else 0)
    if rc < 0:
        _check(rc)
    return rc


def _params(nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda, nr_channels,
            nr_stations, sincos, variant, flags=0) -> Params:
    p = Params()
    p.nr_subgrids = int(nr_subgrids)
    p.grid_size = int(grid_size)
    p.subgrid_size = int(subgrid_size)
    p.image_size = float(image_size)
    p.w_step_in_lambda = float(w_step_in_lambda)
    p.nr_channels = int(nr_channels)
    p.nr_stations = int(nr_stations)
    p.sincos_mode = int(sincos)
    p.variant = int(variant)
    p.flags = int(flags)
    return p


# ------------------------------------------------------------- host-pointer API
def _np(a: Any, dtype, name: str, nbytes: int | None = None) -> np.ndarray:
    if not isinstance(a, np.ndarray):
        raise TypeError(f"{name}: numpy array expected")
    if a.dtype != dtype:
        raise TypeError(f"{name}: dtype {a.dtype}, expected {dtype}")
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError(f"{name}: must be C-contiguous")
    if nbytes is not None and a.nbytes != nbytes:
        raise ValueError(f"{name}: {a.nbytes} bytes, expected {nbytes}")
    return a


def _host_args(nr_subgrids, subgrid_size, nr_channels, nr_stations, uvw, wavenumbers, visibilities,
               spheroidal, aterms, metadata, subgrids, vis_out: bool, sg_out: bool):
    N, C_ = int(subgrid_size), int(nr_channels)
    uvw = _np(uvw, np.float32, "uvw")
    if uvw.size % 3:
        raise ValueError("uvw: size must be a multiple of 3 floats")
    total_timesteps = uvw.size // 3
    _np(wavenumbers, np.float32, "wavenumbers", 4 * C_)
    _np(visibilities, np.complex64, "visibilities", total_timesteps * C_ * NR_CORRELATIONS * 8)
    _np(spheroidal, np.float32, "spheroidal", N * N * 4)
    _np(aterms, np.complex64, "aterms")
    per_slot = int(nr_stations) * N * N * NR_CORRELATIONS
    if aterms.size == 0 or aterms.size % per_slot:
        raise ValueError("aterms: size must be slots * stations * N * N * 4")
    nr_slots = aterms.size // per_slot
    _np(metadata, METADATA_DTYPE, "metadata", 36 * int(nr_subgrids))
    _np(subgrids, np.complex64, "subgrids", int(nr_subgrids) * NR_CORRELATIONS * N * N * 8)
    if vis_out and not visibilities.flags["WRITEABLE"]:
        raise ValueError("visibilities: must be writeable")
    if sg_out and not subgrids.flags["WRITEABLE"]:
        raise ValueError("subgrids: must be writeable")
    ptrs = [a.ctypes.data_as(C.c_void_p) for a in
            (uvw, wavenumbers, visibilities, spheroidal, aterms, metadata, subgrids)]
    return total_timesteps, nr_slots, ptrs


def c_run_gridder(nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda, nr_channels,
                  nr_stations, uvw, wavenumbers, visibilities, spheroidal, aterms, metadata,
                  subgrids, *, sincos: int = SINCOS_FAST, variant: int = 0, flags: int = 0) -> None:
    """cuda::c_run_gridder (tests/gridder_common.cpp:21-30): host arrays in,
    ``subgrids`` [S][4][N][N] complex64 overwritten."""
    tt, slots, ptrs = _host_args(nr_subgrids, subgrid_size, nr_channels, nr_stations, uvw,
                                 wavenumbers, visibilities, spheroidal, aterms, metadata, subgrids,
                                 vis_out=False, sg_out=True)
    p = _params(nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda, nr_channels,
                nr_stations, sincos, variant, flags)
    _check(lib.idgb200_c_run_gridder_ex(C.byref(p), tt, slots, *ptrs))


def c_run_degridder(nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda, nr_channels,
                    nr_stations, uvw, wavenumbers, visibilities, spheroidal, aterms, metadata,
                    subgrids, *, sincos: int = SINCOS_FAST, variant: int = 0, flags: int = 0) -> None:
    """cuda::c_run_degridder (tests/degridder_common.cpp:21-30): host arrays in,
    ``visibilities`` [T][C][4] complex64 overwritten."""
    tt, slots, ptrs = _host_args(nr_subgrids, subgrid_size, nr_channels, nr_stations, uvw,
                                 wavenumbers, visibilities, spheroidal, aterms, metadata, subgrids,
                                 vis_out=True, sg_out=False)
    p = _params(nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda, nr_channels,
                nr_stations, sincos, variant, flags)
    _check(lib.idgb200_c_run_degridder_ex(C.byref(p), tt, slots, *ptrs))


# ----------------------------------------------------------- device-pointer API
def _dev_ptr(t, name: str, min_bytes: int, dtype: str | None = None) -> C.c_void_p:
    import torch

    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name}: CUDA torch tensor expected")
    if dtype is not None and t.dtype != getattr(torch, dtype):
        raise TypeError(f"{name}: dtype {t.dtype}, expected torch.{dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    if t.numel() * t.element_size() < min_bytes:
        raise ValueError(f"{name}: {t.numel() * t.element_size()} bytes, need {min_bytes}")
    return C.c_void_p(t.data_ptr())


def _stream_ptr(stream) -> C.c_void_p:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _device_call(fn, nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda,
                 nr_channels, nr_stations, total_timesteps, uvw, wavenumbers, visibilities,
                 spheroidal, aterms, metadata, subgrids, sincos, variant, stream, flags=0):
    N, C_ = int(subgrid_size), int(nr_channels)
    tt = int(total_timesteps)
    ptrs = [
        _dev_ptr(uvw, "uvw", tt * 12),
        _dev_ptr(wavenumbers, "wavenumbers", C_ * 4),
        _dev_ptr(visibilities, "visibilities", tt * C_ * 32),
        _dev_ptr(spheroidal, "spheroidal", N * N * 4),
        _dev_ptr(aterms, "aterms", int(nr_stations) * N * N * 32),
        _dev_ptr(metadata, "metadata", int(nr_subgrids) * 36),
        _dev_ptr(subgrids, "subgrids", int(nr_subgrids) * N * N * 32),
    ]
    p = _params(nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda, nr_channels,
                nr_stations, sincos, variant, flags)
    _check(fn(C.byref(p), *ptrs, _stream_ptr(stream)))


def gridder(nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda, nr_channels,
            nr_stations, total_timesteps, uvw, wavenumbers, visibilities, spheroidal, aterms,
            metadata, subgrids, *, sincos: int = SINCOS_FAST, variant: int = 0, stream=None,
            flags: int = 0) -> None:
    """One asynchronous gridder launch on device-resident tensors (the kernel launch
    of app/CUDA/util.cpp:163-170).  ``flags=FLAG_FFT_SHIFT`` stores pixel (y, x) at the shifted index."""
    _device_call(lib.idgb200_gridder, nr_subgrids, grid_size, subgrid_size, image_size,
                 w_step_in_lambda, nr_channels, nr_stations, total_timesteps, uvw, wavenumbers,
                 visibilities, spheroidal, aterms, metadata, subgrids, sincos, variant, stream, flags)


def degridder(nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda, nr_channels,
              nr_stations, total_timesteps, uvw, wavenumbers, visibilities, spheroidal, aterms,
              metadata, subgrids, *, sincos: int = SINCOS_FAST, variant: int = 0, stream=None,
              flags: int = 0) -> None:
    _device_call(lib.idgb200_degridder, nr_subgrids, grid_size, subgrid_size, image_size,
                 w_step_in_lambda, nr_channels, nr_stations, total_timesteps, uvw, wavenumbers,
                 visibilities, spheroidal, aterms, metadata, subgrids, sincos, variant, stream, flags)


# ------------------------------------------------------------- performance runs
def _perf(fn) -> dict:
    r = Perf()
    _check(fn(C.byref(r)))
    out = dict(seconds=r.seconds, gflops=r.gflops, gbytes=r.gbytes, mvis=r.mvis,
               nr_subgrids=r.nr_subgrids, iterations=r.iterations,
               mvis_per_s=r.mvis / r.seconds, tflops_per_s=r.gflops / r.seconds * 1e-3, joules=r.joules)
    if r.joules > 0:   # the reference's energy columns (app/common/common.cpp:47-54)
        out.update(watts=r.joules / r.seconds, gflops_per_watt=r.gflops / r.joules, mvis_per_joule=r.mvis / r.joules)
    return out


def p_run_gridder() -> dict:
    """cuda::p_run_gridder (app/CUDA/util.cpp:172-249): shape from the environment."""
    return _perf(lib.idgb200_p_run_gridder)


def p_run_degridder() -> dict:
    return _perf(lib.idgb200_p_run_degridder)


# ------------------------------------------------- synthetic inputs on the device
def init_problem_device(nr_stations=50, nr_timeslots=20, nr_timesteps=128, nr_channels=16,
                        subgrid_size=32, grid_size=1024, image_size=0.01, per_slot_aterms=False,
                        seed=0, device=None, nr_subgrids=None) -> dict:
    """The reference's synthetic inputs (app/common/init.cpp) generated on the GPU.

    ``nr_subgrids`` (optional) keeps only the first that many subgrids of the
    (stations x timeslots) list - used to give each rank its shard.
    Returns a dict of torch CUDA tensors plus the scalar shape."""
    import torch

    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    nr_baselines = nr_stations * (nr_stations - 1) // 2
    S_full = nr_baselines * nr_timeslots
    S = S_full if nr_subgrids is None else min(int(nr_subgrids), S_full)
    T, C_, N = nr_timesteps, nr_channels, subgrid_size
    tt = S * T
    with torch.cuda.device(dev):
        st = _stream_ptr(None)
        uvw = torch.empty((tt, 3), dtype=torch.float32, device=dev)
        wn = torch.empty((C_,), dtype=torch.float32, device=dev)
        vis = torch.empty((tt, C_, NR_CORRELATIONS), dtype=torch.complex64, device=dev)
        sph = torch.empty((N, N), dtype=torch.float32, device=dev)
        at = torch.empty((nr_timeslots, nr_stations, N, N, NR_CORRELATIONS), dtype=torch.complex64,
                         device=dev)
        meta_full = torch.empty((S_full, 9), dtype=torch.int32, device=dev)
        sg = torch.empty((S, NR_CORRELATIONS, N, N), dtype=torch.complex64, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        _check(lib.idgb200_init_uvw(grid_size, S, T, seed, p(uvw), st))
        _check(lib.idgb200_init_wavenumbers(C_, p(wn), st))
        _check(lib.idgb200_init_visibilities(grid_size, image_size, tt, C_, p(uvw), p(vis), st))
        _check(lib.idgb200_init_spheroidal(N, p(sph), st))
        _check(lib.idgb200_init_aterms(nr_timeslots, nr_stations, N, seed, p(at), st))
        _check(lib.idgb200_init_metadata(grid_size, nr_stations, nr_timeslots, T,
                                         int(per_slot_aterms), seed, p(meta_full), st))
        _check(lib.idgb200_init_subgrids(S, N, p(sg), st))
        meta = meta_full[:S].contiguous()
    return dict(nr_subgrids=S, grid_size=grid_size, subgrid_size=N, image_size=image_size,
                w_step_in_lambda=0.0, nr_channels=C_, nr_stations=nr_stations, total_timesteps=tt,
                uvw=uvw, wavenumbers=wn, visibilities=vis, spheroidal=sph, aterms=at,
                metadata=meta, subgrids=sg)
