"""ctypes/numpy access to the CHECKERS (oracle/liboracle.so and, when present,
oracle/_ref/libidgref.so = the reference's own CPU code built by oracle/Makefile).

Test infrastructure only: nothing under ska_sdp_idg_bench_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

# app/common/types.hpp:19-26
METADATA_DTYPE = np.dtype(
    [
        ("baseline_offset", "<i4"),
        ("time_offset", "<i4"),
        ("nr_timesteps", "<i4"),
        ("aterm_index", "<i4"),
        ("station1", "<u4"),
        ("station2", "<u4"),
        ("x", "<i4"),
        ("y", "<i4"),
        ("z", "<i4"),
    ]
)
assert METADATA_DTYPE.itemsize == 36
BASELINE_DTYPE = np.dtype([("station1", "<u4"), ("station2", "<u4")])

IMAGE_SIZE = np.float32(0.01)  # app/common/parameters.hpp:4
W_STEP = np.float32(0.0)  # app/common/parameters.hpp:5


def build_oracle() -> None:
    subprocess.run(["make", "-C", ORACLE_DIR, "--no-print-directory"], check=True,
                   stdout=subprocess.DEVNULL)


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class _Lib:
    """Flat-pointer view over either checker; `prefix` is idgo_ or idgref_."""

    def __init__(self, path: str, prefix: str):
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path)
        self.is_reference = prefix == "idgref_"
        for name in ("flops_gridder", "bytes_gridder"):
            f = getattr(self.lib, prefix + name)
            f.restype = C.c_uint64
            f.argtypes = [C.c_uint64] * 5
        getattr(self.lib, prefix + "max_threads").restype = C.c_int

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # -- threads ---------------------------------------------------------------
    def max_threads(self) -> int:
        return int(self._f("max_threads")())

    def set_threads(self, n: int) -> None:
        self._f("set_threads")(C.c_int(n))

    # -- kernels ---------------------------------------------------------------
    def _run(self, which, p: "Problem", vis, subgrids):
        args = [
            C.c_int(p.nr_subgrids), C.c_int(p.grid_size), C.c_int(p.subgrid_size),
            C.c_float(p.image_size), C.c_float(p.w_step), C.c_int(p.nr_channels),
            C.c_int(p.nr_stations),
        ]
        if self.is_reference:
            args += [C.c_long(p.total_timesteps), C.c_int(p.aterms.shape[0])]
        args += [_ptr(p.uvw), _ptr(p.wavenumbers), _ptr(vis), _ptr(p.spheroidal),
                 _ptr(p.aterms), _ptr(p.metadata), _ptr(subgrids)]
        self._f(which)(*args)

    def gridder(self, p: "Problem") -> np.ndarray:
        out = np.full((p.nr_subgrids, 4, p.subgrid_size, p.subgrid_size), np.nan,
                      np.complex64)
        self._run("gridder", p, p.visibilities, out)
        return out

    def adder(self, p: "Problem", subgrids=None, flags=0) -> np.ndarray:
        """Grid adder (oracle only; SURVEY 8f-1, parity unpinned): complex64 [4][G][G]."""
        assert not self.is_reference, "the reference has no adder"
        sg = np.ascontiguousarray(p.subgrids if subgrids is None else subgrids)
        grid = np.zeros((4, p.grid_size, p.grid_size), np.complex64)
        self._f("adder")(C.c_int(p.nr_subgrids), C.c_int(p.grid_size), C.c_int(p.subgrid_size),
                         C.c_int(flags), _ptr(p.metadata), _ptr(sg), _ptr(grid))
        return grid

    def splitter(self, p: "Problem", grid: np.ndarray, flags=0) -> np.ndarray:
        """Splitter (oracle only; SURVEY 8f-3, parity unpinned): complex64 [S][4][N][N]."""
        assert not self.is_reference, "the reference has no splitter"
        grid = np.ascontiguousarray(grid, np.complex64)
        sg = np.full((p.nr_subgrids, 4, p.subgrid_size, p.subgrid_size), np.nan, np.complex64)
        self._f("splitter")(C.c_int(p.nr_subgrids), C.c_int(p.grid_size), C.c_int(p.subgrid_size),
                            C.c_int(flags), _ptr(p.metadata), _ptr(sg), _ptr(grid))
        return sg

    def subgrid_fft(self, subgrids: np.ndarray, direction=1) -> np.ndarray:
        """Subgrid FFT (oracle only; SURVEY 8f-2, parity unpinned): a transformed copy."""
        assert not self.is_reference, "the reference has no FFT"
        out = np.array(subgrids, np.complex64, order="C", copy=True)
        N = out.shape[-1]
        assert out.shape[-2] == N
        self._f("subgrid_fft")(C.c_long(out.size // (N * N)), C.c_int(N), C.c_int(direction), _ptr(out))
        return out

    def degridder(self, p: "Problem") -> np.ndarray:
        out = np.full((p.total_timesteps, p.nr_channels, 4), np.nan, np.complex64)
        self._run("degridder", p, out, p.subgrids)
        return out

    def gridder_f64(self, p: "Problem") -> np.ndarray:
        assert not self.is_reference
        out = np.zeros((p.nr_subgrids, 4, p.subgrid_size, p.subgrid_size), np.complex128)
        self._run("gridder_f64", p, p.visibilities, out)
        return out

    def degridder_f64(self, p: "Problem") -> np.ndarray:
        assert not self.is_reference
        out = np.zeros((p.total_timesteps, p.nr_channels, 4), np.complex128)
        self._run("degridder_f64", p, out, p.subgrids)
        return out

    # -- metric model ------------------------------------------------------------
    def flops_gridder(self, C_, T, S, N, P=4) -> int:
        return int(self._f("flops_gridder")(C_, T, S, N, P))

    def bytes_gridder(self, C_, T, S, N, P=4) -> int:
        return int(self._f("bytes_gridder")(C_, T, S, N, P))

    def check_error(self, cand: np.ndarray, ref: np.ndarray) -> float:
        assert not self.is_reference
        f = self.lib.idgo_check_error
        f.restype = C.c_double
        cand = np.ascontiguousarray(cand, np.complex64)
        ref = np.ascontiguousarray(ref, np.complex64)
        return float(f(C.c_int(cand.size), _ptr(cand), _ptr(ref)))

    # -- synthetic inputs --------------------------------------------------------
    def make_problem(self, nr_stations=2, nr_timeslots=2, nr_timesteps=128,
                     nr_channels=16, subgrid_size=32, grid_size=1024,
                     image_size=IMAGE_SIZE, w_step=W_STEP, seed=0) -> "Problem":
        """The reference's correctness set-up, in its call order
        (tests/gridder_common.cpp:54-101, tests/degridder_common.cpp:88-100)."""
        nr_baselines = nr_stations * (nr_stations - 1) // 2
        S = nr_baselines * nr_timeslots
        T = nr_timesteps
        N = subgrid_size
        uvw = np.zeros((S, T, 3), np.float32)
        freq = np.zeros(nr_channels, np.float32)
        wn = np.zeros(nr_channels, np.float32)
        vis = np.zeros((S * T, nr_channels, 4), np.complex64)
        baselines = np.zeros(nr_baselines, BASELINE_DTYPE)
        sph = np.zeros((N, N), np.float32)
        aterms = np.zeros((nr_timeslots, nr_stations, N, N, 4), np.complex64)
        meta = np.zeros(S, METADATA_DTYPE)
        sg = np.zeros((S, 4, N, N), np.complex64)

        f = self._f
        f("srand")(C.c_uint(seed))
        # NB the reference passes uvw as Array2D(nr_subgrids, nr_timesteps)
        f("init_uvw")(C.c_uint(grid_size), C.c_int(S), C.c_int(T), _ptr(uvw))
        f("init_frequencies")(C.c_int(nr_channels), _ptr(freq))
        f("init_wavenumbers")(C.c_int(nr_channels), _ptr(freq), _ptr(wn))
        f("init_visibilities")(C.c_uint(grid_size), C.c_float(image_size), C.c_int(S),
                               C.c_int(T), C.c_int(nr_channels), _ptr(freq), _ptr(uvw),
                               _ptr(vis))
        f("init_baselines")(C.c_uint(nr_stations), C.c_int(nr_baselines), _ptr(baselines))
        f("init_spheroidal")(C.c_int(N), _ptr(sph))
        f("init_aterms")(C.c_int(nr_timeslots), C.c_int(nr_stations), C.c_int(N),
                         _ptr(sph), _ptr(aterms))
        f("init_metadata")(C.c_uint(grid_size), C.c_uint(nr_timeslots), C.c_uint(T),
                           C.c_int(nr_baselines), _ptr(baselines), _ptr(meta))
        f("init_subgrids")(C.c_int(S), C.c_int(N), _ptr(sg))
        return Problem(grid_size=grid_size, subgrid_size=N, image_size=float(image_size),
                       w_step=float(w_step), nr_channels=nr_channels,
                       nr_stations=nr_stations, uvw=uvw.reshape(S * T, 3),
                       wavenumbers=wn, visibilities=vis, spheroidal=sph, aterms=aterms,
                       metadata=meta, subgrids=sg, frequencies=freq)


@dataclass
class Problem:
    grid_size: int
    subgrid_size: int
    image_size: float
    w_step: float
    nr_channels: int
    nr_stations: int
    uvw: np.ndarray  # [total_timesteps, 3] f32
    wavenumbers: np.ndarray  # [C] f32
    visibilities: np.ndarray  # [total_timesteps, C, 4] c64
    spheroidal: np.ndarray  # [N, N] f32
    aterms: np.ndarray  # [slots, stations, N, N, 4] c64
    metadata: np.ndarray  # [S] METADATA_DTYPE
    subgrids: np.ndarray  # [S, 4, N, N] c64
    frequencies: np.ndarray | None = None

    @property
    def nr_subgrids(self) -> int:
        return int(self.metadata.shape[0])

    @property
    def total_timesteps(self) -> int:
        return int(self.uvw.shape[0])


def random_problem(seed: int, nr_subgrids=3, subgrid_size=16, nr_channels=5,
                   nr_stations=4, nr_slots=2, max_timesteps=9, grid_size=512,
                   image_size=0.02, w_step=1.3, with_w=True) -> Problem:
    """A ragged, adversarial problem the reference's own init never produces:
    varying nr_timesteps (including 0), non-zero w and w_step, aterm_index != 0,
    unequally spaced wavenumbers, non-zero baseline_offset, z != 0."""
    rng = np.random.default_rng(seed)
    N = subgrid_size
    nts = rng.integers(0, max_timesteps + 1, nr_subgrids)
    nts[rng.integers(0, nr_subgrids)] = max_timesteps
    offs = np.concatenate([[0], np.cumsum(nts)[:-1]])
    T = int(nts.sum())
    base = int(rng.integers(0, 1000))
    meta = np.zeros(nr_subgrids, METADATA_DTYPE)
    meta["baseline_offset"] = base
    meta["time_offset"] = offs
    meta["nr_timesteps"] = nts
    meta["aterm_index"] = rng.integers(0, nr_slots, nr_subgrids)
    st1 = rng.integers(0, nr_stations, nr_subgrids)
    st2 = (st1 + rng.integers(1, nr_stations, nr_subgrids)) % nr_stations
    meta["station1"], meta["station2"] = st1, st2
    meta["x"] = rng.integers(0, grid_size, nr_subgrids)
    meta["y"] = rng.integers(0, grid_size, nr_subgrids)
    meta["z"] = rng.integers(-2, 3, nr_subgrids) if with_w else 0
    uvw = (rng.standard_normal((max(T, 1), 3)) * grid_size / 2).astype(np.float32)
    if not with_w:
        uvw[:, 2] = 0
    wn = np.sort(rng.uniform(2.5, 3.5, nr_channels)).astype(np.float32)

    def cplx(*shape):
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)

    return Problem(grid_size=grid_size, subgrid_size=N, image_size=image_size,
                   w_step=w_step if with_w else 0.0, nr_channels=nr_channels,
                   nr_stations=nr_stations, uvw=uvw[:max(T, 1)], wavenumbers=wn,
                   visibilities=cplx(max(T, 1), nr_channels, 4),
                   spheroidal=rng.uniform(0, 1, (N, N)).astype(np.float32),
                   aterms=cplx(nr_slots, nr_stations, N, N, 4), metadata=meta,
                   subgrids=cplx(nr_subgrids, 4, N, N))


_ORACLE = None
_REF = None


def oracle() -> _Lib:
    global _ORACLE
    if _ORACLE is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        _ORACLE = _Lib(path, "idgo_")
    return _ORACLE


def reference() -> _Lib | None:
    """The reference's own CPU code (None when oracle/_ref was never built)."""
    global _REF
    if _REF is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libidgref.so")
        if not os.path.exists(path):
            return None
        _REF = _Lib(path, "idgref_")
    return _REF


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()
