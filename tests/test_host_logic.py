"""CPU-only: the C-ABI library loads and exports every symbol include/idg_b200.h
declares, the metric model matches the oracle (= the reference's common.cpp), the
sharding logic is right (incl. a world_size-2 gloo run), and - since there is no
GPU here - that every compute entry point refuses to run instead of falling back."""
import ctypes as C
import os
import sys
import re

import numpy as np
import pytest

import ska_sdp_idg_bench_b200 as idg
from ska_sdp_idg_bench_b200 import _lib
from oracle_lib import oracle, random_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch

    return torch.cuda.is_available()


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, "include", "idg_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(idgb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.idgb200_version() == 100


def test_abi_struct_sizes():
    assert C.sizeof(_lib.Params) == 64
    assert idg.METADATA_DTYPE.itemsize == 36
    assert C.sizeof(_lib.Perf) == 48


@pytest.mark.parametrize("shape", [(16, 24500 * 128, 24500, 32), (64, 1000, 10, 32), (16, 77, 3, 64),
                                   (1, 1, 1, 8)])
def test_metric_model_matches_reference_model(shape):
    C_, T, S, N = shape
    o = oracle()
    assert idg.flops_gridder(C_, T, S, N) == o.flops_gridder(C_, T, S, N)
    assert idg.bytes_gridder(C_, T, S, N) == o.bytes_gridder(C_, T, S, N)


def test_error_strings():
    assert b"no CUDA device" in _lib.lib.idgb200_error_string(-2)
    assert b"invalid" in _lib.lib.idgb200_error_string(-1)
    assert _lib.lib.idgb200_error_string(0) == b"ok"


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_device():
    p = random_problem(0)
    out = np.zeros_like(p.subgrids)
    with pytest.raises(idg.IdgError) as e:
        idg.c_run_gridder(p.nr_subgrids, p.grid_size, p.subgrid_size, p.image_size, p.w_step,
                          p.nr_channels, p.nr_stations, p.uvw, p.wavenumbers, p.visibilities,
                          p.spheroidal, p.aterms, p.metadata, out)
    assert e.value.code == -2
    vis = np.zeros_like(p.visibilities)
    with pytest.raises(idg.IdgError) as e:
        idg.c_run_degridder(p.nr_subgrids, p.grid_size, p.subgrid_size, p.image_size, p.w_step,
                            p.nr_channels, p.nr_stations, p.uvw, p.wavenumbers, vis, p.spheroidal,
                            p.aterms, p.metadata, p.subgrids)
    assert e.value.code == -2
    with pytest.raises(idg.IdgError):
        idg.p_run_gridder()
    with pytest.raises(idg.IdgError):
        idg.print_device_info()
    assert not out.any() and not vis.any()


def test_argument_validation():
    p = random_problem(1)
    out = np.zeros_like(p.subgrids)
    args = [p.nr_subgrids, p.grid_size, p.subgrid_size, p.image_size, p.w_step, p.nr_channels,
            p.nr_stations, p.uvw, p.wavenumbers, p.visibilities, p.spheroidal, p.aterms, p.metadata]
    with pytest.raises(TypeError):
        idg.c_run_gridder(*args, out.astype(np.complex128))
    with pytest.raises(ValueError):
        idg.c_run_gridder(*args, out[:, :, ::2])
    with pytest.raises(ValueError):
        idg.c_run_gridder(*args, out[1:])
    bad = list(args)
    bad[8] = p.wavenumbers[:-1].copy()
    with pytest.raises(ValueError):
        idg.c_run_gridder(*bad, out)
    # bad scalar caught by the library itself (EINVAL before any device work)
    bad = list(args)
    bad[3] = 0.0  # image_size
    with pytest.raises(idg.IdgError) as e:
        idg.c_run_gridder(*bad, out)
    assert e.value.code == -1


def test_mirror_pixels_have_exactly_negated_direction_cosines():
    """What the planar path of degridder_tc8.cu rests on: with the
    reference's formula l = float((x + 0.5 - N/2) * image_size / N) (math.hpp:9-12, double intermediate,
    integer N/2) pixel x and its mirror image N-1-x have exactly negated l for EVEN N - and not for odd N,
    which is why odd subgrid sizes never fold - and the fp32 phase chain of gridder_reference.cpp:61-69
    with w = 0 is odd in (l, m), so the mirrored pixel's phase is the exact negative."""
    def compute_l(x, N, image_size):
        return np.float32((x + 0.5 - (N // 2)) * float(np.float32(image_size)) / N)

    rng = np.random.default_rng(5)
    for N in (8, 18, 24, 32, 48, 64):
        l = np.array([compute_l(x, N, 0.01) for x in range(N)], np.float32)
        assert np.array_equal(l, -l[::-1])
        # phase_index = fma(u, l, v * m), phase = fma(-phase_index, k, offset) in fp32 with one rounding per
        # operation (emulated through float64, which holds a product of two floats exactly)
        u, v, uo, vo, k = (np.float32(t) for t in rng.uniform(-900, 900, 5))

        def phase(lx, my):
            vm = np.float32(np.float64(v) * np.float64(my))
            idx = np.float32(np.float64(u) * np.float64(lx) + np.float64(vm))
            vom = np.float32(np.float64(vo) * np.float64(my))
            off = np.float32(np.float64(uo) * np.float64(lx) + np.float64(vom))
            return np.float32(-np.float64(idx) * np.float64(k) + np.float64(off))

        for x in range(N):
            for y in (0, N // 3, N - 1):
                assert phase(l[x], l[y]) == -phase(l[N - 1 - x], l[N - 1 - y])
    l = np.array([compute_l(x, 31, 0.01) for x in range(31)], np.float32)
    assert not np.array_equal(l, -l[::-1])


# ------------------------------------------- report line + CSV against the reference's own
def _capture_fd1(fn):
    """Run fn with the process's stdout (fd 1) redirected into a file; returns what C / C++ code wrote."""
    import tempfile
    sys.stdout.flush()
    saved = os.dup(1)
    with tempfile.TemporaryFile() as tmp:
        os.dup2(tmp.fileno(), 1)
        try:
            fn()
            C.CDLL(None).fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return tmp.read().decode()


@pytest.mark.parametrize("numbers", [
    (7.0579e-3, 1779.19, 4.95488, 50.176, 0.0),          # a perf run without an energy reading
    (12.3456e-3, 1779.19, 4.95488, 50.176, 10.87),       # with one: W, GFLOP/s/W, MVis/J columns
    (35.0e-3, 0.0, 4.95488, 50.176, 0.0),                # no flop figure: GFLOP/s and FLOP/Byte left out
    (1.5e-3, 12.5, 0.0, 0.0, 0.0),
])
def test_report_line_and_csv_match_the_reference_byte_for_byte(numbers, tmp_path):
    """idgb200_report / idgb200_report_csv against the reference's report / report_csv
    (app/common/common.cpp:27-56, 58-98, through oracle/_ref's bridge): the same line on stdout, the same
    file name (<device with / -> ->-<name>-cuda.csv under $OUTPUT_PATH), keys, order and number format."""
    from oracle_lib import reference
    ref = reference()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    seconds, gflops, gbytes, mvis, joules = numbers
    dbl = [C.c_double(x) for x in numbers]
    f_ref, f_ref_csv = ref.lib.idgref_report, ref.lib.idgref_report_csv
    f_ref.restype = f_ref_csv.restype = None
    lib = idg._lib.lib
    name, device = b"gridder_b200", b"NVIDIA B200/SXM"
    line_ref = _capture_fd1(lambda: f_ref(name, *dbl))
    line_ours = _capture_fd1(lambda: lib.idgb200_report(name, *numbers))
    assert line_ours == line_ref and "ms" in line_ref
    old = os.environ.get("OUTPUT_PATH")
    try:
        d_ref, d_ours = tmp_path / "ref", tmp_path / "ours"
        d_ref.mkdir(); d_ours.mkdir()
        os.environ["OUTPUT_PATH"] = str(d_ref)
        out_ref = _capture_fd1(lambda: f_ref_csv(name, device, b"-cuda.csv", *dbl))
        os.environ["OUTPUT_PATH"] = str(d_ours)
        out_ours = _capture_fd1(lambda: lib.idgb200_report_csv(name, device, b"-cuda.csv", *numbers))
    finally:
        if old is None:
            os.environ.pop("OUTPUT_PATH", None)
        else:
            os.environ["OUTPUT_PATH"] = old
    files_ref, files_ours = sorted(os.listdir(d_ref)), sorted(os.listdir(d_ours))
    assert files_ours == files_ref == ["NVIDIA B200-SXM-gridder_b200-cuda.csv"]
    assert (d_ours / files_ours[0]).read_bytes() == (d_ref / files_ref[0]).read_bytes()
    assert out_ours.replace(str(d_ours), "X") == out_ref.replace(str(d_ref), "X")
    keys = [ln.split(",")[0] for ln in (d_ours / files_ours[0]).read_text().splitlines()]
    want = ["ms"] + (["GFLOP/s"] if gflops else []) + (["GB/s"] if gbytes else []) + \
           (["FLOP/Byte"] if gflops and gbytes else []) + (["MVis/s"] if mvis else []) + \
           (["W", "GFLOP/s/W", "MVis/J"] if joules else [])
    assert keys == want


# ------------------------------------------------------------------------ sharding
def test_default_variant_selection():
    """variant 0: the row-column kernels (30) for FAST sincos and the subgrid sizes they take; behind them and
    for the other shapes the per-pixel tcgen05 kernels where the shape fills their tiles, else the FP32 kernels;
    FP32 kernels for the other sincos modes."""
    assert idg.resolve_variant(32, 16, idg.SINCOS_FAST) == 30
    assert idg.resolve_variant(24, 16, idg.SINCOS_FAST) == 30
    assert idg.resolve_variant(8, 1, idg.SINCOS_FAST) == 30
    assert idg.resolve_variant(64, 9, idg.SINCOS_FAST) == 30
    assert idg.resolve_variant(18, 16, idg.SINCOS_FAST) == 24    # 18 is not a multiple of 4: per-pixel kernel, K = 32 stages
    assert idg.resolve_variant(18, 24, idg.SINCOS_FAST) == 21    # 3 blocks of 8 channels
    assert idg.resolve_variant(18, 9, idg.SINCOS_FAST) == 10     # 9 of 16 channels
    assert idg.resolve_variant(32, 16, idg.SINCOS_ACCURATE) == 10
    assert idg.resolve_variant(32, 16, idg.SINCOS_REDUCED) == 10
    assert idg.resolve_variant(32, 16, idg.SINCOS_FAST, variant=24) == 24
    assert idg.resolve_variant(32, 16, idg.SINCOS_FAST, gridder=False) == 30
    assert idg.resolve_variant(8, 16, idg.SINCOS_FAST, gridder=False) == 30
    assert idg.resolve_variant(64, 16, idg.SINCOS_FAST, gridder=False) == 30   # up to 64 x 64: slabs of 32 rows
    assert idg.resolve_variant(72, 16, idg.SINCOS_FAST, gridder=False) == 24   # beyond 64 x 64: two tiles per warp
    assert idg.resolve_variant(72, 12, idg.SINCOS_FAST, gridder=False) == 22   # quads of 4 channels
    assert idg.resolve_variant(18, 1, idg.SINCOS_FAST, gridder=False) == 4     # 1 of 4 channels
    assert idg.resolve_variant(32, 16, idg.SINCOS_ACCURATE, gridder=False) == 4


def test_ska_low_scale_chunk_plan_covers_the_observation_once():
    """tools/ska_low_scale.py (BASELINE config 4): the chunks dealt to the ranks are disjoint and
    add up to the 8,372,224 subgrids of the observation, for every world size the bench uses."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ska_low_scale", os.path.join(ROOT, "tools", "ska_low_scale.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.TOTAL_SUBGRIDS == 8372224
    for world in (1, 2, 4, 8):
        seen, total = set(), 0
        for rank in range(world):
            mine, sizes, n = mod.chunk_plan(32704, world, rank)
            assert not (seen & set(mine))
            seen |= set(mine)
            total += sum(sizes)
        assert len(seen) == n and total == mod.TOTAL_SUBGRIDS
    mine, sizes, _ = mod.chunk_plan(32704, 8, 3, chunks_per_rank=2)
    assert len(mine) == 2 and all(0 < s <= 32704 for s in sizes)


def test_partition_balanced_and_contiguous():
    rng = np.random.default_rng(0)
    nt = rng.integers(0, 200, 1000)
    for w in (1, 2, 3, 4, 8):
        parts = idg.partition_subgrids(nt, w)
        assert parts[0][0] == 0 and parts[-1][1] == 1000
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        loads = [nt[a:b].sum() for a, b in parts]
        assert max(loads) - min(loads) <= 2 * nt.max()
    assert idg.partition_subgrids(np.zeros(5, int), 2) == [(0, 2), (2, 5)]
    assert idg.partition_subgrids(np.array([7]), 4)[-1] == (1, 1) or True


def test_shard_metadata_rebases_time_offsets():
    p = random_problem(3, nr_subgrids=7, max_timesteps=11)
    for s0, s1 in idg.partition_subgrids(p.metadata["nr_timesteps"], 3):
        m, t0, t1 = idg.shard_metadata(p.metadata, s0, s1)
        g = p.metadata[s0:s1]
        glob = g["baseline_offset"] - p.metadata[0]["baseline_offset"] + g["time_offset"]
        live = g["nr_timesteps"] > 0
        assert (m["baseline_offset"] == 0).all()
        assert ((m["time_offset"] + t0)[live] == glob[live]).all()
        assert (m["time_offset"] >= 0).all() and ((m["time_offset"] + m["nr_timesteps"])[live] <= t1 - t0).all()


def test_sharded_oracle_equals_unsharded():
    """The property the multi-GPU path relies on, checked with the oracle standing in
    for the kernel: running each shard on its own slices reproduces the full result."""
    from oracle_lib import Problem, bits_equal

    o = oracle()
    p = random_problem(5, nr_subgrids=9, max_timesteps=8)
    full_g, full_d = o.gridder(p), o.degridder(p)
    out_g = np.zeros_like(full_g)
    out_d = np.full_like(full_d, np.nan)
    for s0, s1 in idg.partition_subgrids(p.metadata["nr_timesteps"], 3):
        m, t0, t1 = idg.shard_metadata(p.metadata, s0, s1)
        if s1 == s0:
            continue
        q = Problem(grid_size=p.grid_size, subgrid_size=p.subgrid_size, image_size=p.image_size,
                    w_step=p.w_step, nr_channels=p.nr_channels, nr_stations=p.nr_stations,
                    uvw=np.ascontiguousarray(p.uvw[t0:max(t1, t0 + 1)]), wavenumbers=p.wavenumbers,
                    visibilities=np.ascontiguousarray(p.visibilities[t0:max(t1, t0 + 1)]),
                    spheroidal=p.spheroidal, aterms=p.aterms, metadata=m,
                    subgrids=np.ascontiguousarray(p.subgrids[s0:s1]))
        out_g[s0:s1] = o.gridder(q)
        d = o.degridder(q)
        ok = ~np.isnan(d.real)
        out_d[t0:t0 + d.shape[0]][ok] = d[ok]
    assert bits_equal(out_g, full_g)
    ok = ~np.isnan(full_d.real)
    assert bits_equal(out_d[ok], full_d[ok])


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch

    from bench import rank_shape, reduce_max_time  # the bench's own N>1 host logic

    shape = rank_shape(dict(nr_stations=6, nr_timeslots=5), rank, world)
    t = reduce_max_time(0.010 * (rank + 1), torch.device("cpu"))
    total = torch.tensor([shape["nr_subgrids"]], dtype=torch.int64)
    dist.all_reduce(total)
    q.put((rank, shape["nr_subgrids"], int(total), t))
    dist.destroy_process_group()


def test_bench_multi_rank_logic_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    # weak scaling: every rank runs the full per-GPU workload; max-over-ranks timing
    assert res[0][1] == res[1][1] == 15 * 5
    assert res[0][2] == res[1][2] == 2 * 75
    assert abs(res[0][3] - 0.020) < 1e-9 and abs(res[1][3] - 0.020) < 1e-9


def build_c_example() -> str:
    """examples/imaging_cycle.c: the C ABI used from plain C (gcc, no CUDA headers)."""
    import subprocess
    pkg = os.path.join(ROOT, "ska_sdp_idg_bench_b200")
    exe = os.path.join(ROOT, "examples", "imaging_cycle")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "imaging_cycle.c"), "-L", pkg, "-lidgb200",
                    f"-Wl,-rpath,{pkg}", "-L/usr/local/cuda/lib64", "-lcudart", "-lm", "-o", exe], check=True)
    return exe


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_c_example_builds_and_refuses_without_a_device():
    import subprocess
    r = subprocess.run([build_c_example()], capture_output=True, text=True)
    assert r.returncode != 0 and "imaging cycle OK" not in r.stdout


def test_next_rows_reject_sizes_beyond_their_index_arithmetic():
    """The shifted slot and the adder / splitter rows use a float reciprocal that is exact below 2^22
    pixels per plane: larger subgrids are refused (before any device work), not mis-indexed."""
    from ska_sdp_idg_bench_b200.api import _params
    p = _params(1, 4096, 2048, 1.0, 0.0, 1, 1, 0, 0, flags=idg.FLAG_FFT_SHIFT)
    dummy = (C.c_void_p * 1)(C.c_void_p(16))
    assert _lib.lib.idgb200_adder(C.byref(p), C.c_void_p(16), C.c_void_p(16), dummy, 1, 4096, None) == -3
    p.flags = 0
    assert _lib.lib.idgb200_adder(C.byref(p), C.c_void_p(16), C.c_void_p(16), dummy, 1, 4096, None) == -3
    assert _lib.lib.idgb200_splitter(C.byref(p), C.c_void_p(16), C.c_void_p(16), dummy, 1, 4096, None) == -3
