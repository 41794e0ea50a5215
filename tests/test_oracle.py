"""CPU-only: pins oracle/idg_oracle.c (the restatement) against
  (1) golden outputs of the reference's own CPU code (tests/golden/*.npz, made by
      tests/golden/make_golden.py from oracle/_ref/libidgref.so),
  (2) the known-answer values SURVEY.md §8c lists for the reference's correctness shape,
  (3) the reference build itself, bit for bit, when oracle/_ref/ is present.
Bar: bit-exact (the restatement spells the reference binary's FMA contractions)."""
import os

import numpy as np
import pytest

from oracle_lib import (METADATA_DTYPE, Problem, bits_equal, oracle, random_problem,
                        reference, ORACLE_DIR)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def config1():
    o = oracle()
    o.set_threads(o.max_threads())
    return o.make_problem()


def test_abi_sizes():
    assert METADATA_DTYPE.itemsize == 36  # types.hpp:19-26


def test_init_known_answers(config1):
    p = config1
    # SURVEY.md §8c, produced by the reference's own initialize_* after srand(0)
    np.testing.assert_allclose(p.uvw[0], [799.010620, 378.322113, 0.0], rtol=0, atol=1e-5)
    assert abs(float(p.wavenumbers[0]) - 3.143767595) < 1e-7
    assert abs(float(p.wavenumbers[15]) - 3.363831282) < 1e-7
    m0, m1 = p.metadata[0], p.metadata[1]
    assert (m0["baseline_offset"], m0["time_offset"], m0["nr_timesteps"], m0["aterm_index"]) == (0, 0, 128, 0)
    assert (m0["station1"], m0["station2"], m0["x"], m0["y"], m0["z"]) == (0, 1, 28, 567, 0)
    assert (m1["time_offset"], m1["x"], m1["y"]) == (128, 325, 600)
    v = p.visibilities.sum(dtype=np.complex128)
    assert abs(v.real - 2.938851e02) < 1e-3 and abs(v.imag - 3.003485e02) < 1e-3
    assert abs(np.abs(p.visibilities).sum(dtype=np.float64) - 1.679360e04) < 0.05
    a = p.aterms.sum(dtype=np.complex128)
    assert abs(a.real - 3.278570e03) < 1e-2 and abs(a.imag) < 1e-3
    assert abs(p.aterms.ravel()[0] - (1.264659 - 0.1j)) < 1e-6
    s = p.subgrids.sum(dtype=np.complex128)
    assert abs(s.real - 41.0) < 1e-3 and abs(s.imag - 1228.8) < 1e-2


def test_gridder_known_answers(config1):
    g = oracle().gridder(config1)
    s = g.sum(dtype=np.complex128)
    assert abs(s.real - -1.431275e04) < 0.05 and abs(s.imag - 5.135324e02) < 0.05
    assert abs(np.abs(g).sum(dtype=np.float64) - 4.025640e05) < 1.0
    assert abs(np.abs(g).max() - 2.780631e03) < 1e-2
    assert abs(g.ravel()[0] - (-5.452361e02 + 1.598559e03j)) < 1e-2
    assert abs(g.ravel()[-1] - (-1.660462e02 + 2.367181e03j)) < 1e-2


def test_degridder_known_answers(config1):
    d = oracle().degridder(config1)
    s = d.sum(dtype=np.complex128)
    assert abs(s.real - 2.761614e02) < 1e-3 and abs(s.imag - 2.391636e02) < 1e-3
    assert abs(np.abs(d).sum(dtype=np.float64) - 1.166473e04) < 0.05
    assert abs(np.abs(d).max() - 2.928413) < 1e-5
    assert abs(d.ravel()[0] - (3.121725e-02 + 4.282365e-01j)) < 1e-6
    assert abs(d.ravel()[-1] - (7.117927e-03 + 4.824635e-01j)) < 1e-6


def test_config1_golden_bit_exact(config1):
    z = np.load(os.path.join(GOLD, "config1.npz"))
    p = config1
    assert bits_equal(p.metadata.view(np.int32).reshape(-1, 9), z["metadata"])
    assert bits_equal(p.wavenumbers, z["wavenumbers"])
    assert bits_equal(p.uvw[:4], z["uvw_head"])
    assert bits_equal(p.aterms.reshape(-1)[:8], z["aterms_head"])
    for k in ("uvw", "visibilities", "spheroidal", "aterms", "subgrids"):
        got = np.asarray(getattr(p, k)).view(np.float32).astype(np.float64).sum()
        assert got == float(z["sum_" + k]), k
    assert bits_equal(oracle().gridder(p), z["gridder"])
    assert bits_equal(oracle().degridder(p), z["degridder"])


def load_golden_problem(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    G, N, C, ST = (int(v) for v in z["scalars"])
    image_size, w_step = (float(v) for v in z["fscalars"])
    meta = np.ascontiguousarray(z["metadata"]).view(METADATA_DTYPE).reshape(-1)
    p = Problem(grid_size=G, subgrid_size=N, image_size=image_size, w_step=w_step,
                nr_channels=C, nr_stations=ST, uvw=z["uvw"], wavenumbers=z["wavenumbers"],
                visibilities=z["visibilities"], spheroidal=z["spheroidal"], aterms=z["aterms"],
                metadata=meta, subgrids=z["subgrids"])
    return p, z["gridder"], z["degridder"]


@pytest.mark.parametrize("name", ["ragged_a", "ragged_b", "ragged_c"])
def test_ragged_golden_bit_exact(name):
    p, g, d = load_golden_problem(name)
    assert bits_equal(oracle().gridder(p), g)
    # rows of subgrids with nr_timesteps == 0 are never written by the degridder
    got = oracle().degridder(p)
    mask = np.zeros(p.total_timesteps, bool)
    for m in p.metadata:
        t0 = int(m["baseline_offset"] - p.metadata[0]["baseline_offset"] + m["time_offset"])
        mask[t0:t0 + int(m["nr_timesteps"])] = True
    assert bits_equal(got[mask], d[mask])


def test_metric_model():
    o = oracle()
    # common.cpp:100-120, figures quoted in SURVEY.md §8d / BASELINE.md §3
    S, T, C, N = 24500, 128, 16, 32
    fl = o.flops_gridder(C, S * T, S, N)
    assert fl == S * T * N * N * (5 + 5 + 2 * C + 32 * C) + S * N * N * 6
    assert abs(fl / (S * T * C) - 35459) < 1
    assert abs(fl * 1e-9 - 1779.19) < 0.01
    by = o.bytes_gridder(C, S * T, S, N)
    assert by == S * T * 12 + S * T * C * 32 + S * N * N * (64 + 64 + 4)
    assert abs(fl / by - 359.08) < 0.01


def test_f64_truth_error_budget(config1):
    """float32 CPU vs the same formula in float64: the error floor SURVEY §8c quotes
    (gridder max|d|/max|v| ~1e-4, rel-RMS ~7e-5; degridder ~2.7e-4 / 1.5e-4)."""
    o = oracle()
    g32, g64 = o.gridder(config1), o.gridder_f64(config1)
    d32, d64 = o.degridder(config1), o.degridder_f64(config1)
    for a32, a64, mx, rms in ((g32, g64, 2e-4, 1.5e-4), (d32, d64, 6e-4, 3e-4)):
        diff = np.abs(a32.astype(np.complex128) - a64)
        assert diff.max() / np.abs(a64).max() < mx
        assert np.sqrt((diff ** 2).sum() / (np.abs(a64) ** 2).sum()) < rms


def test_check_error_metric():
    o = oracle()
    rng = np.random.default_rng(0)
    b = (rng.standard_normal(1000) + 1j * rng.standard_normal(1000)).astype(np.complex64)
    assert o.check_error(b, b) == 0.0
    a = b.copy()
    a[3] += np.complex64(0.5)
    r_max = max(1.0, np.abs(a.real).max())
    assert abs(o.check_error(a, b) - np.sqrt(0.25 / r_max / 1000)) < 1e-6


# ---------------------------------------------------------------- vs reference build
needs_ref = pytest.mark.skipif(reference() is None, reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("shape", [
    dict(),  # the reference's correctness shape
    dict(nr_stations=3, nr_timeslots=2, nr_timesteps=7, nr_channels=3, subgrid_size=24),
    dict(nr_stations=2, nr_timeslots=1, nr_timesteps=16, nr_channels=9, subgrid_size=8, grid_size=256),
])
def test_restatement_equals_reference_build(shape):
    o, r = oracle(), reference()
    po, pr = o.make_problem(**shape), r.make_problem(**shape)
    for k in ("uvw", "wavenumbers", "visibilities", "spheroidal", "aterms", "metadata", "subgrids"):
        assert bits_equal(getattr(po, k), getattr(pr, k)), k
    assert bits_equal(o.gridder(po), r.gridder(pr))
    assert bits_equal(o.degridder(po), r.degridder(pr))
    T, C, S, N = 100, 7, 13, 24
    assert o.flops_gridder(C, T, S, N) == r.flops_gridder(C, T, S, N)
    assert o.bytes_gridder(C, T, S, N) == r.bytes_gridder(C, T, S, N)


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_restatement_equals_reference_build_ragged(seed):
    o, r = oracle(), reference()
    p = random_problem(seed, with_w=(seed % 2 == 0), subgrid_size=8 + 8 * (seed % 3))
    assert bits_equal(o.gridder(p), r.gridder(p))
    a, b = o.degridder(p), r.degridder(p)
    ok = ~np.isnan(b.real)  # never-written rows keep the NaN fill on both sides
    assert bits_equal(np.isnan(a.real), np.isnan(b.real)) and bits_equal(a[ok], b[ok])


@needs_ref
@pytest.mark.skipif(not os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libidgref_native.so")),
                    reason="-march=native twin only exists in the build container")
def test_march_v3_equals_march_native():
    """oracle/Makefile swaps the reference's -march=native for -march=x86-64-v3 so the
    .so runs on the GPU box; on this machine that changes no output bit."""
    from oracle_lib import _Lib

    nat = _Lib(os.path.join(ORACLE_DIR, "_ref", "libidgref_native.so"), "idgref_")
    r = reference()
    p = r.make_problem()
    assert bits_equal(nat.gridder(p), r.gridder(p))
    assert bits_equal(nat.degridder(p), r.degridder(p))
