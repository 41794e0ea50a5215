"""GPU parity of the row-column kernels (csrc/gridder_sep.cu, degridder_sep.cu; variant 30 = what
variant 0 resolves to for FAST sincos) against the oracle, through the C ABI.  Same stated
tolerance as tests/test_gpu_parity.py (FAST: per-pol max|d|/max|ref| <= 1e-3, rel-RMS <= 3e-4)."""
import numpy as np
import pytest

import ska_sdp_idg_bench_b200 as idg
from oracle_lib import oracle, random_problem
from test_gpu_parity import assert_close, covered_rows, per_pol_errors, run_degridder, run_gridder

pytestmark = pytest.mark.gpu

SEP = 30
FAST = idg.SINCOS_FAST

SHAPES = [
    dict(subgrid_size=32, nr_channels=16, max_timesteps=128, nr_subgrids=3),   # the bench shape
    dict(subgrid_size=32, nr_channels=64, max_timesteps=24, nr_subgrids=3, nr_stations=8),   # config 4 in miniature
    dict(subgrid_size=64, nr_channels=16, max_timesteps=20, nr_subgrids=2, nr_stations=6, nr_slots=3),   # config 5
    dict(subgrid_size=8, nr_channels=1, max_timesteps=3),
    dict(subgrid_size=16, nr_channels=70, max_timesteps=5, nr_subgrids=4),
    dict(subgrid_size=24, nr_channels=7, max_timesteps=40, nr_subgrids=6),
    dict(subgrid_size=48, nr_channels=5, max_timesteps=20, nr_subgrids=3),
    dict(subgrid_size=72, nr_channels=9, max_timesteps=11, nr_subgrids=2),     # 3 x 2 tiles, the last column tile 8 wide
    dict(subgrid_size=20, nr_channels=300, max_timesteps=2, nr_subgrids=2),
    dict(subgrid_size=32, nr_channels=24, max_timesteps=50, nr_subgrids=2),    # odd number of channel blocks
    dict(subgrid_size=128, nr_channels=8, max_timesteps=6, nr_subgrids=1),     # 4 x 2 tiles
    dict(subgrid_size=16, nr_channels=3, max_timesteps=500, nr_subgrids=2),    # long subgrids: uvw beyond the staged 384 timesteps
]


def with_linear_channels(p):
    p.wavenumbers[:] = (2.9 + 0.0147 * np.arange(p.nr_channels)).astype(np.float32)
    return p


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("linear", [False, True])
def test_gridder_sep_vs_oracle(shape, linear):
    """w != 0, w_step != 0, image_size 0.02 (oracle_lib.random_problem): every subgrid passes the
    separability check (|gamma| r ~ 1e-5 rad) and is gridded by the row-column kernel."""
    o = oracle()
    p = random_problem(201, **shape)
    if linear:
        with_linear_channels(p)
    ref = o.gridder(p)
    got = run_gridder(p, FAST, SEP)
    mx, rms = assert_close(got, ref, 1, FAST, f"gridder sep {shape}")
    print(f"gridder sep {shape} linear={linear}: max rel {mx.max():.2e}, rel rms {rms.max():.2e}")
    # not the per-pixel kernel's result (the row-column kernel did the work) ...
    fb = run_gridder(p, FAST, 10)
    if np.abs(ref).max() > 0:
        assert not np.array_equal(got, fb)
    # ... and the default for this shape
    assert idg.resolve_variant(p.subgrid_size, p.nr_channels, FAST, 0, gridder=True) == SEP
    assert np.array_equal(run_gridder(p, FAST, 0), got)


def test_gridder_sep_planar_config1():
    o = oracle()
    p = o.make_problem()
    ref, ref64 = o.gridder(p), o.gridder_f64(p)
    got = run_gridder(p, FAST, SEP)
    mx, rms = assert_close(got, ref, 1, FAST, "gridder sep config 1")
    e_gpu = np.abs(got - ref64).max() / np.abs(ref64).max()
    e_cpu = np.abs(ref - ref64).max() / np.abs(ref64).max()
    print(f"gridder sep config 1: max rel {mx}, rel rms {rms}; vs f64 {e_gpu:.2e} (cpu {e_cpu:.2e})")
    assert e_gpu <= 2 * e_cpu + 2e-5


def test_gridder_sep_declines_wide_fields():
    """image_size 0.2 with w ~ N(0, 256): the dropped phase term gamma r reaches ~0.05 rad, so every subgrid
    with timesteps fails the check and is gridded by the per-pixel kernel launched behind the row-column one,
    bit for bit; a subgrid on the plane (w = 0, z such that w_offset = 0 is impossible with w_step != 0, so
    w_step = 0 here) still takes the row-column kernel."""
    o = oracle()
    p = random_problem(202, subgrid_size=32, nr_channels=16, max_timesteps=30, nr_subgrids=5, image_size=0.2, w_step=0.0)
    s = int(np.argmax(p.metadata["nr_timesteps"]))
    t0, nt = int(p.metadata[s]["time_offset"]), int(p.metadata[s]["nr_timesteps"])
    p.uvw[t0:t0 + nt, 2] = 0.0                      # one planar subgrid
    ref = o.gridder(p)
    got = run_gridder(p, FAST, SEP)
    assert_close(got, ref, 1, FAST, "gridder sep wide field")
    per_pixel = run_gridder(p, FAST, 24)
    others = np.arange(p.nr_subgrids) != s
    assert np.array_equal(got[others], per_pixel[others])
    assert not np.array_equal(got[s], per_pixel[s])


def test_list_mode_loops_over_more_subgrids_than_ctas():
    """The kernels behind the row-column ones run in list mode: a fixed number of CTAs (592) loop over the subgrids the
    row-column kernel declined.  declined subgrids make every CTA's loop body run more than once - barriers
    re-initialised, TMEM re-allocated per subgrid - and the result must still be the per-pixel kernel's, bit for bit."""
    o = oracle()
    p = random_problem(205, subgrid_size=32, nr_channels=16, max_timesteps=3, nr_subgrids=1000, image_size=0.2, w_step=0.0)
    ref_g, ref_d = o.gridder(p), o.degridder(p)
    rows = covered_rows(p)
    def declined(q):
        """subgrids (and their visibility rows) that certainly fail the separability check: some |w| >= 20 m (the
        few whose w all happen to be small stay with the row-column kernel)"""
        sub = np.zeros(q.nr_subgrids, bool)
        vis = np.zeros(q.total_timesteps, bool)
        for s in range(q.nr_subgrids):
            t0, nt = int(q.metadata[s]["time_offset"]), int(q.metadata[s]["nr_timesteps"])
            if nt and np.abs(q.uvw[t0:t0 + nt, 2]).max() >= 20.0:
                sub[s] = True
                vis[t0:t0 + nt] = True
        assert sub.sum() > 592
        return sub, vis

    sub, vis = declined(p)
    got = run_gridder(p, FAST, SEP)
    assert_close(got, ref_g, 1, FAST, "gridder, declined subgrids")
    assert np.array_equal(got[sub], run_gridder(p, FAST, 24)[sub])
    got = run_degridder(p, FAST, SEP)
    assert_close(got[rows], ref_d[rows], 2, FAST, "degridder, declined subgrids")
    assert np.array_equal(got[vis], run_degridder(p, FAST, 24)[vis])
    # the FP32 kernels in list mode: a shape whose per-pixel fallback is FP32 (9 of 16 channels)
    q = random_problem(206, subgrid_size=16, nr_channels=9, max_timesteps=3, nr_subgrids=1000, image_size=0.2, w_step=0.0)
    sub, vis = declined(q)
    assert np.array_equal(run_gridder(q, FAST, SEP)[sub], run_gridder(q, FAST, 10)[sub])
    # 9 channels: the quad kernel (degridder_tc.cu) in list mode
    assert np.array_equal(run_degridder(q, FAST, SEP)[vis], run_degridder(q, FAST, 22)[vis])
    r = random_problem(207, subgrid_size=16, nr_channels=1, max_timesteps=3, nr_subgrids=1000, image_size=0.2, w_step=0.0)
    sub, vis = declined(r)
    # 1 channel: the FP32 degridder in list mode
    assert np.array_equal(run_degridder(r, FAST, SEP)[vis], run_degridder(r, FAST, 4)[vis])


def test_gridder_sep_fft_shift_and_empty():
    p = with_linear_channels(random_problem(203, subgrid_size=32, nr_channels=16, max_timesteps=20, nr_subgrids=4))
    plain = run_gridder(p, FAST, SEP)
    out = np.full_like(p.subgrids, np.nan)
    idg.c_run_gridder(p.nr_subgrids, p.grid_size, p.subgrid_size, p.image_size, p.w_step, p.nr_channels,
                      p.nr_stations, p.uvw, p.wavenumbers, p.visibilities, p.spheroidal, p.aterms, p.metadata,
                      out, sincos=FAST, variant=SEP, flags=idg.FLAG_FFT_SHIFT)
    h = p.subgrid_size // 2
    assert np.array_equal(np.roll(out, (-h, -h), axis=(2, 3)), plain)
    p.metadata["nr_timesteps"] = 0
    assert not run_gridder(p, FAST, SEP).any()


def test_gridder_sep_cancellation_guard():
    """Cancellation-adversarial input.  The row-column gridder rounds its A operand to fp16 once
    per term, so a pixel's error is ~2e-4 sqrt(sum |vis|^2) whatever the sum comes to.  Here two subgrids see every
    timestep twice, the second time with the visibilities negated up to 1 %: every pixel sum cancels to 1 % of its
    terms, and an fp16 operand would be off by several per cent of the RESULT.  The kernel notices (all pixel sums of
    the subgrid below GS_CANCEL * sqrt(sum |vis|^2)) and the FP32 kernel launched behind it redoes exactly those
    subgrids; the others keep the row-column result.  Errors are reported relative to the result and to
    sqrt(sum |vis|^2), per subgrid."""
    o = oracle()
    p = with_linear_channels(random_problem(204, subgrid_size=32, nr_channels=16, max_timesteps=128, nr_subgrids=5))
    order = np.argsort(-p.metadata["nr_timesteps"])
    cancelled = sorted(int(s) for s in order[:2])
    for s in cancelled:
        t0, nt = int(p.metadata[s]["time_offset"]), int(p.metadata[s]["nr_timesteps"]) // 2 * 2
        assert nt >= 4
        p.uvw[t0 + 1:t0 + nt:2] = p.uvw[t0:t0 + nt:2]
        p.visibilities[t0 + 1:t0 + nt:2] = -p.visibilities[t0:t0 + nt:2] * np.float32(1.01)
        if int(p.metadata[s]["nr_timesteps"]) % 2:          # an odd last timestep would not cancel
            p.visibilities[t0 + nt] = 0
    ref, ref64 = o.gridder(p), o.gridder_f64(p)
    got = run_gridder(p, FAST, SEP)
    fp32 = run_gridder(p, FAST, 10)
    for s in range(p.nr_subgrids):
        t0, nt = int(p.metadata[s]["time_offset"]), int(p.metadata[s]["nr_timesteps"])
        norm = np.sqrt((np.abs(p.visibilities[t0:t0 + nt].astype(np.complex128)) ** 2).sum())
        peak = np.abs(ref64[s]).max()
        err = np.abs(got[s] - ref64[s]).max()
        print(f"subgrid {s} ({'cancelling' if s in cancelled else 'plain'}, {nt} timesteps): max|result| = {peak:.3e}, "
              f"sqrt(sum|v|^2) = {norm:.3e}, error / max|result| = {err / max(peak, 1e-30):.2e}, "
              f"error / sqrt(sum|v|^2) = {err / max(norm, 1e-30):.2e}, cpu-f32 error / max|result| = "
              f"{np.abs(ref[s] - ref64[s]).max() / max(peak, 1e-30):.2e}")
        if nt == 0:
            continue
        if s in cancelled:
            assert np.array_equal(got[s], fp32[s]), "a cancelling subgrid must be redone by the FP32 kernel"
        else:
            assert not np.array_equal(got[s], fp32[s]), "a plain subgrid must keep the row-column result"
        # per subgrid, relative to its own result: the FAST tolerance
        assert err <= 1e-3 * peak, (s, err / peak)
    assert_close(got, ref, 1, FAST, "gridder sep with cancelling subgrids")


# ------------------------------------------------------------------------------------- degridder
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("linear", [False, True])
def test_degridder_sep_vs_oracle(shape, linear):
    o = oracle()
    p = random_problem(211, **shape)
    if linear:
        with_linear_channels(p)
    ref = o.degridder(p)
    rows = covered_rows(p)
    if p.subgrid_size > 64:   # the row-column degridder takes subgrids of up to 64 x 64 pixels
        with pytest.raises(idg.IdgError):
            run_degridder(p, FAST, SEP)
        assert idg.resolve_variant(p.subgrid_size, p.nr_channels, FAST, 0, gridder=False) != SEP
        assert_close(run_degridder(p, FAST, 0)[rows], ref[rows], 2, FAST, f"degridder default {shape}")
        return
    got = run_degridder(p, FAST, SEP)
    mx, rms = assert_close(got[rows], ref[rows], 2, FAST, f"degridder sep {shape}")
    print(f"degridder sep {shape} linear={linear}: max rel {mx.max():.2e}, rel rms {rms.max():.2e}")
    assert not got[~rows].any(), "rows no subgrid covers must come back as zeros"
    assert np.array_equal(run_degridder(p, FAST, 0), got)


def test_degridder_sep_declines_wide_fields():
    """As test_gridder_sep_declines_wide_fields: subgrids that fail the separability check are degridded by the
    per-pixel kernel launched behind the row-column one, bit for bit."""
    o = oracle()
    p = random_problem(212, subgrid_size=32, nr_channels=16, max_timesteps=30, nr_subgrids=5, image_size=0.2, w_step=0.0)
    s = int(np.argmax(p.metadata["nr_timesteps"]))
    t0, nt = int(p.metadata[s]["time_offset"]), int(p.metadata[s]["nr_timesteps"])
    p.uvw[t0:t0 + nt, 2] = 0.0                      # one planar subgrid
    ref = o.degridder(p)
    rows = covered_rows(p)
    got = run_degridder(p, FAST, SEP)
    assert_close(got[rows], ref[rows], 2, FAST, "degridder sep wide field")
    per_pixel = run_degridder(p, FAST, 24)
    mine = np.zeros(p.total_timesteps, bool)
    mine[t0:t0 + nt] = True
    assert np.array_equal(got[rows & ~mine], per_pixel[rows & ~mine])
    assert not np.array_equal(got[mine], per_pixel[mine])


def test_degridder_sep_fft_shift_and_empty():
    p = with_linear_channels(random_problem(213, subgrid_size=32, nr_channels=16, max_timesteps=20, nr_subgrids=4))
    plain = run_degridder(p, FAST, SEP)
    h = p.subgrid_size // 2
    q = random_problem(213, subgrid_size=32, nr_channels=16, max_timesteps=20, nr_subgrids=4)
    with_linear_channels(q)
    q.subgrids[:] = np.roll(p.subgrids, (h, h), axis=(2, 3))
    out = np.full_like(q.visibilities, np.nan)
    idg.c_run_degridder(q.nr_subgrids, q.grid_size, q.subgrid_size, q.image_size, q.w_step, q.nr_channels,
                        q.nr_stations, q.uvw, q.wavenumbers, out, q.spheroidal, q.aterms, q.metadata, q.subgrids,
                        sincos=FAST, variant=SEP, flags=idg.FLAG_FFT_SHIFT)
    assert np.array_equal(out, plain)
    p.metadata["nr_timesteps"] = 0
    assert not run_degridder(p, FAST, SEP).any()


def test_degridder_sep_config1():
    o = oracle()
    p = o.make_problem()
    ref, ref64 = o.degridder(p), o.degridder_f64(p)
    got = run_degridder(p, FAST, SEP)
    mx, rms = assert_close(got, ref, 2, FAST, "degridder sep config 1")
    e_gpu = np.abs(got - ref64).max() / np.abs(ref64).max()
    e_cpu = np.abs(ref - ref64).max() / np.abs(ref64).max()
    print(f"degridder sep config 1: max rel {mx}, rel rms {rms}; vs f64 {e_gpu:.2e} (cpu {e_cpu:.2e})")
    assert e_gpu <= 2 * e_cpu + 2e-5


@pytest.mark.parametrize("variant", [31, 32])
def test_degridder_sep_operand_scaling(variant):
    """The B operand (P') is scaled per subgrid by a power of two into the range of its fp16 and e4m3 parts
    (degridder_sep.cu: DS_B_EXP), the A operand by 2^8.  The degridder is linear in a subgrid, so
    (a) a power-of-two factor on a subgrid's pixels comes back as exactly that factor on its visibilities, whatever
        the magnitude (2^-100 .. 2^100) and different for every subgrid of a launch,
    (b) a subgrid of zeros gives zeros (no scale), next to scaled ones,
    (c) one pixel 10^6 times larger than the rest (most pixels then sit in the subnormal steps of the 8-bit parts)
        stays inside the tolerance, which is relative to the largest visibility."""
    o = oracle()
    p = with_linear_channels(random_problem(216, subgrid_size=32, nr_channels=16, max_timesteps=24, nr_subgrids=6))
    base = run_degridder(p, FAST, variant)
    exps = [0, -100, 100, 37, -61, None]             # None: the subgrid of zeros
    q = with_linear_channels(random_problem(216, subgrid_size=32, nr_channels=16, max_timesteps=24, nr_subgrids=6))
    for s, e in enumerate(exps):
        q.subgrids[s] = 0 if e is None else p.subgrids[s] * np.float32(2.0) ** e
    got = run_degridder(q, FAST, variant)
    for s, e in enumerate(exps):
        t0, nt = int(p.metadata[s]["time_offset"]), int(p.metadata[s]["nr_timesteps"])
        rows = slice(t0, t0 + nt)      # random_problem: one baseline, time_offset counts from the start of uvw
        if e is None:
            assert not got[rows].any()
        else:
            want = (base[rows].astype(np.complex128) * 2.0 ** e).astype(np.complex64)
            assert np.array_equal(got[rows].view(np.float32), want.view(np.float32)), (s, e)
    # (c) dynamic range inside one subgrid
    r = with_linear_channels(random_problem(217, subgrid_size=32, nr_channels=16, max_timesteps=24, nr_subgrids=3))
    r.subgrids[1, :, 13, 21] *= 1e6
    ref = o.degridder(r)
    rows = covered_rows(r)
    assert_close(run_degridder(r, FAST, variant)[rows], ref[rows], 2, FAST, "degridder sep, one pixel 1e6 x the rest")


# ---------------------------------------------------------- degridder: the pipelined persistent kernel (variant 32)
PIPE, CLASSIC = 32, 31


def assert_same_sum(got, ref):
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("shape", [s for s in SHAPES if s["subgrid_size"] <= 32])
def test_degridder_pipe_equals_one_subgrid_per_cta(shape):
    """degridder_sep.cu holds two kernels with the same arithmetic: one subgrid per CTA (31) and the warp-specialised
    persistent pipeline (32, what 30 selects where its buffers fit).  Same MMAs in the same order, same split of the sum
    over the rows: bit-identical visibilities."""
    p = random_problem(221, **shape)
    for q in (p, with_linear_channels(p)):
        got, ref = run_degridder(q, FAST, PIPE), run_degridder(q, FAST, CLASSIC)
        assert_same_sum(got, ref)
        assert np.array_equal(got, run_degridder(q, FAST, PIPE))


def test_degridder_pipe_many_ragged_subgrids():
    """More subgrids than SMs: every CTA of the persistent kernel walks through several subgrids with different numbers of
    timesteps (0 included), some of them declined (wide field, large w) and left to the per-pixel kernel behind it, so
    that both B buffers, both meta slots and every barrier phase are reused many times."""
    o = oracle()
    p = random_problem(222, subgrid_size=32, nr_channels=16, max_timesteps=70, nr_subgrids=900)
    assert (p.metadata["nr_timesteps"] == 0).any()
    ref = o.degridder(p)
    rows = covered_rows(p)
    got = run_degridder(p, FAST, PIPE)
    assert_close(got[rows], ref[rows], 2, FAST, "degridder pipe, 900 ragged subgrids")
    assert_same_sum(got, run_degridder(p, FAST, CLASSIC))
    assert not got[~rows].any()
    # the same with a wide field: most subgrids are declined, a few (small |w|) stay
    q = random_problem(223, subgrid_size=24, nr_channels=8, max_timesteps=9, nr_subgrids=700, image_size=0.2, w_step=0.0)
    assert_same_sum(run_degridder(q, FAST, PIPE), run_degridder(q, FAST, CLASSIC))
    rows = covered_rows(q)
    assert_close(run_degridder(q, FAST, PIPE)[rows], o.degridder(q)[rows], 2, FAST, "degridder pipe, wide field")


def test_degridder_pipe_refused_where_buffers_do_not_fit():
    """4000 channels: s_wn / s_dw / s_lin no longer fit beside the pipeline's buffers; 32 refuses, 30 takes the
    one-subgrid-per-CTA kernel."""
    p = random_problem(224, subgrid_size=32, nr_channels=4000, max_timesteps=2, nr_subgrids=2)
    with pytest.raises(idg.IdgError):
        run_degridder(p, FAST, PIPE)
    assert np.array_equal(run_degridder(p, FAST, SEP), run_degridder(p, FAST, CLASSIC))
