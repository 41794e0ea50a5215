"""The "next" rows of SURVEY.md 8f around the gridder / degridder: grid adder (8f-1), subgrid FFT
(8f-2), splitter (8f-3) and the subgrid FFT shift flag.  Oracle statements vs numpy on the CPU,
CUDA kernels vs the oracle on the GPU through the C ABI.  Parity of these rows is UNPINNED - the
reference has none of them (oracle/idg_next_oracle.c); the shift flag of the gridder / degridder
is checked against the pinned path (same values, permuted)."""
import numpy as np
import pytest

from oracle_lib import oracle, random_problem

SHIFT = 1   # IDGO_FFT_SHIFT == IDGB200_FLAG_FFT_SHIFT


def numpy_adder(p, shift=False):
    G, N = p.grid_size, p.subgrid_size
    grid = np.zeros((4, G, G), np.complex128)
    for s in range(p.nr_subgrids):
        x0, y0 = int(p.metadata["x"][s]), int(p.metadata["y"][s])
        sg = np.roll(p.subgrids[s], (-(N // 2), -(N // 2)), axis=(1, 2)) if shift else p.subgrids[s]
        ys, xs = np.arange(N) + y0, np.arange(N) + x0
        my, mx = (ys >= 0) & (ys < G), (xs >= 0) & (xs < G)
        grid[:, ys[my][:, None], xs[mx][None, :]] += sg[:, my][:, :, mx]
    return grid


def numpy_splitter(p, grid, shift=False):
    G, N = p.grid_size, p.subgrid_size
    out = np.zeros((p.nr_subgrids, 4, N, N), np.complex64)
    for s in range(p.nr_subgrids):
        x0, y0 = int(p.metadata["x"][s]), int(p.metadata["y"][s])
        ys, xs = np.arange(N) + y0, np.arange(N) + x0
        my, mx = (ys >= 0) & (ys < G), (xs >= 0) & (xs < G)
        sg = np.zeros((4, N, N), np.complex64)
        sg[:, np.nonzero(my)[0][:, None], np.nonzero(mx)[0][None, :]] = grid[:, ys[my][:, None], xs[mx][None, :]]
        out[s] = np.roll(sg, (N // 2, N // 2), axis=(1, 2)) if shift else sg
    return out


def edge_problem(seed, **kw):
    p = random_problem(seed, **kw)
    G, N = p.grid_size, p.subgrid_size
    # overhang every edge and a corner, one subgrid fully outside
    p.metadata["x"][:5] = [-N // 2, G - N // 3, 5, G + 3, -N - 1][: min(5, p.nr_subgrids)]
    p.metadata["y"][:5] = [7, -N // 4, G - 1, 2, 9][: min(5, p.nr_subgrids)]
    return p


def tiled_problem(seed, **kw):
    """subgrids placed side by side inside the grid: no overlap, nothing clipped"""
    p = random_problem(seed, **kw)
    G, N = p.grid_size, p.subgrid_size
    per_row = G // N
    assert p.nr_subgrids <= per_row * per_row
    for s in range(p.nr_subgrids):
        p.metadata["x"][s] = (s % per_row) * N
        p.metadata["y"][s] = (s // per_row) * N
    return p


def cplx_planes(seed, *shape):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


# ------------------------------------------------------------------ CPU: oracle vs numpy
@pytest.mark.parametrize("flags", [0, SHIFT])
def test_adder_oracle_matches_numpy(flags):
    o = oracle()
    for seed, N in ((1, 16), (2, 10)):
        p = edge_problem(seed, nr_subgrids=12, subgrid_size=N, grid_size=96)
        got, ref = o.adder(p, flags=flags), numpy_adder(p, bool(flags))
        assert np.allclose(got, ref, rtol=0, atol=1e-5 * np.abs(ref).max())
        assert np.abs(got).sum() > 0


@pytest.mark.parametrize("flags", [0, SHIFT])
def test_splitter_oracle_matches_numpy(flags):
    o = oracle()
    for seed, N in ((3, 16), (4, 10)):
        p = edge_problem(seed, nr_subgrids=12, subgrid_size=N, grid_size=96)
        grid = cplx_planes(seed, 4, 96, 96)
        got = o.splitter(p, grid, flags=flags)
        assert np.array_equal(got, numpy_splitter(p, grid, bool(flags)))
        assert (got[4] == 0).all()            # the subgrid that lies outside the grid


@pytest.mark.parametrize("flags", [0, SHIFT])
def test_splitter_inverts_adder_on_disjoint_subgrids(flags):
    o = oracle()
    p = tiled_problem(5, nr_subgrids=9, subgrid_size=16, grid_size=64)
    assert np.array_equal(o.splitter(p, o.adder(p, flags=flags), flags=flags), p.subgrids)


@pytest.mark.parametrize("N", [1, 5, 8, 24, 32])
def test_fft_oracle_matches_numpy(N):
    o = oracle()
    a = cplx_planes(N, 3, 4, N, N)
    fwd = np.fft.fft2(a.astype(np.complex128), axes=(-2, -1))
    bwd = np.fft.ifft2(a.astype(np.complex128), axes=(-2, -1))
    assert np.allclose(o.subgrid_fft(a, 1), fwd, rtol=0, atol=1e-6 * np.abs(fwd).max())
    assert np.allclose(o.subgrid_fft(a, -1), bwd, rtol=0, atol=1e-6 * np.abs(bwd).max())
    back = o.subgrid_fft(o.subgrid_fft(a, 1), -1)
    assert np.allclose(back, a, rtol=0, atol=1e-6 * np.abs(a).max())


# ------------------------------------------------------------------ GPU: kernels vs oracle
def _dev_problem(p):
    import torch
    dev = torch.device("cuda", 0)
    meta = torch.from_numpy(np.ascontiguousarray(p.metadata).view(np.int32).reshape(-1, 9)).to(dev)
    return dev, meta, torch.from_numpy(p.subgrids).to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, SHIFT])
@pytest.mark.parametrize("shape", [dict(nr_subgrids=40, subgrid_size=32, grid_size=256),
                                   dict(nr_subgrids=9, subgrid_size=24, grid_size=100),
                                   dict(nr_subgrids=3, subgrid_size=64, grid_size=64)])
def test_adder_gpu_vs_oracle(shape, flags):
    import torch

    import ska_sdp_idg_bench_b200 as idg
    o = oracle()
    p = edge_problem(11, **shape)
    ref = o.adder(p, flags=flags)
    dev, meta, sg = _dev_problem(p)
    G, N = p.grid_size, p.subgrid_size
    tol = 1e-5 * np.abs(ref).max()
    grid = torch.zeros((4, G, G), dtype=torch.complex64, device=dev)
    idg.adder(p.nr_subgrids, G, N, meta, sg, grid, flags=flags)
    assert np.allclose(grid.cpu().numpy(), ref, rtol=0, atol=tol)
    # the same grid cut into row blocks behind separate pointers (what the multi-GPU path uses)
    for nr_parts in (2, 3):
        rpp = (G + nr_parts - 1) // nr_parts
        parts = [torch.zeros((4, rpp, G), dtype=torch.complex64, device=dev) for _ in range(nr_parts)]
        idg.adder(p.nr_subgrids, G, N, meta, sg, parts, rows_per_part=rpp, flags=flags)
        whole = torch.cat(parts, dim=1)[:, :G].cpu().numpy()
        assert np.allclose(whole, ref, rtol=0, atol=tol)
    # accumulates: a second call doubles the grid
    idg.adder(p.nr_subgrids, G, N, meta, sg, grid, flags=flags)
    assert np.allclose(grid.cpu().numpy(), 2 * ref, rtol=0, atol=2 * tol)
    with pytest.raises(idg.IdgError):
        idg.adder(p.nr_subgrids, G, N, meta, sg, [grid], rows_per_part=G // 2)   # parts do not cover the grid
    with pytest.raises(idg.IdgError):
        idg.adder(p.nr_subgrids, G, N, meta, sg, grid, flags=6)                  # unknown flag bits


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, SHIFT])
@pytest.mark.parametrize("shape", [dict(nr_subgrids=40, subgrid_size=32, grid_size=256),
                                   dict(nr_subgrids=9, subgrid_size=24, grid_size=100),
                                   dict(nr_subgrids=3, subgrid_size=64, grid_size=64)])
def test_splitter_gpu_vs_oracle_bitwise(shape, flags):
    import torch

    import ska_sdp_idg_bench_b200 as idg
    o = oracle()
    p = edge_problem(12, **shape)
    G, N = p.grid_size, p.subgrid_size
    grid_h = cplx_planes(7, 4, G, G)
    ref = o.splitter(p, grid_h, flags=flags)
    dev, meta, _ = _dev_problem(p)
    grid = torch.from_numpy(grid_h).to(dev)
    out = torch.full((p.nr_subgrids, 4, N, N), float("nan"), dtype=torch.complex64, device=dev)
    idg.splitter(p.nr_subgrids, G, N, meta, out, grid, flags=flags)
    assert np.array_equal(out.cpu().numpy(), ref)      # a gather: bit for bit
    for nr_parts in (2, 3):
        rpp = (G + nr_parts - 1) // nr_parts
        padded = torch.zeros((4, nr_parts * rpp, G), dtype=torch.complex64, device=dev)
        padded[:, :G] = grid
        parts = [padded[:, r * rpp:(r + 1) * rpp].contiguous() for r in range(nr_parts)]
        out.fill_(float("nan"))
        idg.splitter(p.nr_subgrids, G, N, meta, out, parts, rows_per_part=rpp, flags=flags)
        assert np.array_equal(out.cpu().numpy(), ref)
    # round trip through the GPU adder on disjoint subgrids
    q = tiled_problem(13, nr_subgrids=4, subgrid_size=N, grid_size=2 * N)
    dev, meta, sg = _dev_problem(q)
    g2 = torch.zeros((4, 2 * N, 2 * N), dtype=torch.complex64, device=dev)
    idg.adder(q.nr_subgrids, 2 * N, N, meta, sg, g2, flags=flags)
    back = torch.empty_like(sg)
    idg.splitter(q.nr_subgrids, 2 * N, N, meta, back, g2, flags=flags)
    assert torch.equal(back, sg)
    # a subgrid array that is only 8-byte aligned (a complex64 view at an odd element offset) takes the 8-byte
    # stores instead of faulting; the adder likewise reads it
    flat = torch.full((sg.numel() + 1,), float("nan"), dtype=torch.complex64, device=dev)
    odd = flat[1:].view(sg.shape)
    assert odd.data_ptr() % 16 == 8
    idg.splitter(q.nr_subgrids, 2 * N, N, meta, odd, g2, flags=flags)
    assert torch.equal(odd, sg)
    g3 = torch.zeros_like(g2)
    idg.adder(q.nr_subgrids, 2 * N, N, meta, odd, g3, flags=flags)
    assert torch.equal(g3, g2)
    # argument validation of the Python mirror: wrong device, dtype, size, layout raise instead of reaching the kernel
    with pytest.raises(TypeError):
        idg.splitter(q.nr_subgrids, 2 * N, N, meta.cpu(), back, g2)
    with pytest.raises(TypeError):
        idg.splitter(q.nr_subgrids, 2 * N, N, meta.to(torch.int64), back, g2)
    with pytest.raises(ValueError):
        idg.adder(q.nr_subgrids + 1, 2 * N, N, meta, sg, g2)
    with pytest.raises(ValueError):
        idg.adder(q.nr_subgrids, 2 * N, N, meta, sg, g2[:, : N])          # short (and non-contiguous) grid part
    with pytest.raises(ValueError):
        idg.reduce_parts([g2, g2[:, :N].contiguous()], torch.empty_like(g2))


@pytest.mark.gpu
@pytest.mark.parametrize("N", [8, 16, 24, 32, 48, 64, 20, 7, 1])
def test_subgrid_fft_gpu_vs_oracle(N):
    """Stated tolerance: max|d| <= 2e-6 * max|ref| * log2(N^2) + tiny (fp32 butterflies vs the
    float64 sums of the oracle), both directions; backward(forward(x)) returns x."""
    import torch

    import ska_sdp_idg_bench_b200 as idg
    o = oracle()
    S = 11                                   # 44 planes: not a multiple of the planes per CTA
    a = cplx_planes(100 + N, S, 4, N, N)
    dev = torch.device("cuda", 0)
    tol = 2e-6 * max(1.0, np.log2(N * N))
    for direction in (1, -1):
        ref = o.subgrid_fft(a, direction)
        t = torch.from_numpy(a).to(dev)
        idg.subgrid_fft(S, N, t, direction)
        got = t.cpu().numpy()
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err <= tol, f"N={N} direction={direction}: {err:.3e} > {tol:.3e}"
    t = torch.from_numpy(a).to(dev)
    idg.subgrid_fft(S, N, t, 1)
    idg.subgrid_fft(S, N, t, -1)
    assert np.abs(t.cpu().numpy() - a).max() <= 2 * tol * np.abs(a).max()
    idg.subgrid_fft(0, N, t, 1)              # empty batch
    with pytest.raises(idg.IdgError):
        idg.subgrid_fft(S, N, t, 0)


@pytest.mark.gpu
def test_subgrid_fft_full_size_parseval_and_linearity():
    """BASELINE config 2 size (24,500 subgrids of 32 x 32 x 4): properties instead of an oracle run.
    Parseval per plane, linearity, and the DC term equals the plane sum."""
    import torch

    import ska_sdp_idg_bench_b200 as idg
    S, N = 24500, 32
    dev = torch.device("cuda", 0)
    g = torch.Generator(dev).manual_seed(5)
    a = torch.view_as_complex(torch.randn((S, 4, N, N, 2), device=dev, generator=g))
    b = torch.view_as_complex(torch.randn((S, 4, N, N, 2), device=dev, generator=g))
    fa, fb, fab = a.clone(), b.clone(), (a + 2 * b)
    for t in (fa, fb, fab):
        idg.subgrid_fft(S, N, t, 1)
    e_in = (a.abs() ** 2).sum(dim=(-2, -1)).double()
    e_out = (fa.abs() ** 2).sum(dim=(-2, -1)).double() / (N * N)
    assert float(((e_out - e_in).abs() / e_in).max()) < 1e-5
    scale = float(fab.abs().max())
    assert float((fab - (fa + 2 * fb)).abs().max()) < 1e-5 * scale
    assert float((fa[..., 0, 0] - a.sum(dim=(-2, -1))).abs().max()) < 1e-5 * scale
    # against torch's FFT (cuFFT) as a second opinion at full size
    ref = torch.fft.fft2(a)
    assert float((fa - ref).abs().max()) < 2e-5 * float(ref.abs().max())


def _run(fn, p, out_vis, sg, sincos, flags):
    import ska_sdp_idg_bench_b200 as idg
    fn = idg.c_run_gridder if fn == "gridder" else idg.c_run_degridder
    fn(p.nr_subgrids, p.grid_size, p.subgrid_size, p.image_size, p.w_step, p.nr_channels, p.nr_stations,
       p.uvw, p.wavenumbers, out_vis, p.spheroidal, p.aterms, p.metadata, sg, sincos=sincos, flags=flags)


@pytest.mark.gpu
@pytest.mark.parametrize("sincos", [0, 1])         # FAST -> tcgen05 kernels, REDUCED -> FP32 kernels
@pytest.mark.parametrize("shape", [dict(nr_subgrids=5, subgrid_size=32, nr_channels=16, max_timesteps=128),
                                   dict(nr_subgrids=4, subgrid_size=24, nr_channels=5, max_timesteps=37),
                                   dict(nr_subgrids=2, subgrid_size=64, nr_channels=8, max_timesteps=64)])
def test_fft_shift_flag_permutes_the_pinned_result(shape, sincos):
    """IDGB200_FLAG_FFT_SHIFT: the gridder stores, the degridder reads pixel (y, x) at
    ((y + N/2) % N, (x + N/2) % N): the same values as the pinned, unshifted path, bit for bit."""
    p = random_problem(21, **shape)
    N = p.subgrid_size
    plain = np.full_like(p.subgrids, np.nan)
    shifted = np.full_like(p.subgrids, np.nan)
    _run("gridder", p, p.visibilities, plain, sincos, 0)
    _run("gridder", p, p.visibilities, shifted, sincos, SHIFT)
    assert np.isfinite(plain.view(np.float32)).all()
    assert np.array_equal(shifted, np.roll(plain, (N // 2, N // 2), axis=(2, 3)))
    v_plain = np.full_like(p.visibilities, np.nan)
    v_shift = np.full_like(p.visibilities, np.nan)
    _run("degridder", p, v_plain, p.subgrids, sincos, 0)
    _run("degridder", p, v_shift, np.ascontiguousarray(np.roll(p.subgrids, (N // 2, N // 2), axis=(2, 3))),
         sincos, SHIFT)
    assert np.array_equal(v_plain.view(np.uint32), v_shift.view(np.uint32))


@pytest.mark.gpu
def test_imaging_round_trip_pipeline():
    """gridder -> FFT -> adder -> splitter -> inverse FFT -> degridder on the GPU equals the same
    chain of oracle statements (gridder / degridder pinned, the middle unpinned)."""
    import torch

    import ska_sdp_idg_bench_b200 as idg
    o = oracle()
    p = tiled_problem(31, nr_subgrids=9, subgrid_size=32, nr_channels=8, max_timesteps=64, grid_size=128)
    S, G, N = p.nr_subgrids, p.grid_size, p.subgrid_size
    # oracle chain
    sg = o.gridder(p)
    grid = o.adder(p, o.subgrid_fft(sg, 1), flags=SHIFT)
    back = o.subgrid_fft(o.splitter(p, grid, flags=SHIFT), -1)
    q = random_problem(31, nr_subgrids=9, subgrid_size=32, nr_channels=8, max_timesteps=64, grid_size=128)
    q.metadata[:] = p.metadata
    q.subgrids[:] = back
    ref_vis = o.degridder(q)
    # GPU chain, device resident
    dev = torch.device("cuda", 0)
    T = p.total_timesteps
    dt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    uvw, wn, vis, sph, at = dt(p.uvw), dt(p.wavenumbers), dt(p.visibilities), dt(p.spheroidal), dt(p.aterms)
    meta = dt(np.ascontiguousarray(p.metadata).view(np.int32).reshape(-1, 9))
    d_sg = torch.empty((S, 4, N, N), dtype=torch.complex64, device=dev)
    d_grid = torch.zeros((4, G, G), dtype=torch.complex64, device=dev)
    d_vis = torch.empty_like(vis)
    args = (S, G, N, p.image_size, p.w_step, p.nr_channels, p.nr_stations, T, uvw, wn)
    idg.gridder(*args, vis, sph, at, meta, d_sg, sincos=idg.SINCOS_REDUCED)
    idg.subgrid_fft(S, N, d_sg, 1)
    idg.adder(S, G, N, meta, d_sg, d_grid, flags=SHIFT)
    d_sg.fill_(float("nan"))
    idg.splitter(S, G, N, meta, d_sg, d_grid, flags=SHIFT)
    idg.subgrid_fft(S, N, d_sg, -1)
    idg.degridder(*args, d_vis, sph, at, meta, d_sg, sincos=idg.SINCOS_REDUCED)
    got = d_vis.cpu().numpy()
    err = np.abs(got - ref_vis).max() / np.abs(ref_vis).max()
    assert err < 1e-4, err
    # and the round trip through the grid reproduced the gridder's subgrids (disjoint tiles)
    assert np.abs(back - sg).max() <= 1e-5 * np.abs(sg).max()


@pytest.mark.gpu
def test_reduce_parts_is_the_ordered_sum():
    """idgb200_reduce_parts: out = sources[0] + sources[1] + ... in that order, bit for bit (the
    second half of the multi-GPU adder; here all sources live on one GPU)."""
    import torch

    import ska_sdp_idg_bench_b200 as idg
    dev = torch.device("cuda", 0)
    g = torch.Generator(dev).manual_seed(3)
    n = 4 * 37 * 50                                     # even, not a multiple of the CTA size
    src = [torch.view_as_complex(torch.randn((n, 2), device=dev, generator=g)) for _ in range(5)]
    out = torch.full((n,), float("nan"), dtype=torch.complex64, device=dev)
    idg.reduce_parts(src, out)
    ref = src[0].clone()
    for t in src[1:]:
        ref = ref + t
    assert torch.equal(out, ref)
    idg.reduce_parts(src[:1], out)
    assert torch.equal(out, src[0])
    with pytest.raises(idg.IdgError):
        idg.reduce_parts(src, out[:n - 1])              # odd element count
    with pytest.raises(idg.IdgError):
        idg.reduce_parts([t for t in src] * 4, out)     # more than 16 sources
