"""Grid adder (the first "next" row): oracle statement vs numpy on the CPU, CUDA
kernel vs the oracle on the GPU.  Parity is UNPINNED - the reference has no adder to compare with
(oracle/idg_adder_oracle.c)."""
import numpy as np
import pytest

from oracle_lib import oracle, random_problem


def numpy_adder(p):
    G, N = p.grid_size, p.subgrid_size
    grid = np.zeros((4, G, G), np.complex128)
    for s in range(p.nr_subgrids):
        x0, y0 = int(p.metadata["x"][s]), int(p.metadata["y"][s])
        ys, xs = np.arange(N) + y0, np.arange(N) + x0
        my, mx = (ys >= 0) & (ys < G), (xs >= 0) & (xs < G)
        grid[:, ys[my][:, None], xs[mx][None, :]] += p.subgrids[s][:, my][:, :, mx]
    return grid


def edge_problem(seed, **kw):
    p = random_problem(seed, **kw)
    G, N = p.grid_size, p.subgrid_size
    # overhang every edge and a corner, one subgrid fully outside
    p.metadata["x"][:5] = [-N // 2, G - N // 3, 5, G + 3, -N - 1][: min(5, p.nr_subgrids)]
    p.metadata["y"][:5] = [7, -N // 4, G - 1, 2, 9][: min(5, p.nr_subgrids)]
    return p


def test_adder_oracle_matches_numpy():
    o = oracle()
    for seed in (1, 2):
        p = edge_problem(seed, nr_subgrids=12, subgrid_size=16, grid_size=96)
        got, ref = o.adder(p), numpy_adder(p)
        assert np.allclose(got, ref, rtol=0, atol=1e-5 * np.abs(ref).max())
        assert np.abs(got).sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [dict(nr_subgrids=40, subgrid_size=32, grid_size=256),
                                   dict(nr_subgrids=9, subgrid_size=24, grid_size=100),
                                   dict(nr_subgrids=3, subgrid_size=64, grid_size=64)])
def test_adder_gpu_vs_oracle(shape):
    import torch

    import ska_sdp_idg_bench_b200 as idg
    o = oracle()
    p = edge_problem(11, **shape)
    ref = o.adder(p)
    dev = torch.device("cuda", 0)
    meta = torch.from_numpy(np.ascontiguousarray(p.metadata).view(np.int32).reshape(-1, 9)).to(dev)
    sg = torch.from_numpy(p.subgrids).to(dev)
    G, N = p.grid_size, p.subgrid_size
    tol = 1e-5 * np.abs(ref).max()
    grid = torch.zeros((4, G, G), dtype=torch.complex64, device=dev)
    idg.adder(p.nr_subgrids, G, N, meta, sg, grid)
    assert np.allclose(grid.cpu().numpy(), ref, rtol=0, atol=tol)
    # the same grid cut into row blocks behind separate pointers (what the multi-GPU path uses)
    for nr_parts in (2, 3):
        rpp = (G + nr_parts - 1) // nr_parts
        parts = [torch.zeros((4, rpp, G), dtype=torch.complex64, device=dev) for _ in range(nr_parts)]
        idg.adder(p.nr_subgrids, G, N, meta, sg, parts, rows_per_part=rpp)
        whole = torch.cat(parts, dim=1)[:, :G].cpu().numpy()
        assert np.allclose(whole, ref, rtol=0, atol=tol)
    # accumulates: a second call doubles the grid
    idg.adder(p.nr_subgrids, G, N, meta, sg, grid)
    assert np.allclose(grid.cpu().numpy(), 2 * ref, rtol=0, atol=2 * tol)
    with pytest.raises(idg.IdgError):
        idg.adder(p.nr_subgrids, G, N, meta, sg, [grid], rows_per_part=G // 2)   # parts do not cover the grid
