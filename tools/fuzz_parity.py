"""Randomised parity run (not part of the pytest suite): many random ragged shapes through the
default FAST kernels (tcgen05 where the shape qualifies) and a few explicit variants, against the
oracle.  Usage: python tools/fuzz_parity.py [n_cases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ska_sdp_idg_bench_b200 as idg  # noqa: E402
from oracle_lib import oracle, random_problem  # noqa: E402
from test_gpu_parity import TOL, covered_rows, per_pol_errors, run_degridder, run_gridder  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
o = oracle()
o.set_threads(o.max_threads())
tmx, trms = TOL[idg.SINCOS_FAST]
worst = 0.0
for case in range(n_cases):
    N = int(rng.choice([16, 20, 24, 32, 40, 48, 64]))
    C = int(rng.choice([4, 8, 12, 16, 24, 32, 33, 64]))
    T = int(rng.integers(1, 70))
    S = int(rng.integers(1, 6))
    if N <= 32 and rng.integers(0, 4) == 0:      # more subgrids than SMs: the persistent degridder's CTAs walk through several
        S, T = int(rng.integers(150, 500)), min(T, 24)
    linear = bool(rng.integers(0, 2))
    planar = bool(rng.integers(0, 2))     # w = 0 everywhere: the pixel-pair folded paths
    p = random_problem(int(rng.integers(1 << 30)), subgrid_size=N, nr_channels=C, max_timesteps=T, nr_subgrids=S,
                       nr_stations=int(rng.integers(2, 7)), nr_slots=int(rng.integers(1, 4)), with_w=not planar)
    if linear:
        # every other case with a spacing that is exact in fp32 (one spacing for all blocks bit for bit: the
        # regular-case gridder loops, folded or not)
        dw = float(rng.choice([2.0 ** -6, 2.0 ** -7])) if rng.integers(0, 2) else float(rng.uniform(0.001, 0.02))
        p.wavenumbers[:] = (2.5 + dw * np.arange(C)).astype(np.float32)
    if planar and S > 1 and rng.integers(0, 3) == 0:   # one subgrid off the plane in an otherwise planar launch
        s = int(np.argmax(p.metadata["nr_timesteps"]))
        p.uvw[int(p.metadata[s]["time_offset"]), 2] = 2.5
    ref_g, ref_d = o.gridder(p), o.degridder(p)
    rows = covered_rows(p)
    for gv in (0, 22):
        mx, rms = per_pol_errors(run_gridder(p, idg.SINCOS_FAST, gv), ref_g, 1)
        ok = (mx <= tmx).all() and (rms <= trms).all()
        worst = max(worst, rms.max() / trms, mx.max() / tmx)
        if not ok:
            print("FAIL gridder", gv, dict(N=N, C=C, T=T, S=S, linear=linear, planar=planar), mx, rms)
            sys.exit(1)
    if rows.any():
        for dv in (0, 23):
            got = run_degridder(p, idg.SINCOS_FAST, dv)
            mx, rms = per_pol_errors(got[rows], ref_d[rows], 2)
            ok = (mx <= tmx).all() and (rms <= trms).all() and not got[~rows].any()
            worst = max(worst, rms.max() / trms, mx.max() / tmx)
            if not ok:
                print("FAIL degridder", dv, dict(N=N, C=C, T=T, S=S, linear=linear, planar=planar), mx, rms)
                sys.exit(1)
    print(f"case {case}: N={N} C={C} T<={T} S={S} linear={linear} planar={planar} gridder v{idg.resolve_variant(N, C, 0)} "
          f"degridder v{idg.resolve_variant(N, C, 0, gridder=False)} ok, worst error / tolerance so far {worst:.3f}")
print(f"{n_cases} cases passed; worst error / tolerance = {worst:.3f}")

# ---- the next rows: subgrid FFT, adder, splitter on random shapes and positions
import torch  # noqa: E402

dev = torch.device("cuda", 0)
worst_fft = worst_add = 0.0
for case in range(n_cases):
    N = int(rng.choice([8, 12, 16, 20, 24, 32, 40, 48, 64]))
    G = int(rng.choice([64, 100, 256, 1000]))
    S = int(rng.integers(1, 40))
    flags = int(rng.integers(0, 2))
    p = random_problem(int(rng.integers(1 << 30)), subgrid_size=N, nr_subgrids=S, grid_size=G, max_timesteps=1,
                       nr_channels=1)
    p.metadata["x"] = rng.integers(-N, G + N, S)        # overhang on every side, some fully outside
    p.metadata["y"] = rng.integers(-N, G + N, S)
    meta = torch.from_numpy(np.ascontiguousarray(p.metadata).view(np.int32).reshape(-1, 9)).to(dev)
    sg = torch.from_numpy(p.subgrids).to(dev)
    for direction in (1, -1):
        t = sg.clone()
        idg.subgrid_fft(S, N, t, direction)
        ref = o.subgrid_fft(p.subgrids, direction)
        err = float(np.abs(t.cpu().numpy() - ref).max() / np.abs(ref).max())
        worst_fft = max(worst_fft, err / (2e-6 * max(1.0, np.log2(N * N))))
    grid = torch.zeros((4, G, G), dtype=torch.complex64, device=dev)
    idg.adder(S, G, N, meta, sg, grid, flags=flags)
    ref = o.adder(p, flags=flags)
    scale = max(float(np.abs(ref).max()), 1e-30)
    worst_add = max(worst_add, float(np.abs(grid.cpu().numpy() - ref).max() / scale) / 1e-5)
    out = torch.full_like(sg, float("nan"))
    idg.splitter(S, G, N, meta, out, grid, flags=flags)
    if not np.array_equal(out.cpu().numpy(), o.splitter(p, grid.cpu().numpy(), flags=flags)):
        print("FAIL splitter", dict(N=N, G=G, S=S, flags=flags))
        sys.exit(1)
    if worst_fft > 1 or worst_add > 1:
        print("FAIL fft/adder", dict(N=N, G=G, S=S, flags=flags), worst_fft, worst_add)
        sys.exit(1)
print(f"{n_cases} next-row cases passed (splitter bit-exact); worst error / tolerance: fft {worst_fft:.3f}, "
      f"adder {worst_add:.3f}")
