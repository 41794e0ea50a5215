"""CPU prototype (numpy, float64) of next round's candidate for the general path: pixel-pair folding for
w != 0.  For a pixel q and its mirror image q' the phase splits into an odd part alpha (the u, v
terms) and an even part beta = n (w_offset - w k) that both share, so with gamma_v = w_offset - w_v k_c
    D[q], D[q'] = sum_m (i n)^m / m!  (E_m +- i F_m),   E_m = sum_v gamma_v^m vis_v cos(alpha_v),
                                                        F_m = sum_v gamma_v^m vis_v sin(alpha_v):
each order m is the folded GEMM of gridder_fold.cu with the B rows scaled by gamma_v^m (in the kernel:
by (gamma_v / gamma_max)^m, the power of gamma_max going to the epilogue).  This script measures the
truncation error against the direct sum on random ragged problems of the test suite's generator, per order.
Usage: python tools/fold_w_taylor.py [seed]"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import random_problem  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for shape in (dict(subgrid_size=32, nr_channels=16, max_timesteps=64, nr_subgrids=3),
              dict(subgrid_size=64, nr_channels=16, max_timesteps=32, nr_subgrids=2)):
    p = random_problem(seed, **shape)
    N, G = p.subgrid_size, p.grid_size
    x = np.arange(N)
    l1 = (x + 0.5 - N // 2) * p.image_size / N
    L, M = np.meshgrid(l1, l1)                      # L[y][x] = l(x), M[y][x] = l(y)
    tmp = L * L + M * M
    Nn = tmp / (1 + np.sqrt(1 - tmp))
    k = p.wavenumbers.astype(np.float64)
    worst = {}
    beta_max = 0.0
    for m_ in p.metadata:
        nt, t0 = int(m_["nr_timesteps"]), int(m_["time_offset"])
        if nt == 0:
            continue
        u, v, w = (p.uvw[t0:t0 + nt, i].astype(np.float64) for i in range(3))
        vis = p.visibilities[t0:t0 + nt, :, 0].astype(np.complex128)            # one polarisation is enough
        uo = (int(m_["x"]) + N // 2 - G // 2) * (2 * np.pi / p.image_size)
        vo = (int(m_["y"]) + N // 2 - G // 2) * (2 * np.pi / p.image_size)
        wo = 2 * np.pi * p.w_step * (int(m_["z"]) + 0.5)
        # direct sum (gridder_reference.cpp:61-78 in float64)
        idx = u[:, None, None] * L + v[:, None, None] * M + w[:, None, None] * Nn          # [t][y][x]
        off = uo * L + vo * M + wo * Nn
        phase = off[None, None] - idx[:, None] * k[None, :, None, None]                    # [t][c][y][x]
        D = (vis[:, :, None, None] * np.exp(1j * phase)).sum(axis=(0, 1))
        # folded: odd part alpha, even part beta = n * gamma_v
        alpha = (uo * L + vo * M)[None, None] - (u[:, None, None] * L + v[:, None, None] * M)[:, None] * k[None, :, None, None]
        gamma = wo - w[:, None] * k[None, :]                                               # [t][c]
        beta_max = max(beta_max, float(np.abs(gamma).max() * Nn.max()))
        for order in (1, 2, 3, 4, 5, 6, 8):
            Dq = np.zeros_like(D)
            for mm in range(order):
                gv = vis * gamma ** mm
                E = (gv[:, :, None, None] * np.cos(alpha)).sum(axis=(0, 1))
                F = (gv[:, :, None, None] * np.sin(alpha)).sum(axis=(0, 1))
                Dq += (1j * Nn) ** mm / math.factorial(mm) * (E + 1j * F)
            # only the first half of the pixels is computed this way in the kernel; the mirror half uses E - iF
            # with the same coefficients, which is the same expression evaluated at (-l, -m): check both halves
            err = np.abs(Dq - D).max() / np.abs(D).max()
            worst[order] = max(worst.get(order, 0.0), float(err))
    print(f"{shape}: max |beta| = {beta_max:.3f} rad; max|d|/max|D| by number of orders: " +
          ", ".join(f"{o}: {e:.1e}" for o, e in worst.items()))
