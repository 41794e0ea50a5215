"""CPU prototype of the row-column (separable) formulation of the IDG gridder / degridder
(csrc/gridder_sep.cu, degridder_sep.cu).  Test infrastructure: numpy, float64
containers with the kernels' fp32 phase arithmetic and fp16 / e4m3 operand roundings emulated.

    phase(y, x, v) = phase_y(y, v) + phase_x(x, v) + gamma_v * r(y, x)
    phase_x = (u_off l_x + w_off nx_x) - (u l_x + w nx_x) k_c ,  nx_x = f(l_x^2),  f(s) = s / (1 + sqrt(1 - s))
    r = n(l, m) - f(l^2) - f(m^2) ~ l^2 m^2 / 4          (dropped: the separability condition)

so that  sum_v vis_v e^{i phase} = sum_v [Y_v(y) vis_v] X_v(x)  is a GEMM with K = visibilities.
Prints the error of the formulation against the oracle's float64 sum and the fp32 CPU code.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import oracle_lib  # noqa: E402


def f_n(s):
    return s / (1.0 + np.sqrt(1.0 - s))


def geometry(p):
    N = p.subgrid_size
    l = ((np.arange(N) + 0.5 - N // 2) * float(p.image_size) / N).astype(np.float32)
    nx = f_n(l.astype(np.float64) ** 2).astype(np.float32)
    return l, nx


def ctx(p, s):
    m = p.metadata[s]
    N, G = p.subgrid_size, p.grid_size
    scale = 2 * np.pi / float(np.float32(p.image_size))
    u_off = np.float32((int(m["x"]) + N // 2 - G // 2) * scale)
    v_off = np.float32((int(m["y"]) + N // 2 - G // 2) * scale)
    w_off = np.float32(2 * np.pi * float(np.float32(float(np.float32(p.w_step)) * (int(m["z"]) + 0.5))))
    t0 = int(m["baseline_offset"]) - int(p.metadata[0]["baseline_offset"]) + int(m["time_offset"])
    return u_off, v_off, w_off, t0, int(m["nr_timesteps"])


def half_phasors(p, s, sign):
    """X[t, c, x], Y[t, c, y] (complex128 of fp32-evaluated phases); sign = +1 gridder, -1 degridder."""
    l, nx = geometry(p)
    u_off, v_off, w_off, t0, nt = ctx(p, s)
    uvw = p.uvw[t0:t0 + nt]
    k = p.wavenumbers.astype(np.float32)
    f32 = np.float32

    def half(a_off, a):
        off = (f32(a_off) * l + f32(w_off) * nx).astype(f32)                      # [N]
        idx = (a[:, None] * l[None, :] + uvw[:, 2:3] * nx[None, :]).astype(f32)    # [t, N]
        ph = (off[None, None, :] - idx[:, None, :] * k[None, :, None]).astype(f32)
        return np.exp(1j * sign * ph.astype(np.float64))

    return half(u_off, uvw[:, 0]), half(v_off, uvw[:, 1]), t0, nt


def r16(z):
    """round the real and imaginary parts to fp16 (after the kernels' power-of-two scaling: relative)"""
    return z.real.astype(np.float16).astype(np.float64) + 1j * z.imag.astype(np.float16).astype(np.float64)


def scaled16(z, split):
    a = np.abs(np.concatenate([z.real.ravel(), z.imag.ravel()])).max() if z.size else 0.0
    sc = 2.0 ** (13 - np.floor(np.log2(a))) if a > 0 else 1.0
    hi = r16(z * sc)
    if split:
        hi = hi + r16(z * sc - hi)
    return hi / sc


def e4m3(v):
    """round to e4m3 (4 significant bits, normal from 2^-6, subnormal step 2^-9, saturating at 448), half to even"""
    v = np.asarray(v, np.float64)
    a = np.abs(v)
    e = np.maximum(np.floor(np.log2(np.where(a > 0, a, 1.0))), -6.0)
    step = 2.0 ** (e - 3)
    return np.sign(v) * np.minimum(np.round(a / step) * step, 448.0)


def c8(z):
    return e4m3(z.real) + 1j * e4m3(z.imag)


def aterm_g(p, s, pix):
    m = p.metadata[s]
    a1 = p.aterms[m["aterm_index"], m["station1"]].astype(np.complex128).reshape(p.subgrid_size, p.subgrid_size, 2, 2)
    a2 = p.aterms[m["aterm_index"], m["station2"]].astype(np.complex128).reshape(p.subgrid_size, p.subgrid_size, 2, 2)
    P = pix.transpose(1, 2, 0).reshape(p.subgrid_size, p.subgrid_size, 2, 2)
    out = np.conj(a1.transpose(0, 1, 3, 2)) @ P @ a2
    return (out * p.spheroidal[:, :, None, None]).reshape(p.subgrid_size, p.subgrid_size, 4).transpose(2, 0, 1)


def aterm_d(p, s):
    m = p.metadata[s]
    N = p.subgrid_size
    a1 = p.aterms[m["aterm_index"], m["station1"]].astype(np.complex128).reshape(N, N, 2, 2)
    a2 = p.aterms[m["aterm_index"], m["station2"]].astype(np.complex128).reshape(N, N, 2, 2)
    P = (p.subgrids[s].astype(np.complex128) * p.spheroidal[None]).transpose(1, 2, 0).reshape(N, N, 2, 2)
    out = a1 @ P @ np.conj(a2.transpose(0, 1, 3, 2))
    return out.reshape(N, N, 4)   # [y][x][pol]


def gridder_sep(p, emulate=True):
    out = np.zeros((p.nr_subgrids, 4, p.subgrid_size, p.subgrid_size), np.complex128)
    for s in range(p.nr_subgrids):
        X, Y, t0, nt = half_phasors(p, s, +1)
        vis = p.visibilities[t0:t0 + nt].astype(np.complex128)          # [t, c, pol]
        A = Y[:, :, :, None] * vis[:, :, None, :]                        # [t, c, y, pol]  (fp32 products)
        if emulate:
            A = scaled16(A.astype(np.complex64).astype(np.complex128), split=False)   # operand A: fp16
            X = r16(X) + r16(X - r16(X))                                 # operand B: fp16 hi + lo
        pix = np.einsum("tcyp,tcx->pyx", A, X)
        out[s] = aterm_g(p, s, pix)
    return out


def degridder_sep(p, emulate=True):
    """emulate: False = exact operands, "f16x3" = fp16 hi + lo on both operands and three products (hi hi, lo hi, hi lo:
    round 2's first form), "f16x1" = one fp16 product, True / "fp8" = what degridder_sep.cu does: hi hi in fp16 and the two
    cross products as one e4m3 MMA, A scaled by 2^8 and B to [2^7, 2^8) so that both parts sit in e4m3's range"""
    out = np.zeros((p.total_timesteps, p.nr_channels, 4), np.complex128)
    for s in range(p.nr_subgrids):
        X, Y, t0, nt = half_phasors(p, s, -1)
        P = aterm_d(p, s)                                                # [y][x][pol]
        if emulate in (True, "fp8"):
            P = P.astype(np.complex64).astype(np.complex128)
            a = np.abs(np.concatenate([P.real.ravel(), P.imag.ravel()])).max() if P.size else 0.0
            sc = 2.0 ** (7 - np.floor(np.log2(a))) if a > 0 else 1.0
            B, A = P * sc, X * 256.0
            Bh, Ah = r16(B), r16(A)
            Q = (np.einsum("tcx,yxp->tcyp", Ah, Bh) + np.einsum("tcx,yxp->tcyp", c8(A - Ah), c8(Bh)) +
                 np.einsum("tcx,yxp->tcyp", c8(Ah), c8(B - Bh))) / (sc * 256.0)
        else:
            if emulate == "f16x3":
                P = scaled16(P.astype(np.complex64).astype(np.complex128), split=True)
                X = r16(X) + r16(X - r16(X))
            elif emulate == "f16x1":
                P = scaled16(P.astype(np.complex64).astype(np.complex128), split=False)
                X = r16(X)
            Q = np.einsum("tcx,yxp->tcyp", X, P)                         # tensor core, fp32 accumulate
        if emulate:
            Q = Q.astype(np.complex64).astype(np.complex128)
        out[t0:t0 + nt] = np.einsum("tcy,tcyp->tcp", Y, Q)
    return out


def report(name, got, ref32, ref64):
    d64 = np.abs(got - ref64).max() / np.abs(ref64).max()
    c64 = np.abs(ref32 - ref64).max() / np.abs(ref64).max()
    d32 = np.abs(got - ref32).max() / np.abs(ref32).max()
    rms = np.sqrt((np.abs(got - ref32) ** 2).sum() / (np.abs(ref32) ** 2).sum())
    print(f"{name:34s} vs f64 {d64:.2e} (cpu-f32 vs f64 {c64:.2e})   vs cpu-f32 max-rel {d32:.2e} rel-rms {rms:.2e}")


def main():
    o = oracle_lib.oracle()
    cases = [("config 1", o.make_problem())]
    for seed in range(4):
        cases.append((f"ragged seed {seed} (w, 0.02)", oracle_lib.random_problem(seed, nr_subgrids=3, subgrid_size=32,
                                                                                  nr_channels=8, max_timesteps=40)))
    cases.append(("w ~ N(0, 256), image 0.1", oracle_lib.random_problem(7, subgrid_size=32, nr_channels=8, max_timesteps=40,
                                                                        image_size=0.1)))
    for name, p in cases:
        g32, g64 = o.gridder(p), o.gridder_f64(p)
        d32, d64 = o.degridder(p), o.degridder_f64(p)
        cov = np.zeros(p.total_timesteps, bool)      # degridder: only rows some subgrid covers
        for s in range(p.nr_subgrids):
            _, _, _, t0, nt = ctx(p, s)
            cov[t0:t0 + nt] = True
        for emu in (False, True):
            tag = "fp16 operands" if emu else "exact operands"
            report(f"{name} gridder {tag}", gridder_sep(p, emu), g32, g64)
        for emu, tag in ((False, "exact operands"), ("f16x1", "one fp16 product"), ("f16x3", "three fp16 products"),
                         ("fp8", "fp16 + e4m3 cross products")):
            report(f"{name} degridder {tag}", degridder_sep(p, emu)[cov], d32[cov], d64[cov])


if __name__ == "__main__":
    main()
