"""Timeline of the pipelined degridder's hand-offs (degridder_sep.cu, DP_TRACE build):
    KERNEL=degridder_sep tools/build_ab.sh trace:"-DDP_TRACE"  &&  python tools/pipe_trace.py [name]
CTA 0 records clock64() at every hand-off of the 64 tiles of four steady-state subgrids, per warp; this prints, per
tile, the events relative to the first one (SM clocks) and the averages: who waits for whom."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ska_sdp_idg_bench_b200._lib import Perf  # noqa: E402

os.environ["NR_ITERATIONS"] = "1"
os.environ["IDGB200_SINCOS"] = "0"
os.environ["IDGB200_VARIANT"] = "32"
os.environ.setdefault("IDGB200_ENERGY_SECONDS", "0")
name = sys.argv[1] if len(sys.argv) > 1 else "trace"
lib = C.CDLL(os.path.join(ROOT, "tools", "bin", f"libidgb200_ab_{name}.so"))
fn = lib.idgb200_p_run_degridder
fn.restype = C.c_int
fn.argtypes = [C.POINTER(Perf)]
perf = Perf()
assert fn(C.byref(perf)) == 0
EV, T = 56, 64
buf = np.zeros(EV * T, np.int64)
assert lib.idgb200_dp_trace_read(buf.ctypes.data_as(C.POINTER(C.c_longlong))) == 0
t = buf.reshape(EV, T).astype(np.float64)
t[t == 0] = np.nan                      # a consumer group only sees the tiles of its accumulator
P0, P1 = t[0:16:2], t[1:16:2]            # producers: A buffer free, a_full arrived       [warp][tile]
C0, C1 = t[16:48:2], t[17:48:2]          # consumers: awake after mma_done, d_empty arrived
I0, I1 = t[48], t[49]                    # issuer: waits done, committed
t0 = I0[0]
print(f"kernel {perf.seconds * 1e3:.3f} ms; clocks relative to the issuer's first traced tile")
print("tile |  P start min/max   P end min/max |  I start    I end |  C wake min/max  C d_empty min/max | d_empty(it-2) a_full(it) -> I start")
for i in range(T):
    de = np.nanmax(C1[:, i - 2]) - t0 if i >= 2 else float("nan")
    print(f"{i:4d} | {P0[:, i].min() - t0:8.0f} {P0[:, i].max() - t0:8.0f} {P1[:, i].min() - t0:8.0f} {P1[:, i].max() - t0:8.0f} |"
          f" {I0[i] - t0:8.0f} {I1[i] - t0:8.0f} | {np.nanmin(C0[:, i]) - t0:8.0f} {np.nanmax(C0[:, i]) - t0:8.0f} {np.nanmin(C1[:, i]) - t0:8.0f} {np.nanmax(C1[:, i]) - t0:8.0f} |"
          f" {de:8.0f} {P1[:, i].max() - t0:8.0f} {I0[i] - t0:8.0f}")
print("setup start / end per subgrid:", [(float(t[50, 16 * k] - t0), float(t[51, 16 * k] - t0)) for k in range(4)])
for i in (20, 21, 22, 23):
    print(f"tile {i}: consumer warps awake after commit {(C0[:, i] - I1[i]).round(0)}")
    print(f"         d_empty after awake               {(C1[:, i] - C0[:, i]).round(0)}")
    print(f"         producers: start after commit(it-2) {(P0[:, i] - I1[i - 2]).round(0)}, busy {(P1[:, i] - P0[:, i]).round(0)}")
sl = slice(2, None)
print(f"\nmean clocks per tile (issuer commit to commit): {float(np.mean(np.diff(I1))):.0f}")
print(f"producer warp: A buffer free -> a_full, mean per warp  {np.mean(P1[:, sl] - P0[:, sl], axis=1).round(0)}")
print(f"consumer warp: awake -> d_empty, mean per warp         {np.nanmean(C1[:, sl] - C0[:, sl], axis=1).round(0)}")
print(f"issuer: waits done -> committed                        {float(np.mean(I1 - I0)):7.0f}")
print(f"committed -> first / last consumer awake               {float(np.mean(np.nanmin(C0, 0) - I1)):7.0f} / {float(np.mean(np.nanmax(C0, 0) - I1)):7.0f}")
print(f"mma_done(it-2) seen by first / last producer after commit(it-2)   {float(np.mean(P0[:, 2:].min(0) - I1[:-2])):7.0f} / {float(np.mean(P0[:, 2:].max(0) - I1[:-2])):7.0f}")
print(f"issuer start after the later of a_full(it), d_empty(it-2)       {float(np.mean(I0[2:] - np.maximum(P1[:, 2:].max(0), np.nanmax(C1[:, :-2], 0)))):7.0f}")
print(f"  of which a_full(it) was the later one in {int(np.sum(P1[:, 2:].max(0) > np.nanmax(C1[:, :-2], 0)))} of {T - 2} tiles")
