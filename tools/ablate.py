"""What bounds the two default tensor-core kernels?  Times ablation builds (tools/build_ablate.sh) of
gridder_tc.cu / degridder_tc8.cu at the default perf shape through each library's own
idgb200_p_run_* (CUDA events, NR_ITERATIONS launches): the same kernel without its MMAs, with the
operand stores predicated off at run time (they still issue, nothing reaches shared memory), and both.  Results of an ablated kernel are
garbage by construction; only the time means something.
Usage: python tools/ablate.py [0,1,2,3]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ska_sdp_idg_bench_b200._lib import Perf  # noqa: E402  (struct layout only)

WHAT = {0: "full kernel (untuned build)", 1: "no MMAs", 2: "operand stores predicated off",
        3: "no MMAs, operand stores predicated off"}
os.environ.setdefault("NR_ITERATIONS", "5")
os.environ["IDGB200_SINCOS"] = "0"
os.environ["IDGB200_VARIANT"] = "0"
os.environ.setdefault("IDGB200_ENERGY_SECONDS", "0")
rows = []
for n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2,3").split(",")]:
    lib = C.CDLL(os.path.join(ROOT, "tools", "bin", f"libidgb200_ablate{n}.so"))
    row = dict(ablate=n, what=WHAT.get(n, ""))
    for kind in ("gridder", "degridder"):
        fn = getattr(lib, f"idgb200_p_run_{kind}")
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(Perf)]
        perf = Perf()
        rc = fn(C.byref(perf))
        assert rc == 0, (kind, n, rc)
        row[kind + "_ms"] = perf.seconds * 1e3
    rows.append(row)
    print(f"ablate {n} ({row['what']:40s}): gridder {row['gridder_ms']:7.2f} ms   degridder {row['degridder_ms']:7.2f} ms", flush=True)
print(json.dumps(rows))
