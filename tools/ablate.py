"""What bounds a tensor-core kernel?  Times ablation builds (tools/build_ablate.sh; its header lists the bits and the
kernels that honour them) at the default perf shape through each library's own
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
        3: "no MMAs, operand stores predicated off", 4: "no sum over the rows", 8: "no A rows", 12: "no A rows, no sum",
        13: "no A rows, no sum, no MMAs", 16: "no B operand",
        45: "no A rows, no sum, no MMAs, no P' (hand-offs only)", 64: "producers poll instead of sleeping"}
os.environ.setdefault("NR_ITERATIONS", "5")
os.environ["IDGB200_SINCOS"] = "0"
os.environ["IDGB200_VARIANT"] = "0"
os.environ.setdefault("IDGB200_ENERGY_SECONDS", "0")
rows = []
if len(sys.argv) > 1 and sys.argv[1] == "--ab":      # A/B builds of tools/build_ab.sh: every libidgb200_ab_<name>.so
    import glob
    kinds = (sys.argv[2],) if len(sys.argv) > 2 else ("gridder", "degridder")
    libs = [(os.path.basename(f)[len("libidgb200_ab_"):-3], f) for f in sorted(glob.glob(os.path.join(ROOT, "tools", "bin", "libidgb200_ab_*.so")))]
else:
    kinds = ("gridder", "degridder")
    libs = [(int(x), os.path.join(ROOT, "tools", "bin", f"libidgb200_ablate{int(x)}.so"))
            for x in (sys.argv[1] if len(sys.argv) > 1 else "0,1,4,8,12,13,16").split(",")]
for n, path in libs:
    lib = C.CDLL(path)
    row = dict(ablate=n, what=WHAT.get(n, ""))
    for kind in kinds:
        fn = getattr(lib, f"idgb200_p_run_{kind}")
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(Perf)]
        perf = Perf()
        rc = fn(C.byref(perf))
        assert rc == 0, (kind, n, rc)
        row[kind + "_ms"] = perf.seconds * 1e3
    rows.append(row)
    print(f"ablate {n} ({row['what']:40s}): " + "   ".join(f"{k} {row[k + '_ms']:7.2f} ms" for k in kinds), flush=True)
print(json.dumps(rows))
