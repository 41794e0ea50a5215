"""Times every kernel variant / sincos mode at the default perf shape through the
library's own p_run_* (CUDA-event timing, env-driven like the reference's perf mode).
Usage: python tools/sweep.py [gridder variants] [degridder variants], e.g. 0,1,2 0,3"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ska_sdp_idg_bench_b200 as idg  # noqa: E402

gv = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2,3,4").split(",") if x != ""]
dv = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2,3").split(",") if x != ""]
modes = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0").split(",")]
os.environ.setdefault("NR_ITERATIONS", "5")
PEAK = 148 * 128 * 2 * 1.965e9 * 1e-12
rows = []
for mode in modes:
    os.environ["IDGB200_SINCOS"] = str(mode)
    for kind, variants, fn in (("gridder", gv, idg.p_run_gridder), ("degridder", dv, idg.p_run_degridder)):
        for v in variants:
            os.environ["IDGB200_VARIANT"] = str(v)
            r = fn()
            rows.append(dict(kernel=kind, variant=v, sincos=mode, ms=r["seconds"] * 1e3,
                             mvis_per_s=r["mvis_per_s"], tflops=r["tflops_per_s"],
                             fp32_frac=r["tflops_per_s"] / PEAK))
print("kernel     variant sincos      ms    MVis/s  TFLOP/s  fp32_frac")
for r in rows:
    print(f"{r['kernel']:10s} {r['variant']:7d} {r['sincos']:6d} {r['ms']:7.2f} {r['mvis_per_s']:9.1f} "
          f"{r['tflops']:8.2f} {r['fp32_frac']:9.3f}")
print(json.dumps(rows))
