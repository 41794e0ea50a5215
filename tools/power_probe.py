"""Board power and SM clock of the row-column kernels under sustained load, per ablation build (tools/build_ablate.sh):
NR_ITERATIONS launches (default 600, ~3 s) through each library's idgb200_p_run_*, nvidia-smi sampled every 100 ms
meanwhile.  Energy per launch = median power x time per launch: both kernels sit at the 1 kW cap, so what a role costs in
JOULES is what it costs in time.  Usage: python tools/power_probe.py [0,1,4,8] [gridder|degridder ...]"""
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ska_sdp_idg_bench_b200._lib import Perf  # noqa: E402


class Sampler:
    def __init__(self):
        self.lines = []
        self.proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                      "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def between(self, i0, i1):
        v = []
        for ln in self.lines[i0 + 2 * (i1 - i0) // 3:i1]:     # the last third of the region: nvidia-smi's power is a ~1 s average
            try:
                c, p = (float(x) for x in ln.split(","))
                v.append((c, p))
            except ValueError:
                pass
        if not v:
            return None, None
        return statistics.median(c for c, _ in v), statistics.median(p for _, p in v)


def main():
    os.environ.setdefault("NR_ITERATIONS", "600")
    os.environ["IDGB200_SINCOS"] = "0"
    os.environ["IDGB200_VARIANT"] = "0"
    os.environ.setdefault("IDGB200_ENERGY_SECONDS", "0")
    which = (sys.argv[1] if len(sys.argv) > 1 else "0,1,4,8").split(",")
    kinds = sys.argv[2:] or ["degridder"]
    smp = Sampler()
    time.sleep(0.5)
    rows = []
    for n in which:
        path = os.path.join(ROOT, "tools", "bin", f"libidgb200_ablate{int(n)}.so") if n.isdigit() else n
        lib = C.CDLL(path)
        for kind in kinds:
            fn = getattr(lib, f"idgb200_p_run_{kind}")
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(Perf)]
            perf = Perf()
            i0 = len(smp.lines)
            assert fn(C.byref(perf)) == 0
            i1 = len(smp.lines)
            mhz, watt = smp.between(i0, i1)
            row = dict(lib=n, kind=kind, ms=perf.seconds * 1e3, sm_mhz=mhz, power_w=watt,
                       joule_per_launch=(watt * perf.seconds if watt else None))
            rows.append(row)
            print(json.dumps(row), flush=True)
            time.sleep(1.0)
    smp.proc.terminate()


if __name__ == "__main__":
    main()
