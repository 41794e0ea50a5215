"""Why did round 1's driver bench (BENCH_r01: 11.65 ms/step) and its scaling run at N = 1 (SCALE_r01: 7.06 ms)
disagree on the same commit?  Both timed `--steps 20 --warmup 5` of the folded gridder (variant 29) as the very
first kernel work of a fresh process: 0.14 - 0.23 s, one CUDA-event pair around all 20 steps.  Two suspects:
  (A) the timed region starts while the GPU is still ramping up from its idle clocks,
  (B) variant 29's list ring is allocated lazily: every one of the first 32 launches calls cudaMalloc.
Each case below runs in a fresh process after the GPU has idled for --idle seconds and reports the old
protocol's number (mean of 20 steps after 5 warm-up steps) for three back-to-back repetitions, the per-step
times of the first repetition, and the SM clock NVML reports before the first launch and after the last:
  v24        no ring involved            -> (A) alone
  v29        ring filled lazily          -> (A) + (B): repetitions 1 and 2 allocate, 3 does not
  v30        the row-column kernel (preallocated, event-guarded scratch)
    python tools/bench_split_probe.py            # driver: runs the cases, prints JSON lines

The recorded run (profiles/r02_bench_split_probe.jsonl) is from commit 3d9a16f, the last one that still built
variant 29; at HEAD that kernel is gone (superseded by the row-column gridder) and its cases report an error line.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:
        return None


def case(variant: int, idle: float):
    sys.path.insert(0, ROOT)
    import torch

    import ska_sdp_idg_bench_b200 as idg
    dev = torch.device("cuda", 0)
    p = idg.init_problem_device(device=dev)
    S, tt, C, N = p["nr_subgrids"], p["total_timesteps"], p["nr_channels"], p["subgrid_size"]
    scal = (S, p["grid_size"], N, p["image_size"], 0.0, C, p["nr_stations"], tt)

    def step():
        idg.gridder(*scal, p["uvw"], p["wavenumbers"], p["visibilities"], p["spheroidal"], p["aterms"], p["metadata"],
                    p["subgrids"], sincos=idg.SINCOS_FAST, variant=variant)

    try:
        step()
    except idg.IdgError as e:
        print(json.dumps({"variant": variant, "error": str(e)}), flush=True)
        return
    torch.cuda.synchronize()
    time.sleep(idle)
    out = {"variant": variant, "idle_s": idle, "sm_mhz_before": sm_clock(), "repetitions": []}
    for rep in range(3):
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        t0 = time.perf_counter()
        ev[0].record()
        for i in range(20):
            step()
            ev[i + 1].record()
        host_enqueue_ms = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        per = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(20)]
        r = {"old_protocol_ms_per_step": ev[0].elapsed_time(ev[20]) / 20, "host_enqueue_ms_total": host_enqueue_ms,
             "sm_mhz_after": sm_clock()}
        if rep == 0:
            r["per_step_ms"] = per
        else:
            r["min_ms"], r["max_ms"] = min(per), max(per)
        out["repetitions"].append(r)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--case":
        case(int(sys.argv[2]), float(sys.argv[3]) if len(sys.argv) > 3 else 8.0)
    else:
        for v in (24, 29, 30, 29, 24):
            subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(v), "8"], check=False)
