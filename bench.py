#!/usr/bin/env python
"""Benchmark of the IDG gridder / degridder hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]/[2]): the reference's default performance shape
(app/CUDA/util.cpp:177-183): 50 stations x 20 timeslots -> 24,500 subgrids of 32x32
pixels x 4 polarisations, 128 timesteps x 16 channels each = 50.176 MVis per step,
synthetic inputs from the reference's generators (app/common/init.cpp) evaluated on
the device.  A step is one gridder pass over all subgrids; the degridder is timed the
same way right after and reported under "degridder".  With N GPUs every rank runs the
full per-GPU workload on its own shard (weak scaling, no collective on the data path).

One JSON line on stdout (rank 0).  `value` = whole-job gridder MVis/s with inputs
resident in HBM; `e2e` = the same metric through the host-pointer C ABI
(idgb200_c_run_gridder_ex) with pinned host buffers, copies inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_SHAPE = dict(nr_stations=50, nr_timeslots=20, nr_timesteps=128, nr_channels=16,
                     subgrid_size=32, grid_size=1024, image_size=0.01)
SM_FP32_LANES = 128   # FP32 FMA lanes per SM (B200)
SM_XU_LANES = 16      # MUFU lanes per SM
CONSTANTS_PATH = os.path.join(ROOT, "profiles", "roofline_constants.json")   # written by tools/ncu_summary.py --json


# ----------------------------------------------------------------- shared helpers
def shape_counts(shape: dict) -> dict:
    nbl = shape["nr_stations"] * (shape["nr_stations"] - 1) // 2
    S = nbl * shape["nr_timeslots"]
    tt = S * shape["nr_timesteps"]
    return dict(nr_subgrids=S, total_timesteps=tt, mvis=1e-6 * tt * shape["nr_channels"])


def rank_shape(overrides: dict, rank: int, world: int) -> dict:
    """Per-rank workload under weak scaling: the full shape on every rank, with a
    rank-specific seed so the shards differ like different baselines would."""
    shape = dict(DEFAULT_SHAPE)
    shape.update(overrides)
    shape.update(shape_counts(shape))
    shape["seed"] = rank
    shape["rank"], shape["world"] = rank, world
    return shape


def reduce_max_time(seconds: float, device) -> float:
    """Max over ranks of a locally measured duration."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self) -> int:
        """index of the next sample: brackets a timed region"""
        return len(self.lines)

    def between(self, i0: int, i1: int):
        """SM clock (median / min / max) and board power (median) of the under-load samples taken between two marks
        (None without samples): the clock one kernel's region ran at - the power cap bites the tensor-core kernels
        harder than the FP32 ones"""
        sm, pw = [], []
        for ln in self.lines[i0:max(i1, i0 + 1) + 1]:      # + the sample in flight when the region ended
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[1])); pw.append(float(f[3]))
            except (ValueError, IndexError):
                continue
        if not sm:
            return None
        keep = [i for i, x in enumerate(sm) if x >= 0.5 * max(sm)]
        hi = [sm[i] for i in keep]
        return {"sm_mhz": statistics.median(hi), "sm_mhz_min": min(hi), "sm_mhz_max": max(hi),
                "power_w": statistics.median(pw[i] for i in keep), "samples": len(hi)}

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [x for x in sm if x >= 0.5 * max(sm)]  # under-load samples
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# -------------------------------------------------------------------- CPU arm
def load_cpu_checker():
    """(lib, kind): the reference's own CPU code when oracle/_ref was built, else the
    oracle restatement ("port").  Only used as the reported CPU baseline / checker."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    ref = oracle_lib.reference()
    if ref is not None:
        return ref, "reference", oracle_lib
    return oracle_lib.oracle(), "port", oracle_lib


def cpu_problem(oracle_lib, shape: dict, nr_subgrids: int):
    """The first nr_subgrids subgrids of the workload, made by the checker's own
    restatement of app/common/init.cpp (host rand()), for the CPU arm."""
    lib, _, _ = load_cpu_checker()
    nbl_needed = max(1, -(-nr_subgrids // shape["nr_timeslots"]))
    # smallest station count whose baseline list covers the sample
    st = 2
    while st * (st - 1) // 2 < nbl_needed:
        st += 1
    p = lib.make_problem(nr_stations=st, nr_timeslots=shape["nr_timeslots"],
                         nr_timesteps=shape["nr_timesteps"], nr_channels=shape["nr_channels"],
                         subgrid_size=shape["subgrid_size"], grid_size=shape["grid_size"])
    S = min(nr_subgrids, p.nr_subgrids)
    T = shape["nr_timesteps"]
    return oracle_lib.Problem(
        grid_size=p.grid_size, subgrid_size=p.subgrid_size, image_size=p.image_size,
        w_step=p.w_step, nr_channels=p.nr_channels, nr_stations=p.nr_stations,
        uvw=np.ascontiguousarray(p.uvw[:S * T]), wavenumbers=p.wavenumbers,
        visibilities=np.ascontiguousarray(p.visibilities[:S * T]), spheroidal=p.spheroidal,
        aterms=p.aterms, metadata=np.ascontiguousarray(p.metadata[:S]),
        subgrids=np.ascontiguousarray(p.subgrids[:S]))


def reference_gpu_kernels() -> dict | None:
    """The reference's own fastest CUDA kernels (app/CUDA/kernels/gridder_v8.cu, degridder_v6.cu with
    its own host runner and main, compiled in place for sm_100a by oracle/Makefile `refcuda`) in their
    performance mode on this box: the same-box GPU comparator of BASELINE.md 1.  Not the `--impl
    reference` arm (that is the CPU path); reported beside it.  None if the binaries are absent."""
    import re
    import subprocess
    import tempfile
    out = {}
    for key, exe in (("gridder", "refcuda-gridder_v8"), ("degridder", "refcuda-degridder_v6")):
        path = os.path.join(ROOT, "oracle", "_ref", exe)
        if not os.path.exists(path):
            return None
        with tempfile.TemporaryDirectory() as tmp:
            env = dict(os.environ, OUTPUT_PATH=tmp)
            for k in ("NR_STATIONS", "NR_TIMESLOTS", "NR_CHANNELS", "SUBGRID_SIZE", "GRID_SIZE",
                      "NR_TIMESTEPS_SUBGRID", "NR_ITERATIONS", "NR_WARM_UP_RUNS"):
                env.pop(k, None)     # the reference's default perf shape = bench config 2
            try:
                txt = subprocess.run([path], env=env, capture_output=True, text=True, timeout=300).stdout
            except (OSError, subprocess.TimeoutExpired):
                return None
        m = re.search(r"(\w+):\s+([0-9.]+) ms,\s+([0-9.]+) GFLOP/s,.*?([0-9.]+) MVis/s", txt)
        if not m:
            return None
        out[key] = {"kernel": m.group(1), "ms_per_launch": float(m.group(2)),
                    "tflops": float(m.group(3)) * 1e-3, "mvis_per_s": float(m.group(4))}
    out["what"] = ("the reference's unmodified gridder_v8 / degridder_v6 (sources compiled where they lie, "
                   "sm_100a) through its own p_run_* at its default shape (= this workload), 5 launches")
    return out


def host_threads() -> int:
    """Cores this process may use.  torchrun exports OMP_NUM_THREADS=1, which is not what
    "all the host threads it can use" means, so the CPU arm sets the count explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


_REAL_STDOUT = None


def quiet_stdout() -> None:
    """Anything a library prints to stdout (NCCL's version banner ...) goes to stderr, so
    that the JSON line is the only thing on stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj: dict) -> None:
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def time_cpu(lib, prob, which: str) -> float:
    t0 = time.perf_counter()
    (lib.gridder if which == "gridder" else lib.degridder)(prob)
    return time.perf_counter() - t0


def run_reference_arm(args) -> None:
    """--impl reference: the reference's CPU gridder (oracle/_ref when it compiled,
    else the port) on all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib, kind, oracle_lib = load_cpu_checker()
    cores = host_threads()
    lib.set_threads(cores)
    shape = rank_shape({}, 0, 1)
    # calibrate, then size one step to ~3 s
    probe = cpu_problem(oracle_lib, shape, max(cores, 4))
    dt = time_cpu(lib, probe, "gridder")
    per_subgrid = dt / probe.nr_subgrids
    n = int(min(shape["nr_subgrids"], max(cores, 3.0 / per_subgrid)))
    prob = cpu_problem(oracle_lib, shape, n)
    n = prob.nr_subgrids
    mvis_step = 1e-6 * n * shape["nr_timesteps"] * shape["nr_channels"]
    for _ in range(args.warmup):
        time_cpu(lib, prob, "gridder")
    t = [time_cpu(lib, prob, "gridder") for _ in range(args.steps)]
    td = [time_cpu(lib, prob, "degridder") for _ in range(max(1, min(args.steps, 3)))]
    sec = sum(t) / len(t)
    value = mvis_step / sec
    sample = (f"first {n} of {shape['nr_subgrids']} subgrids per step, same shape "
              f"(N=32, 128 timesteps x 16 channels), all {cores} host threads (OpenMP)")
    out = {
        "impl": "reference", "metric": "gridder_mvis_per_s", "value": value, "unit": "MVis/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(shape, args.gpus),
        "cpu_baseline": {"value": value, "unit": "MVis/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "degridder": {"value": mvis_step / (sum(td) / len(td)), "unit": "MVis/s"},
        "e2e": {"value": value, "unit": "MVis/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def workload_config(shape: dict, n_gpus: int) -> dict:
    return {
        "workload": "ska-sdp-idg-bench default perf shape (app/CUDA/util.cpp:177-183): "
                    "50 stations x 20 timeslots = 24500 subgrids/GPU, subgrid 32, 4 pols, "
                    "128 timesteps x 16 channels, grid 1024, image_size 0.01, w_step 0",
        "subgrids_per_gpu": shape["nr_subgrids"], "mvis_per_step_per_gpu": shape["mvis"],
        "sharding": f"subgrid list, {n_gpus} independent shard(s), no collective",
        "l2": "inputs per step (1.6 GB visibilities + 0.8 GB subgrids) exceed the 126 MB L2; no flush",
    }


# --------------------------------------------------------------------- GPU arm
def load_constants() -> dict:
    """Per-kernel figures from the committed ncu captures (tools/ncu_summary.py --json): DRAM bytes per subgrid,
    pipe utilisations.  The roofline fraction itself needs none of them (analytic flops / live time / measured peak)."""
    try:
        return json.load(open(CONSTANTS_PATH))
    except Exception:
        return {}


def stage_flops(shape: dict, kind: str) -> dict:
    """tcgen05 flops of one launch of the row-column kernels (2 M N K per MMA), counted from the shape.
    gridder_sep.cu: one M=128 (4 pols x 32 rows) x N=4*XT x K=16 MMA per (tile, timestep, 8-channel block); of the
    N columns half are the fp16 lo part of the column phasors.  degridder_sep.cu: per tile of 128 visibilities
    and K step of 8 columns one M=128 x N=8*Ny x K=16 fp16 MMA (hi hi) and one K=32 e4m3 MMA (lo hi + hi lo), which
    holds the tensor pipe as long as the fp16 one.
    `useful` counts the contraction the formulation needs once in the operands' nominal precision (no lo parts,
    no padding of rows / channels / visibilities); `executed` everything issued, in fp16-equivalent flops (an e4m3
    MMA counted at the pipe time it takes = half its nominal flops: what the fp16 peak can be compared with)."""
    N, C, S, T = shape["subgrid_size"], shape["nr_channels"], shape["nr_subgrids"], shape["nr_timesteps"]
    vis = S * T * C
    useful = 2.0 * vis * (4 * N) * (2 * N) * 2       # complex MAC of 4N rows x N columns (gridder) = 4N x N x 8 flop
    if kind == "gridder":
        ytiles, ncb = -(-N // 32), -(-C // 8)
        executed = 0.0
        for x0 in range(0, N, 64):
            xt = min(64, N - x0)
            executed += 2.0 * 128 * (4 * xt) * 16 * ytiles * S * T * ncb
    else:
        kc = -(-(N // 4) // 2)                         # K steps of 16 (8 columns)
        tiles = S * -(-(T * C) // 128)
        executed = 2.0 * 128 * (8 * N) * 16 * 2 * kc * tiles
    return {"useful": useful, "executed": executed}


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import ska_sdp_idg_bench_b200 as idg

    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    strong = args.scaling == "strong"
    shape = rank_shape({}, rank, world)
    sincos = {"fast": idg.SINCOS_FAST, "reduced": idg.SINCOS_REDUCED,
              "accurate": idg.SINCOS_ACCURATE}[args.sincos]
    # weak scaling: every rank the whole config-2 observation with its own seed.  strong scaling: ONE observation
    # (seed 0 on every rank) cut by shard.partition_subgrids; a rank keeps its contiguous range of the subgrid
    # list with the uvw / visibility rows it covers (shard.shard_metadata), nothing is exchanged.
    prob = idg.init_problem_device(
        nr_stations=shape["nr_stations"], nr_timeslots=shape["nr_timeslots"],
        nr_timesteps=shape["nr_timesteps"], nr_channels=shape["nr_channels"],
        subgrid_size=shape["subgrid_size"], grid_size=shape["grid_size"],
        image_size=shape["image_size"], seed=0 if strong else shape["seed"], device=dev)
    S_total = prob["nr_subgrids"]
    if strong and world > 1:
        meta_np = np.ascontiguousarray(prob["metadata"].cpu().numpy()).view(idg.METADATA_DTYPE).reshape(-1)
        s0, s1 = idg.partition_subgrids(meta_np["nr_timesteps"], world)[rank]
        m_loc, t0, t1 = idg.shard_metadata(meta_np, s0, s1)
        prob["metadata"] = torch.from_numpy(m_loc.view(np.int32).reshape(-1, 9).copy()).to(dev)
        prob["uvw"] = prob["uvw"][t0:t1].contiguous()
        prob["visibilities"] = prob["visibilities"][t0:t1].contiguous()
        prob["subgrids"] = prob["subgrids"][s0:s1].contiguous()
        prob["nr_subgrids"], prob["total_timesteps"] = s1 - s0, t1 - t0
    S, tt, C_, N = prob["nr_subgrids"], prob["total_timesteps"], prob["nr_channels"], prob["subgrid_size"]
    my_mvis = 1e-6 * tt * C_
    scal = (S, prob["grid_size"], N, prob["image_size"], 0.0, C_, prob["nr_stations"], tt)
    sub_in = prob["subgrids"].clone()  # degridder input (the gridder overwrites subgrids)
    vis_out = torch.empty_like(prob["visibilities"])
    # the same observation off the plane: w ~ N(0, 256 m) on every timestep (the reference's generator writes
    # w = 0, init.cpp:4-25; a real observation does not)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    uvw_w = prob["uvw"].clone()
    uvw_w[:, 2] = torch.randn(uvw_w.shape[0], device=dev, generator=gen) * 256.0

    def gridder_step(variant, uvw=None):
        u = prob["uvw"] if uvw is None else uvw
        return lambda: idg.gridder(*scal, u, prob["wavenumbers"], prob["visibilities"], prob["spheroidal"],
                                   prob["aterms"], prob["metadata"], prob["subgrids"], sincos=sincos, variant=variant)

    def degridder_step(variant, uvw=None):
        u = prob["uvw"] if uvw is None else uvw
        return lambda: idg.degridder(*scal, u, prob["wavenumbers"], vis_out, prob["spheroidal"], prob["aterms"],
                                     prob["metadata"], sub_in, sincos=sincos, variant=variant)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x: float) -> float:
        return reduce_max_time(x, dev)

    sampler = None      # the clock sampler of the device-resident timings (set below; timed() brackets its region in it)

    def timed(step, steps, warmup, min_seconds=None, max_rounds=64, adaptive=True):
        """Warm-up: at least `warmup` steps, then on until two consecutive steps agree to 2 % (allocations,
        module loading and clock ramps stay out of the timed region).  Timed: rounds of EXACTLY `steps` steps,
        back to back, every step between its own pair of CUDA events on the launching (torch current) stream,
        repeated until the region is >= min_seconds; barrier + synchronize on both sides.  Seconds per step =
        whole region / steps in it, max over ranks (so a host-side stall that starves the GPU shows up - compare
        with the per-step median, which does not see it).  adaptive=False: exactly `warmup` warm-up steps - for steps
        that synchronise the ranks with each other, which every rank must call the same number of times."""
        if min_seconds is None:
            min_seconds = args.min_seconds
        last, n_warm = None, 0
        e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        while True:
            e_a.record(); step(); e_b.record(); e_b.synchronize()
            ms = e_a.elapsed_time(e_b)
            n_warm += 1
            if n_warm >= warmup and (not adaptive or (last is not None and abs(ms - last) <= 0.02 * last)
                                     or n_warm >= warmup + 40):
                break
            last = ms
        rounds = int(min(max_rounds, max(1, -(-min_seconds // max(steps * ms * 1e-3, 1e-6)))))
        rounds = int(all_max(float(rounds)))
        n = rounds * steps
        barrier()
        c0 = sampler.mark() if sampler is not None else 0
        l0 = idg.launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        ev[0].record()
        for i in range(n):
            step()
            ev[i + 1].record()
        barrier()
        c1 = sampler.mark() if sampler is not None else 0
        per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
        sec = ev[0].elapsed_time(ev[n]) * 1e-3
        return {"sec_per_step": all_max(sec) / n, "rounds": rounds, "timed_steps": n, "timed_region_s": sec, "clock_marks": (c0, c1),
                "warmup_steps": n_warm, "launches_per_step": (idg.launch_count() - l0) / n,
                "step_ms": {"min": per[0], "median": per[n // 2], "max": per[-1]}}

    def rate(t):
        """whole-job MVis/s from the max-over-ranks seconds per step"""
        total = world * shape["mvis"] if not strong else shape["mvis"]
        return total / t["sec_per_step"]

    g_variant = idg.resolve_variant(N, C_, sincos, args.variant, gridder=True)
    d_variant = idg.resolve_variant(N, C_, sincos, args.degridder_variant, gridder=False)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_g = timed(gridder_step(args.variant), args.steps, args.warmup)
    t_d = timed(degridder_step(args.degridder_variant), args.steps, args.warmup)
    t_gw = timed(gridder_step(args.variant, uvw_w), args.steps, args.warmup)
    t_dw = timed(degridder_step(args.degridder_variant, uvw_w), args.steps, args.warmup)
    # comparators on the same data, a single round each: the per-pixel tcgen05 kernels (last round's defaults)
    # and the FP32 / SFU kernels (north_star: tensor cores only if they win)
    cmp_t = {}
    if sincos == idg.SINCOS_FAST and not args.no_compare:
        for name, mk, v in (("gridder_per_pixel", gridder_step, 24), ("degridder_per_pixel", degridder_step, 28),
                            ("gridder_fp32", gridder_step, 10), ("degridder_fp32", degridder_step, 4)):
            try:
                cmp_t[name] = (v, timed(mk(v, uvw_w), args.steps, args.warmup, min_seconds=0.0))
            except idg.IdgError:
                pass
    gridder_step(args.variant)()   # leave the default kernels' results (planar data) for the parity sample
    degridder_step(args.degridder_variant)()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else {}

    # ---- e2e through the host-pointer C ABI, pinned host buffers; the copy floor measured the same way
    e2e = None
    if not args.no_e2e:
        host = {}
        for k in ("uvw", "wavenumbers", "visibilities", "spheroidal", "aterms", "metadata", "subgrids"):
            t = prob[k]
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            host[k] = h
        torch.cuda.synchronize()
        npy = {k: v.numpy() for k, v in host.items()}
        meta_np = np.ascontiguousarray(npy["metadata"]).view(idg.METADATA_DTYPE).reshape(-1)
        a = (S, prob["grid_size"], N, prob["image_size"], 0.0, C_, prob["nr_stations"],
             npy["uvw"], npy["wavenumbers"], npy["visibilities"], npy["spheroidal"],
             npy["aterms"], meta_np, npy["subgrids"])

        def host_timed(fn, n):
            for _ in range(2):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()            # returns after the D2H of the result completed
            torch.cuda.synchronize()
            return all_max(time.perf_counter() - t0) / n

        e_steps = max(1, min(args.steps, 5))
        sec_e = host_timed(lambda: idg.c_run_gridder(*a, sincos=sincos, variant=args.variant), e_steps)
        sec_ed = host_timed(lambda: idg.c_run_degridder(*a, sincos=sincos, variant=args.degridder_variant), e_steps)
        h2d_g = sum(npy[k].nbytes for k in ("uvw", "wavenumbers", "visibilities", "spheroidal", "aterms", "metadata"))
        h2d_d = sum(npy[k].nbytes for k in ("uvw", "wavenumbers", "spheroidal", "aterms", "metadata", "subgrids"))
        # copy floor at this N: the step's own bytes, H2D and D2H at the same time on two streams, all ranks at
        # once (they share the host's memory and PCIe root), no kernel
        d_in = torch.empty(h2d_g, dtype=torch.uint8, device=dev)
        d_outb = torch.empty(npy["subgrids"].nbytes, dtype=torch.uint8, device=dev)
        h_in = torch.empty(h2d_g, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(npy["subgrids"].nbytes, dtype=torch.uint8, pin_memory=True)
        s1_, s2_ = torch.cuda.Stream(), torch.cuda.Stream()

        def copy_only():
            with torch.cuda.stream(s1_):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2_):
                h_out.copy_(d_outb, non_blocking=True)
            s1_.synchronize(); s2_.synchronize()

        sec_copy = host_timed(copy_only, e_steps)
        del d_in, d_outb, h_in, h_out
        total = (world if not strong else 1) * shape["mvis"]
        e2e = {"value": total / sec_e, "unit": "MVis/s", "h2d_bytes_per_step": int(h2d_g),
               "d2h_bytes_per_step": int(npy["subgrids"].nbytes), "ms_per_step": sec_e * 1e3, "steps": e_steps,
               "api": "idgb200_c_run_gridder_ex (host pointers, pinned; chunked H2D / kernel / D2H on 3 streams)",
               "copy_floor_ms": sec_copy * 1e3, "e2e_over_copy_floor": sec_copy / sec_e,
               "copy_floor_note": "the same bytes per rank, H2D and D2H concurrently on two streams, every rank at once, "
                                  "no kernel: what the host's memory and PCIe give at this N; e2e_over_copy_floor = "
                                  "floor / step (1 = the kernels are fully hidden behind the copies)",
               "h2d_gbs_in_floor": h2d_g / sec_copy * 1e-9,
               "degridder": {"value": total / sec_ed, "unit": "MVis/s", "ms_per_step": sec_ed * 1e3,
                             "h2d_bytes_per_step": int(h2d_d), "d2h_bytes_per_step": int(npy["visibilities"].nbytes),
                             "api": "idgb200_c_run_degridder_ex (host pointers, pinned)"}}
        del host, npy

    # ---- the "next" rows around the two kernels: subgrid FFT, adder, splitter - byte movers
    next_rows = None
    if world == 1 and not args.no_compare:
        G = prob["grid_size"]
        work = sub_in.clone()
        grid = torch.zeros((4, G, G), dtype=torch.complex64, device=dev)
        sg_bytes = work.numel() * 8
        grid_bytes = min(grid.numel() * 8, sg_bytes)
        hbm = float(measured_peaks().get("hbm_gbs") or 0.0)

        def nr(step, nbytes):
            t = timed(step, args.steps, args.warmup, min_seconds=0.0)
            gbs = nbytes / t["sec_per_step"] * 1e-9
            return {"ms": t["sec_per_step"] * 1e3, "gb_per_s": gbs, "hbm_frac": gbs / hbm if hbm else None,
                    "algorithmic_bytes": int(nbytes)}

        next_rows = {
            "subgrid_fft": nr(lambda: idg.subgrid_fft(S, N, work, 1), 2 * sg_bytes),
            "adder": nr(lambda: idg.adder(S, G, N, prob["metadata"], work, grid), sg_bytes + grid_bytes),
            "splitter": nr(lambda: idg.splitter(S, G, N, prob["metadata"], work, grid), sg_bytes + grid_bytes),
            "note": "HBM-bound byte movers (hbm_frac against MEASURED_PEAKS.json hbm_gbs); the reference has none "
                    "of them (parity unpinned, oracle/idg_next_oracle.c)",
        }
        del work, grid

    # ---- N > 1: the grid adder with the grid row-scattered over the ranks (BASELINE config 5): the one
    # step of the pipeline with a real exchange.  Every rank adds its (FFT'd) subgrids; push / pull / nccl as in
    # ska_sdp_idg_bench_b200/grid_adder_rs.py, `auto` = the library's rule (idgb200_adder_rs_mode)
    adder_rs = None
    if world > 1 and not args.no_compare:
        G = prob["grid_size"]
        work = sub_in.clone()
        idg.subgrid_fft(S, N, work, 1)
        rs = idg.GridAdderRS(G, dev)
        adder_rs = {"grid_size": G, "subgrid_size": N, "subgrids_per_gpu": S, "auto_mode": idg.adder_rs_mode(S, N, G),
                    "what": "memset + adder + reduce-scatter of the 4 x G x G grid by rows, ms per call, max over ranks"}
        for mode in ("push", "pull", "nccl"):
            t = timed(lambda: rs.add(S, N, prob["metadata"], work, mode=mode, flags=idg.FLAG_FFT_SHIFT),
                      args.steps, args.warmup, min_seconds=0.0, adaptive=False)
            adder_rs[mode + "_ms"] = t["sec_per_step"] * 1e3
        del work, rs

    # ---- CPU baseline + parity sample, rank 0 at N=1 only
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        lib, kind, oracle_lib = load_cpu_checker()
        cores = host_threads()
        lib.set_threads(cores)
        n = max(64, 4 * cores)
        T = shape["nr_timesteps"]

        def sample(uvw):
            return oracle_lib.Problem(
                grid_size=prob["grid_size"], subgrid_size=N, image_size=prob["image_size"], w_step=0.0,
                nr_channels=C_, nr_stations=prob["nr_stations"],
                uvw=uvw[:n * T].cpu().numpy(), wavenumbers=prob["wavenumbers"].cpu().numpy(),
                visibilities=prob["visibilities"][:n * T].cpu().numpy(),
                spheroidal=prob["spheroidal"].cpu().numpy(), aterms=prob["aterms"].cpu().numpy(),
                metadata=np.ascontiguousarray(prob["metadata"][:n].cpu().numpy()).view(
                    oracle_lib.METADATA_DTYPE).reshape(-1),
                subgrids=sub_in[:n].cpu().numpy())

        cp = sample(prob["uvw"])
        dt = time_cpu(lib, cp, "gridder")          # also a warm-up
        reps = int(max(1, min(20, 12.0 / max(dt, 1e-3))))
        tg = [time_cpu(lib, cp, "gridder") for _ in range(reps)]
        td = [time_cpu(lib, cp, "degridder") for _ in range(max(1, reps // 2))]
        mv = 1e-6 * n * T * C_
        cpu_baseline = {
            "value": mv / (sum(tg) / len(tg)), "unit": "MVis/s", "cores": cores, "kind": kind,
            "sample": f"first {n} subgrids of the same device-generated workload, {reps} repeats, "
                      f"OpenMP over subgrids on all {cores} host threads",
            "degridder_value": mv / (sum(td) / len(td)),
        }

        def errs(got, ref, axis):
            return [float(np.abs(np.take(got, p, axis) - np.take(ref, p, axis)).max() /
                          np.abs(np.take(ref, p, axis)).max()) for p in range(4)]

        # parity of the bench-size run on that sample (per-pol max|d|/max|ref|), planar and off the plane
        parity = {"sample_subgrids": n, "sincos": args.sincos, "tolerance": 1e-3,
                  "gridder_max_rel_per_pol": errs(prob["subgrids"][:n].cpu().numpy(), lib.gridder(cp), 1),
                  "degridder_max_rel_per_pol": errs(vis_out[:n * T].cpu().numpy(), lib.degridder(cp), 2)}
        cpw = sample(uvw_w)
        gridder_step(args.variant, uvw_w)()
        degridder_step(args.degridder_variant, uvw_w)()
        torch.cuda.synchronize()
        parity["general_w"] = {"gridder_max_rel_per_pol": errs(prob["subgrids"][:n].cpu().numpy(), lib.gridder(cpw), 1),
                               "degridder_max_rel_per_pol": errs(vis_out[:n * T].cpu().numpy(), lib.degridder(cpw), 2)}

    ref_gpu = reference_gpu_kernels() if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        peaks = measured_peaks()
        consts = load_constants()
        sms = idg.sm_count()
        f_max = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        p_fp32 = sms * SM_FP32_LANES * 2 * f_max * 1e6 * 1e-12       # TFLOP/s
        p_xu = sms * SM_XU_LANES * f_max * 1e6                       # MUFU/s
        full = shape if not strong else shape
        shape_launch = dict(shape, nr_subgrids=S)                    # what one launch of this rank processes
        flops = idg.flops_gridder(C_, tt, S, N)
        nbytes = idg.bytes_gridder(C_, tt, S, N)
        hbm_peak = float(peaks.get("hbm_gbs") or 6650.0)
        tensor_peak = float(peaks.get("bf16_tflops") or 2250.0)
        tensor_sustained = float(peaks.get("bf16_tflops_sustained") or 0.0)    # cuBLAS back to back for seconds: at the power cap
        items = float(N) * N * tt * C_                               # (pixel, visibility) pairs per launch

        def roof(t, kind, variant):
            """fp32_frac = flops_gridder / t / FP32 peak, sfu_frac = 2 N^2 vis/s / XU peak, stated plainly
            for every kernel.  The row-column kernels (variant 30) do the model's MACs as a GEMM on tcgen05 and make
            64 instead of 1024 phasors per visibility, so those two exceed 1 and bound nothing; their roofline is the
            tensor pipe: `achieved` = the flops of the contraction the formulation needs (stage_flops: useful) / t against the
            measured dense tensor peak (MEASURED_PEAKS.json bf16_tflops = the fp16 rate), `executed_frac` with the
            fp16 lo operand parts and padding it actually issues.  The FP32 kernels are bound by FP32 issue."""
            sec = t["sec_per_step"]
            tf = flops / sec * 1e-12
            r = {"kernel_ms": sec * 1e3, "fp32_frac": tf / p_fp32, "sfu_frac": 2.0 * items / sec / p_xu,
                 "fp32_model_tflops": tf,
                 "hbm": {"achieved": nbytes / sec * 1e-9, "peak": hbm_peak, "unit": "GB/s", "frac": nbytes / sec * 1e-9 / hbm_peak,
                         "algorithmic_bytes": int(nbytes)},
                 "peaks": {"fp32_tflops": p_fp32, "xu_tmufu": p_xu * 1e-12, "tensor_tflops": tensor_peak,
                           "source": f"{sms} SMs x 128 FP32 lanes x 2 / x 16 MUFU lanes x {f_max:.0f} MHz (sm_max_mhz); "
                                     "tensor and HBM: MEASURED_PEAKS.json (burst: kernel timed alone)"}}
            c = consts.get(f"{kind}_sep" if variant == 30 else f"{kind}_v{variant}", {})
            per_sg = c.get("dram_bytes_per_subgrid")
            r["traffic"] = per_sg * S if per_sg else None
            if c:
                r["ncu"] = {k: c[k] for k in c if k.endswith("_pct") or k in ("source", "subgrids", "duration_ms")}
            if variant == 30:
                fl = stage_flops(shape_launch, kind)
                r.update({"bound": "tensor", "achieved": fl["useful"] / sec * 1e-12, "peak": tensor_peak, "unit": "TFLOP/s",
                          "frac": fl["useful"] / sec * 1e-12 / tensor_peak,
                          "executed_tflops": fl["executed"] / sec * 1e-12,
                          "executed_frac": fl["executed"] / sec * 1e-12 / tensor_peak,
                          "flops_per_launch": fl})
                if tensor_sustained:    # these kernels are timed back to back for >= 1 s and sit at the board's power cap, like
                    # the sustained cuBLAS figure (MEASURED_PEAKS.json: bf16_tflops_sustained, SM clock ~1360 MHz under it)
                    r["vs_sustained_peak"] = {"peak": tensor_sustained, "frac": fl["useful"] / sec * 1e-12 / tensor_sustained,
                                              "executed_frac": fl["executed"] / sec * 1e-12 / tensor_sustained}
            else:
                r.update({"bound": "fp32" if variant in (10, 4) else "issue", "achieved": tf, "peak": p_fp32,
                          "unit": "TFLOP/s (reference flop model)", "frac": tf / p_fp32})
            ck = sampler.between(*t["clock_marks"]) if rank == 0 and "clock_marks" in t else None
            mhz = (ck or {}).get("sm_mhz") or clocks.get("sm_mhz")
            if mhz:     # the clock THIS kernel's region ran at (falls back to the whole run's median)
                r["sm_mhz"] = mhz
                if ck:
                    r["clocks_in_region"] = ck
                r["frac_at_measured_clock"] = r["frac"] * f_max / mhz
            return r

        def entry(t, kind, variant, what=None):
            e = {"value": rate(t), "unit": "MVis/s", "variant": variant, "ms_per_step": t["sec_per_step"] * 1e3,
                 "tflops": (world if not strong else 1) * idg.flops_gridder(C_, shape["total_timesteps"], S_total, N)
                           / t["sec_per_step"] * 1e-12 if not strong else
                           idg.flops_gridder(C_, shape["total_timesteps"], S_total, N) / t["sec_per_step"] * 1e-12,
                 "step_ms": t["step_ms"], "timed_steps": t["timed_steps"], "rounds": t["rounds"],
                 "timed_region_s": t["timed_region_s"], "warmup_steps": t["warmup_steps"],
                 "gpu_launches_per_step": t["launches_per_step"], "roofline": roof(t, kind, variant)}
            if what:
                e["what"] = what
            return e

        top = entry(t_g, "gridder", g_variant)
        out = {
            "metric": "gridder_mvis_per_s", "value": top["value"],
            "unit": "MVis/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": top["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None,
            "dtype": "f32 phasors and accumulation; fp16 (+ fp16 / e4m3 error-compensation parts where stated) tcgen05 operands" if g_variant >= 20 else "f32",
            "data": "synthetic",
            "config": dict(workload_config(shape, world), sincos=args.sincos,
                           gridder_variant=g_variant, degridder_variant=d_variant,
                           sharding=(f"one observation of {S_total} subgrids cut into {world} contiguous ranges balanced by "
                                     f"timesteps (shard.py), no collective" if strong else
                                     f"subgrid list, {world} independent shard(s) of {S_total} subgrids, no collective"),
                           timing="per-step CUDA events; rounds of `steps` steps until the region is >= 1 s; warm-up until two "
                                  "consecutive steps agree to 2 %",
                           gridder_kernel="gridder_sep.cu: row-column form, one tcgen05 GEMM per subgrid with the visibilities as K "
                                          "(A = fp16(Y vis), B = fp16 hi + lo column phasors), per-subgrid separability check on the "
                                          "device, per-pixel kernel behind it" if g_variant == 30 else f"variant {g_variant}",
                           degridder_kernel="degridder_sep.cu: row-column form, tcgen05 GEMM over the columns (hi hi in fp16, the "
                                            "two cross products as one e4m3 MMA), the sum over the rows on the CUDA cores out of TMEM; persistent warp-specialised pipeline, one CTA "
                                            "per SM (producers / issuer / consumers / next-subgrid setup)" if d_variant == 30
                                            else f"variant {d_variant}"),
            "tflops": top["tflops"], "step_ms": top["step_ms"], "timed_steps": top["timed_steps"], "rounds": top["rounds"],
            "timed_region_s": top["timed_region_s"], "warmup_steps": top["warmup_steps"],
            "roofline": top["roofline"],
            "degridder_value": rate(t_d), "degridder_ms_per_step": t_d["sec_per_step"] * 1e3,
            "degridder": entry(t_d, "degridder", d_variant),
            "general_w": {"what": "the same observation with w ~ N(0, 256 m) on every timestep (the reference's generator writes "
                                  "w = 0): the default kernels take the same path",
                          "gridder": entry(t_gw, "gridder", g_variant), "degridder": entry(t_dw, "degridder", d_variant)},
            "comparators": {name: entry(t, name.split("_")[0], v, "same data (w != 0), one round") for name, (v, t) in cmp_t.items()},
            "reference_gpu": ref_gpu,
            "cpu_baseline": cpu_baseline, "parity": parity, "e2e": e2e, "next_rows": next_rows, "adder_rs": adder_rs,
            "gpu_launches": int(round(t_g["launches_per_step"] * args.steps)),
            # (kernels per `steps` steps: the row-column gridder + the two list-mode kernels behind it; the whole timed region,
            # `rounds` x `steps` steps, launched gpu_launches_in_timed_region)
            "gpu_launches_in_timed_region": int(round(t_g["launches_per_step"] * t_g["timed_steps"])),
            "clocks": clocks, "device": idg.device_name(),
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--sincos", default="fast", choices=["fast", "reduced", "accurate"])
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--degridder-variant", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-compare", action="store_true")
    ap.add_argument("--min-seconds", type=float, default=1.0,
                    help="rounds of --steps steps are repeated until the timed region is this long (0: one round; "
                         "for runs under ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
