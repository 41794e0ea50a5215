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
# tensor-core kernel variants -> (MUFU per (pixel, visibility) item, dispatch cycles per warp-item executed,
# dispatch cycles per warp-item of the formulation's minimal instruction mix).
# Dispatch cycles = warp instructions + 1 per packed fp32x2 instruction (FFMA2 / FMUL2 / FADD2 hold the
# sub-partition's dispatch port for two cycles: DESIGN.md 3).  "Executed" is counted by ncu on the committed
# captures under profiles/ (gridder 24; degridder 22; the other variants' figures are older captures or
# estimates) and includes loop control, barrier polls, MMA issue and the builder warp.  "Minimal" counts only
# what the formulation needs per item: gridder 24 = F2FP 1 + recurrence FFMA2 2 x 6/8 +
# (first rotation 4 + first-channel sincos 4) / 8 + STS.128 1/4 + per-timestep (phase index 12 + rotation
# sincos 16) / 64 = 4.19; degridder 22 per pixel and channel quad = phase index 4 + two sincos 8 + 2 cos 1 +
# rotation 4 + 2 recurrence steps 4 + 4 x (F2FP + 2 FHFMA + F2FP) 16 + STS 2 + LDS 1 = 40 / 4 = 10;
# degridder 24 per pixel and group of 8 channels = LDS 1 + phase index 4 + two sincos 8 + 2 cos 1 + rotation 4
# + 6 recurrence steps 12 + 8 x 4 = 32 + STS 4 = 66 / 8 = 8.25.
TC_GRIDDER = {11: (2.0, 10.1, None), 12: (1.5, 12.6, None), 13: (1.375, 12.9, None), 14: (1.25, 13.3, None),
              15: (1.0, 14.2, None), 21: (0.375, 10.0, None), 22: (0.375, 15.2, None), 23: (2.0, 17.1, None),
              24: (0.382, 6.91, 4.19), 26: (0.382, 6.91, 4.19)}
TC_DEGRIDDER = {11: (2.0, 11.5, None), 12: (1.5, 13.5, None), 13: (1.25, 13.2, None), 14: (1.0, 15.5, None),
                21: (1.0, 10.5, None), 22: (1.006, 14.63, 10.0), 23: (2.0, 18.5, None),
                24: (0.505, 11.95, 8.25), 28: (0.505, 11.95, 8.25)}
# planar launches (pixel-pair folding): per NOMINAL (pixel, visibility) item, i.e. half the general path's
# figures plus the epilogue / builder share; executed counts from profiles/r01_fold_ncu_full.txt
# (2.79 / 5.02 warp instructions per nominal warp-item; the packed share taken from the general path's captures)
TC_GRIDDER[29] = (0.191, 3.30, 2.10)
TC_DEGRIDDER_FOLDED = (0.253, 5.52, 4.13)


# tcgen05.mma instructions (M=128, N=16, K=16, both operands in shared memory) per (pixel, visibility) item:
# one MMA covers 128 rows x 8 K-pairs = 1024 items; the hi + lo phasor kernels issue two per block.
TC_MMA_PER_ITEM = {"gridder": {21: 1 / 1024, 24: 1 / 1024, 26: 1 / 1024, 22: 2 / 1024, 23: 2 / 1024, 29: 0.5 / 1024},
                   "degridder": {21: 1 / 1024, 22: 2 / 1024, 23: 2 / 1024, 24: 2 / 1024, 25: 2 / 1024, 28: 2 / 1024}}
# measured on the B200 (tools/smem_mix.cu, profiles/r01_smem_mix_microbench.log): tcgen05.mma of that shape take
# 40.0 clocks each on an SM when the pipe stays fed (256 between two commits; 51-55 with a drain every 16) - the
# 4 KB A tile is fetched from shared memory at ~100 B/clk, 5x the 8 clocks of its math - and concurrent STS.128
# traffic of the LSU does not slow them (the two paths do not share one crossbar)
MMA_SS_CLOCKS = 40.0


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

    shape = rank_shape({}, rank, world)
    sincos = {"fast": idg.SINCOS_FAST, "reduced": idg.SINCOS_REDUCED,
              "accurate": idg.SINCOS_ACCURATE}[args.sincos]
    prob = idg.init_problem_device(
        nr_stations=shape["nr_stations"], nr_timeslots=shape["nr_timeslots"],
        nr_timesteps=shape["nr_timesteps"], nr_channels=shape["nr_channels"],
        subgrid_size=shape["subgrid_size"], grid_size=shape["grid_size"],
        image_size=shape["image_size"], seed=shape["seed"], device=dev)
    S, tt, C_, N = prob["nr_subgrids"], prob["total_timesteps"], prob["nr_channels"], prob["subgrid_size"]
    scal = (S, prob["grid_size"], N, prob["image_size"], 0.0, C_, prob["nr_stations"], tt)
    tens = (prob["uvw"], prob["wavenumbers"], prob["visibilities"], prob["spheroidal"],
            prob["aterms"], prob["metadata"], prob["subgrids"])
    sub_in = prob["subgrids"].clone()  # degridder input (the gridder overwrites subgrids)
    vis_out = torch.empty_like(prob["visibilities"])

    def step_gridder():
        idg.gridder(*scal, *tens, sincos=sincos, variant=args.variant)

    def step_degridder():
        idg.degridder(*scal, prob["uvw"], prob["wavenumbers"], vis_out, prob["spheroidal"],
                      prob["aterms"], prob["metadata"], sub_in, sincos=sincos,
                      variant=args.degridder_variant)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps, warmup):
        """K steps between CUDA events on the launching (torch current) stream,
        barrier + synchronize on both sides; returns (max-over-ranks seconds, launches)."""
        for _ in range(warmup):
            step()
        barrier()
        l0 = idg.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        sec = e0.elapsed_time(e1) * 1e-3
        return reduce_max_time(sec, dev), idg.launch_count() - l0

    # which kernels variant 0 selects for this shape (idgb200_resolve_variant)
    g_variant = idg.resolve_variant(N, C_, sincos, args.variant, gridder=True)
    d_variant = idg.resolve_variant(N, C_, sincos, args.degridder_variant, gridder=False)

    def step_gridder_fp32():
        idg.gridder(*scal, *tens, sincos=sincos, variant=10)

    def step_degridder_fp32():
        idg.degridder(*scal, prob["uvw"], prob["wavenumbers"], vis_out, prob["spheroidal"],
                      prob["aterms"], prob["metadata"], sub_in, sincos=sincos, variant=4)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sec_g, launches_g = timed(step_gridder, args.steps, args.warmup)
    sec_d, launches_d = timed(step_degridder, args.steps, args.warmup)
    # the FP32/SFU gridder next to the tensor-core one (north_star: tensor cores only if they win)
    # The synthetic observation is the reference's own (init.cpp:4-25): every w is 0, so the default kernels
    # take their planar paths (mirror-image pixel pairs share one phasor row: gridder_fold.cu, degridder_tc8.cu).
    # The same kernels' general path (any w) is timed beside them on the same data: variants 24 / 28.
    planar = not bool(prob["uvw"][:, 2].any().item())
    sec_gw = sec_dw = None
    if planar and g_variant == 29:
        sec_gw = timed(lambda: idg.gridder(*scal, *tens, sincos=sincos, variant=24), args.steps, args.warmup)[0]
    if planar and d_variant == 24:
        sec_dw = timed(lambda: idg.degridder(*scal, prob["uvw"], prob["wavenumbers"], vis_out, prob["spheroidal"],
                                             prob["aterms"], prob["metadata"], sub_in, sincos=sincos, variant=28),
                       args.steps, args.warmup)[0]
    sec_g32 = timed(step_gridder_fp32, args.steps, args.warmup)[0] if g_variant in TC_GRIDDER else None
    sec_d32 = timed(step_degridder_fp32, args.steps, args.warmup)[0] if d_variant in TC_DEGRIDDER else None
    step_gridder()   # leave the default kernels' results in prob["subgrids"] / vis_out for the parity sample
    step_degridder()
    clocks = sampler.stop() if rank == 0 else {}

    # ---- e2e through the host-pointer C ABI, pinned host buffers
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

        def e2e_step():
            idg.c_run_gridder(*a, sincos=sincos, variant=args.variant)

        e_steps = max(1, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()          # returns after the D2H of the result completed
        torch.cuda.synchronize()
        sec_e = reduce_max_time(time.perf_counter() - t0, dev)
        h2d = sum(npy[k].nbytes for k in ("uvw", "wavenumbers", "visibilities", "spheroidal",
                                          "aterms", "metadata"))
        # what the link gives: one pinned 1 GiB host->device copy on this rank, CUDA events
        probe_h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
        probe_d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        probe_d.copy_(probe_h, non_blocking=True)
        torch.cuda.synchronize()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        probe_d.copy_(probe_h, non_blocking=True)
        pe1.record()
        torch.cuda.synchronize()
        h2d_gbs = (1 << 30) / (pe0.elapsed_time(pe1) * 1e-3) * 1e-9
        del probe_h, probe_d
        # the degridder the same way: subgrids host -> device, visibilities device -> host
        def e2e_step_d():
            idg.c_run_degridder(*a, sincos=sincos, variant=args.degridder_variant)

        for _ in range(2):
            e2e_step_d()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step_d()
        torch.cuda.synchronize()
        sec_ed = reduce_max_time(time.perf_counter() - t0, dev)
        e2e_degridder = {"value": world * shape["mvis"] * e_steps / sec_ed, "unit": "MVis/s",
                         "ms_per_step": sec_ed / e_steps * 1e3,
                         "h2d_bytes_per_step": int(sum(npy[k].nbytes for k in ("uvw", "wavenumbers", "spheroidal", "aterms",
                                                                               "metadata", "subgrids"))),
                         "d2h_bytes_per_step": int(npy["visibilities"].nbytes),
                         "api": "idgb200_c_run_degridder_ex (host pointers, pinned)"}
        e2e = {"value": world * shape["mvis"] * e_steps / sec_e, "unit": "MVis/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(npy["subgrids"].nbytes),
               "ms_per_step": sec_e / e_steps * 1e3, "steps": e_steps,
               "api": "idgb200_c_run_gridder_ex (host pointers, pinned; chunked H2D/kernel/D2H on 3 streams)",
               "pcie_h2d_gbs_measured": h2d_gbs, "h2d_floor_ms": h2d / (h2d_gbs * 1e9) * 1e3,
               "note": "copy-bound: the step cannot be faster than its host->device bytes over the measured link",
               "degridder": e2e_degridder}

    # ---- the "next" rows around the two kernels: subgrid FFT, adder, splitter - byte movers,
    # timed with the same protocol on a scratch copy; algorithmic bytes as in tools/next_rows_bench.py
    next_rows = None
    if world == 1:
        G = prob["grid_size"]
        work = sub_in.clone()
        grid = torch.zeros((4, G, G), dtype=torch.complex64, device=dev)
        sg_bytes = work.numel() * 8
        grid_bytes = min(grid.numel() * 8, sg_bytes)
        hbm = float(measured_peaks().get("hbm_gbs") or 0.0)

        def nr(step, nbytes):
            sec, _ = timed(step, args.steps, args.warmup)
            gbs = nbytes * args.steps / sec * 1e-9
            return {"ms": sec / args.steps * 1e3, "gb_per_s": gbs, "hbm_frac": gbs / hbm if hbm else None,
                    "algorithmic_bytes": int(nbytes)}

        next_rows = {
            "subgrid_fft": nr(lambda: idg.subgrid_fft(S, N, work, 1), 2 * sg_bytes),
            "adder": nr(lambda: idg.adder(S, G, N, prob["metadata"], work, grid), sg_bytes + grid_bytes),
            "splitter": nr(lambda: idg.splitter(S, G, N, prob["metadata"], work, grid), sg_bytes + grid_bytes),
            "note": "HBM-bound byte movers (hbm_frac against MEASURED_PEAKS.json hbm_gbs); the reference has none "
                    "of them (parity unpinned, oracle/idg_next_oracle.c); together < 4 % of a gridder launch",
        }
        del work, grid

    # ---- CPU baseline + parity sample, rank 0 at N=1 only
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        lib, kind, oracle_lib = load_cpu_checker()
        cores = host_threads()
        lib.set_threads(cores)
        n = max(64, 4 * cores)
        # the same workload: copy the first n subgrids of the device-generated inputs
        T = shape["nr_timesteps"]
        cp = oracle_lib.Problem(
            grid_size=prob["grid_size"], subgrid_size=N, image_size=prob["image_size"], w_step=0.0,
            nr_channels=C_, nr_stations=prob["nr_stations"],
            uvw=prob["uvw"][:n * T].cpu().numpy(), wavenumbers=prob["wavenumbers"].cpu().numpy(),
            visibilities=prob["visibilities"][:n * T].cpu().numpy(),
            spheroidal=prob["spheroidal"].cpu().numpy(), aterms=prob["aterms"].cpu().numpy(),
            metadata=np.ascontiguousarray(prob["metadata"][:n].cpu().numpy()).view(
                oracle_lib.METADATA_DTYPE).reshape(-1),
            subgrids=sub_in[:n].cpu().numpy())
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
        # parity of the bench-size run on that sample (per-pol max|d|/max|ref|)
        ref_g = lib.gridder(cp)
        got_g = prob["subgrids"][:n].cpu().numpy()
        ref_d = lib.degridder(cp)
        got_d = vis_out[:n * T].cpu().numpy()
        parity = {
            "sample_subgrids": n, "sincos": args.sincos,
            "gridder_max_rel_per_pol": [float(np.abs(got_g[:, p] - ref_g[:, p]).max() /
                                              np.abs(ref_g[:, p]).max()) for p in range(4)],
            "degridder_max_rel_per_pol": [float(np.abs(got_d[..., p] - ref_d[..., p]).max() /
                                                np.abs(ref_d[..., p]).max()) for p in range(4)],
        }

    ref_gpu = reference_gpu_kernels() if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        peaks = measured_peaks()
        sms = idg.sm_count()
        f_max = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        p_fp32 = sms * SM_FP32_LANES * 2 * f_max * 1e6 * 1e-12       # TFLOP/s
        p_xu = sms * SM_XU_LANES * f_max * 1e6                       # MUFU/s
        flops = idg.flops_gridder(C_, tt, S, N)
        nbytes = idg.bytes_gridder(C_, tt, S, N)
        hbm_peak = float(peaks.get("hbm_gbs") or 6650.0)

        def roof(sec, steps, traffic_per_subgrid, tc, mma_per_item=None):
            """tc None: FP32 kernel (bound = the FP32 issue port); else (MUFU per item, dispatch
            cycles per warp-item) of a tensor-core kernel, whose MACs run on tcgen05 and whose roof
            is the XU (MUFU) pipe or the instruction dispatch port, whichever is busier."""
            t = sec / steps
            tf = flops / t * 1e-12
            items = float(N) * N * tt * C_                     # (pixel, visibility) pairs per launch
            hbm = {"achieved": nbytes / t * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                   "frac": nbytes / t * 1e-9 / hbm_peak,
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6650"}
            common = {"traffic": traffic_per_subgrid * S,
                      "traffic_note": "DRAM bytes/launch (dram__bytes_read+write) from the ncu --set full capture "
                                      "under profiles/, scaled by subgrid count; algorithmic bytes "
                                      "(bytes_gridder) = %.3f GB/launch" % (nbytes * 1e-9),
                      "kernel_ms": t * 1e3, "hbm": hbm,
                      "fp32_model_tflops": tf, "fp32_model_frac": tf / p_fp32,
                      "fp32_model_note": f"reference flop model (flops_gridder) / ({sms} SMs x 128 FP32 lanes x 2 x "
                                         f"{f_max:.0f} MHz): the north_star's 'fraction of the FP32+SFU roofline'"}
            if tc is None:
                r = dict({"bound": "fp32", "achieved": tf, "peak": p_fp32, "unit": "TFLOP/s", "frac": tf / p_fp32,
                          "bound_note": "CUDA-core FP32 issue bound (359 flop/B): neither HBM nor tensor is the roof",
                          "peak_source": f"{sms} SMs x 128 FP32 lanes x 2 x {f_max:.0f} MHz (sm_max_mhz of "
                                         "MEASURED_PEAKS.json; the file holds no FP32 figure)",
                          "sfu_frac": 2.0 * items / t / p_xu}, **common)
            else:
                mufu_per_item, inst_per_item, floor_per_item = tc
                mufu = mufu_per_item * items / t               # MUFU.SIN + MUFU.COS executed per second
                p_issue = sms * 4 * f_max * 1e6                 # dispatch cycles / s (1 per SMSP and clock)
                issue = inst_per_item * items / 32.0 / t
                sfu_frac, issue_frac = mufu / p_xu, issue / p_issue
                # the same port, counting only the minimal instruction mix of the formulation: what
                # fraction of the dispatch roof does useful work (the kernel-quality figure)
                useful_frac = None if floor_per_item is None else floor_per_item * items / 32.0 / t / p_issue
                note = ("tcgen05 kernel: the complex MACs run on the tensor pipe (6-8 %% busy); what bounds it is "
                        "generating the phasor operand: %.3f MUFU and %.2f dispatch cycles (instructions + 1 per "
                        "packed fp32x2 instruction) per (pixel, visibility) against the XU pipe and the "
                        "sub-partition's dispatch port (fp32_model_frac > 1 is the flop model's MACs and most of "
                        "its sincos having left the FP32 / XU pipes)" % (mufu_per_item, inst_per_item))
                if sfu_frac >= issue_frac:
                    r = dict({"bound": "sfu", "achieved": mufu * 1e-12, "peak": p_xu * 1e-12, "unit": "TMUFU/s",
                              "frac": sfu_frac, "peak_source": f"{sms} SMs x 16 MUFU lanes x {f_max:.0f} MHz"})
                else:
                    useful = useful_frac if useful_frac is not None else issue_frac
                    r = dict({"bound": "issue", "achieved": useful * p_issue * 1e-12, "peak": p_issue * 1e-12,
                              "unit": "T dispatch-cycles/s", "frac": useful,
                              "frac_note": "useful dispatch cycles (minimal instruction mix of the formulation, "
                                           "%.2f per warp-item) / peak; issue_frac beside it counts every executed "
                                           "instruction (%.2f per warp-item, ncu)" % (floor_per_item or inst_per_item,
                                                                                      inst_per_item),
                              "peak_source": f"{sms} SMs x 4 sub-partitions x {f_max:.0f} MHz; dispatch cycles per "
                                             "item from the ncu capture under profiles/"})
                if mma_per_item:
                    # the busiest resource: the tensor core's shared-memory operand fetch (one small-N MMA per
                    # 1024 items still reads its whole 4 KB A tile)
                    p_mma = sms * f_max * 1e6 / MMA_SS_CLOCKS
                    mma = mma_per_item * items / t
                    # ... but only when it is busier than the dispatch port with everything it executes: the
                    # ablation builds (tools/ablate.py, profiles/r01_ablation.log) show the
                    # instruction stream that makes the phasors to be 82 % / 77 % of the default kernels' time,
                    # the MMAs and operand stores adding the rest by interference
                    if mma / p_mma >= max(sfu_frac, issue_frac):
                        r = {"bound": "tensor", "achieved": mma * 1e-9, "peak": p_mma * 1e-9,
                             "unit": "G tcgen05.mma/s (M=128, N=16, K=16, operands in shared memory)",
                             "frac": mma / p_mma, "useful_issue_frac": useful_frac,
                             "peak_source": f"{sms} SMs x {f_max:.0f} MHz / {MMA_SS_CLOCKS} clocks per MMA, measured "
                                            "back to back on this GPU type (tools/smem_mix.cu, "
                                            "profiles/r01_smem_mix_microbench.log)",
                             "frac_note": "share of the time the SM's tensor pipe needs for the launch's MMAs at the "
                                          "measured rate: with N = 16 an MMA is bound by fetching its 4 KB A tile from "
                                          "shared memory, not by its math (pipe_tensor_cycles_active 11-13 % in ncu)"}
                r.update({"bound_note": note, "sfu_frac": sfu_frac, "issue_frac": issue_frac,
                          "useful_issue_frac": useful_frac,
                          "bound_evidence": "ablation builds of the general-path kernels (no MMAs, operand stores "
                                            "predicated off): 10.53 of 12.85 ms (gridder) and 17.61 of 22.80 ms "
                                            "(degridder) are the instruction stream alone, at 90 % / 94 % of the "
                                            "dispatch port (profiles/r01_ablation.log); the tensor pipe sustains "
                                            "41.7 clocks per N = 16 MMA from 12-24 issuing warps "
                                            "(profiles/r01_mma_commit_microbench.log); tensor_frac counts the folded "
                                            "gridder's N = 32 MMAs at that rate",
                          "tensor_frac": (mma_per_item * items / t) / (sms * f_max * 1e6 / MMA_SS_CLOCKS)
                          if mma_per_item else None}, **common)
            if clocks.get("sm_mhz"):
                r["frac_at_measured_clock"] = r["frac"] * f_max / clocks["sm_mhz"]
            return r

        tc_g, tc_d = TC_GRIDDER.get(g_variant), TC_DEGRIDDER.get(d_variant)
        mma_g, mma_d = TC_MMA_PER_ITEM["gridder"].get(g_variant), TC_MMA_PER_ITEM["degridder"].get(d_variant)
        if planar and d_variant == 24:
            tc_d, mma_d = TC_DEGRIDDER_FOLDED, mma_d / 2
        if g_variant == 29 and not planar:
            tc_g, mma_g = TC_GRIDDER[24], TC_MMA_PER_ITEM["gridder"][24]
        total_mvis = world * shape["mvis"]

        def general_w(sec, variant, tc, mma, traffic):
            if sec is None:
                return None
            return {"value": total_mvis * args.steps / sec, "unit": "MVis/s", "variant": variant,
                    "ms_per_step": sec / args.steps * 1e3, "tflops": world * flops * args.steps / sec * 1e-12,
                    "what": "the same launch through the kernel's general path (any w: no pixel-pair folding), "
                            "which is what a subgrid with a single w != 0 takes",
                    "roofline": roof(sec, args.steps, traffic, tc, mma)}
        out = {
            "metric": "gridder_mvis_per_s", "value": total_mvis * args.steps / sec_g,
            "unit": "MVis/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec_g / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32 (phasors and accumulation), fp16 x fp16 -> f32 tcgen05 operands" if tc_g is not None
                     else "f32",
            "data": "synthetic",
            "config": dict(workload_config(shape, world), sincos=args.sincos,
                           gridder_variant=g_variant, degridder_variant=d_variant,
                           gridder_kernel=("tcgen05: fp16 phasor tile x fp16 hi+lo visibilities, f32 accumulate in "
                                           "TMEM" + ("; phasors of equally spaced channels from the first channel of each "
                                                     "8-channel block by one rotation and the three-term recurrence (the "
                                                     "reference's gridder_v8 rotates every channel)"
                                                     if g_variant in (21, 24, 26, 29) else "")
                                           + ("; planar subgrids (every w = 0, checked per subgrid on the device): one phasor row per "
                                              "mirror-image pixel pair, N = 32 MMAs, E +- i F recombined in the epilogue "
                                              "(gridder_fold.cu)" if g_variant == 29 and planar else "")
                                           if tc_g is not None else "FP32 FFMA2 + MUFU"),
                           degridder_kernel=(("tcgen05: fp16 hi+lo phasor tile x fp16 hi+lo pixels, f32 accumulate in "
                                              "TMEM; phasors of equally spaced channel quads by rotation"
                                              if d_variant == 22 else
                                              "tcgen05, two M-tiles per warp: fp16 hi+lo phasor tile x fp16 hi+lo pixels, f32 "
                                              "accumulate in TMEM; phasors of equally spaced groups of 8 channels by the "
                                              "three-term recurrence" + ("; planar subgrids (every w = 0, checked per "
                                              "subgrid): summed over mirror-image pixel pairs, half the K dimension"
                                              if planar else "") if d_variant == 24 else "tcgen05, opt-in variant")
                                             if tc_d is not None else "FP32 FFMA2 + MUFU")),
            "tflops": world * flops * args.steps / sec_g * 1e-12,
            # ncu --set full, 3675-subgrid launches: 343.1 MB (tcgen05 gridder), 314.7 MB (tcgen05 degridder), FP32 kernels from the
            # 1740-subgrid captures (profiles/)
            "roofline": roof(sec_g, args.steps, (339.1e6 if g_variant == 29 and planar else 343.1e6) / 3675 if tc_g is not None else 145.165e6 / 1740,
                             tc_g, mma_g),
            "planar": planar,
            "gridder_general_w": general_w(sec_gw, 24, TC_GRIDDER[24], TC_MMA_PER_ITEM["gridder"][24], 343.1e6 / 3675),
            "degridder": {"value": total_mvis * args.steps / sec_d, "unit": "MVis/s",
                          "ms_per_step": sec_d / args.steps * 1e3,
                          "tflops": world * flops * args.steps / sec_d * 1e-12,
                          "roofline": roof(sec_d, args.steps, 314.7e6 / 3675 if tc_d is not None
                                           else 128.329e6 / 1740, tc_d, mma_d)},
            "degridder_general_w": general_w(sec_dw, 28, TC_DEGRIDDER[24], TC_MMA_PER_ITEM["degridder"][24],
                                             314.7e6 / 3675),
            "gridder_fp32": None if sec_g32 is None else {
                "value": total_mvis * args.steps / sec_g32, "unit": "MVis/s", "variant": 10,
                "ms_per_step": sec_g32 / args.steps * 1e3,
                "roofline": roof(sec_g32, args.steps, 145.165e6 / 1740, None)},
            "degridder_fp32": None if sec_d32 is None else {
                "value": total_mvis * args.steps / sec_d32, "unit": "MVis/s", "variant": 4,
                "ms_per_step": sec_d32 / args.steps * 1e3,
                "roofline": roof(sec_d32, args.steps, 128.329e6 / 1740, None)},
            "reference_gpu": ref_gpu,
            "cpu_baseline": cpu_baseline, "parity": parity, "e2e": e2e, "next_rows": next_rows,
            "gpu_launches": int(launches_g), "degridder_gpu_launches": int(launches_d),
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
    ap.add_argument("--sincos", default="fast", choices=["fast", "reduced", "accurate"])
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--degridder-variant", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
