"""Per-step CUDA-event timing of gridder / degridder variants at the bench shape (config 2), optionally with
w ~ N(0, sigma): python tools/quick_time.py [--w-sigma 256] [--g 30,29,24] [--d 30,24,28] [--steps 30]
                                           [--channels 16] [--subgrid 32] [--stations 50]
Prints one JSON line per kernel: min / median / max ms per launch after a warm-up that lasts until two
consecutive launches agree to 2 %."""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ska_sdp_idg_bench_b200 as idg  # noqa: E402


def per_step(step, steps, min_warm=3, max_warm=200):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(max_warm + 1)]
    last, warm = None, 0
    for i in range(max_warm):
        ev[i].record(); step(); ev[i + 1].record(); ev[i + 1].synchronize()
        ms = ev[i].elapsed_time(ev[i + 1]); warm += 1
        if i + 1 >= min_warm and last is not None and abs(ms - last) <= 0.02 * last:
            break
        last = ms
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(steps):
        step(); ev[i + 1].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return {"min_ms": min(t), "median_ms": statistics.median(t), "max_ms": max(t), "mean_ms": ev[0].elapsed_time(ev[steps]) / steps,
            "warmup_steps": warm}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--w-sigma", type=float, default=0.0)
    ap.add_argument("--g", default="30,24")
    ap.add_argument("--d", default="0,28")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--channels", type=int, default=16)
    ap.add_argument("--subgrid", type=int, default=32)
    ap.add_argument("--stations", type=int, default=50)
    ap.add_argument("--timeslots", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    p = idg.init_problem_device(nr_stations=a.stations, nr_timeslots=a.timeslots, nr_timesteps=128, nr_channels=a.channels,
                                subgrid_size=a.subgrid, grid_size=1024, image_size=0.01, seed=0, device=dev)
    if a.w_sigma > 0:
        g = torch.Generator(device=dev).manual_seed(5)
        p["uvw"][:, 2] = torch.randn(p["uvw"].shape[0], device=dev, generator=g) * a.w_sigma
    S, tt, C, N = p["nr_subgrids"], p["total_timesteps"], p["nr_channels"], p["subgrid_size"]
    scal = (S, p["grid_size"], N, p["image_size"], 0.0, C, p["nr_stations"], tt)
    sub_in = p["subgrids"].clone()
    vis_out = torch.empty_like(p["visibilities"])
    mvis = 1e-6 * tt * C
    ref = {}
    for kind, variants in (("gridder", a.g), ("degridder", a.d)):
        for v in [int(x) for x in variants.split(",") if x != ""]:
            if kind == "gridder":
                def step():
                    idg.gridder(*scal, p["uvw"], p["wavenumbers"], p["visibilities"], p["spheroidal"], p["aterms"],
                                p["metadata"], p["subgrids"], sincos=idg.SINCOS_FAST, variant=v)
                out = p["subgrids"]
            else:
                def step():
                    idg.degridder(*scal, p["uvw"], p["wavenumbers"], vis_out, p["spheroidal"], p["aterms"], p["metadata"],
                                  sub_in, sincos=idg.SINCOS_FAST, variant=v)
                out = vis_out
            try:
                r = per_step(step, a.steps)
            except idg.IdgError as e:
                print(json.dumps({"kernel": kind, "variant": v, "error": str(e)})); continue
            r.update(kernel=kind, variant=v, resolved=idg.resolve_variant(N, C, idg.SINCOS_FAST, v, gridder=kind == "gridder"),
                     mvis_per_s=mvis / (r["median_ms"] * 1e-3), w_sigma=a.w_sigma, subgrids=S, channels=C, subgrid=N)
            torch.cuda.synchronize()
            # bit-level fingerprint of the whole output: equal across builds <=> the same bits
            bits = torch.view_as_real(out).contiguous().view(torch.int32).to(torch.int64)
            r["fingerprint"] = int((bits * (torch.arange(bits.numel(), device=bits.device).view(bits.shape) % 1000003 + 1)).sum().item() & 0xFFFFFFFFFFFF)
            o = out[: 64].clone() if kind == "gridder" else out[: 64 * 128].clone()
            if kind in ref:   # agreement with the first variant listed (different kernels, same answer)
                d = (o - ref[kind]).abs().max().item() / ref[kind].abs().max().item()
                r["max_rel_diff_vs_first"] = d
            else:
                ref[kind] = o
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
