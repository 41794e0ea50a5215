"""Throughput of the "next" rows (subgrid FFT, adder, splitter) at the default perf
shape (24,500 subgrids of 32 x 32 x 4) and at subgrid 64, against the HBM roof: all three are byte
movers.  Algorithmic bytes per launch: FFT 16 B per pixel (read + write in place); adder 8 B per
pixel read + the touched grid cells once (the reductions land in L2; at most one cell per subgrid
pixel); splitter 8 B per pixel written + the touched grid cells.
CUDA events on the launching stream, 3 warm-up + 10 timed launches, one JSON line.

    python tools/next_rows_bench.py [--subgrid-size 32] [--grid-size 1024]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid-size", type=int, default=1024)
    ap.add_argument("--subgrid-size", type=int, default=32)
    ap.add_argument("--stations", type=int, default=50)
    ap.add_argument("--timeslots", type=int, default=20)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()

    import torch

    import ska_sdp_idg_bench_b200 as idg

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    G, N = args.grid_size, args.subgrid_size
    prob = idg.init_problem_device(nr_stations=args.stations, nr_timeslots=args.timeslots, nr_timesteps=1,
                                   nr_channels=1, subgrid_size=N, grid_size=G, seed=7, device=dev)
    S, meta = prob["nr_subgrids"], prob["metadata"]
    sg = torch.view_as_complex(torch.randn((S, 4, N, N, 2), device=dev))
    grid = torch.zeros((4, G, G), dtype=torch.complex64, device=dev)
    peak = None
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    sg_bytes = S * 4 * N * N * 8
    grid_bytes = min(4 * G * G * 8, sg_bytes)      # touched cells: at most one per subgrid pixel
    rows = {}

    def row(name, ms, nbytes):
        gbs = nbytes / ms * 1e-6
        rows[name] = {"ms": round(ms, 4), "algorithmic_mbytes": round(nbytes * 1e-6, 1), "gb_per_s": round(gbs, 1),
                      "hbm_frac": round(gbs / peak, 3) if peak else None}

    row("subgrid_fft_forward", timed(lambda: idg.subgrid_fft(S, N, sg, 1)), 2 * sg_bytes)
    row("subgrid_fft_backward", timed(lambda: idg.subgrid_fft(S, N, sg, -1)), 2 * sg_bytes)
    ref = torch.view_as_complex(torch.randn((S, 4, N, N, 2), device=dev))
    row("torch_fft2_cufft_out_of_place", timed(lambda: torch.fft.fft2(ref)), 2 * sg_bytes)
    sg = torch.view_as_complex(torch.randn((S, 4, N, N, 2), device=dev))
    row("adder", timed(lambda: idg.adder(S, G, N, meta, sg, grid)), sg_bytes + grid_bytes)
    row("adder_fft_shift", timed(lambda: idg.adder(S, G, N, meta, sg, grid, flags=idg.FLAG_FFT_SHIFT)),
        sg_bytes + grid_bytes)
    row("splitter", timed(lambda: idg.splitter(S, G, N, meta, sg, grid)), sg_bytes + grid_bytes)
    row("copy_subgrids_torch", timed(lambda: ref.copy_(sg)), 2 * sg_bytes)
    print(json.dumps({"what": "next rows (SURVEY 8f) on one B200", "subgrids": S, "subgrid_size": N, "grid_size": G,
                      "hbm_peak_gbs": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs", "rows": rows,
                      "device": idg.device_name()}))


if __name__ == "__main__":
    main()
