"""BASELINE config 5 as ONE pipeline on N GPUs: subgrid 64, per-station A-terms that change with the timeslot
(aterm_index = timeslot), gridder -> subgrid FFT -> grid adder with the grid row-scattered over the ranks
(ska_sdp_idg_bench_b200.GridAdderRS: push / pull / nccl).

ONE observation (seed 0 on every rank) is cut by shard.partition_subgrids; every rank grids and transforms
its range and the three reduce-scatter variants must give the slices of the grid a single GPU computes from
the whole observation (rank 0 does that too, as the check).  One JSON line on rank 0.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/config5_pipeline.py
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stations", type=int, default=16)
    ap.add_argument("--timeslots", type=int, default=8)
    ap.add_argument("--timesteps", type=int, default=64)
    ap.add_argument("--channels", type=int, default=16)
    ap.add_argument("--subgrid-size", type=int, default=64)
    ap.add_argument("--grid-size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import ska_sdp_idg_bench_b200 as idg

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    dist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)

    G, N, C = args.grid_size, args.subgrid_size, args.channels
    full = idg.init_problem_device(nr_stations=args.stations, nr_timeslots=args.timeslots, nr_timesteps=args.timesteps,
                                   nr_channels=C, subgrid_size=N, grid_size=G, per_slot_aterms=True, seed=0, device=dev)
    meta_np = np.ascontiguousarray(full["metadata"].cpu().numpy()).view(idg.METADATA_DTYPE).reshape(-1)
    assert len(set(meta_np["aterm_index"].tolist())) == args.timeslots       # per-slot A-term planes in use
    s0, s1 = idg.partition_subgrids(meta_np["nr_timesteps"], world)[rank]
    m_loc, t0, t1 = idg.shard_metadata(meta_np, s0, s1)
    meta = torch.from_numpy(m_loc.view(np.int32).reshape(-1, 9).copy()).to(dev)
    uvw, vis = full["uvw"][t0:t1].contiguous(), full["visibilities"][t0:t1].contiguous()
    S = s1 - s0
    sg = torch.empty((S, 4, N, N), dtype=torch.complex64, device=dev)

    def grid_and_transform():
        idg.gridder(S, G, N, full["image_size"], 0.0, C, full["nr_stations"], t1 - t0, uvw, full["wavenumbers"], vis,
                    full["spheroidal"], full["aterms"], meta, sg)
        idg.subgrid_fft(S, N, sg, 1)

    rs = idg.GridAdderRS(G, dev)

    def timed(fn):
        for _ in range(2):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ms_front = timed(grid_and_transform)
    slices, ms = {}, {}
    for mode in ("push", "pull", "nccl"):
        ms[mode] = timed(lambda: rs.add(S, N, meta, sg, mode=mode, flags=idg.FLAG_FFT_SHIFT))
        slices[mode] = rs.add(S, N, meta, sg, mode=mode, flags=idg.FLAG_FFT_SHIFT).clone()
    auto = idg.adder_rs_mode(S, N, G)
    ms_pipeline = timed(lambda: (grid_and_transform(), rs.add(S, N, meta, sg, mode="auto", flags=idg.FLAG_FFT_SHIFT)))

    # the single-GPU answer: the whole observation through the same kernels on this rank's GPU
    S_all = full["nr_subgrids"]
    sg_all = torch.empty((S_all, 4, N, N), dtype=torch.complex64, device=dev)
    idg.gridder(S_all, G, N, full["image_size"], 0.0, C, full["nr_stations"], full["total_timesteps"], full["uvw"],
                full["wavenumbers"], full["visibilities"], full["spheroidal"], full["aterms"], full["metadata"], sg_all)
    idg.subgrid_fft(S_all, N, sg_all, 1)
    # sharding changes nothing: this rank's subgrids are bit-identical to its range of the single-GPU run
    shard_bitwise = bool(torch.equal(sg_all[s0:s1], sg))
    grid = torch.zeros((4, G, G), dtype=torch.complex64, device=dev)
    idg.adder(S_all, G, N, full["metadata"], sg_all, grid, flags=idg.FLAG_FFT_SHIFT)
    rows = rs.rows
    want = torch.zeros((4, rows, G), dtype=torch.complex64, device=dev)
    hi = min(G, (rank + 1) * rows)
    if hi > rank * rows:
        want[:, : hi - rank * rows] = grid[:, rank * rows:hi]
    scale = float(grid.abs().max())
    errs = torch.tensor([float((slices[m] - want).abs().max()) / scale for m in ("push", "pull", "nccl")] +
                        [0.0 if shard_bitwise else 1.0], dtype=torch.float64, device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        mvis = 1e-6 * full["total_timesteps"] * C
        print(json.dumps({
            "what": "BASELINE config 5: subgrid 64, per-timeslot A-terms, gridder -> subgrid FFT -> grid adder reduce-scatter",
            "n_gpus": world, "subgrids": S_all, "subgrid_size": N, "grid_size": G, "channels": C, "mvis": mvis,
            "gridder_plus_fft_ms": ms_front, "adder_rs_ms": ms, "auto_mode": auto, "pipeline_ms": ms_pipeline,
            "pipeline_mvis_per_s": mvis / (ms_pipeline * 1e-3),
            "max_abs_diff_over_max_vs_single_gpu": {"push": float(errs[0]), "pull": float(errs[1]), "nccl": float(errs[2])},
            "sharded_subgrids_bitwise_equal_single_gpu": bool(float(errs[3]) == 0.0)}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
