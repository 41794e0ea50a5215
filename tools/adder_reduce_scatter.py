"""Grid adder on N GPUs (BASELINE config 5 names the step): every rank holds its
shard of the subgrids, the sum of all ranks' adders is wanted row-scattered - rank r ends up with
grid rows [r * rows, (r + 1) * rows).

  baseline  each rank adds into a full local grid, then ncclReduceScatter (torch.distributed)
  fused     the owners' slices live in symmetric (peer-mapped) memory; every rank's adder kernel
            reduces its subgrids straight into the owners' slices with system-scope
            red.global.add.v4.f32 over NVLink / NVSwitch: no partial grid, no second pass
  pull      each rank adds into its own local grid, which lives in symmetric memory; after a barrier
            every rank sums ITS slice out of all ranks' local grids with 16-byte peer loads over
            NVLink (idgb200_reduce_parts): a one-shot reduce-scatter, rank-ordered sums
            (bit-reproducible), no staging buffers, no NCCL

All give the same slices (checked here to fp32 summation order).  Times are CUDA events on the
launching stream with a barrier on both sides, max over ranks.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/adder_reduce_scatter.py
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
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem

    import ska_sdp_idg_bench_b200 as idg

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)

    G, N = args.grid_size, args.subgrid_size
    prob = idg.init_problem_device(nr_stations=args.stations, nr_timeslots=args.timeslots, nr_timesteps=1,
                                   nr_channels=1, subgrid_size=N, grid_size=G, seed=100 + rank, device=dev)
    S, meta = prob["nr_subgrids"], prob["metadata"]
    sg = torch.randn((S, 4, N, N, 2), device=dev, generator=torch.Generator(dev).manual_seed(rank + 1))
    sg = torch.view_as_complex(sg.contiguous())
    rpp = (G + world - 1) // world

    def timed(fn):
        for _ in range(3):
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

    # ---- baseline: local grid (already cut into the owners' row blocks) + ncclReduceScatter
    local_grid = torch.zeros((world, 4, rpp, G), dtype=torch.complex64, device=dev)
    out_nccl = torch.zeros((4, rpp, G), dtype=torch.complex64, device=dev)
    local_parts = [local_grid[r] for r in range(world)]

    def baseline():
        local_grid.zero_()
        idg.adder(S, G, N, meta, sg, local_parts, rows_per_part=rpp)
        dist.reduce_scatter_tensor(torch.view_as_real(out_nccl), torch.view_as_real(local_grid))

    ms_base = timed(baseline)

    # ---- fused: peer-mapped slices, the adder's atomics are the reduce-scatter
    slice_f = symm_mem.empty((4, rpp, G, 2), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(slice_f, dist.group.WORLD)
    peer_ptrs = [int(p) for p in hdl.buffer_ptrs]

    def fused():
        slice_f.zero_()
        hdl.barrier(channel=0)                       # every owner's slice is zero before anyone adds into it
        idg.adder(S, G, N, meta, sg, peer_ptrs, rows_per_part=rpp)
        hdl.barrier(channel=1)                       # every rank's reductions have landed

    ms_fused = timed(fused)

    # ---- pull: local grid in symmetric memory, every rank reduces its slice out of all local grids
    part_elems = 4 * rpp * G
    sym_grid_f = symm_mem.empty((world, 4, rpp, G, 2), dtype=torch.float32, device=dev)
    hdl2 = symm_mem.rendezvous(sym_grid_f, dist.group.WORLD)
    sym_grid = torch.view_as_complex(sym_grid_f)
    sym_parts = [sym_grid[r] for r in range(world)]
    my_part_of = [int(p) + rank * part_elems * 8 for p in hdl2.buffer_ptrs]   # rank q's copy of MY part
    out_pull = torch.zeros((4, rpp, G), dtype=torch.complex64, device=dev)

    def pull():
        sym_grid_f.zero_()
        idg.adder(S, G, N, meta, sg, sym_parts, rows_per_part=rpp)
        hdl2.barrier(channel=0)                      # every rank's local grid is complete
        idg.reduce_parts(my_part_of, out_pull)
        hdl2.barrier(channel=1)                      # nobody zeroes a grid that is still being read

    ms_pull = timed(pull)
    fused()
    baseline()
    pull()
    torch.cuda.synchronize()
    a, b = torch.view_as_complex(slice_f), out_nccl
    err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    err_pull = float((out_pull - b).abs().max() / b.abs().max().clamp_min(1e-30))
    errs = torch.tensor([err, err_pull], dtype=torch.float64, device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        grid_bytes = 4 * G * G * 8
        print(json.dumps({
            "what": "grid adder + reduce-scatter by grid rows (SURVEY 8f-1)", "n_gpus": world, "grid_size": G,
            "subgrid_size": N, "subgrids_per_gpu": S, "grid_mbytes": grid_bytes * 1e-6,
            "nccl_ms": ms_base, "fused_peer_atomics_ms": ms_fused, "pull_peer_loads_ms": ms_pull,
            "speedup_push_vs_nccl": ms_base / ms_fused, "speedup_pull_vs_nccl": ms_base / ms_pull,
            "max_rel_difference_fused_vs_nccl": float(errs[0]), "max_rel_difference_pull_vs_nccl": float(errs[1]),
            "overlap_subgrid_pixels_per_grid_cell": world * S * N * N / (G * G),
            "note": "nccl = memset + adder into a full local grid + ncclReduceScatter; fused (push) = memset of the "
                    "own slice + barrier + adder with system-scope red.v4.f32 into the owners' slices + barrier; "
                    "pull = memset + adder into a local grid in symmetric memory + barrier + idgb200_reduce_parts "
                    "(16-byte peer loads of the own slice from every rank) + barrier"}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
