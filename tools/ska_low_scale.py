"""BASELINE config 4: the SKA1-Low-scale synthetic observation (512 stations, 64 channels, 64
timeslots x 128 timesteps per baseline, subgrid 32 = 8,372,224 subgrids, 68.6 GVis, 2.19 TB of
visibilities), sharded by subgrid over the GPUs of one box with no collective on the kernel path.

The visibilities cannot be resident, so every rank loops over device-resident chunks of its share:
a chunk = `--chunk-subgrids` subgrids of one 512-station timeslot, generated on the device by the
library's init kernels with the chunk's own seed, then gridded and degridded.  Only
the kernels are timed (CUDA events on the launching stream, inputs resident - the definition of
bench.py's `value`); the wall clock including the input generation is reported beside it.

  python tools/ska_low_scale.py --chunks-per-rank 2                       # quick look, 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \\
      tools/ska_low_scale.py --full                                       # the whole observation
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NR_STATIONS, NR_TIMESLOTS, NR_TIMESTEPS, NR_CHANNELS, SUBGRID = 512, 64, 128, 64, 32
NR_BASELINES = NR_STATIONS * (NR_STATIONS - 1) // 2
TOTAL_SUBGRIDS = NR_BASELINES * NR_TIMESLOTS          # 8,372,224


def chunk_plan(chunk_subgrids: int, world: int, rank: int, chunks_per_rank=None):
    """(timeslot, first-seed) pairs of this rank: the list of all chunks of the observation,
    dealt round-robin; `chunks_per_rank` truncates it for a partial run."""
    per_slot = (NR_BASELINES + chunk_subgrids - 1) // chunk_subgrids
    chunks = [(ts, q) for ts in range(NR_TIMESLOTS) for q in range(per_slot)]
    mine = chunks[rank::world]
    if chunks_per_rank is not None:
        mine = mine[:chunks_per_rank]
    sizes = [min(chunk_subgrids, NR_BASELINES - q * chunk_subgrids) for _, q in mine]
    return mine, sizes, len(chunks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk-subgrids", type=int, default=32704)
    ap.add_argument("--chunks-per-rank", type=int, default=2)
    ap.add_argument("--full", action="store_true", help="the whole observation (all chunks)")
    ap.add_argument("--sincos", default="fast", choices=["fast", "reduced", "accurate"])
    ap.add_argument("--w-sigma", type=float, default=0.0,
                    help="w ~ N(0, sigma) metres on every timestep (the reference's generator writes w = 0)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import ska_sdp_idg_bench_b200 as idg

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sincos = {"fast": idg.SINCOS_FAST, "reduced": idg.SINCOS_REDUCED, "accurate": idg.SINCOS_ACCURATE}[args.sincos]
    mine, sizes, nchunks_total = chunk_plan(args.chunk_subgrids, world, rank,
                                            None if args.full else args.chunks_per_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ms_g = ms_d = 0.0
    vis_done = 0
    barrier()
    t_wall = time.perf_counter()
    for (ts, q), S in zip(mine, sizes):
        prob = idg.init_problem_device(nr_stations=NR_STATIONS, nr_timeslots=1, nr_timesteps=NR_TIMESTEPS,
                                       nr_channels=NR_CHANNELS, subgrid_size=SUBGRID, nr_subgrids=S,
                                       seed=1 + ts * 1000 + q, device=dev)
        if args.w_sigma > 0:
            gen = torch.Generator(device=dev).manual_seed(7 + ts * 1000 + q)
            prob["uvw"][:, 2] = torch.randn(prob["uvw"].shape[0], device=dev, generator=gen) * args.w_sigma
        scal = (S, prob["grid_size"], SUBGRID, prob["image_size"], 0.0, NR_CHANNELS, NR_STATIONS,
                prob["total_timesteps"])
        sub_in = prob["subgrids"].clone()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        idg.gridder(*scal, prob["uvw"], prob["wavenumbers"], prob["visibilities"], prob["spheroidal"],
                    prob["aterms"], prob["metadata"], prob["subgrids"], sincos=sincos)
        e[1].record()
        idg.degridder(*scal, prob["uvw"], prob["wavenumbers"], prob["visibilities"], prob["spheroidal"],
                      prob["aterms"], prob["metadata"], sub_in, sincos=sincos)
        e[2].record()
        torch.cuda.synchronize()
        ms_g += e[0].elapsed_time(e[1])
        ms_d += e[1].elapsed_time(e[2])
        vis_done += S * NR_TIMESTEPS * NR_CHANNELS
        del prob, sub_in
    barrier()
    wall = time.perf_counter() - t_wall

    stats = torch.tensor([ms_g, ms_d, float(vis_done), wall], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = stats
    if rank == 0:
        total_vis = float(sm[2])
        out = {
            "workload": "BASELINE config 4: SKA1-Low-scale synthetic observation, 512 stations x 64 timeslots "
                        "x 128 timesteps, 64 channels, subgrid 32",
            "n_gpus": world, "chunk_subgrids": args.chunk_subgrids, "chunks_total": nchunks_total,
            "chunks_run": len(mine) * world if not args.full else nchunks_total,
            "fraction_of_observation": total_vis / (TOTAL_SUBGRIDS * NR_TIMESTEPS * NR_CHANNELS),
            "gvis": total_vis * 1e-9,
            "gridder": {"kernel_seconds_max_rank": float(mx[0]) * 1e-3,
                        "mvis_per_s": total_vis / (float(mx[0]) * 1e-3) * 1e-6,
                        "variant": idg.resolve_variant(SUBGRID, NR_CHANNELS, sincos, 0, gridder=True)},
            "degridder": {"kernel_seconds_max_rank": float(mx[1]) * 1e-3,
                          "mvis_per_s": total_vis / (float(mx[1]) * 1e-3) * 1e-6,
                          "variant": idg.resolve_variant(SUBGRID, NR_CHANNELS, sincos, 0, gridder=False)},
            "wall_seconds_incl_input_generation": float(mx[3]),
            "sincos": args.sincos, "w_sigma_m": args.w_sigma, "data": "synthetic, generated on device per chunk",
            "sharding": "chunks dealt round-robin over ranks, no collective on the kernel path",
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
