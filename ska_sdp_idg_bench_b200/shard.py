"""Partition of the subgrid list across GPUs.

Subgrids are independent, so the list is cut into contiguous ranges balanced by
sum(nr_timesteps); each rank gets metadata[s0:s1] together with the slice of
uvw / visibilities that range covers and subgrids[s0:s1].  No collective is
needed on the kernel path and the per-subgrid arithmetic is unchanged, so the
N-GPU result is bit-identical to the 1-GPU result.
"""
from __future__ import annotations

import numpy as np


def partition_subgrids(nr_timesteps: np.ndarray, world_size: int) -> list[tuple[int, int]]:
    """Contiguous [s0, s1) ranges, one per rank, balanced by work (= timesteps)."""
    nt = np.asarray(nr_timesteps, dtype=np.int64)
    S = int(nt.shape[0])
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    total = int(nt.sum())
    if total == 0:
        cuts = np.linspace(0, S, world_size + 1).round().astype(int)
        return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world_size)]
    csum = np.concatenate([[0], np.cumsum(nt)])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        s = int(np.searchsorted(csum, target, side="left"))
        # pick the closer of the two neighbouring cut points
        if s > 0 and abs(csum[s - 1] - target) <= abs(csum[min(s, S)] - target):
            s -= 1
        bounds.append(min(max(s, bounds[-1]), S))
    bounds.append(S)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def shard_metadata(metadata: np.ndarray, s0: int, s1: int) -> tuple[np.ndarray, int, int]:
    """metadata[s0:s1] rebased so that it indexes rank-local uvw / visibility slices.

    Returns (local_metadata, t0, t1): the rank needs rows [t0, t1) of the global
    uvw / visibility arrays and its local time offsets are relative to t0.  The
    kernels compute time_offset = (baseline_offset - metadata[0].baseline_offset)
    + time_offset (gridder_reference.cpp:16-25), so the local copy folds the
    baseline term into time_offset and zeroes baseline_offset.
    """
    m = np.array(metadata[s0:s1], copy=True)
    if m.shape[0] == 0:
        return m, 0, 0
    base0 = int(metadata[0]["baseline_offset"])
    start = (m["baseline_offset"].astype(np.int64) - base0) + m["time_offset"].astype(np.int64)
    live = m["nr_timesteps"] > 0
    if not live.any():
        t0 = t1 = 0
    else:
        t0 = int(start[live].min())
        t1 = int((start[live] + m["nr_timesteps"][live]).max())
    m["baseline_offset"] = 0
    m["time_offset"] = np.where(live, start - t0, 0).astype(np.int32)
    return m, t0, t1
