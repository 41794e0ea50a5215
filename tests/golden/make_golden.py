"""Regenerates tests/golden/*.npz from the REFERENCE's own CPU code
(oracle/_ref/libidgref.so, built by `make -C oracle` from /root/reference).
Run in the build container only:  python tests/golden/make_golden.py

  config1.npz   the reference's correctness shape (tests/gridder_common.cpp:54-64:
                2 stations, 2 timeslots, 128 timesteps, 16 channels, N=32, G=1024),
                inputs made by the reference's initialize_* after srand(0);
                stores the full gridder/degridder outputs + input checksums.
  ragged_*.npz  adversarial problems (ragged nr_timesteps incl. 0, w != 0,
                w_step != 0, aterm_index != 0, unequal wavenumbers); inputs are
                stored too because numpy's RNG streams are not a stable ABI.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import random_problem, reference  # noqa: E402

FIELDS = ("uvw", "wavenumbers", "visibilities", "spheroidal", "aterms", "metadata", "subgrids")


def main():
    ref = reference()
    assert ref is not None, "build oracle/_ref first (make -C oracle)"
    ref.set_threads(ref.max_threads())

    p = ref.make_problem()
    cks = {f"sum_{k}": np.array(np.asarray(getattr(p, k)).view(np.float32).astype(np.float64).sum()
                                 if k != "metadata" else 0.0) for k in FIELDS if k != "metadata"}
    np.savez_compressed(
        os.path.join(HERE, "config1.npz"),
        gridder=ref.gridder(p), degridder=ref.degridder(p),
        metadata=p.metadata.view(np.int32).reshape(-1, 9), wavenumbers=p.wavenumbers,
        uvw_head=p.uvw[:4], aterms_head=p.aterms.reshape(-1)[:8], **cks)

    cases = {
        "ragged_a": dict(seed=11, nr_subgrids=5, subgrid_size=16, nr_channels=5, max_timesteps=9),
        "ragged_b": dict(seed=12, nr_subgrids=4, subgrid_size=24, nr_channels=3, max_timesteps=13,
                         nr_stations=5, nr_slots=3),
        "ragged_c": dict(seed=13, nr_subgrids=3, subgrid_size=8, nr_channels=17, max_timesteps=6,
                         with_w=False),
    }
    for name, kw in cases.items():
        seed = kw.pop("seed")
        q = random_problem(seed, **kw)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            gridder=ref.gridder(q), degridder=ref.degridder(q),
            scalars=np.array([q.grid_size, q.subgrid_size, q.nr_channels, q.nr_stations], np.int64),
            fscalars=np.array([q.image_size, q.w_step], np.float32),
            **{k: (getattr(q, k) if k != "metadata" else q.metadata.view(np.int32).reshape(-1, 9))
               for k in FIELDS})
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
