"""BASELINE config 5 (subgrid 64, per-timeslot A-terms, grid-adder reduce-scatter over NVLink) as one
pipeline on two GPUs: tools/config5_pipeline.py under torchrun.  Skipped with fewer than two devices."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_config5_pipeline_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "tools", "config5_pipeline.py"), "--steps", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    print(line)
    assert out["n_gpus"] == 2 and out["subgrid_size"] == 64
    assert out["sharded_subgrids_bitwise_equal_single_gpu"]
    # sums of the same fp32 terms in a different order: agreement to fp32 rounding of the grid's largest cell
    for mode, err in out["max_abs_diff_over_max_vs_single_gpu"].items():
        assert err < 2e-6, (mode, err)


def test_adder_rs_rule():
    """idgb200_adder_rs_mode: push when a rank has fewer subgrid pixels than the grid has cells."""
    import ska_sdp_idg_bench_b200 as idg
    assert idg.adder_rs_mode(24500, 32, 1024) == "pull"      # 24 pixels per cell
    assert idg.adder_rs_mode(24500, 32, 8192) == "push"      # 0.4 per cell
    assert idg.adder_rs_mode(255, 64, 1024) == "push"
    assert idg.adder_rs_mode(256, 64, 1024) == "pull"
