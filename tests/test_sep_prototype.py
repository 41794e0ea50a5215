"""CPU checks of tools/sep_prototype.py, the numpy statement of the row-column formulation that the choice of operand
precisions in csrc/degridder_sep.cu rests on: its e4m3 rounding is the hardware's (torch's
float8_e4m3fn conversion is the reference here), and on the reference's own correctness shape the degridder with an fp16
product plus e4m3 cross products is as close to the float64 sums as the reference's fp32 code, while a single fp16 product
is not."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import sep_prototype as sp  # noqa: E402
from oracle_lib import oracle  # noqa: E402


def test_e4m3_emulation_matches_torch():
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.standard_normal(20000) * 10.0 ** rng.uniform(-4, 2.6, 20000),
                        [0.0, 448.0, 447.9, 1.0625, 1.1875, 2.0 ** -9, 2.0 ** -10, 3 * 2.0 ** -10, 0.017578125, -0.0009765625]])
    v = v[np.abs(v) <= 448.0].astype(np.float32)
    want = torch.from_numpy(v).to(torch.float8_e4m3fn).to(torch.float32).numpy()
    assert np.array_equal(sp.e4m3(v).astype(np.float32), want)


def test_degridder_operand_precisions_on_config1():
    o = oracle()
    p = o.make_problem()
    ref32, ref64 = o.degridder(p), o.degridder_f64(p)

    def err(x):
        return np.abs(x - ref64).max() / np.abs(ref64).max()

    e_cpu, e_one, e_fp8 = err(ref32), err(sp.degridder_sep(p, "f16x1")), err(sp.degridder_sep(p, "fp8"))
    print(f"config 1 degridder vs float64: cpu fp32 {e_cpu:.2e}, one fp16 product {e_one:.2e}, fp16 + e4m3 {e_fp8:.2e}")
    assert e_fp8 <= 1.25 * e_cpu
    assert e_one >= 2.5 * e_cpu
