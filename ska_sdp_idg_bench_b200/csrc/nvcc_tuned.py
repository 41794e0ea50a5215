#!/usr/bin/env python
"""nvcc -c with one extra step: sass_tune.py on the cubin between ptxas and fatbinary.

    nvcc_tuned.py [--tune-kernels REGEX] [--no-tune] -- <nvcc> <nvcc arguments of a -c compile>

nvcc has no hook between its internal steps, so this replays the command list that
`nvcc -dryrun` prints (preprocess, cicc, ptxas, fatbinary, cudafe++, host compile) and
runs csrc/sass_tune.py on the .cubin right after ptxas wrote it.  Everything else is
nvcc's own pipeline, so the object file is what `nvcc -c` would have produced with a
few control bits of some FFMA2 instructions changed (see sass_tune.py).

The bit positions sass_tune.py edits were established for ptxas 12.9 (TUNED_TOOLKITS).  With any other
toolkit, or when the replay or the tuning step fails for any reason, the object is built by the plain
nvcc command instead, with a warning: the tuning only buys a few percent on the FP32 kernels and the
untuned object computes the same bits (tests/test_gpu_parity.py::test_sass_tuned_equals_untuned_bitwise).
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TUNED_TOOLKITS = ("12.9",)   # CUDA releases whose ptxas encodes FFMA2 as sass_tune.py expects


def toolkit_release(nvcc: str) -> str | None:
    try:
        out = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout
    except OSError:
        return None
    m = re.search(r"release (\d+\.\d+)", out)
    return m.group(1) if m else None


def main(argv) -> int:
    if "--" not in argv:
        print(__doc__)
        return 2
    split = argv.index("--")
    opts, cmd = argv[1:split], argv[split + 1:]
    tune = "--no-tune" not in opts and os.environ.get("IDGB200_NO_SASS_TUNE", "0") != "1"
    if tune and toolkit_release(cmd[0]) not in TUNED_TOOLKITS:
        sys.stderr.write(f"nvcc_tuned: CUDA {toolkit_release(cmd[0])} is not in {TUNED_TOOLKITS}: building untuned\n")
        tune = False
    if not tune:
        return subprocess.run(cmd).returncode
    rc = tuned(opts, cmd)
    if rc != 0:
        sys.stderr.write("nvcc_tuned: WARNING: the tuned build failed, falling back to the plain nvcc object\n")
        return subprocess.run(cmd).returncode
    return 0


def tuned(opts, cmd) -> int:
    tune = True
    kernels = opts[opts.index("--tune-kernels") + 1] if "--tune-kernels" in opts else "."

    dry = subprocess.run(cmd + ["-dryrun"], capture_output=True, text=True)
    if dry.returncode != 0:
        sys.stderr.write(dry.stderr)
        return dry.returncode
    lines = [ln[3:] for ln in dry.stderr.splitlines() if ln.startswith("#$ ")]
    env = dict(os.environ)
    for ln in lines:
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)=(.*)$", ln)
        if m:  # nvcc's variable definitions (only $CICC_PATH is referenced by later steps)
            env[m.group(1)] = m.group(2).strip().strip('"')
            continue
        r = subprocess.run(["bash", "-c", ln], env=env, stderr=subprocess.DEVNULL if ln.startswith("rm ") else None)
        if r.returncode != 0 and not ln.startswith("rm "):  # nvcc's own clean-up steps may find nothing
            sys.stderr.write(f"nvcc_tuned: step failed: {ln[:200]}\n")
            return r.returncode
        if tune and re.match(r"^\s*ptxas\s", ln):
            m = re.search(r'-o\s+"([^"]+\.cubin)"', ln)
            if not m:
                sys.stderr.write("nvcc_tuned: could not find the cubin in the ptxas step\n")
                return 1
            r = subprocess.run([sys.executable, os.path.join(HERE, "sass_tune.py"), m.group(1),
                                "--kernels", kernels])
            if r.returncode != 0:
                return r.returncode
    # nvcc removes its intermediates itself; the replay has to do it by hand
    import glob
    for prefix in {m for ln in lines for m in re.findall(r"(/tmp/tmpxft_[0-9a-f]+_[0-9a-f]+)-", ln)}:
        for f in glob.glob(prefix + "-*"):
            try:
                os.remove(f)
            except OSError:
                pass
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
