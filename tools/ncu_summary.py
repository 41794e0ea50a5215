"""Extracts the metrics DESIGN.md / bench.py quote from an .ncu-rep.

  python tools/ncu_summary.py <file.ncu-rep> > profiles/<name>.txt           text summary of every kernel in it
  python tools/ncu_summary.py --json profiles/roofline_constants.json name=file.ncu-rep[:subgrids] ...
        the per-kernel constants bench.py reads (DRAM bytes per subgrid, pipe utilisations, instruction counts);
        `subgrids` = subgrids the captured launch processed (default: its grid size).  Deleting the JSON and
        re-running this on the .ncu-rep files kept under profiles/ reproduces it exactly."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
    "launch__shared_mem_per_block", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_thread_inst_executed_op_ffma", "smsp__sass_thread_inst_executed_op_fp32",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__inst_executed_op_shared",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
]
ALSO = ["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
PCT = {
    "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "lsu_smem_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "tensor_smem_wavefronts_pct": "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_base(value, unit):
    v = float(value.replace(",", ""))
    unit = unit.split("/")[0]
    for pre, f in (("G", 1e9), ("M", 1e6), ("K", 1e3), ("k", 1e3), ("m", 1e-3), ("u", 1e-6), ("n", 1e-9)):
        if unit.startswith(pre) and unit not in ("ms",) and len(unit) > 1 and unit[1:] in ("byte", "s", "second", "hz", "Hz"):
            return v * f
    if unit == "ms":
        return v * 1e-3
    return v


if len(sys.argv) > 1 and sys.argv[1] == "--json":
    import json
    import os
    dest, out_json = sys.argv[2], {}
    for spec in sys.argv[3:]:
        name, rest = spec.split("=", 1)
        rep, _, sub = rest.partition(":")
        hdr, units, rows = raw_rows(rep)
        r = rows[-1]

        def get(metric):
            i = hdr.index(metric)
            return to_base(r[i], units[i])

        grid = int(r[hdr.index("Grid Size")].strip("()").split(",")[0])
        subgrids = int(sub) if sub else grid
        dram = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
        e = {"source": os.path.relpath(rep, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
             "kernel": r[hdr.index("Kernel Name")][:80], "grid": grid, "subgrids": subgrids,
             "duration_ms": get("gpu__time_duration.sum") * 1e3,
             "dram_bytes": dram, "dram_bytes_per_subgrid": dram / subgrids,
             "inst_executed": get("smsp__inst_executed.sum"),
             "registers_per_thread": get("launch__registers_per_thread"),
             "shared_mem_per_block": get("launch__shared_mem_per_block")}
        for k, m in PCT.items():
            e[k] = get(m)
        out_json[name] = e
    json.dump(out_json, open(dest, "w"), indent=1, sort_keys=True)
    print(json.dumps(out_json, indent=1, sort_keys=True))
    sys.exit(0)

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(f"== {name[:110]}")
    print(f"   grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
    for h, u, v in zip(hdr, units, r):
        if h in ALSO or any(h.startswith(w) for w in WANT) and not h.endswith((".max", ".min")) and ".max." not in h \
                and ".min." not in h and ".sum.pct" not in h:
            print(f"   {h:92s} {v:>16s} {u}")
