"""Extracts the metrics DESIGN.md / bench.py quote from an .ncu-rep into a small text file.
Usage: python tools/ncu_summary.py <file.ncu-rep> > profiles/<name>.txt"""
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
