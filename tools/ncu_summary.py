"""Prints / saves the key counters of an .ncu-rep (run where ncu is installed; no GPU needed).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [profiles/out.json]"""
import csv
import json
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smsp__inst_executed_op_shared_atom.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__cycles_active.min",
        "sm__cycles_active.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (r[i] + " " + units[i]).strip()
    for i, name in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in name and "not_issued" not in name:
            d[name.replace("smsp__pcsamp_warps_issue_", "")] = r[i]
    out.append(d)
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
for d in out:
    for k, v in d.items():
        print(f"{k:70s} {v}")
    print()
