"""Builds profiles/kernel_models.json: the executed-work model of each hot kernel's OWN algorithm, from the committed ncu
summaries (tools/ncu_summary.py output).  bench.py multiplies `lane_ops_per_unit` by the measured units/s to get the
achieved useful lane-ops/s it reports against the measured integer-issue peak (roofline.frac <= 1).

  python tools/kernel_model.py            (no GPU needed; rerun after committing a new ncu summary)

lane_ops_per_unit = smsp__inst_executed.sum x smsp__thread_inst_executed_per_inst_executed.ratio / units per profiled launch
(thread-level instructions of active lanes: predicated-off and idle lanes are not counted as work)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# kernel -> (summary file candidates newest first, kernel-name substring, units in the profiled launch, unit)
SOURCES = {
    "k_equity_mc_smem": (["ncu_mc_r02.json", "ncu_mc_r02_lockstep.json", "ncu_mc_r01.json"], "k_equity_mc_smem", 4096 * 100000, "rollout"),   # tools/prof_driver.py mc
    "k_equity_mc_pool": (["ncu_mc_r02_pool.json"], "k_equity_mc_pool", 4096 * 100000, "rollout"),                     # the decoupled experiment (-DPHE_MC_POOL)
    "k_enum7": (["ncu_enum7_r02.json", "ncu_enum7_r01.json"], "k_enum7", 133784560, "hand"),                          # tools/prof_driver.py enum
    "k_eval7_smem": (["ncu_eval7_r02.json", "ncu_eval7_r01.json"], "k_eval7_smem", 1 << 25, "hand"),
}


def num(d, key):
    try:
        return float(str(d[key]).split()[0])
    except Exception:
        return None


def kb(d, key):
    v = d.get(key)
    if v is None:
        return 0.0
    x, unit = float(str(v).split()[0]), (str(v).split() + [""])[1].lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)


out = {}
for kernel, (files, sub, units, unit) in SOURCES.items():
    for f in files:
        path = os.path.join(ROOT, "profiles", f)
        if not os.path.exists(path):
            continue
        rows = [r for r in json.load(open(path)) if sub in r.get("Kernel Name", "")]
        if not rows:
            continue
        r = rows[-1]
        inst, lanes = num(r, "smsp__inst_executed.sum"), num(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
        wf = num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        out[kernel] = {
            "unit": unit, "units_in_profiled_launch": units, "source": f"profiles/{f}",
            "lane_ops_per_unit": inst * lanes / units, "warp_inst_per_32_units": inst * 32 / units,
            "smem_wavefronts_per_32_units": wf * 32 / units if wf else None,
            "dram_bytes_per_launch": kb(r, "dram__bytes_read.sum") + kb(r, "dram__bytes_write.sum"),
            "hw": {"duration": r.get("gpu__time_duration.sum"), "registers": r.get("launch__registers_per_thread"),
                   "smem_per_cta": r.get("launch__shared_mem_per_block_dynamic"),
                   "issue_slot_util_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                   "fma_pipe_pct": num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                   "xu_pipe_pct": num(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                   "active_lanes_per_inst": lanes, "smem_wavefronts": wf,
                   "smem_bank_conflict_wavefronts": num(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                   "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")},
        }
        break
path = os.path.join(ROOT, "profiles", "kernel_models.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
for k, v in out.items():
    print(f"{k:20s} {v['lane_ops_per_unit']:.1f} lane-ops/{v['unit']}  {v['warp_inst_per_32_units']:.1f} warp-inst/32  ({v['source']})")
print("wrote", path)
