"""bench.py contract pieces that can be checked without a GPU: the reference arm (the reference's own code from baseline/_ref on
the host cores) prints one JSON line with the driver's keys, and the roofline helpers find the kernel models they ask for."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line(reference_tree):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--skip-extras"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mc_equity_rollouts_per_s" and d["unit"] == "rollouts/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "batched_mc_equity" and d["config"]["states"] == 4096 and d["config"]["players"] == 6
    assert d["cpu_baseline"]["kind"] == "scaffolding+shim" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "rollouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = d["counts"]
    assert c["win"] + c["tie"] + c["loss"] > 0 and 0.05 < c["equity"] < 0.5          # six-max: a random hand wins about one time in six


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=120,
                         cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_roofline_helpers_and_kernel_models():
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    for kernel in ("k_equity_mc_smem", "k_enum7", "k_eval7_smem"):
        km = bench.kernel_model(kernel)
        assert km.get("lane_ops_per_unit", 0) > 0 and km.get("source", "").startswith("profiles/ncu_"), kernel
    r = bench.int_roofline("k_equity_mc_smem", 3.9e10, "rollout", 105.0, 1295.7)
    assert r["bound"] == "int_issue" and 0 < r["frac"] <= 1.0 and r["frac_vs_reference_model"] > 1.0 and r["unit"] == "T lane-ops/s"
    states = np.array([[0b11 | (0b111 << 16), 0b11]], dtype=np.int64)       # hero + a three-card flop: d = 2 + 10 dealt cards for five opponents
    assert bench.w_rollout_reference_model(states, 5) == 6 * 150 + 70 * 3 + 12 * 12 + 4 * 6
