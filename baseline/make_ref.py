"""Recipe for baseline/_ref/: an UNMODIFIED copy of the reference's Python packages, made from /root/reference where
it lies (build container only).  baseline/_ref/ is git-ignored (no reference source enters the history) but is not
gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist.

    python baseline/make_ref.py          (also run by __graft_entry__.build() when /root/reference is present)

What uses it (never the product):
  * bench.py --impl reference      times the reference's own get_winner_set_v2 / simulate_processes / SingleThreadMCTS
  * tests/test_insitu.py           runs the reference's own GameEnv / Env / MCTS with this library patched in
Both put oracle/shim (the restated phevaluator package, objgraph / mem_top stubs) on sys.path next to it: the PyPI
wheel `phevaluator` is not installable offline.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("env", "rl", "utils", "tools", "config")      # what rl/MCTS.py, env/workflow.py and their imports need
KEEP = (".py", ".yml", ".yaml")


def make(force=False):
    if not os.path.isdir(SRC):
        return None
    stamp = os.path.join(DST, "MANIFEST.json")
    if os.path.exists(stamp) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for pkg in PACKAGES:
        for dirpath, _, files in os.walk(os.path.join(SRC, pkg)):
            for f in sorted(files):
                if not f.endswith(KEEP):
                    continue
                src = os.path.join(dirpath, f)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                with open(src, "rb") as fh:
                    manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(stamp, "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return DST


def paths():
    """sys.path entries that make `import env.game`, `import rl.MCTS` ... resolve to the unmodified reference."""
    return [os.path.join(ROOT, "oracle", "shim"), DST]


if __name__ == "__main__":
    out = make(force=True)
    print("reference absent, nothing done" if out is None else f"copied {SRC} -> {out}")
    sys.exit(0)
