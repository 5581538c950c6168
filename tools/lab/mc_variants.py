"""A/B timing of Monte-Carlo kernel variants in ONE GPU call.

  python tools/lab/mc_variants.py build            (build container: nvcc cross-compiles every variant into lib/variants/)
  python tools/lab/mc_variants.py run [rollouts]   (GPU box: times each variant in its own process, prints rate + checksum)

Every variant keeps the dealing stream, so the win-count checksum of the config-3 batch must not change."""
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
VARIANTS = {
    "default": [],                                                   # lock-step kernel: board-shared flush test, spare slot row, shared base in a register, items interleaved over the states
    "pool": ["-DPHE_MC_POOL"],                                       # k_equity_mc_pool: dealing decoupled from evaluation
    "pool_skip_reads": ["-DPHE_MC_POOL", "-DPHE_POOL_SKIP_READS"],   # ... finished lanes skip the pair-table read (fewer wavefronts, more instructions)
    "remat_base": ["-DPHE_MC_REMAT_BASE"],                           # let ptxas rebuild the shared-window base at every use
    "generic_eval": ["-DPHE_MC_GENERIC_EVAL"],                       # per-hand flush detection (round-1 evaluator)
    "round1_like": ["-DPHE_MC_GENERIC_EVAL", "-DPHE_MC_REMAT_BASE"],
    "pipe_balance": ["-DPHE_PIPE_BALANCE"],                          # disjoint ORs / cursor adds as IMAD on the FMA pipe
}
b = importlib.import_module("texasholdempocker-rl_b200.build")
VDIR = os.path.join(b.LIBDIR, "variants")


def path(name):
    return os.path.join(VDIR, name, "libphe_b200.so")


if sys.argv[1] == "build":
    for name, flags in VARIANTS.items():
        print(name, b.build_library(force=True, extra_flags=flags, out=path(name)))
else:
    R = sys.argv[2] if len(sys.argv) > 2 else "200000"
    names = sys.argv[3:] or list(VARIANTS)
    for name in names:
        env = dict(os.environ, PHE_B200_LIB=path(name))
        script = os.path.join(ROOT, "tools", "lab", os.environ["PHE_MC_SCRIPT"]) if os.environ.get("PHE_MC_SCRIPT") else os.path.join(ROOT, "tools", "mc_time.py")
        r = subprocess.run([sys.executable, script, R], env=env, capture_output=True, text=True)
        print(json.dumps({"variant": name, "flags": VARIANTS.get(name), "out": r.stdout.strip().splitlines()[-1:] or r.stderr.strip().splitlines()[-3:]}))
