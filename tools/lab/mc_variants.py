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
    "default": [],                                                   # k_equity_mc_pool: dealing decoupled from evaluation
    "pool_skip_reads": ["-DPHE_POOL_SKIP_READS"],                    # ... finished lanes skip the pair-table read (fewer wavefronts, more instructions)
    "v1": ["-DPHE_MC_V1"],                                           # lock-step kernel: board-shared flush test, spare slot row, shared base in a register
    "v1_round1_like": ["-DPHE_MC_V1", "-DPHE_MC_GENERIC_EVAL", "-DPHE_MC_REMAT_BASE"],
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
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mc_time.py"), R], env=env, capture_output=True, text=True)
        print(json.dumps({"variant": name, "flags": VARIANTS.get(name), "out": r.stdout.strip().splitlines()[-1:] or r.stderr.strip().splitlines()[-3:]}))
