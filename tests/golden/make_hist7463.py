"""Writes tests/golden/hist7463.json: the rank histogram of all C(52,7) hands, computed by the CPU oracle
(oracle/phe_oracle.c, restated phevaluator algorithm validated against the first-principles evaluator).
Run from the repo root: python tests/golden/make_hist7463.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

h9, h = O.enum7()
assert h9.tolist() == O.K1_HIST9
json.dump({"hist9": h9.tolist(), "hist7463": h.tolist()}, open(os.path.join(ROOT, "tests", "golden", "hist7463.json"), "w"))
print("ok", int(h.sum()))
