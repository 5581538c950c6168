"""Generates tests/golden/plan_run.json: the reference's own simulate_processes (env/workflow.py:102-227, unmodified;
phevaluator shim as in make_golden.py) driven with NON-default simulation_recurrent_params shapes, to pin how plan rows
are interpreted (segments, ordered restarts, the pruning test, a random last step).  Build container only."""
import json
import os
import queue
import random
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")

from env.game import deck  # noqa: E402
import env.workflow as workflow  # noqa: E402
from tools import interrupt  # noqa: E402

# round -> list of (plan rows, expected deal count) ; every plan reaches nine cards
CUSTOM = {
    3: [([(1, 0, True), (1, 0, True)], 45 * 44),                       # ordered villain cards: two single-card segments
        ([(2, 60, True)], 60)],                                        # pure Monte Carlo
    2: [([(1, 0, False), (1, 0, False), (1, 0, True)], 15180),         # one 3-card segment: C(46,3) unordered {river, v1, v2}
        ([(1, 0, True), (2, 30, True)], 46 * 30),                      # enumerate river, sample villain
        ([(1, 0, False), (1, 0, True), (1, 0, True)], 1035 * 44)],     # unordered (river, v1) pair then ordered v2
    1: [([(1, 0, False), (1, 0, True), (2, 7, True)], 1081 * 7),       # unordered (turn, river), sample villain
        ([(4, 500, True)], 500)],
    0: [([(1, 0, True), (6, 40, True)], 50 * 40)],
}
rng = random.Random(5)
random.seed(3)
out = []
for rnd, plans in CUSTOM.items():
    n_exist = {0: 2, 1: 5, 2: 6, 3: 7}[rnd]
    for plan, n_expected in plans:
        for rep in range(3):
            ids = rng.sample(range(52), n_exist)
            hand = [deck[x] for x in ids[0:2]]
            flop = [deck[x] for x in ids[2:5]] if rnd >= 1 else None
            turn = [deck[ids[5]]] if rnd >= 2 else None
            river = [deck[ids[6]]] if rnd >= 3 else None
            q_in, q_out = queue.Queue(), queue.Queue()
            th = threading.Thread(target=workflow.simulate_processes, args=(q_in, q_out, {rnd: ([list(s) for s in plan], n_expected)}, 0, "WARNING"), daemon=True)
            th.start()
            q_in.put((0, "p", rnd, flop, turn, river, hand))
            r = q_out.get(timeout=600)
            interrupt._interrupted = True
            th.join(timeout=5)
            interrupt._interrupted = False
            out.append({"round": rnd, "ids": ids, "plan": [[g, k, int(e)] for g, k, e in plan], "n": r[3], "sum": r[4]})
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "plan_run.json"), "w"))
print(len(out), "tasks")
