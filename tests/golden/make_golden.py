"""Generates tests/golden/reference_run.json by running the REFERENCE'S OWN CODE in the build container.

  python tests/golden/make_golden.py      (needs /root/reference; not runnable on the GPU box)

What runs unmodified from /root/reference: env/game.py (GameEnv.get_winner_set_v2, generate_game_result, the
legacy in-tree evaluator GameEnv.get_winner_set), env/workflow.py (simulate_processes), env/cards.py,
utils/pokerhandevaluator_utils.py.  What is substituted: the un-installable PyPI package `phevaluator`
(oracle/shim/phevaluator, restated algorithm) and the import-only stubs objgraph / mem_top.

Sections written:
  showdowns   random 2..6-player showdowns: winner sets from get_winner_set_v2 (reference scaffolding + shim)
              and from get_winner_set (legacy v1, no shim involved -- an independent pin; v1 mis-ranks
              wheel straights, SURVEY 8c K5, so those deals are flagged)
  results     generate_game_result outputs for the same deals
  tasks       simulate_processes fed through queue.Queue with the default and debug plans: result tuples
  card_ids    card_to_evaluator(deck[i]) for the 52 deck cards
"""
import json
import os
import queue
import random
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from env.game import GameEnv, deck  # noqa: E402
from env.constants import GamePlayerResult  # noqa: E402
from utils.pokerhandevaluator_utils import card_to_evaluator  # noqa: E402
import env.workflow as workflow  # noqa: E402
from tools import interrupt  # noqa: E402

rng = random.Random(20261019)
out = {"card_ids": [card_to_evaluator(c) for c in deck]}

# ---- showdowns ----
shows = []
for i in range(3000):
    P = rng.choice([2, 2, 2, 3, 4, 5, 6])
    ids = rng.sample(range(52), 5 + 2 * P)
    if i % 10 == 0:   # force some wheel / straight / flush heavy boards
        ids = rng.choice([[48, 1, 6, 11, 12], [3, 7, 11, 15, 30], [32, 37, 42, 47, 0]]) + ids
        seen, ded = set(), []
        for x in ids:
            if x not in seen:
                seen.add(x); ded.append(x)
        ids = ded[: 5 + 2 * P]
    flop = sorted(deck[x] for x in ids[0:3])
    turn, river = [deck[ids[3]]], [deck[ids[4]]]
    players = {p: sorted(deck[x] for x in ids[5 + 2 * p: 7 + 2 * p]) for p in range(P)}
    w2 = GameEnv.get_winner_set_v2(flop, turn, river, players)
    try:
        w1 = GameEnv.get_winner_set(flop, turn, river, players)
        w1m = sum(1 << p for p in w1)
    except Exception as e:  # noqa: BLE001
        w1m = -1
    res = {p: GamePlayerResult.LOSE for p in range(P)}
    GameEnv.generate_game_result(w2, res)
    shows.append({"ids": ids, "P": P, "v2": sum(1 << p for p in w2), "v1": w1m, "result": [res[p].value for p in range(P)]})
out["showdowns"] = shows

# ---- simulate_processes ----
DEFAULT = {0: ([(1, 0, False), (1, 0, False), (1, 0, True), (4, 5, True)], 98000),
           1: ([(1, 0, True), (1, 0, True), (2, 45, True)], 97290),
           2: ([(1, 0, True), (1, 0, False), (1, 0, True)], 45540),
           3: ([(1, 0, False), (1, 0, True)], 990)}
DEBUG = {0: ([(7, 100, True)], 100), 1: ([(4, 100, True)], 100), 2: ([(3, 100, True)], 100), 3: ([(2, 100, True)], 100)}


def run_tasks(plans, tasks):
    q_in, q_out = queue.Queue(), queue.Queue()
    th = threading.Thread(target=workflow.simulate_processes, args=(q_in, q_out, {k: ([list(s) for s in v[0]], v[1]) for k, v in plans.items()}, 0, "WARNING"), daemon=True)
    th.start()
    for t in tasks:
        q_in.put(t)
    res = [q_out.get(timeout=600) for _ in tasks]
    interrupt._interrupted = True
    th.join(timeout=5)
    interrupt._interrupted = False
    return res


def make_task(gid, rnd, ids):
    hand = [deck[x] for x in ids[0:2]]
    flop = [deck[x] for x in ids[2:5]] if rnd >= 1 else None
    turn = [deck[ids[5]]] if rnd >= 2 else None
    river = [deck[ids[6]]] if rnd >= 3 else None
    return (gid, "p", rnd, flop, turn, river, hand)


tasks, metas = [], []
fixed = [(3, [51, 50, 0, 21, 32, 6, 31]), (3, [22, 26, 30, 34, 0, 1, 47]), (2, [51, 50, 0, 21, 32, 6]), (1, [51, 50, 0, 21, 32]), (0, [51, 50])]
for rnd, ids in fixed:
    metas.append({"round": rnd, "ids": ids})
for rnd, cnt in ((3, 40), (2, 6), (1, 2), (0, 2)):
    for _ in range(cnt):
        metas.append({"round": rnd, "ids": rng.sample(range(52), {0: 2, 1: 5, 2: 6, 3: 7}[rnd])})
for i, m in enumerate(metas):
    tasks.append(make_task(i, m["round"], m["ids"]))
random.seed(1)
for r in run_tasks(DEFAULT, tasks):
    metas[r[0]]["default"] = {"n": r[3], "sum": r[4]}
for r in run_tasks(DEBUG, tasks):
    metas[r[0]]["debug"] = {"n": r[3], "sum": r[4]}
out["tasks"] = metas
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "reference_run.json"), "w"))
bad = [s for s in shows if s["v1"] != s["v2"]]
print("showdowns", len(shows), "v1!=v2:", len(bad), "tasks", len(metas))
