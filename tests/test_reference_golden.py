"""Golden vectors produced by the reference's own code (tests/golden/make_golden.py, run in the build
container): winner sets of GameEnv.get_winner_set_v2 / generate_game_result, the legacy in-tree evaluator
GameEnv.get_winner_set (independent of the phevaluator shim), and simulate_processes result tuples.
CPU part: the oracle reproduces them.  GPU part (marked gpu): the CUDA path reproduces them."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

GOLD = json.load(open(os.path.join(GOLDEN, "reference_run.json")))
N_EXIST = {0: 2, 1: 5, 2: 6, 3: 7}


def has_wheel(cards):
    return {12, 0, 1, 2, 3} <= {c >> 2 for c in cards}


def holes_of(s):
    return [s["ids"][5 + 2 * p: 7 + 2 * p] for p in range(s["P"])]


def test_card_ids_match_reference_deck(phe):
    assert GOLD["card_ids"] == list(range(52))          # deck[i] has id i (env/game.py:14-18)
    assert [phe.card_to_evaluator(c) for c in phe.DECK] == GOLD["card_ids"]


def test_oracle_matches_reference_v2_winner_sets(oracle):
    for s in GOLD["showdowns"]:
        assert oracle.winner_mask(s["ids"][:5], holes_of(s)) == s["v2"]


def test_oracle_matches_legacy_in_tree_evaluator(oracle):
    """get_winner_set (v1) shares no code with the shim.  It agrees with the oracle everywhere except for its two
    known defects: A-2-3-4-5 is not a straight for it and it ignores the second kicker of trips."""
    differ = 0
    for s in GOLD["showdowns"]:
        w, ranks = oracle.winner_mask(s["ids"][:5], holes_of(s), with_ranks=True)
        if w != s["v1"]:
            differ += 1
            wheel = any(has_wheel(s["ids"][:5] + h) for h in holes_of(s))
            trips = 1610 <= int(ranks.min()) <= 2467
            assert wheel or trips, s
    assert differ <= 0.03 * len(GOLD["showdowns"])


def test_generate_game_result_matches_reference(phe):
    vals = {-1.0: phe.GamePlayerResult.LOSE, 0.0: phe.GamePlayerResult.EVEN, 1.0: phe.GamePlayerResult.WIN}
    for s in GOLD["showdowns"][:500]:
        res = {p: phe.GamePlayerResult.LOSE for p in range(s["P"])}
        phe.GameEnv.generate_game_result({p for p in range(s["P"]) if s["v2"] >> p & 1}, res)
        assert [res[p] for p in range(s["P"])] == [vals[v] for v in s["result"]]


def _check_tasks(solver):
    """solver(exist_ids, plan) -> (win, tie, loss).  Exhaustive plans must match the reference tuple exactly;
    plans with a random step must match n exactly and the sum within 4 sigma of the two estimators."""
    from oracle import oracle as O
    for t in GOLD["tasks"]:
        for name, plans in (("default", O.DEFAULT_PLANS), ("debug", O.DEBUG_PLANS)):
            plan, n = plans[t["round"]]
            w, ti, l = solver(t["ids"], plan)
            ref = t[name]
            assert w + ti + l == n == ref["n"]
            if all(step[0] == 1 for step in plan):
                assert float(w - l) == ref["sum"], (t, name)
            else:
                sigma = (2.0 / n) ** 0.5        # each deal scores -1/0/+1: variance <= 1 for either estimator
                assert abs((w - l) / n - ref["sum"] / n) <= 4 * sigma + 1e-9, (t, name)


def test_oracle_matches_reference_simulate_processes(oracle):
    _check_tasks(lambda ids, plan: oracle.equity_plan(ids, plan, seed=7)[:3])


def test_pyport_matches_reference_scaffolding(oracle):
    """The pure-Python port timed by bench.py --impl reference returns the reference's numbers."""
    from oracle import pyport
    for t in GOLD["tasks"]:
        if t["round"] == 3:
            assert pyport.run_plan(oracle.DEFAULT_PLANS[3][0], list(t["ids"])) == (t["default"]["sum"], 990)
    for s in GOLD["showdowns"][:300]:
        board = s["ids"][:5]
        w = pyport.get_winner_set_v2(board[:3], board[3:4], board[4:5], {p: h for p, h in enumerate(holes_of(s))})
        assert sum(1 << p for p in w) == s["v2"]


@pytest.mark.gpu
def test_gpu_matches_reference_winner_sets(phe):
    D = phe.DECK
    games = []
    for s in GOLD["showdowns"]:
        ids = s["ids"]
        games.append(([D[i] for i in ids[:3]], [D[ids[3]]], [D[ids[4]]], {p: [D[i] for i in h] for p, h in enumerate(holes_of(s))}))
    got = phe.winner_sets(games)
    for s, w in zip(GOLD["showdowns"], got):
        assert sum(1 << p for p in w) == s["v2"]
    s = GOLD["showdowns"][0]
    assert phe.GameEnv.get_winner_set_v2(*games[0]) == {p for p in range(s["P"]) if s["v2"] >> p & 1}


@pytest.mark.gpu
def test_gpu_matches_reference_simulate_processes(phe):
    def solve(ids, plan):
        return tuple(int(x) for x in phe.equity_plan(np.array([ids], dtype=np.uint8), plan, seed=7)[0])
    _check_tasks(solve)


@pytest.mark.gpu
def test_gpu_service_returns_reference_tuples(phe):
    import queue
    from oracle import oracle as O
    D = phe.DECK
    q_in, q_out = queue.Queue(), queue.Queue()
    exact = [(i, t) for i, t in enumerate(GOLD["tasks"]) if t["round"] >= 2]
    for i, t in exact:
        ids = t["ids"]
        q_in.put((i, "p", t["round"], [D[x] for x in ids[2:5]], [D[ids[5]]], [D[ids[6]]] if t["round"] == 3 else None, [D[x] for x in ids[:2]]))
    q_in.put(None)
    phe.simulate_processes(q_in, q_out, {k: v for k, v in O.DEFAULT_PLANS.items()}, 0, "WARNING")
    got = {}
    while not q_out.empty():
        r = q_out.get()
        got[r[0]] = r
    for i, t in exact:
        assert got[i] == (i, "p", t["round"], t["default"]["n"], t["default"]["sum"])


# ---- unusual plan shapes, run through the reference's own simulate_processes (tests/golden/make_plan_golden.py) ----
PLAN_GOLD = json.load(open(os.path.join(GOLDEN, "plan_run.json")))


def _check_custom_plans(solver, num_deals):
    for t in PLAN_GOLD:
        plan = [tuple(r) for r in t["plan"]]
        assert num_deals(len(t["ids"]), plan) == t["n"]
        w, ti, l = solver(t["ids"], plan)
        assert w + ti + l == t["n"]
        if all(r[0] == 1 for r in plan):
            assert float(w - l) == t["sum"], t            # exhaustive, any segmentation: exact
        else:
            assert abs((w - l) - t["sum"]) / t["n"] <= 4 * (2.0 / t["n"]) ** 0.5, t


def test_oracle_matches_reference_on_custom_plans(oracle):
    _check_custom_plans(lambda ids, plan: oracle.equity_plan(ids, plan, seed=3)[:3],
                        lambda n_exist, plan: oracle.equity_plan(list(range(n_exist)), plan)[3])


def test_plan_parser_counts_match_reference_on_custom_plans(phe):
    for t in PLAN_GOLD:
        assert phe.plan_num_deals(len(t["ids"]), [tuple(r) for r in t["plan"]]) == t["n"]


@pytest.mark.gpu
def test_gpu_matches_reference_on_custom_plans(phe):
    _check_custom_plans(lambda ids, plan: tuple(int(x) for x in phe.equity_plan(np.array([ids], dtype=np.uint8), plan, seed=3)[0]),
                        phe.plan_num_deals)
