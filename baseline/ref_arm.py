"""Reference arm: the reference's OWN CPU code for the hot path, timed on the box's host cores.

Everything evaluated here is the unmodified checkout in baseline/_ref (baseline/make_ref.py) -- GameEnv.get_winner_set_v2
(env/game.py:663-681), simulate_processes (env/workflow.py:102-227), SingleThreadMCTS (rl/MCTS.py:410-473) -- with
oracle/shim on sys.path for the one thing that cannot be installed offline: the PyPI package `phevaluator`
(restated pure-Python algorithm).  kind = "scaffolding+shim".  Nothing of the product is on this path;
only the synthetic-state generator is shared so that both arms roll the same states.

The reference has no N-player equity loop (simulate_processes.cal_game_result is heads-up, env/workflow.py:130-148), so
the 6-max rollout below is its heads-up loop generalised in its own idiom: `random.sample` of the unseen cards
(env/workflow.py:190, env/game.py:139-140), the N-player generic `GameEnv.get_winner_set_v2` for evaluation + winner
logic, and the +1 / 0 / -1 mapping of env/workflow.py:142-147.
"""
import importlib.util
import logging
import multiprocessing as mp
import os
import random
import signal
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def setup_paths():
    spec = importlib.util.spec_from_file_location("make_ref", os.path.join(HERE, "make_ref.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    if not os.path.exists(os.path.join(mk.DST, "MANIFEST.json")) and mk.make() is None:
        raise RuntimeError("baseline/_ref is missing and /root/reference is not present: run baseline/make_ref.py in the build container")
    for p in reversed(mk.paths()):
        if p not in sys.path:
            sys.path.insert(0, p)
    if ROOT not in sys.path:
        sys.path.append(ROOT)
    return mk.DST


def state_cards(known, hero):
    """u64 card sets (bit 16*suit + rank) -> (hero ids, board ids) with id = rank*4 + suit (deck[id], env/game.py:14-18)."""
    def ids(m):
        return [((p & 15) << 2) | (p >> 4) for p in range(64) if (m >> p) & 1]
    return ids(hero), ids(known & ~hero)


# ---- 6-max Monte-Carlo rollouts on the reference's winner logic ---------------------------------------------------------
def _mc_chunk(job):
    states, rollouts, n_opp, seed = job
    from env.game import GameEnv, deck
    random.seed(seed)
    w = t = l = 0
    for known, hero in states:
        h, b = state_cards(known, hero)
        hero_cards = [deck[i] for i in h]
        board = sorted(deck[i] for i in b)
        seen = set(h + b)
        unknown = [c for i, c in enumerate(deck) if i not in seen]
        need = 5 - len(board) + 2 * n_opp
        for _ in range(rollouts):
            new = random.sample(unknown, need)
            players = {0: hero_cards}
            for o in range(n_opp):
                players[o + 1] = new[2 * o: 2 * o + 2]
            full = board + new[2 * n_opp:]
            winners = GameEnv.get_winner_set_v2(full[:3], full[3:4], full[4:5], players)
            if 0 in winners:
                if len(winners) == 1:
                    w += 1
                else:
                    t += 1
            else:
                l += 1
    return w, t, l


def _init_worker():
    setup_paths()
    logging.disable(logging.CRITICAL)


class Pool:
    def __init__(self, procs):
        self.procs = procs
        self.pool = mp.get_context("fork").Pool(procs, initializer=_init_worker)

    def mc6(self, states, rollouts_per_state, n_opp, seed):
        """One step: `rollouts_per_state` rollouts of every state, split over the workers.  Returns (w, t, l)."""
        per = (len(states) + self.procs - 1) // self.procs
        jobs = [(states[i: i + per], rollouts_per_state, n_opp, seed + i) for i in range(0, len(states), per)]
        w = t = l = 0
        for a, b, c in self.pool.map(_mc_chunk, jobs):
            w, t, l = w + a, t + b, l + c
        return w, t, l

    def map(self, fn, jobs):
        return self.pool.map(fn, jobs)

    def close(self):
        self.pool.close()
        self.pool.join()


# ---- GameEnv.get_winner_set_v2 in a loop (evals/s = 2 * calls/s) ---------------------------------------------
def _winner_chunk(job):
    n, seed = job
    from env.game import GameEnv, deck
    rng = random.Random(seed)
    deals = [rng.sample(deck, 9) for _ in range(256)]
    t0 = time.perf_counter()
    for i in range(n):
        d = deals[i & 255]
        GameEnv.get_winner_set_v2(d[4:7], d[7:8], d[8:9], {0: d[0:2], 1: d[2:4]})
    return n, time.perf_counter() - t0


def winner_set_rate(pool, calls_per_worker, workers):
    res = pool.map(_winner_chunk, [(calls_per_worker, 7 + i) for i in range(workers)])
    return 2 * sum(n for n, _ in res) / max(dt for _, dt in res)


# ---- simulate_processes itself, as main.py:52-53 starts it -----------------------------------------------------------
def simulate_processes_rate(tasks, plans, workers, timeout=600):
    """Unmodified env.workflow.simulate_processes worker processes fed through multiprocessing queues.
    Returns (results, seconds)."""
    from env.workflow import simulate_processes
    ctx = mp.get_context("fork")
    in_q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=simulate_processes, args=(in_q, out_q, plans, i, logging.ERROR), daemon=True) for i in range(workers)]
    for p in procs:
        p.start()
    time.sleep(0.3)
    t0 = time.perf_counter()
    for task in tasks:
        in_q.put(task)
    out = []
    try:
        while len(out) < len(tasks):
            out.append(out_q.get(timeout=timeout))
    finally:
        dt = time.perf_counter() - t0
        for p in procs:
            if p.pid:
                os.kill(p.pid, signal.SIGTERM)          # tools/interrupt.py turns SIGTERM into a clean return
        for p in procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()
    return out, dt


def random_tasks(n_games, rounds, seed):
    """Heads-up equity tasks in the reference's wire format (env/env.py:240)."""
    from env.game import deck
    rng = random.Random(seed)
    tasks = []
    for g in range(n_games):
        d = rng.sample(deck, 9)
        for p in range(2):
            hand = sorted(d[2 * p: 2 * p + 2])
            flop, turn, river = sorted(d[4:7]), d[7:8], d[8:9]
            for r in rounds:
                tasks.append((g, f"player_{p}", r, flop if r >= 1 else None, turn if r >= 2 else None, river if r >= 3 else None, hand))
    return tasks


# ---- the reference's own MCTS (BASELINE config 4) on the CPU ----------------------------------------------------------
def reference_params():
    import yaml
    from tools.param_parser import update_concurrent
    cfg = os.path.join(setup_paths(), "config")
    default = yaml.safe_load(open(os.path.join(cfg, "default.yml"), encoding="utf-8"))
    debug = yaml.safe_load(open(os.path.join(cfg, "debug.yml"), encoding="utf-8"))
    return update_concurrent(default.copy(), debug.copy())


def build_model_and_env(seed=0):
    """config/debug.yml over config/default.yml: the reference's model (random init) and a heads-up 25/50 Env."""
    import torch
    from env.env import Env
    from rl.AlphaGoZero import AlphaGoZero
    params = reference_params()
    mp_ = dict(params["model_param_dict"], save_train_data=False)
    torch.manual_seed(seed)
    model = AlphaGoZero(**mp_).model.eval()
    env = Env(None, mp_["num_output_class"], mp_["historical_action_per_round"], small_blind=params["small_blind"], big_blind=params["big_blind"],
              ignore_all_async_tasks=True)
    return params, mp_, model, env


def make_search(kind, params, mp_, model, n_simulation, **kw):
    """kind 'reference': rl.MCTS.SingleThreadMCTS as py_poker_engine/alpha_go_zero_player.py:56-68 builds it."""
    from env.constants import ChoiceMethod
    from rl.MCTS import SingleThreadMCTS
    assert kind == "reference"
    return SingleThreadMCTS(mp_["num_output_class"], is_root=True, apply_dirichlet_noice=False, small_blind=params["small_blind"], model=model,
                            n_simulation=n_simulation, c_puct=params["mcts_c_puct"], tau=params["mcts_tau"], model_Q_epsilon=params["mcts_model_Q_epsilon"],
                            choice_method=ChoiceMethod(params["mcts_choice_method"]), log_to_file=False, pid=0)


def mcts_moves(env, search_factory, n_moves, seed):
    """Plays `n_moves` searched moves of self-play (new hands as needed).  Returns (simulations, seconds, moves)."""
    import numpy as np
    from env.constants import ChoiceMethod
    random.seed(seed)
    np.random.seed(seed)
    sims, moves, hand = 0, 0, 0
    t0 = time.perf_counter()
    obs = env.reset(hand, seed=seed)
    while moves < n_moves:
        search = search_factory()
        probs = search.simulate(obs, env)
        action_bin, action, _ = search.get_action(probs, env=env, choice_method=ChoiceMethod.PROBABILITY)
        obs, _, done, _ = env.step(action, action_bin)
        sims += search.n_simulation
        moves += 1
        if done:
            hand += 1
            obs = env.reset(hand, seed=seed + hand)
    return sims, time.perf_counter() - t0, moves
