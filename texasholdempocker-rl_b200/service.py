"""GPU equity service speaking the reference's task / result tuples.

Replaces the pool of ``simulate_processes`` workers (env/workflow.py:102-227): same queue protocol, same
``simulation_recurrent_param_dict`` plans, but tasks are drained in batches and each (round, batch) is one
``phe_equity_plan`` launch.

  task   (game_id, player_name, current_round, flop_cards, turn_cards, river_cards, player_hand_card)
         (env/env.py:240; board entries are None for streets not reached)
  result (game_id, player_name, current_round, num_estimation, game_result_sum)   (env/workflow.py:225)
"""
import logging
import queue as _queue

import numpy as np

from .cardset import card_id
from .equity import equity_plan, plan_num_deals

_N_EXIST = {0: 2, 1: 5, 2: 6, 3: 7}


def task_exist_ids(task):
    """Known cards of a task in the reference's order (env/workflow.py:212-221)."""
    _, _, rnd, flop, turn, river, hand = task
    ids = [card_id(c) for c in hand]
    if rnd > 0:
        ids += [card_id(c) for c in flop]
    if rnd > 1:
        ids += [card_id(c) for c in turn]
    if rnd > 2:
        ids += [card_id(c) for c in river]
    if len(ids) != _N_EXIST[rnd] or len(set(ids)) != len(ids):
        raise ValueError(f"task for round {rnd} has {len(ids)} known cards / duplicates")
    return ids


def solve_tasks(tasks, simulation_recurrent_param_dict, seed=0, stream_base=0):
    """Evaluate a list of task tuples; returns result tuples in task order.  Tasks are grouped by round (one launch per
    round); sampled plan steps draw from Philox streams stream_base .. stream_base + len(tasks) - 1 of key `seed`, numbered
    along the grouped order, so no two tasks of a call share a stream whichever round they belong to."""
    results = [None] * len(tasks)
    by_round = {}
    for i, t in enumerate(tasks):
        by_round.setdefault(t[2], []).append(i)
    first_stream = stream_base
    for rnd, idxs in by_round.items():
        plan, num_estimation = simulation_recurrent_param_dict[rnd]
        exist = np.array([task_exist_ids(tasks[i]) for i in idxs], dtype=np.uint8)
        wtl = equity_plan(exist, plan, seed=seed, stream_base=first_stream & 0xFFFFFFFF)
        first_stream += len(idxs)
        # the reference asserts this too (env/workflow.py:223)
        assert (wtl.sum(1) == num_estimation).all(), f"plan for round {rnd} does not generate {num_estimation} deals"
        sums = (wtl[:, 0] - wtl[:, 2]).astype(np.float64).tolist()
        for row, i in enumerate(idxs):
            t = tasks[i]
            results[i] = (t[0], t[1], rnd, num_estimation, sums[row])
    return results


def winning_probability(game_result_sum, num_estimation):
    """env/workflow.py:247-251"""
    return (game_result_sum / num_estimation) * 0.5 + 0.5


def check_plans(simulation_recurrent_param_dict):
    for rnd, (plan, num_estimation) in simulation_recurrent_param_dict.items():
        n = plan_num_deals(_N_EXIST[rnd], plan)
        if n != num_estimation:
            raise ValueError(f"round {rnd}: plan yields {n} deals but num_estimation is {num_estimation}")


def worker_seed(pid):
    """Philox key of one equity worker: the reference's workers draw from the unseeded global `random` (env/workflow.py:190),
    so no two of them repeat each other; here the key mixes the worker id, the OS process id and the start time."""
    import os
    import time
    x = (int(pid) * 0x9E3779B97F4A7C15 + os.getpid() * 0xC2B2AE3D27D4EB4F + time.time_ns()) & (2**64 - 1)
    x ^= x >> 29
    return (x * 0xBF58476D1CE4E5B9) & (2**64 - 1)


def simulate_processes(in_queue, out_queue, simulation_recurrent_param_dict, pid, log_level,
                       interrupt_callback=None, max_batch=4096, seed=None):
    """Drop-in body for the reference's equity worker (same first five parameters, env/workflow.py:102).

    Blocks on in_queue with a 0.1 s timeout like the reference, drains up to max_batch tasks, solves them on
    the GPU and puts the result tuples on out_queue.  Returns when interrupt_callback() is true or a ``None``
    task arrives.  ``seed=None`` (default) gives every worker process its own Philox key (``worker_seed``) and every
    task it serves its own stream, so sampled plans are uncorrelated across tasks, workers and runs -- like the
    reference's unseeded ``random``; pass an integer for reproducible results.
    """
    if seed is None:
        seed = worker_seed(pid)
    logging.getLogger().setLevel(log_level)
    logging.info(f"simulate_processes (B200) {pid} started")
    check_plans(simulation_recurrent_param_dict)
    served = 0
    while True:
        if interrupt_callback is not None and interrupt_callback():
            return served
        try:
            first = in_queue.get(block=True, timeout=0.1)
        except _queue.Empty:
            continue
        if first is None:
            return served
        batch = [first]
        stop = False
        while len(batch) < max_batch:
            try:
                t = in_queue.get(block=False)
            except _queue.Empty:
                break
            if t is None:
                stop = True
                break
            batch.append(t)
        for r in solve_tasks(batch, simulation_recurrent_param_dict, seed=seed, stream_base=served):
            out_queue.put(r)
        served += len(batch)
        if stop:
            return served
