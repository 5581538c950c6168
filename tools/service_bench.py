"""Throughput of the GPU equity service on the reference's task tuples and default plans
(simulate_processes replacement, env/workflow.py:102-227).  usage: python tools/service_bench.py [n_games]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import texasholdempocker_rl_b200 as phe
from texasholdempocker_rl_b200 import equity_plan

DEFAULT = {0: ([(1, 0, False), (1, 0, False), (1, 0, True), (4, 5, True)], 98000),
           1: ([(1, 0, True), (1, 0, True), (2, 45, True)], 97290),
           2: ([(1, 0, True), (1, 0, False), (1, 0, True)], 45540),
           3: ([(1, 0, False), (1, 0, True)], 990)}
n_games = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rng = np.random.default_rng(0)
D = phe.DECK
tasks = []
for g in range(n_games):           # a finished game emits 4 rounds x 2 players = 8 tasks (env/env.py:119-147)
    ids = rng.permutation(52)[:9].tolist()
    for p in range(2):
        hand = [D[i] for i in ids[2 * p: 2 * p + 2]]
        flop, turn, river = [D[i] for i in ids[4:7]], [D[ids[7]]], [D[ids[8]]]
        for rnd in range(4):
            tasks.append((g, f"p{p}", rnd, flop if rnd >= 1 else None, turn if rnd >= 2 else None, river if rnd >= 3 else None, hand))
torch.cuda.set_device(0)
phe.solve_tasks(tasks[:8], DEFAULT)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = phe.solve_tasks(tasks, DEFAULT)
dt = time.perf_counter() - t0
deals = sum(r[3] for r in res)
print(f"service: {len(tasks)} tasks ({n_games} games), {deals} deals in {dt*1e3:.1f} ms -> {len(tasks)/dt:.0f} tasks/s, {deals/dt:.3e} deals/s (host tuples in, tuples out)")
# kernel-only per round
for rnd in range(4):
    n_exist = {0: 2, 1: 5, 2: 6, 3: 7}[rnd]
    T = 2 * n_games
    exist = torch.from_numpy(np.argsort(rng.random((T, 52)), axis=1)[:, :n_exist].astype(np.uint8)).cuda()
    equity_plan(exist, DEFAULT[rnd][0])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        w = equity_plan(exist, DEFAULT[rnd][0])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"round {rnd}: {T} tasks x {DEFAULT[rnd][1]} deals: {ms:.3f} ms -> {T*DEFAULT[rnd][1]/(ms*1e-3):.3e} deals/s")
