"""Single-call latency of the drop-in entry points (host buffers in, host result out)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import texasholdempocker_rl_b200 as phe

torch.cuda.set_device(0)
D = phe.DECK
ids = [51, 50, 0, 21, 32, 10, 9]
phe.evaluate_cards(*ids)


def t(fn, n=2000):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


print(f"evaluate_cards(7 ids)            {t(lambda: phe.evaluate_cards(*ids)):8.1f} us/call")
flop, turn, river = [D[0], D[21], D[32]], [D[10]], [D[9]]
players = {0: [D[51], D[50]], 1: [D[47], D[46]]}
print(f"get_winner_set_v2 (2 players)    {t(lambda: phe.GameEnv.get_winner_set_v2(flop, turn, river, players)):8.1f} us/call")
m = np.array([phe.cards_to_mask(ids)], dtype=np.int64)
print(f"eval7(numpy[1])                  {t(lambda: phe.eval7(m)):8.1f} us/call")
big = np.repeat(m, 4096)
print(f"eval7(numpy[4096])               {t(lambda: phe.eval7(big), 500):8.1f} us/call")
st = phe.pack_states([["As", "Ah"]], [["2c", "7d", "Tc"]])
print(f"equity_mc(1 state, 10k rollouts) {t(lambda: phe.equity_mc(st, 1, 10000, seed=1), 300):8.1f} us/call")
ex = np.array([ids], dtype=np.uint8)
print(f"equity_plan(river task)          {t(lambda: phe.equity_plan(ex, [(1, 0, False), (1, 0, True)]), 500):8.1f} us/call")
