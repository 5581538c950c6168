"""Monte-Carlo throughput per street (0 / 3 / 4 / 5 board cards) and opponent count, for kernel A/B work.
usage: python tools/lab/mc_streets.py [rollouts]   (PHE_B200_LIB selects a variant build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import texasholdempocker_rl_b200 as phe
R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
torch.cuda.set_device(0)
out = []
for n_opp in (5, 1):
    for street in (0, 3, 4, 5):
        states = torch.from_numpy(phe.random_states(2048, 5, streets=(street,))).cuda()
        w = phe.equity_mc(states, n_opp, R, seed=1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            w = phe.equity_mc(states, n_opp, R, seed=1)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        out.append(f"opp={n_opp} board={street}: {2048 * R / (ms * 1e-3):.3e}/s chk {int(w[:, 0].sum())}")
print(" | ".join(out))
