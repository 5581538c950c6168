"""Times the config-3 Monte-Carlo batch (4096 six-max states) for kernel A/B work.  usage: python tools/mc_time.py [rollouts]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import texasholdempocker_rl_b200 as phe

R = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
torch.cuda.set_device(0)
states = torch.from_numpy(phe.random_states(4096, 20261019)).cuda()
w = phe.equity_mc(states, 5, R, seed=1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    w = phe.equity_mc(states, 5, R, seed=1)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(f"{4096 * R / (ms * 1e-3):.4e} rollouts/s  {ms:.2f} ms  checksum {int(w[:, 0].sum())}")
