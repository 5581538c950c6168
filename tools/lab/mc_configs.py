"""Monte-Carlo throughput on the BASELINE shapes: config 3 (4096 six-max states), config 5 shape (64 six-max states, many
chunks per state), a heads-up batch.  usage: python tools/lab/mc_configs.py   (PHE_B200_LIB selects a variant build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import texasholdempocker_rl_b200 as phe
torch.cuda.set_device(0)
out = []
for name, S, seed, n_opp, R in (("config3", 4096, 20261019, 5, 1_000_000), ("config5_shape", 64, 20261020, 5, 60_000_000), ("headsup", 4096, 20261019, 1, 2_000_000)):
    states = torch.from_numpy(phe.random_states(S, seed)).cuda()
    w = phe.equity_mc(states, n_opp, R // 10, seed=0x5EED20261019)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
        w = phe.equity_mc(states, n_opp, R, seed=0x5EED20261019)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 2
    out.append(f"{name}: {S * R / (ms * 1e-3):.4e}/s chk {int(w[:, 0].sum())}")
print(" | ".join(out))
