"""Small single-GPU driver for ncu captures (one workload per invocation, few launches).
usage: python tools/prof_driver.py {enum|mc|eval7|plan} [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import texasholdempocker_rl_b200 as phe

what = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.cuda.set_device(0)
if what == "enum":
    for _ in range(reps):
        h9, h = phe.enumerate_c52_7(device_out=True)
    torch.cuda.synchronize()
    print(h9.tolist())
elif what == "mc":
    states = torch.from_numpy(phe.random_states(4096, 20261019)).cuda()
    for _ in range(reps):
        w = phe.equity_mc(states, 5, 100000, seed=1)
    torch.cuda.synchronize()
    print(int(w.sum()))
elif what == "eval7":
    g = torch.Generator().manual_seed(1)
    ids = torch.rand(1 << 18, 52, generator=g).argsort(1)[:, :7].numpy().astype(np.uint8)
    masks = torch.from_numpy(phe.ids_to_masks(ids)).cuda().repeat(128)
    for _ in range(reps):
        r = phe.eval7(masks)
    torch.cuda.synchronize()
    print(int(r.sum()))
elif what == "plan":
    rng = np.random.default_rng(5)
    exist = torch.from_numpy(np.argsort(rng.random((512, 52)), axis=1)[:, :6].astype(np.uint8)).cuda()
    plan = [(1, 0, True), (1, 0, False), (1, 0, True)]
    for _ in range(reps):
        w = phe.equity_plan(exist, plan)
    torch.cuda.synchronize()
    print(int(w.sum()))
