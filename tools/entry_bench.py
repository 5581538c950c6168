"""Device-resident throughput of every batched entry point (CUDA events, 1 GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import texasholdempocker_rl_b200 as phe


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


torch.cuda.set_device(0)
g = torch.Generator().manual_seed(1)
N = 1 << 24
ids = torch.rand(1 << 18, 52, generator=g).argsort(1)[:, :7].to(torch.uint8)
ids_d = ids.cuda().repeat(N >> 18, 1)
masks_d = torch.from_numpy(phe.ids_to_masks(ids.numpy())).cuda().repeat(N >> 18)
out = torch.empty(N, dtype=torch.int16, device="cuda")
ms = timeit(lambda: phe.eval7(masks_d, out=out)); print(f"eval7 masks   N=2^24: {ms:.3f} ms  {N/ms/1e6:.1f} G evals/s  ({N*10/ms/1e6:.0f} GB/s)")
ms = timeit(lambda: phe.eval7(ids_d, out=out)); print(f"eval7 ids     N=2^24: {ms:.3f} ms  {N/ms/1e6:.1f} G evals/s  ({N*9/ms/1e6:.0f} GB/s)")
for P in (2, 6):
    G = 1 << 22
    deals = torch.rand(1 << 16, 52, generator=g).argsort(1)[:, :5 + 2 * P].numpy()
    boards = torch.from_numpy(phe.ids_to_masks(deals[:, :5])).cuda().repeat(G >> 16)
    holes = torch.from_numpy(phe.ids_to_masks(deals[:, 5:].reshape(-1, P, 2))).cuda().repeat(G >> 16, 1)
    ms = timeit(lambda: phe.showdown(boards, holes)); print(f"showdown P={P} G=2^22: {ms:.3f} ms  {G/ms/1e6:.2f} G games/s  {G*P/ms/1e6:.1f} G evals/s")
known = masks_d[: 1 << 22]
ms = timeit(lambda: phe.deal(known, 9, seed=3)); print(f"deal need=9  B=2^22: {ms:.3f} ms  {(1<<22)/ms/1e6:.2f} G deals/s")
# settlement: 2^20 heads-up river showdowns with one all-in layer
G = 1 << 20
deals = torch.rand(1 << 16, 52, generator=g).argsort(1)[:, :9].numpy()
boards = torch.from_numpy(phe.ids_to_masks(deals[:, :5])).cuda().repeat(G >> 16)
holes = torch.from_numpy(phe.ids_to_masks(deals[:, 5:].reshape(-1, 2, 2))).cuda().repeat(G >> 16, 1)
act = torch.full((G, 4, 2), 2, dtype=torch.int8, device="cuda")
val = torch.tensor([[50, 50], [100, 100], [200, 150], [0, 0]], dtype=torch.int32, device="cuda").repeat(G, 1, 1)
nr = torch.full((G,), 4, dtype=torch.int32, device="cuda"); bt = torch.zeros(G, dtype=torch.int32, device="cuda")
ob = torch.full((G,), 1, dtype=torch.int32, device="cuda")
ms = timeit(lambda: phe.settle(boards, holes, act, val, nr, bt, ob, 25)); print(f"settle P=2   G=2^20: {ms:.3f} ms  {G/ms/1e6:.3f} G games/s")
for rnd, T in ((3, 65536), (2, 2048), (1, 1024), (0, 1024)):
    n_exist = {0: 2, 1: 5, 2: 6, 3: 7}[rnd]
    plan = {0: [(1, 0, False), (1, 0, False), (1, 0, True), (4, 5, True)], 1: [(1, 0, True), (1, 0, True), (2, 45, True)],
            2: [(1, 0, True), (1, 0, False), (1, 0, True)], 3: [(1, 0, False), (1, 0, True)]}[rnd]
    exist = torch.rand(T, 52, generator=g).argsort(1)[:, :n_exist].to(torch.uint8).cuda()
    n = phe.plan_num_deals(n_exist, plan)
    ms = timeit(lambda: phe.equity_plan(exist, plan), reps=3); print(f"plan round {rnd}  T={T}: {ms:.3f} ms  {T*n/ms/1e6:.2f} G deals/s  {T/ms*1e3:.0f} tasks/s")
