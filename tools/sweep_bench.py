"""BASELINE config 5: sharded Monte-Carlo equity sweep -- S = 64 six-max states, R rollouts per state (default 1e9),
the global sample range split over the ranks, one all_reduce(SUM) of int64[64, 3] at the end.

  python tools/sweep_bench.py [R]                                                             (one GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/sweep_bench.py [R]

Prints one JSON line on rank 0: rollouts/s (whole job, device time = max over ranks), the all-reduce time, and a
checksum of the counts -- the counts are integer sums over a counter-based stream, so the checksum is identical for
every G."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import texasholdempocker_rl_b200 as phe

R = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
S, SEED = 64, 0x5EED20261019
states = torch.from_numpy(phe.random_states(S, 20261019 + 1)).to(dev)
b, e = phe.shard_range(R, rank, world)
wtl = torch.zeros(S, 3, dtype=torch.int64, device=dev)


def sweep():
    wtl.zero_()
    phe.equity_mc(states, 5, e - b, seed=SEED, sample_offset=b, out=wtl)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


small = torch.zeros(S, 3, dtype=torch.int64, device=dev)
phe.equity_mc(states, 5, 100000, seed=SEED, out=small)              # warm-up: tables, kernels
if world > 1:
    dist.all_reduce(small)
barrier()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
sweep()
e1.record()
if world > 1:
    dist.all_reduce(wtl)
e2.record()
barrier()
t = torch.tensor([e0.elapsed_time(e2), e1.elapsed_time(e2)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
total_ms, ar_ms = (float(x) for x in t.tolist())
counts = wtl.cpu()
assert bool((counts.sum(1) == R).all()), "win + tie + loss must equal the rollouts of every state"
if rank == 0:
    print(json.dumps({"workload": "sharded_mc_equity_sweep", "states": S, "players": 6, "rollouts_per_state": R, "n_gpus": world,
                      "ms": total_ms, "rollouts_per_s": S * R / (total_ms * 1e-3), "all_reduce_ms": ar_ms if world > 1 else 0.0,
                      "checksum_wins": int(counts[:, 0].sum()), "checksum_ties": int(counts[:, 1].sum()),
                      "equity_state0": float(phe.win_probability(counts[0]))}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
