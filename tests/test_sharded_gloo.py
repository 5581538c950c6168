"""N>1 host logic on CPU: two gloo ranks shard the sample / combination range, compute their shard with the
oracle standing in for the kernel (the product path has no CPU compute), and all-reduce the counts."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import texasholdempocker_rl_b200 as phe
    from oracle import oracle as O
    states = phe.random_states(6, 42)

    def local_mc(st, n_opp, rollouts, seed, sample_offset=0):
        return np.stack([O.equity_mc(int(st[s, 0]), int(st[s, 1]), n_opp, rollouts, seed, sample_offset=sample_offset, state_idx=s)
                         for s in range(len(st))])

    counts = phe.sharded_equity_mc(states, 5, 5001, seed=9, group=None, local_fn=local_mc)
    h9, h = phe.sharded_enum7(local_fn=lambda b, e: O.enum7(b, e, nthreads=1), comb_begin=1000, comb_end=2_001_001)
    if rank == 0:
        ret["counts"] = np.asarray(counts).copy()
        ret["h9"] = np.asarray(h9).copy()
        ret["h"] = np.asarray(h).copy()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_rank():
    from oracle import oracle as O
    import texasholdempocker_rl_b200 as phe
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    states = phe.random_states(6, 42)
    single = np.stack([O.equity_mc(int(states[s, 0]), int(states[s, 1]), 5, 5001, 9, state_idx=s) for s in range(6)])
    assert (ret["counts"] == single).all() and (ret["counts"].sum(1) == 5001).all()
    o9, oh = O.enum7(1000, 2_001_001, nthreads=2)
    assert (ret["h9"] == o9).all() and (ret["h"] == oh).all()
