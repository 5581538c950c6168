"""Fused enumeration + peer-memory all-reduce: correctness against the golden histogram and timing against the
kernel + NCCL all-reduce version.  Run under torchrun (or plain python for one GPU)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import texasholdempocker_rl_b200 as phe

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "hist7463.json")))
nvls = {None: None, '0': False, '1': True}[os.environ.get('PHE_FUSED_NVLS')]   # default: NVLS when the system offers a multicast mapping
fe = phe.FusedEnum7(nvls=nvls)  # fe.check() at the end raises if a peer missed an exchange
for it in range(5):
    h9, h = fe.run()
    torch.cuda.synchronize()
    ok = h9.tolist() == gold["hist9"] and h.tolist() == gold["hist7463"]
    assert ok, f"rank {rank} iteration {it}: fused result differs from the golden histogram"
# a sub-range too (odd size, exercises ranks with few tiles)
h9, h = fe.run(1000, 300000)
torch.cuda.synchronize()
from oracle import oracle as O
o9, oh = O.enum7(1000, 300000, nthreads=2)
assert (h.cpu().numpy() == oh).all() and (h9.cpu().numpy() == o9).all()


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    fe.run()
t_fused = timed(g.replay)
h9, h = fe.out[:9], fe.out[9:]
assert h9.tolist() == gold["hist9"] and h.tolist() == gold["hist7463"]
hist = torch.zeros(7472, dtype=torch.int64, device="cuda")


def nccl_step():
    hist.zero_()
    phe.sharded_enum7  # noqa: B018
    import ctypes as C
    phe._native.check(phe._native.lib().phe_enum7_part(C.c_void_p(hist.data_ptr()), C.c_void_p(hist.data_ptr() + 72), 0, 133784560, rank, world,
                                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    if world > 1:
        dist.all_reduce(hist)


nccl_step(); torch.cuda.synchronize()
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    nccl_step()
t_nccl = timed(g2.replay)
# a library built with -DPHE_ENUM_PROF leaves %globaltimer stamps of the last launch behind the status words
nbytes = int(phe._native.lib().phe_enum7_allreduce_workspace())
st = fe.ws[nbytes - 128 + 16: nbytes - 128 + 16 + 40].view(torch.int64).cpu().tolist()
if st[0]:
    names = ["enumerate+flush (CTA 0 start -> last CTA)", "push payload to peers", "signal + wait for peers", "sum slots"]
    print(f"rank {rank}: " + ", ".join(f"{n} {(st[i + 1] - st[i]) / 1e3:.1f} us" for i, n in enumerate(names)), flush=True)
if rank == 0:
    print(f"world {world} ({'NVLS multimem' if fe.nvls else 'peer stores'}): fused {t_fused:.1f} us/step ({133784560/t_fused*1e6:.3e} evals/s)   kernel+NCCL {t_nccl:.1f} us/step")
if world > 1:
    del g, g2
    torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
