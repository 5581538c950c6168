"""Where the fixed cost of one k_enum7 launch goes: %globaltimer stamps of a library built with -DPHE_ENUM_PROF
(PHE_B200_LIB=.../libphe_b200_enumprof.so).  Per-CTA stamps (start, tables staged, loop done, flush done) come from
phe_debug_enum_prof; the last CTA's epilogue stamps through the one-rank fused entry point.  Lab tool, one GPU."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import texasholdempocker_rl_b200 as phe

torch.cuda.set_device(0)
L = phe._native.lib()
fe = phe.FusedEnum7()
nbytes = int(L.phe_enum7_allreduce_workspace())
T = 133_784_560


def show(tag, run):
    for _ in range(6):
        run()
        torch.cuda.synchronize()
    buf = np.zeros(1024, dtype=np.uint64)
    assert L.phe_debug_enum_prof(C.c_void_p(buf.ctypes.data)) == 0
    p = buf.reshape(256, 4)[:148].astype(np.int64)
    t0 = p[:, 0].min()
    q = lambda col, f: (f(p[:, col]) - t0) / 1e3
    print(f"{tag}: start {q(0, np.min):.1f}..{q(0, np.max):.1f}  staged {q(1, np.min):.1f}..{q(1, np.max):.1f}  loop done min {q(2, np.min):.1f} "
          f"p25 {(np.percentile(p[:, 2], 25) - t0) / 1e3:.1f} median {(np.median(p[:, 2]) - t0) / 1e3:.1f} p75 {(np.percentile(p[:, 2], 75) - t0) / 1e3:.1f} "
          f"max {q(2, np.max):.1f}  flush done max {q(3, np.max):.1f} us", flush=True)
    return t0


for n in (10_000, 1_000_000, T // 8, T // 2, T):
    t0 = show(f"fused world=1 [0,{n})", lambda: fe.run(0, n))
    st = fe.ws[nbytes - 128 + 16: nbytes - 128 + 16 + 64].view(torch.int64).cpu().tolist()
    print(f"      last CTA: epilogue +{(st[1]-t0)/1e3:.1f}  pushed +{(st[2]-t0)/1e3:.1f}  waited +{(st[3]-t0)/1e3:.1f}  thread 0 summed +{(st[5]-t0)/1e3:.1f}  "
          f"all summed +{(st[6]-t0)/1e3:.1f}  end +{(st[4]-t0)/1e3:.1f} us")
hist = torch.zeros(7472, dtype=torch.int64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for nparts in (1, 2, 4, 8):
    show(f"plain part 0 of {nparts}", lambda: L.phe_enum7_part(C.c_void_p(hist.data_ptr()), C.c_void_p(hist.data_ptr() + 72), 0, T, 0, nparts, st))
