"""Kernel time of one interleaved part of the C(52,7) enumeration vs the number of parts (fixed-cost probe)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import texasholdempocker_rl_b200 as phe

L = phe._native.lib()
torch.cuda.set_device(0)
hist = torch.zeros(9 + 7463, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for nparts in (1, 2, 4, 8, 16, 64):
    ts = []
    for rep in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hist.zero_()
        a.record()
        phe._native.check(L.phe_enum7_part(C.c_void_p(hist.data_ptr()), C.c_void_p(hist.data_ptr() + 72), 0, 133784560, 0, nparts, C.c_void_p(st)))
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"nparts {nparts:3d}  part 0: median {ts[len(ts)//2]*1e3:8.1f} us  min {ts[0]*1e3:8.1f} us")
