import sys; sys.path.insert(0,'/root/repo')
import torch, texasholdempocker_rl_b200 as phe
torch.cuda.set_device(0)
for n in (10_000, 300_000, 1_000_000, 5_000_000, 20_000_000, 66_892_280, 133_784_560):
    for _ in range(3): phe.enumerate_c52_7(0, n, device_out=True)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): phe.enumerate_c52_7(0, n, device_out=True)
    e1.record(); torch.cuda.synchronize()
    print(n, "%.1f us per call (incl. output alloc+zero)" % (e0.elapsed_time(e1)/20*1e3))
