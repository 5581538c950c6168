import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
import texasholdempocker_rl_b200 as phe
pv = phe.policy_value
for (B, H, D) in ((1024, 10, 64), (4096, 6, 16)):
    qkv = torch.randn(B, 98, 3, H, D, device="cuda", dtype=torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    def t(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20 * 1e3
    mine = t(lambda: pv._attention(q, k, v))
    ref = t(lambda: F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)))
    byts = B * H * 98 * D * 2 * 4
    print(f"B={B} H={H} D={D}: phe_nn_attention {mine:.1f} us ({byts / mine / 1e3:.0f} GB/s algorithmic)   torch sdpa {ref:.1f} us")
