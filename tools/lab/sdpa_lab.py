"""Which attention kernel for [B*H, S=98, hd=64] bf16?  (lab; B=1024, H=10)"""
import torch, time
import torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend
B, H, S, D = 1024, 10, 98, 64
qkv = torch.randn(B, S, 3, H, D, device="cuda", dtype=torch.bfloat16)
q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
def timeit(fn, name):
    try:
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): o = fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:28s} {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us")
        return o
    except Exception as e:
        print(f"{name:28s} failed: {str(e)[:100]}")
ref = timeit(lambda: F.scaled_dot_product_attention(q, k, v), "sdpa default")
for be in (SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.CUDNN_ATTENTION, SDPBackend.MATH):
    def f(be=be):
        with sdpa_kernel(be):
            return F.scaled_dot_product_attention(q, k, v)
    o = timeit(f, str(be))
    if o is not None: print("   max diff vs default", float((o.float() - ref.float()).abs().max()))
try:
    from flash_attn import flash_attn_func, flash_attn_qkvpacked_func
    o = timeit(lambda: flash_attn_qkvpacked_func(qkv), "flash_attn qkvpacked")
    print("   max diff", float((o.transpose(1, 2).float() - ref.float()).abs().max()))
except Exception as e:
    print("flash_attn import failed", e)
