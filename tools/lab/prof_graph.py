"""Kernel breakdown of one CUDA-graph replay of the policy / value predictor (default.yml dims).  usage: python tools/lab/prof_graph.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
import texasholdempocker_rl_b200 as phe
pv = phe.policy_value
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pred = pv.BatchedPredictor(pv.fill_parameters(pv.PolicyValueNet(**pv.DEFAULT_MODEL_PARAMS), 1), dtype=torch.bfloat16, max_batch=B, buckets=[B])
pred.warmup()
obs = torch.from_numpy(pv.random_observations(B, seed=3)).cuda()
for _ in range(3): pred.predict_batch(obs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): pred.predict_batch(obs)
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per batch of {B}")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): pred.predict_batch(obs)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
