import sys, os
sys.path.insert(0, "/root/repo")
import torch
import texasholdempocker_rl_b200 as phe
pv = phe.policy_value
from torch.profiler import profile, ProfilerActivity
m = pv.fill_parameters(pv.PolicyValueNet(**pv.DEFAULT_MODEL_PARAMS), 1).cuda().to(torch.bfloat16).eval()
import sys
BATCH = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
obs = torch.from_numpy(pv.random_observations(BATCH, seed=3)).cuda()
with torch.no_grad():
    for _ in range(3): m.policy_value(obs)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3): m.policy_value(obs)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
