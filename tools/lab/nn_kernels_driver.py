"""A few launches of the memory-bound glue kernels at bench size, for ncu (tools/lab)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import texasholdempocker_rl_b200 as phe
pv = phe.policy_value
torch.cuda.set_device(0)
net = pv.fill_parameters(pv.PolicyValueNet(**pv.DEFAULT_MODEL_PARAMS), 1).to("cuda", torch.bfloat16).eval()
obs = torch.from_numpy(pv.random_observations(1024, seed=3)).cuda()
with torch.no_grad():
    for _ in range(2):
        x = net.embed(obs)                                                     # k_embed_obs: [1024, 98, 640]
        y = pv._add_layernorm(x, x, net.transform_encoder.layers[0].norm1)    # k_add_layernorm: [100352, 640]
enc = phe.ObsEncoder()
b = enc.empty_batch(1 << 16)
rng = np.random.default_rng(0)
b["values"][:] = rng.integers(1, 4000, b["values"].shape) * 25
b["bins"][:] = rng.integers(0, 400, b["bins"].shape) * 25
d = {k: torch.from_numpy(v).cuda() for k, v in b.items()}
for _ in range(2):
    o = enc.encode(d)                                                          # k_encode_obs: 65536 observations
torch.cuda.synchronize()
print(float(y.float().abs().mean()), int(o.sum()))
