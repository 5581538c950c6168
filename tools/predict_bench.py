"""Throughput of the batched policy / value predictor on one GPU.

  python tools/predict_bench.py [default|debug] [batch ...]

Per batch size: device-resident replay (CUDA events around K graph replays) and end to end from host numpy
observations (pinned H2D + replay + D2H, wall clock), plus the plain eager fp32 module for comparison."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import texasholdempocker_rl_b200 as phe

pv = phe.policy_value


def flops_per_obs(p, n_out=2):
    d, L, S, ff = p["embedding_dim"] + p["positional_embedding_dim"], p["num_layers"], 98, 2048
    full = 2 * S * (4 * d * d + 2 * d * ff) + 4 * S * S * d
    last = 2 * (S * 2 * d * d + n_out * (2 * d * d + 2 * d * ff)) + 4 * n_out * S * d
    return (L - 1) * full + last, L * full


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "default"
    params = pv.DEFAULT_MODEL_PARAMS if which == "default" else pv.DEBUG_MODEL_PARAMS
    batches = [int(a) for a in sys.argv[2:]] or [1, 16, 256, 1024, 4096]
    torch.cuda.set_device(0)
    model = pv.fill_parameters(pv.PolicyValueNet(**params), 1)
    pruned, full = flops_per_obs(params)
    print(f"{which}: d={model.d_model} layers={params['num_layers']}  {pruned / 1e9:.3f} GFLOP/obs (unpruned {full / 1e9:.3f})")
    for dtype in (torch.bfloat16, torch.float32):
        pred = pv.BatchedPredictor(model, dtype=dtype, max_batch=max(batches), buckets=batches)
        pred.warmup()
        for B in batches:
            obs = pv.random_observations(B, seed=B)
            dobs = torch.from_numpy(obs).cuda()
            pred.predict_batch(dobs)
            torch.cuda.synchronize()
            K = max(3, min(50, 20000 // B))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                pred.predict_batch(dobs)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            t0 = time.perf_counter()
            for _ in range(K):
                pred.predict_batch(obs)
            hs = (time.perf_counter() - t0) / K
            print(f"  {str(dtype)[6:]:9s} B={B:5d}: device {ms * 1e3:9.1f} us  {B / ms * 1e3:12.0f} obs/s  {B * pruned / ms / 1e9:7.1f} TFLOP/s | "
                  f"host e2e {hs * 1e6:9.1f} us  {B / hs:12.0f} obs/s")
        del pred
    # eager fp32 reference-style module call (what SingleThreadMCTS.predict does per node)
    m = pv.fill_parameters(pv.PolicyValueNet(**params), 1).cuda().eval()
    obs1 = torch.from_numpy(pv.random_observations(1, seed=1)).cuda()
    with torch.no_grad():
        for _ in range(5):
            m.policy_value(obs1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            p, v = m.policy_value(obs1)
            p.cpu(); v.cpu()
        print(f"  eager fp32 B=1 (no graph): {(time.perf_counter() - t0) / 50 * 1e6:.0f} us per call")


if __name__ == "__main__":
    main()
