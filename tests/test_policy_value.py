"""Policy / value inference against outputs of the reference's own TransformerAlphaGoZeroModel
(tests/golden/model_run.json, made by tests/golden/make_model_golden.py from /root/reference/rl/models.py).

Tolerances (floating point): fp32 module vs the reference's fp32 outputs 2e-4 absolute on logits / values (different
but equivalent operation order: folded constants, fused QKV, pruned last layer); bf16 predictor 3e-2 on action
probabilities and 6e-2 on the value (bf16 has 8 mantissa bits; logits here are O(1))."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

import texasholdempocker_rl_b200 as phe

pv = phe.policy_value


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(GOLDEN, "model_run.json")))["cases"]


def _model(case):
    return pv.fill_parameters(pv.PolicyValueNet(**case["params"]), case["seed"]).eval()


def test_field_layout_matches_reference_constants():
    starts, sizes, rows = pv.field_layout(num_output_class=12, num_bins=10, historical_action_per_round=3)
    # field_start_idx_array printed from the reference model built with config/default.yml's parameters
    assert rows == 220 and starts.shape == (93,)
    assert starts[:9].tolist() == [5, 9, 12, 12, 12, 12, 12, 12, 12]
    assert starts[16:26].tolist() == [33, 42, 51, 60, 69, 77, 85, 102, 110, 118]
    assert starts[26] == 135 and starts[36] == 155 and starts[46] == 164 and starts[56] == 173
    assert starts[66:69].tolist() == [182, 186, 196] and set(starts[69:].tolist()) == {207}
    assert int(sizes[-1]) == 13 and int(starts[-1] + sizes[-1]) == rows


@pytest.mark.parametrize("idx", [0, 1])
def test_module_matches_reference_outputs_fp32(golden, idx):
    case = golden[idx]
    m = _model(case)
    obs = torch.tensor(case["obs"], dtype=torch.int32)
    with torch.no_grad():
        out = m(obs)
        probs, value = m.policy_value(obs)
    for got, want in zip(out, case["outputs"]):
        np.testing.assert_allclose(got.double().numpy(), np.asarray(want), atol=2e-4, rtol=0)
    np.testing.assert_allclose(probs.double().numpy(), np.asarray(case["action_probs"]), atol=1e-4, rtol=0)
    np.testing.assert_allclose(value.double().numpy(), np.asarray(case["outputs"][1]), atol=2e-4, rtol=0)


def test_state_dict_round_trip_and_wrapper_prefix(golden):
    case = golden[0]
    m = _model(case)
    sd = {"model." + k: v.clone() for k, v in m.state_dict().items()}        # keys as saved from the AlphaGoZero wrapper
    m2 = pv.PolicyValueNet(**case["params"]).eval()
    m2.load_reference_state_dict(sd)
    obs = torch.tensor(case["obs"][:4], dtype=torch.int32)
    with torch.no_grad():
        a, b = m(obs), m2(obs)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_last_layer_pruning_is_exact_per_row(golden):
    """Rows of the pruned last layer equal the same rows of the full layer (attention rows are independent)."""
    case = golden[0]
    m = _model(case)
    obs = torch.tensor(case["obs"][:8], dtype=torch.int32)
    with torch.no_grad():
        h = m.embed(obs)
        for layer in m.transform_encoder.layers[:-1]:
            h = layer(h)
        full = m.transform_encoder.layers[-1](h)
        part = m.transform_encoder.layers[-1](h, n_query=5)
    np.testing.assert_allclose(part.numpy(), full[:, :5].numpy(), atol=1e-5, rtol=0)


def test_bad_observations_are_rejected(golden):
    case = golden[0]
    m = _model(case)
    obs = np.asarray(case["obs"][:2], dtype=np.int32)
    m.check_observations(obs)
    bad = obs.copy()
    bad[1, -1] = 13                                 # last field: 13 rows, index 13 leaves the table
    with pytest.raises(ValueError):
        m.check_observations(bad)
    bad = obs.copy()
    bad[0, 3] = -1
    with pytest.raises(ValueError):
        m.check_observations(bad)
    with pytest.raises(ValueError):
        m.check_observations(obs[:, :-1])
    with pytest.raises(ValueError):
        pv.PolicyValueNet(**dict(case["params"], num_acting_player_fields=9))      # field count mismatch


def test_predictor_needs_cuda(golden):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(phe._native.PheError):
        pv.BatchedPredictor(_model(golden[0]))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,ap,av", [(torch.float32, 1e-3, 2e-3), (torch.bfloat16, 3e-2, 6e-2)])
def test_predictor_matches_reference_gpu(golden, dtype, ap, av):
    for case in golden:
        pred = pv.BatchedPredictor(_model(case), dtype=dtype, max_batch=32)
        obs = np.asarray(case["obs"], dtype=np.int32)
        probs, value = pred.predict_batch(obs)                               # 48 obs -> buckets 32 + 16, 6 obs -> bucket 16
        assert probs.dtype == np.float32 and probs.shape == (obs.shape[0], 12) and value.shape == (obs.shape[0],)
        np.testing.assert_allclose(probs, np.asarray(case["action_probs"]), atol=ap, rtol=0)
        np.testing.assert_allclose(value, np.asarray(case["outputs"][1]), atol=av, rtol=0)
        np.testing.assert_allclose(probs.sum(1), 1.0, atol=1e-5)
        p1, v1 = pred.predict(obs[3])                                         # MCTS.predict drop-in: one observation
        assert p1.shape == (12,) and np.ndim(v1) == 0
        np.testing.assert_allclose(p1, probs[3], atol=ap, rtol=0)
        dp, dv = pred.predict_batch(torch.from_numpy(obs).cuda())            # device-resident path
        torch.cuda.synchronize()
        np.testing.assert_allclose(dp.cpu().numpy(), probs, atol=ap, rtol=0)
        np.testing.assert_allclose(dv.cpu().numpy(), value, atol=av, rtol=0)
        assert pred.replays >= 3


@pytest.mark.gpu
def test_predictor_threads_gpu(golden):
    """Many Python threads call predict() concurrently (env/workflow.py:848-878): results equal the batched ones."""
    import threading
    case = golden[0]
    pred = pv.BatchedPredictor(_model(case), dtype=torch.float32, max_batch=16)
    obs = np.asarray(case["obs"], dtype=np.int32)
    want, _ = pred.predict_batch(obs)
    got = [None] * obs.shape[0]

    def work(lo, hi):
        for i in range(lo, hi):
            got[i] = pred.predict(obs[i])[0]

    ts = [threading.Thread(target=work, args=(k * 12, (k + 1) * 12)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    np.testing.assert_allclose(np.stack(got), want, atol=1e-4, rtol=0)


@pytest.mark.gpu
def test_predictor_follows_weight_updates_gpu(golden):
    """The reference's predict_batch_process reloads weights while it serves (update_model_param_queue, env/workflow.py:274-593).
    Captured graphs hold the folded embedding tables by address: an update -- through the predictor or in place behind its
    back -- must reach the replays (tables refolded into the same storage), checked against a fresh fp32 module."""
    obs = pv.random_observations(48, seed=3)
    pred = pv.BatchedPredictor(pv.fill_parameters(pv.PolicyValueNet(**pv.DEBUG_MODEL_PARAMS), 1), dtype=torch.float32, max_batch=64)
    pred.warmup()
    graphs = dict(pred._graphs)

    def want(seed):
        with torch.no_grad():
            p, v = pv.fill_parameters(pv.PolicyValueNet(**pv.DEBUG_MODEL_PARAMS), seed).eval().policy_value(torch.from_numpy(obs))
        return p.numpy(), v.numpy()
    for seed, how in ((1, None), (2, "api"), (3, "in_place"), (4, "api")):
        if how == "api":
            pred.load_reference_state_dict(pv.fill_parameters(pv.PolicyValueNet(**pv.DEBUG_MODEL_PARAMS), seed).state_dict())
        elif how == "in_place":
            pv.fill_parameters(pred.model, seed)
        p, v = pred.predict_batch(obs)
        wp, wv = want(seed)
        assert np.abs(p - wp).max() < 1e-3 and np.abs(v - wv).max() < 2e-3, f"stale weights after update {seed} ({how})"
    assert all(pred._graphs[b][0] is g[0] for b, g in graphs.items()), "graphs were re-captured although the tables kept their storage"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,d", [(1, 96), (37, 96), (1000, 640), (333, 2048 if True else 0), (5, 8), (64, 1024)])
def test_add_layernorm_kernel_gpu(dtype, rows, d):
    """phe_nn_add_layernorm against a plain fp32 PyTorch LayerNorm(x + r) of the same (rounded) inputs.
    Tolerance: fp32 1e-5; bf16 output rounding 2^-8 relative on O(1..4) values -> 2e-2 absolute."""
    if dtype == torch.float32 and d > 1024:
        pytest.skip("f32 rows are limited to 1024 elements")
    g = torch.Generator(device="cpu").manual_seed(rows * 131 + d)
    x = (torch.randn(rows, d, generator=g) * 1.5).to(dtype).cuda()
    r = (torch.randn(rows, d, generator=g) + 0.3).to(dtype).cuda()
    norm = torch.nn.LayerNorm(d, eps=1e-5)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(d, generator=g))
        norm.bias.copy_(0.1 * torch.randn(d, generator=g))
    norm = norm.to("cuda", dtype)
    before = phe._native.launch_count()
    with torch.no_grad():
        got = pv._add_layernorm(x, r, norm)
        want = torch.nn.functional.layer_norm(x.float() + r.float(), (d,), norm.weight.float(), norm.bias.float(), 1e-5)
    assert phe._native.launch_count() == before + 1            # the native kernel ran, not the fallback expression
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert got.dtype == dtype
    np.testing.assert_allclose(got.float().cpu().numpy(), want.cpu().numpy(), atol=tol, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_embed_obs_kernel_gpu(golden, dtype):
    """phe_nn_embed_obs is a gather-copy of folded tables: bit-identical to the PyTorch expression on the same tables."""
    for case in golden:
        m = _model(case).to("cuda", dtype)
        obs = torch.tensor(case["obs"], dtype=torch.int32)
        with torch.no_grad():
            before = phe._native.launch_count()
            got = m.embed(obs.cuda())
            assert phe._native.launch_count() == before + 1
            pos, starters, field = m._fold()
            idx = obs.cuda().long() + m.env_embedding.field_start_idx_array.long()
            tok = torch.cat([field[idx], pos.unsqueeze(0).expand(obs.shape[0], -1, -1)], dim=2)
            want = torch.cat([starters.unsqueeze(0).expand(obs.shape[0], -1, -1), tok], dim=1)
        assert got.shape == (obs.shape[0], 98, m.d_model)
        assert torch.equal(got, want)
        # and against the reference formulation (unfolded tables, scale applied last) within rounding
        cpu = _model(case)
        with torch.no_grad():
            ref = cpu.embed(obs)
        np.testing.assert_allclose(got.float().cpu().numpy(), ref.numpy(), atol=1e-6 if dtype == torch.float32 else 3e-2, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("H,D", [(10, 64), (6, 16)])
@pytest.mark.parametrize("Sq,S", [(98, 98), (2, 98), (5, 98), (1, 1), (17, 33), (112, 112), (16, 97)])
def test_attention_kernel_gpu(H, D, Sq, S):
    """phe_nn_attention against a plain fp32 PyTorch softmax(q k^T / sqrt(D)) v of the same bf16 inputs, on views into
    packed projections (the strides the model uses).  Tolerance: probabilities and outputs are rounded to bf16
    (2^-8 relative) -> 2e-2 absolute on O(1) outputs."""
    B = 7
    g = torch.Generator(device="cpu").manual_seed(H * 1000 + Sq * 10 + S)
    qkv = torch.randn(B, S, 3, H, D, generator=g).to(torch.bfloat16).cuda()
    qsrc = torch.randn(B, Sq, H, D, generator=g).to(torch.bfloat16).cuda()
    k, v = qkv[:, :, 1], qkv[:, :, 2]
    q = qkv[:, :, 0] if Sq == S else qsrc
    before = phe._native.launch_count()
    got = pv._attention(q, k, v)
    assert phe._native.launch_count() == before + 1
    qf, kf, vf = q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)
    want = torch.softmax(qf @ kf.transpose(-1, -2) / D ** 0.5, dim=-1) @ vf
    want = want.transpose(1, 2).reshape(B, Sq, H * D)
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    np.testing.assert_allclose(got.float().cpu().numpy(), want.cpu().numpy(), atol=2e-2, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("shift", [-40.0, -200.0, 30.0])
def test_attention_kernel_strongly_shifted_logits_gpu(shift):
    """Softmax is shift invariant, so a checkpoint may produce rows whose real logits are all far below zero (or far above).
    Padded key columns (98 -> 112) must not take part in the row maximum or the row sum: every query row sees logits
    q.k/sqrt(D) = shift + noise here (a constant component in q and k), against PyTorch's SDPA in fp32."""
    B, H, D, S = 5, 10, 64, 98
    g = torch.Generator(device="cpu").manual_seed(int(abs(shift)))
    q = torch.randn(B, S, H, D, generator=g) * 0.3
    k = torch.randn(B, S, H, D, generator=g) * 0.3
    v = torch.randn(B, S, H, D, generator=g)
    a = (abs(shift) * D ** 0.5) ** 0.5                       # q0 * k0 / sqrt(D) = shift on the first dimension
    q[..., 0] = a
    k[..., 0] = a if shift > 0 else -a
    q, k, v = (t.to(torch.bfloat16).cuda().contiguous() for t in (q, k, v))
    got = pv._attention(q, k, v)
    want = torch.nn.functional.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2))
    want = want.transpose(1, 2).reshape(B, S, H * D)
    assert torch.isfinite(got.float()).all()
    np.testing.assert_allclose(got.float().cpu().numpy(), want.cpu().numpy(), atol=3e-2, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("B", [1, 3, 64, 1000])
def test_policy_value_heads_kernel_gpu(golden, dtype, tol, B):
    """phe_nn_policy_value_heads against the PyTorch expression it replaces (Linear + softmax, Linear) on the same
    encoder rows, computed in fp32 from the same (rounded) weights."""
    m = _model(golden[0]).to("cuda", dtype).eval()
    g = torch.Generator(device="cpu").manual_seed(B)
    h = torch.randn(B, 2, m.d_model, generator=g).to(dtype).cuda()
    before = phe._native.launch_count()
    with torch.no_grad():
        out = m.policy_value_packed(h)
        wp = torch.softmax(h[:, 0].float() @ m.action_prob_dense.weight.float().t() + m.action_prob_dense.bias.float(), dim=1)
        wv = h[:, 1].float() @ m.reward_value_dense.weight.float().t() + m.reward_value_dense.bias.float()
    assert phe._native.launch_count() == before + 1
    assert out.dtype == torch.float32 and out.shape == (B, 13)
    np.testing.assert_allclose(out[:, :12].cpu().numpy(), wp.cpu().numpy(), atol=tol, rtol=0)
    np.testing.assert_allclose(out[:, 12].cpu().numpy(), wv[:, 0].cpu().numpy(), atol=tol * 4, rtol=0)
    np.testing.assert_allclose(out[:, :12].sum(1).cpu().numpy(), 1.0, atol=1e-5)
