"""Batched policy / value inference on the device the equity kernels run on.

Replaces, for inference only:
  * ``rl/models.py:12-115``  ``EnvEmbedding``  (field + position embeddings of the int32[93] observation)
  * ``rl/models.py:402-481`` ``TransformerAlphaGoZeroModel`` (post-LN ``nn.TransformerEncoder`` + five heads)
  * ``rl/MCTS.py:464-473``   ``SingleThreadMCTS.predict(observation) -> (action_prob float32[A], value float32)``
  * ``env/workflow.py:274-593`` ``predict_batch_process`` + ``utils/tensorrt_utils.py:18-290`` (TensorRT engines with one
    set of static buffers per batch size) -> one CUDA graph per batch bucket, pinned staging buffers, PyTorch kernels.

``PolicyValueNet`` keeps the reference's parameter names and shapes, so a reference checkpoint's ``state_dict`` loads
unchanged (``load_reference_state_dict`` also accepts the ``AlphaGoZero`` wrapper's ``model.``-prefixed keys).  It is a
plain PyTorch module -- north_star: "utils/tensorrt_utils.py is replaced by PyTorch on the same device" -- re-laid out
for inference: the position half of every token and the starter tokens are constants folded at load time, and the
last encoder layer only computes the query rows whose heads are read (the five starter tokens; two of them for
``predict``), which removes 1/num_layers of the work for the full model and almost half of it for config/debug.yml's
two layers.  Training, optimiser, replay memory and the queue plumbing stay in the reference.
"""
import ctypes as C
import math
import threading

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _native as N
from ._native import PheError

# ---- observation layout ---------------------------------------------------------------------------------------
# env/constants.py: MAX_PLAYER_NUMBER (:4), CardFigure 13 (:14-28), CardDecor 4 (:31-36), PlayerStatus 4 (:47-51),
# CUTTER_BINS_LIST 8 (:108), CUTTER_SELF_BINS_LIST 17 (:109), CUTTER_DEFAULT_LIST 9 (:110), CUTTER_SELF_DEFAULT_LIST 19 (:111);
# utils/workflow_utils.py:179-184 get_num_feature_bins_v2() = 10 with the shipped ACTION_BINS_DICT.
MAX_PLAYER_NUMBER = 2
N_FIGURES, N_DECORS, N_PLAYER_STATUS = 13, 4, 4
N_CUT_BINS, N_CUT_SELF_BINS, N_CUT_DEFAULT, N_CUT_SELF_DEFAULT = 8, 17, 9, 19
N_ACTION_FEATURE_BINS = 10
NUM_STARTERS = 5


def field_dims(num_output_class=12, num_bins=10, historical_action_per_round=3, n_action_bins=N_ACTION_FEATURE_BINS,
               max_players=MAX_PLAYER_NUMBER):
    """[(number of consecutive observation fields, vocabulary size)] in observation order (rl/models.py:39-67)."""
    return [(1, 4), (1, max_players + 1), (7, N_FIGURES + 2), (7, N_DECORS + 2),
            (1, N_CUT_DEFAULT), (1, N_CUT_DEFAULT), (1, N_CUT_DEFAULT),
            (1, N_CUT_DEFAULT), (1, N_CUT_BINS), (1, N_CUT_BINS), (1, N_CUT_SELF_BINS),
            (1, N_CUT_BINS), (1, N_CUT_BINS), (1, N_CUT_SELF_BINS),
            (n_action_bins, N_CUT_SELF_DEFAULT + 1), (n_action_bins, N_CUT_BINS + 1), (n_action_bins, N_CUT_BINS + 1),
            (n_action_bins, N_CUT_BINS + 1),
            (max_players - 1, N_PLAYER_STATUS), (max_players - 1, num_bins), (max_players - 1, num_bins + 1),
            (4 * historical_action_per_round * max_players, num_output_class + 1)]


def field_layout(**kw):
    """(field_start int32[F], vocabulary size per field int32[F], total embedding rows) -- rl/models.py:68-76."""
    starts, sizes, cur = [], [], NUM_STARTERS
    for n, dim in field_dims(**kw):
        starts += [cur] * n
        sizes += [dim] * n
        cur += dim
    return np.asarray(starts, np.int32), np.asarray(sizes, np.int32), cur


def random_observations(n, seed=0, **kw):
    """Synthetic observations: every field uniform over its own vocabulary (valid input for the embedding)."""
    _, sizes, _ = field_layout(**kw)
    rng = np.random.default_rng(seed)
    return (rng.random((n, sizes.shape[0])) * sizes[None, :]).astype(np.int32)


# ---- the network ------------------------------------------------------------------------------------------------
class _EnvEmbedding(nn.Module):
    """Parameter container with the reference's names (rl/models.py:68-98)."""

    def __init__(self, embedding_dim, positional_embedding_dim, starts, n_rows):
        super().__init__()
        seq = NUM_STARTERS + starts.shape[0]
        self.register_buffer("starter_idx_array", torch.arange(NUM_STARTERS, dtype=torch.int32))
        self.register_buffer("field_start_idx_array", torch.from_numpy(starts.copy()))
        self.register_buffer("position_embedding_idx", torch.arange(seq, dtype=torch.int32).unsqueeze(0))
        self.field_embedding = nn.Embedding(n_rows, embedding_dim)
        self.position_embedding = nn.Embedding(seq, positional_embedding_dim)


class _SelfAttention(nn.Module):
    def __init__(self, d, heads):
        super().__init__()
        self.heads = heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)


_DTYPES = {torch.float32: 0, torch.bfloat16: 1}        # PHE_DTYPE_F32 / PHE_DTYPE_BF16


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _native_ok(*tensors):
    return all(t.is_cuda and t.dtype in _DTYPES and t.dtype == tensors[0].dtype for t in tensors)


def _add_layernorm(x, residual, norm):
    """LayerNorm(x + residual): phe_nn_add_layernorm on the device (one fused pass), plain ops on the host."""
    d = x.shape[-1]
    if _native_ok(x, residual, norm.weight) and (d * x.element_size()) % 16 == 0:
        x, residual = x.contiguous(), residual.contiguous()
        out = torch.empty_like(x)
        N.check(N.lib().phe_nn_add_layernorm(_ptr(x), _ptr(residual), _ptr(norm.weight), _ptr(norm.bias), _ptr(out), x.numel() // d, d,
                                             float(norm.eps), _DTYPES[x.dtype], _stream()))
        return out
    return F.layer_norm(x + residual, (d,), norm.weight, norm.bias, norm.eps)


def _attention(q, k, v):
    """softmax(q k^T / sqrt(D)) v for q [B, Sq, H, D], k / v [B, S, H, D] (views into the packed projections) -> [B, Sq, H*D].
    bf16 on the device with S <= 112 and D in (16, 64): phe_nn_attention (one CTA per (batch, head), csrc/phe_nn.cu);
    otherwise PyTorch's scaled_dot_product_attention."""
    B, Sq, H, D = q.shape
    S = k.shape[1]
    if (q.is_cuda and q.dtype == torch.bfloat16 and D in (16, 64) and S <= 112 and Sq <= 112 and k.stride() == v.stride()
            and q.stride(3) == 1 and k.stride(3) == 1 and q.stride(2) == D and k.stride(2) == D
            and all(st % 8 == 0 for st in (q.stride(0), q.stride(1), k.stride(0), k.stride(1)))
            and all(t.data_ptr() % 16 == 0 for t in (q, k, v))):
        out = torch.empty((B, Sq, H * D), dtype=q.dtype, device=q.device)
        N.check(N.lib().phe_nn_attention(_ptr(q), _ptr(k), _ptr(v), _ptr(out), B, H, Sq, S, D, q.stride(0), q.stride(1), k.stride(0), k.stride(1),
                                         _DTYPES[q.dtype], _stream()))
        return out
    a = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    return a.transpose(1, 2).reshape(B, Sq, H * D)


def _linear_relu(x, lin):
    """relu(x @ W^T + b) with bias and ReLU in the GEMM epilogue (cuBLASLt) on the device; plain ops on the host."""
    if x.is_cuda:
        shp = x.shape
        return torch._addmm_activation(lin.bias, x.reshape(-1, shp[-1]), lin.weight.t(), use_gelu=False).view(*shp[:-1], lin.out_features)
    return F.relu(F.linear(x, lin.weight, lin.bias))


class _EncoderLayer(nn.Module):
    """Post-LN encoder layer, ReLU feed-forward (nn.TransformerEncoderLayer defaults as used at rl/models.py:446-447)."""

    def __init__(self, d, heads, dim_feedforward=2048, eps=1e-5):
        super().__init__()
        self.self_attn = _SelfAttention(d, heads)
        self.linear1 = nn.Linear(d, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d)
        self.norm1 = nn.LayerNorm(d, eps=eps)
        self.norm2 = nn.LayerNorm(d, eps=eps)

    def forward(self, x, n_query=None):
        """x [B, S, d]; with n_query only the first n_query rows are produced (keys / values still see all S)."""
        B, S, d = x.shape
        sa, H = self.self_attn, self.self_attn.heads
        if n_query is None or n_query >= S:
            qkv = F.linear(x, sa.in_proj_weight, sa.in_proj_bias).view(B, S, 3, H, d // H)
            q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
            xq = x
        else:
            xq = x[:, :n_query]
            q = F.linear(xq, sa.in_proj_weight[:d], sa.in_proj_bias[:d]).view(B, n_query, H, d // H)
            kv = F.linear(x, sa.in_proj_weight[d:], sa.in_proj_bias[d:]).view(B, S, 2, H, d // H)
            k, v = kv[:, :, 0], kv[:, :, 1]
        a = _attention(q, k, v)
        y = _add_layernorm(sa.out_proj(a), xq, self.norm1)
        return _add_layernorm(self.linear2(_linear_relu(y, self.linear1)), y, self.norm2)


class _Encoder(nn.Module):
    def __init__(self, d, heads, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayer(d, heads) for _ in range(num_layers)])


class PolicyValueNet(nn.Module):
    """Inference form of ``TransformerAlphaGoZeroModel`` (rl/models.py:402-481); same state_dict keys."""

    HEADS = ("action_prob", "reward_value", "card_result_value", "player_winning_prob", "opponent_winning_prob")

    def __init__(self, num_bins=10, num_winning_prob_bins=10, num_output_class=12, embedding_dim=512, positional_embedding_dim=128,
                 num_layers=6, transformer_head_dim=64, historical_action_per_round=6, num_acting_player_fields=9,
                 num_other_player_fields=3, **_ignored):
        super().__init__()
        d = int(embedding_dim + positional_embedding_dim)
        if d % transformer_head_dim:
            raise ValueError(f"embedding_dim + positional_embedding_dim ({d}) must be divisible by transformer_head_dim ({transformer_head_dim})")
        self.d_model, self.embedding_dim, self.num_output_class = d, embedding_dim, num_output_class
        self.scale = math.sqrt(d)
        starts, sizes, n_rows = field_layout(num_output_class=num_output_class, num_bins=num_bins,
                                             historical_action_per_round=historical_action_per_round)
        # rl/models.py:428: the reference's sequence length formula must agree with the field list
        expect = 2 + 2 * 7 + num_acting_player_fields + num_other_player_fields * (MAX_PLAYER_NUMBER - 1) + 4 * historical_action_per_round * MAX_PLAYER_NUMBER
        if expect != starts.shape[0]:
            raise ValueError(f"observation has {starts.shape[0]} fields but the model parameters describe {expect}")
        self.n_fields, self.seq_len = int(starts.shape[0]), NUM_STARTERS + int(starts.shape[0])
        self.register_buffer("_field_sizes", torch.from_numpy(sizes.copy()), persistent=False)
        self.env_embedding = _EnvEmbedding(embedding_dim, positional_embedding_dim, starts, n_rows)
        self.transform_encoder = _Encoder(d, d // transformer_head_dim, num_layers)
        self.action_prob_dense = nn.Linear(d, num_output_class)
        self.reward_value_dense = nn.Linear(d, 1)
        self.card_result_value_dense = nn.Linear(d, 1)
        self.player_winning_prob_dense = nn.Linear(d, num_winning_prob_bins)
        self.opponent_winning_prob_dense = nn.Linear(d, num_winning_prob_bins)
        self._folded = None

    # -- loading ---------------------------------------------------------------------------------------------------
    def load_reference_state_dict(self, state_dict):
        """Accepts the state_dict of the reference model or of its ``AlphaGoZero`` wrapper (keys prefixed ``model.``)."""
        sd = {(k[6:] if k.startswith("model.") else k): v for k, v in state_dict.items()}
        return self.load_state_dict(sd, strict=True)        # in place; _fold() notices the new tensor versions

    def _fold(self):
        """Inference-time tables: position half of every token, the five starter tokens and the sqrt(d) scaling folded
        together.  They live in PERSISTENT buffers: when the weights change in place (load_state_dict, an optimiser step)
        the tables are recomputed into the same storage, so CUDA graphs that captured their addresses stay valid; the
        buffers are only re-allocated when the weights move to another device or dtype (BatchedPredictor re-captures then)."""
        w = self.env_embedding.field_embedding.weight
        pw = self.env_embedding.position_embedding.weight
        place, version = (w.device, w.dtype), (w._version, pw._version, w.data_ptr(), pw.data_ptr())
        if self._folded is not None and self._folded[0] == place and self._folded[1] == version:
            return self._folded[2:]
        with torch.no_grad():
            pos = (pw * self.scale).to(w.dtype)                                                       # [S, dp]
            fresh = (pos[NUM_STARTERS:], torch.cat([w[:NUM_STARTERS] * self.scale, pos[:NUM_STARTERS]], dim=1), w * self.scale)
            if self._folded is not None and self._folded[0] == place:
                for buf, val in zip(self._folded[2:], fresh):
                    buf.copy_(val)
                self._folded = (place, version) + self._folded[2:]
            else:
                self._folded = (place, version) + tuple(t.contiguous().clone() for t in fresh)
        return self._folded[2:]

    # -- forward ---------------------------------------------------------------------------------------------------
    def embed(self, x):
        """int32/int64 [B, F] observation -> [B, S, d] encoder input (rl/models.py:100-114, :462-466)."""
        B = x.shape[0]
        pos, starters, field = self._fold()
        if x.is_cuda and field.dtype in _DTYPES and (self.embedding_dim * field.element_size()) % 16 == 0 and (pos.shape[1] * field.element_size()) % 16 == 0:
            # one gather-copy kernel (csrc/phe_nn.cu) instead of gather, scale and two concatenations
            xi = x.to(torch.int32).contiguous()
            out = torch.empty((B, self.seq_len, self.d_model), dtype=field.dtype, device=x.device)
            N.check(N.lib().phe_nn_embed_obs(_ptr(xi), _ptr(self.env_embedding.field_start_idx_array), _ptr(field), _ptr(pos), _ptr(starters),
                                             _ptr(out), B, self.n_fields, NUM_STARTERS, self.embedding_dim, pos.shape[1], field.shape[0],
                                             _DTYPES[field.dtype], _stream()))
            return out
        idx = x.to(torch.int64) + self.env_embedding.field_start_idx_array.to(torch.int64)
        tok = F.embedding(idx, self.env_embedding.field_embedding.weight) * self.scale          # [B, F, de]
        tok = torch.cat([tok, pos.unsqueeze(0).expand(B, -1, -1)], dim=2)
        return torch.cat([starters.unsqueeze(0).expand(B, -1, -1), tok], dim=1)

    def encode(self, x, n_out=NUM_STARTERS):
        """Encoder output rows [B, n_out, d] of the first n_out tokens (the heads read tokens 0..4 only)."""
        h = self.embed(x)
        layers = self.transform_encoder.layers
        for layer in layers[:-1]:
            h = layer(h)
        return layers[-1](h, n_query=n_out)

    def forward(self, x):
        """Same five outputs as the reference forward (rl/models.py:462-481)."""
        h = self.encode(x)
        return (self.action_prob_dense(h[:, 0]), self.reward_value_dense(h[:, 1]).squeeze(1),
                torch.sigmoid(self.card_result_value_dense(h[:, 2])).squeeze(1),
                self.player_winning_prob_dense(h[:, 3]), self.opponent_winning_prob_dense(h[:, 4]))

    def policy_value(self, x):
        """What MCTS consumes (rl/MCTS.py:464-473): softmax action probabilities float32[B, A] and value float32[B]."""
        h = self.encode(x, n_out=2)
        out = self.policy_value_packed(h)
        if out is not None:
            return out[:, : self.num_output_class], out[:, self.num_output_class]
        return (torch.softmax(self.action_prob_dense(h[:, 0]).float(), dim=1), self.reward_value_dense(h[:, 1]).squeeze(1).float())

    def policy_value_packed(self, h, out=None):
        """Encoder rows [B, 2, d] -> float32 [B, A + 1] (probabilities, then the value) in one kernel
        (phe_nn_policy_value_heads); None when the tensors are not on the device / not a supported type."""
        w = self.action_prob_dense.weight
        if not (h.is_cuda and h.dtype in _DTYPES and w.dtype == h.dtype and self.num_output_class <= 32 and (self.d_model * h.element_size()) % 16 == 0):
            return None
        h = h.contiguous()
        B = h.shape[0]
        if out is None:
            out = torch.empty((B, self.num_output_class + 1), dtype=torch.float32, device=h.device)
        N.check(N.lib().phe_nn_policy_value_heads(_ptr(h), _ptr(w), _ptr(self.action_prob_dense.bias), _ptr(self.reward_value_dense.weight),
                                                  _ptr(self.reward_value_dense.bias), _ptr(out), B, self.d_model, self.num_output_class,
                                                  self.d_model, _DTYPES[h.dtype], _stream()))
        return out

    def check_observations(self, x):
        """ValueError if an observation would index outside the embedding table.

        Quirk kept from the reference: a cutter with n thresholds emits n + 1 bins (n + 2 with the invalid bin,
        tools/biner.py:2-5) but rl/models.py:39-67 allots the field only n (n + 1) rows, so the top bin of every ratio
        feature reads the first row of the NEXT field's range.  Env.get_obs produces such values routinely
        (tests/golden/obs_run.json.gz), so they are accepted here as they are by the reference; only indices that
        leave the table are rejected."""
        x = torch.as_tensor(x)
        if x.dim() != 2 or x.shape[1] != self.n_fields:
            raise ValueError(f"observations must be [B, {self.n_fields}], got {tuple(x.shape)}")
        idx = x.to(torch.int64) + self.env_embedding.field_start_idx_array.to(x.device, torch.int64)
        bad = (x < 0) | (idx >= self.env_embedding.field_embedding.num_embeddings)
        if bool(bad.any()):
            b, f = [int(v) for v in bad.nonzero()[0]]
            raise ValueError(f"observation {b}: field {f} = {int(x[b, f])} indexes outside the embedding table")


# ---- the predictor -------------------------------------------------------------------------------------------------
class BatchedPredictor:
    """CUDA-graph replayed ``policy_value`` over static buffers, one graph per batch bucket.

    ``predict(observation)`` is the drop-in for ``SingleThreadMCTS.predict`` (rl/MCTS.py:464-473);
    ``predict_batch(observations)`` is what a lock-step tree driver or the reference's ``predict_batch_process`` loop
    calls with everything that is waiting (env/workflow.py:274-593).  Host arrays go through pinned staging buffers
    (one H2D copy of B*F int32, one D2H copy of B*(A+1) float32 per call); device tensors are used in place.
    There is no CPU path: constructing the predictor without a CUDA device raises ``PheError``.
    """

    def __init__(self, model, device=None, dtype=torch.bfloat16, max_batch=1024, buckets=None):
        if not torch.cuda.is_available():
            raise PheError("BatchedPredictor needs a CUDA device (there is no CPU inference path)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.model = model.to(self.device, dtype).eval()
        for p in self.model.parameters():
            p.requires_grad_(False)
        self.dtype = dtype
        self.n_fields, self.n_actions = model.n_fields, model.num_output_class
        if buckets is None:
            buckets, b = [], 1
            while b < max_batch:
                buckets.append(b)
                b *= 4
            buckets.append(max_batch)
        self.buckets = sorted(set(int(b) for b in buckets))
        self.max_batch = self.buckets[-1]
        self._lock = threading.Lock()
        self._graphs = {}
        self._fold_token = None
        self._stream = torch.cuda.Stream(device=self.device)
        self._pin_in = torch.empty((self.max_batch, self.n_fields), dtype=torch.int32).pin_memory()
        self._pin_out = torch.empty((self.max_batch, self.n_actions + 1), dtype=torch.float32).pin_memory()
        self.replays = 0

    def _bucket(self, n):
        for b in self.buckets:
            if n <= b:
                return b
        return self.max_batch

    def _graph(self, b):
        g = self._graphs.get(b)
        if g is None:
            x = torch.zeros((b, self.n_fields), dtype=torch.int32, device=self.device)
            out = torch.empty((b, self.n_actions + 1), dtype=torch.float32, device=self.device)

            def step():
                h = self.model.encode(x, n_out=2)
                if self.model.policy_value_packed(h, out=out) is None:         # not reached on a CUDA device with f32 / bf16 weights
                    p, v = self.model.policy_value(x)
                    out[:, : self.n_actions] = p
                    out[:, self.n_actions] = v

            with torch.no_grad(), torch.cuda.stream(self._stream):
                for _ in range(3):          # warm up allocator, cuBLAS handles and attention kernel selection
                    step()
                self._stream.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=self._stream):
                    step()
            g = self._graphs[b] = (graph, x, out)
        return g

    def _refresh_tables(self):
        """Called on the predictor's stream before every replay: refolds the embedding tables if the weights changed in
        place (same storage: the captured graphs stay valid) and drops the graphs if the tables were re-allocated."""
        token = tuple(t.data_ptr() for t in self.model._fold())
        if token != self._fold_token:
            self._graphs.clear()
            self._fold_token = token

    def load_reference_state_dict(self, state_dict):
        """Weight update of a live predictor (the reference's predict_batch_process reloads checkpoints through
        update_model_param_queue, env/workflow.py:274-593): parameters are overwritten IN PLACE on the predictor's stream,
        the folded tables follow in place, and the captured graphs keep replaying with the new weights."""
        with self._lock, torch.cuda.stream(self._stream), torch.no_grad():
            res = self.model.load_reference_state_dict(state_dict)
            self._refresh_tables()
            self._stream.synchronize()
        return res

    update_weights = load_reference_state_dict

    def warmup(self):
        with self._lock, torch.cuda.stream(self._stream):
            self._refresh_tables()
            for b in self.buckets:
                self._graph(b)
        torch.cuda.synchronize(self.device)

    @torch.no_grad()
    def predict_batch(self, observations, check=False):
        """observations int32 [B, F] (numpy / host tensor / device tensor) -> (action_probs float32[B, A], value float32[B]),
        numpy for host input, device tensors for device input."""
        on_device = isinstance(observations, torch.Tensor) and observations.is_cuda
        obs = observations if isinstance(observations, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(observations, dtype=np.int32))
        if obs.dim() != 2 or obs.shape[1] != self.n_fields:
            raise ValueError(f"observations must be [B, {self.n_fields}], got {tuple(obs.shape)}")
        if check:
            self.model.check_observations(obs)
        B = obs.shape[0]
        probs = torch.empty((B, self.n_actions), dtype=torch.float32, device=self.device if on_device else "cpu")
        value = torch.empty((B,), dtype=torch.float32, device=probs.device)
        caller = torch.cuda.current_stream(self.device)
        with self._lock, torch.cuda.stream(self._stream):
            if on_device:
                self._stream.wait_stream(caller)
            self._refresh_tables()
            for lo in range(0, B, self.max_batch):
                n = min(self.max_batch, B - lo)
                graph, x, out = self._graph(self._bucket(n))
                if on_device:
                    x[:n].copy_(obs[lo: lo + n], non_blocking=True)
                else:
                    self._pin_in[:n].copy_(obs[lo: lo + n])
                    x[:n].copy_(self._pin_in[:n], non_blocking=True)
                graph.replay()
                self.replays += 1
                if on_device:
                    probs[lo: lo + n].copy_(out[:n, : self.n_actions], non_blocking=True)
                    value[lo: lo + n].copy_(out[:n, self.n_actions], non_blocking=True)
                else:
                    self._pin_out[:n].copy_(out[:n], non_blocking=True)
                    self._stream.synchronize()
                    probs[lo: lo + n].copy_(self._pin_out[:n, : self.n_actions])
                    value[lo: lo + n].copy_(self._pin_out[:n, self.n_actions])
            if on_device:
                caller.wait_stream(self._stream)
        if on_device:
            return probs, value
        return probs.numpy(), value.numpy()

    def predict(self, observation):
        """``SingleThreadMCTS.predict`` (rl/MCTS.py:464-473): one observation -> (action_prob float32[A], value float32 scalar)."""
        p, v = self.predict_batch(np.asarray(observation, dtype=np.int32).reshape(1, -1))
        return p[0], v[0]


# parameter sets of the reference's configurations (config/default.yml:44-66, config/debug.yml:24-30 over the defaults)
DEFAULT_MODEL_PARAMS = dict(num_bins=10, num_winning_prob_bins=10, num_output_class=12, embedding_dim=512, positional_embedding_dim=128,
                            num_layers=8, transformer_head_dim=64, historical_action_per_round=3, num_acting_player_fields=50,
                            num_other_player_fields=3)
DEBUG_MODEL_PARAMS = dict(DEFAULT_MODEL_PARAMS, embedding_dim=64, positional_embedding_dim=32, transformer_head_dim=16, num_layers=2)


def fill_parameters(module, seed, std=0.05):
    """Deterministic synthetic weights: every floating tensor of the state_dict, in sorted key order, from one seeded CPU
    generator (normal * std; LayerNorm weights around 1).  tests/golden/make_model_golden.py fills the REFERENCE model
    with this same routine, so fixtures need not carry weights."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = module.state_dict()
    with torch.no_grad():
        for k in sorted(sd):
            v = sd[k]
            if not v.dtype.is_floating_point:
                continue
            r = torch.randn(v.shape, generator=g, dtype=torch.float32) * std
            if ".norm" in k and k.endswith("weight"):
                r += 1.0
            v.copy_(r.to(v.dtype))
    return module
