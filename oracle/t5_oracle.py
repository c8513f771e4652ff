"""CPU oracle for the umT5 text encoder that feeds the DiT's cross-attention.  TEST INFRASTRUCTURE ONLY.

A from-scratch functional restatement (torch, fp32 on CPU) of the reference's `WanT5EncoderModel.forward`
(videox_fun/models/wan_text_encoder.py:256-304) and everything under it; SURVEY.md §8f rank 3.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this; the product (videocof_b200/) never does.

Pinned: tools/gen_golden_t5.py runs the UNMODIFIED reference file from /root/reference (tools/ref_loader.py) on the
deterministic parameters of `make_t5_params` and commits the outputs (tests/golden/t5_*.npz);
tests/test_t5_oracle.py checks this file against them.  The reference ships no tests of its own.

Line numbers below are those of videox_fun/models/wan_text_encoder.py.

`emulate_bf16=True` rounds where the reference's bf16 eager path (`text_encoder.to(torch.bfloat16)`,
fast_infer.py:308-312) holds bf16 tensors AND the CUDA kernels do too (layer inputs / outputs, projections,
probabilities, the gated product); with False it is the fp32 gold.
"""
import math

import torch

__all__ = ["T5Config", "make_t5_params", "relative_position_bucket", "position_bias", "t5_forward",
           "t5_layer_norm", "t5_attention", "gelu_tanh"]


class T5Config:
    """Constructor arguments of WanT5EncoderModel (:257-266).  umT5-XXL (Wan-2.1's text encoder, config.json of the
    checkpoint): vocab 256384, dim 4096, dim_attn 4096, dim_ffn 10240, 64 heads, 24 layers, 32 buckets, per-layer
    position embeddings (shared_pos False)."""

    def __init__(self, vocab=256384, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=24,
                 num_buckets=32, shared_pos=False, max_dist=128, eps=1e-6):
        self.vocab, self.dim, self.dim_attn, self.dim_ffn = vocab, dim, dim_attn, dim_ffn
        self.num_heads, self.num_layers, self.num_buckets = num_heads, num_layers, num_buckets
        self.shared_pos, self.max_dist, self.eps = shared_pos, max_dist, eps

    @property
    def head_dim(self):
        return self.dim_attn // self.num_heads

    def to_kwargs(self):
        return dict(vocab=self.vocab, dim=self.dim, dim_attn=self.dim_attn, dim_ffn=self.dim_ffn,
                    num_heads=self.num_heads, num_layers=self.num_layers, num_buckets=self.num_buckets,
                    shared_pos=self.shared_pos, dropout=0.0)


def _rb(x, on):
    return x.to(torch.bfloat16).to(torch.float32) if on else x


def make_t5_params(cfg, seed=0, bf16_exact=True):
    """Deterministic parameters keyed by the reference's state-dict names.  Scales follow init_weights (:21-36)
    except the position-embedding table, which gets unit-scale entries so the bias visibly shapes the softmax."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=1.0, mean=0.0):
        t = mean + torch.randn(*shape, generator=g, dtype=torch.float32) * std
        return t.to(torch.bfloat16).to(torch.float32) if bf16_exact else t

    D, A, Fd, H = cfg.dim, cfg.dim_attn, cfg.dim_ffn, cfg.num_heads
    p = {"token_embedding.weight": rnd(cfg.vocab, D)}
    if cfg.shared_pos:
        p["pos_embedding.embedding.weight"] = rnd(cfg.num_buckets, H)
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        p[b + "norm1.weight"] = rnd(D, std=0.05, mean=1.0)
        p[b + "attn.q.weight"] = rnd(A, D, std=(D * cfg.head_dim) ** -0.5 * 4)   # x4: logits of O(1), not O(0.1)
        p[b + "attn.k.weight"] = rnd(A, D, std=D ** -0.5)
        p[b + "attn.v.weight"] = rnd(A, D, std=D ** -0.5)
        p[b + "attn.o.weight"] = rnd(D, A, std=A ** -0.5)
        p[b + "norm2.weight"] = rnd(D, std=0.05, mean=1.0)
        p[b + "ffn.gate.0.weight"] = rnd(Fd, D, std=D ** -0.5)
        p[b + "ffn.fc1.weight"] = rnd(Fd, D, std=D ** -0.5)
        p[b + "ffn.fc2.weight"] = rnd(D, Fd, std=Fd ** -0.5)
        if not cfg.shared_pos:
            p[b + "pos_embedding.embedding.weight"] = rnd(cfg.num_buckets, H)
    p["norm.weight"] = rnd(D, std=0.05, mean=1.0)
    return p


def relative_position_bucket(rel_pos, num_buckets=32, max_dist=128, bidirectional=True):
    """:224-247 — T5's log-spaced bucketing of (key index - query index); int64 tensor in, int64 buckets out.
    The fp32 log / division / truncation order is the reference's, so bucket boundaries fall identically."""
    if bidirectional:
        nb = num_buckets // 2
        buckets = (rel_pos > 0).long() * nb
        rel_pos = torch.abs(rel_pos)
    else:
        nb = num_buckets
        buckets = torch.zeros_like(rel_pos)
        rel_pos = -torch.min(rel_pos, torch.zeros_like(rel_pos))
    max_exact = nb // 2
    large = max_exact + (torch.log(rel_pos.float() / max_exact) / math.log(max_dist / max_exact)
                         * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(rel_pos < max_exact, rel_pos, large)


def position_bias(table, lq, lk, num_buckets=32, max_dist=128):
    """:207-222 — bias[h, i, j] = table[bucket(j - i), h]; table is the nn.Embedding weight [buckets, heads]."""
    rel = torch.arange(lk).unsqueeze(0) - torch.arange(lq).unsqueeze(1)
    b = relative_position_bucket(rel, num_buckets, max_dist, True)
    return table.float()[b].permute(2, 0, 1).contiguous()


def gelu_tanh(x):
    """:39-42 — the module spells the tanh approximation out."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def t5_layer_norm(x, w, eps, emu=False):
    """:45-57 — x * rsqrt(mean(x^2) + eps) in fp32, cast to the weight dtype, then * weight."""
    y = x.float() * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    return _rb(w.float() * _rb(y, emu), emu)


def t5_attention(p, pre, x, mask, bias, cfg, emu=False):
    """:60-112 — bias-free q/k/v/o projections, NO 1/sqrt(d) scaling, additive position bias, keys with
    mask == 0 get the most negative bf16 value instead of their bias (masked_fill_, :98), fp32 softmax.
    x [L, dim]; mask [L] (0/1) or None; bias [heads, L, L]."""
    n, c = cfg.num_heads, cfg.head_dim
    lin = torch.nn.functional.linear
    q = _rb(lin(x, p[pre + "q.weight"].float()), emu).view(-1, n, c)
    k = _rb(lin(x, p[pre + "k.weight"].float()), emu).view(-1, n, c)
    v = _rb(lin(x, p[pre + "v.weight"].float()), emu).view(-1, n, c)
    ab = bias.clone()
    if mask is not None:
        ab.masked_fill_((mask == 0).view(1, 1, -1), torch.finfo(torch.bfloat16).min)
    s = torch.einsum("inc,jnc->nij", q, k) + ab
    a = _rb(torch.softmax(s.float(), dim=-1), emu)
    o = _rb(torch.einsum("nij,jnc->inc", a, v).reshape(-1, n * c), emu)
    return _rb(lin(o, p[pre + "o.weight"].float()), emu)


def t5_forward(p, cfg, ids, mask=None, emulate_bf16=False):
    """:281-294 (+ block :153-158, ffn :128-133).  ids int64 [B, L]; mask [B, L] of 0/1 or None -> [B, L, dim] fp32.
    Dropout is inert (eval)."""
    emu = emulate_bf16
    lin = torch.nn.functional.linear
    outs = []
    for b in range(ids.shape[0]):
        x = p["token_embedding.weight"].float()[ids[b]]
        L = x.shape[0]
        m = None if mask is None else mask[b]
        shared = position_bias(p["pos_embedding.embedding.weight"], L, L, cfg.num_buckets, cfg.max_dist) \
            if cfg.shared_pos else None
        for i in range(cfg.num_layers):
            pre = f"blocks.{i}."
            e = shared if cfg.shared_pos else position_bias(p[pre + "pos_embedding.embedding.weight"], L, L,
                                                            cfg.num_buckets, cfg.max_dist)
            x = _rb(x + t5_attention(p, pre + "attn.", t5_layer_norm(x, p[pre + "norm1.weight"], cfg.eps, emu),
                                     m, e, cfg, emu), emu)
            h = t5_layer_norm(x, p[pre + "norm2.weight"], cfg.eps, emu)
            g = _rb(gelu_tanh(_rb(lin(h, p[pre + "ffn.gate.0.weight"].float()), emu)), emu)
            u = _rb(_rb(lin(h, p[pre + "ffn.fc1.weight"].float()), emu) * g, emu)
            x = _rb(x + _rb(lin(u, p[pre + "ffn.fc2.weight"].float()), emu), emu)
        outs.append(t5_layer_norm(x, p["norm.weight"], cfg.eps, emu))
    return torch.stack(outs)
