"""CPU oracle for the Wan-2.1 DiT forward of VideoCoF.  TEST INFRASTRUCTURE ONLY.

A from-scratch restatement (functional torch, fp32 on CPU) of the reference's
`WanTransformer3DModel.forward` and everything under it.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this; the product path (videocof_b200/) never does.

Pinned: tools/gen_golden.py runs the UNMODIFIED reference files from /root/reference
(under a small diffusers shim) on the same deterministic parameters and inputs and commits
the outputs under tests/golden/; tests/test_oracle_golden.py checks this file against them.
The reference itself ships no tests or golden vectors.

Every function cites the reference lines it restates (paths relative to the reference
repo, videox_fun/models/wan_transformer3d.py unless another file is named).

`emulate_bf16=True` inserts bf16 roundings where the reference's CUDA path (bf16 weights
under torch.autocast(bf16)) produces bf16 tensors (SURVEY.md §3.3 dtype trace), so that
the CUDA kernels can be compared with a tight tolerance; with False it is the fp32 gold.
"""
import math

import torch

__all__ = ["DiTConfig", "dit_forward", "block_forward", "rope_table", "temporal_positions",
           "rope_apply", "sinusoidal_embedding", "make_dit_params", "make_block_params", "attention_ref"]


class DiTConfig:
    """Subset of the reference constructor arguments (:578-604) that the T2V path uses."""

    def __init__(self, dim=1536, ffn_dim=8960, num_heads=12, num_layers=30, in_dim=16, out_dim=16,
                 freq_dim=256, text_dim=4096, text_len=512, patch_size=(1, 2, 2), eps=1e-6):
        self.dim, self.ffn_dim, self.num_heads, self.num_layers = dim, ffn_dim, num_heads, num_layers
        self.in_dim, self.out_dim, self.freq_dim = in_dim, out_dim, freq_dim
        self.text_dim, self.text_len, self.patch_size, self.eps = text_dim, text_len, patch_size, eps

    @property
    def head_dim(self):
        return self.dim // self.num_heads

    @staticmethod
    def wan_1_3b():
        return DiTConfig(1536, 8960, 12, 30)

    @staticmethod
    def wan_14b():
        return DiTConfig(5120, 13824, 40, 40)

    def to_kwargs(self):
        return dict(model_type="t2v", patch_size=self.patch_size, text_len=self.text_len,
                    in_dim=self.in_dim, dim=self.dim, ffn_dim=self.ffn_dim, freq_dim=self.freq_dim,
                    text_dim=self.text_dim, out_dim=self.out_dim, num_heads=self.num_heads,
                    num_layers=self.num_layers, eps=self.eps)


def _rb(x, on):
    """Round through bf16 (what a bf16 tensor would hold) when emulating the CUDA path."""
    return x.to(torch.bfloat16).to(torch.float32) if on else x


class _ParamGen:
    def __init__(self, seed, dtype, bf16_exact):
        self.g = torch.Generator().manual_seed(seed)
        self.dtype, self.bf16_exact = dtype, bf16_exact
        self.p = {}

    def rnd(self, *shape, std=1.0, mean=0.0):
        t = mean + torch.randn(*shape, generator=self.g, dtype=torch.float32) * std
        if self.bf16_exact:
            t = t.to(torch.bfloat16).to(torch.float32)
        return t.to(self.dtype)

    def lin(self, name, n_out, n_in, std=None, bias_std=0.02):
        std = std if std is not None else math.sqrt(2.0 / (n_in + n_out))
        self.p[name + ".weight"] = self.rnd(n_out, n_in, std=std)
        self.p[name + ".bias"] = self.rnd(n_out, std=bias_std)

    def block(self, cfg, i):
        C, Fd = cfg.dim, cfg.ffn_dim
        b = f"blocks.{i}."
        self.p[b + "modulation"] = self.rnd(1, 6, C, std=1.0 / math.sqrt(C))
        for attn in ("self_attn", "cross_attn"):
            for nm in ("q", "k", "v", "o"):
                self.lin(b + f"{attn}.{nm}", C, C)
            self.p[b + f"{attn}.norm_q.weight"] = self.rnd(C, std=0.05, mean=1.0)
            self.p[b + f"{attn}.norm_k.weight"] = self.rnd(C, std=0.05, mean=1.0)
        self.p[b + "norm3.weight"] = self.rnd(C, std=0.05, mean=1.0)
        self.p[b + "norm3.bias"] = self.rnd(C, std=0.02)
        self.lin(b + "ffn.0", Fd, C)
        self.lin(b + "ffn.2", C, Fd)


def make_block_params(cfg, i=0, seed=0, dtype=torch.float32, bf16_exact=True):
    """Parameters of ONE block only (bench.py cpu_baseline: a 14B block without the embeddings)."""
    pg = _ParamGen(seed, dtype, bf16_exact)
    pg.block(cfg, i)
    return pg.p


def make_dit_params(cfg, seed=0, dtype=torch.float32, bf16_exact=True):
    """Deterministic random parameters keyed by the reference's state-dict names.

    Scales follow the reference's init_weights (:1133-1155) and block modulation init (:462), except
    head.head.weight, which the reference zero-initialises (the output would be identically zero).
    With bf16_exact the values are bf16-representable so a bf16 model and the fp32 oracle share
    exact weights.
    """
    pg = _ParamGen(seed, dtype, bf16_exact)
    C = cfg.dim
    kin = cfg.in_dim * math.prod(cfg.patch_size)
    pg.p["patch_embedding.weight"] = pg.rnd(C, cfg.in_dim, *cfg.patch_size, std=math.sqrt(2.0 / (kin + C)))
    pg.p["patch_embedding.bias"] = pg.rnd(C, std=0.02)
    pg.lin("text_embedding.0", C, cfg.text_dim, std=0.02)
    pg.lin("text_embedding.2", C, C, std=0.02)
    pg.lin("time_embedding.0", C, cfg.freq_dim, std=0.02)
    pg.lin("time_embedding.2", C, C, std=0.02)
    pg.lin("time_projection.1", 6 * C, C)
    for i in range(cfg.num_layers):
        pg.block(cfg, i)
    pg.p["head.modulation"] = pg.rnd(1, 2, C, std=1.0 / math.sqrt(C))
    pg.lin("head.head", cfg.out_dim * math.prod(cfg.patch_size), C, std=0.02)
    return pg.p


# --------------------------------------------------------------------------------------
# embeddings
# --------------------------------------------------------------------------------------
def sinusoidal_embedding(dim, position):
    """:31-41 — float64 cos|sin table of the timestep."""
    half = dim // 2
    position = position.to(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half, dtype=torch.float64) / half))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def rope_angles(max_len, dim, theta=10000.0):
    """:44-52 — angle table [max_len, dim/2] in float64 (the reference stores exp(i*angle))."""
    return torch.outer(torch.arange(max_len, dtype=torch.float64),
                       1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64) / dim))


def rope_table(head_dim):
    """:692-699 — concatenated (temporal | height | width) angle table [1024, head_dim/2]."""
    d = head_dim
    return torch.cat([rope_angles(1024, d - 4 * (d // 6)), rope_angles(1024, 2 * (d // 6)),
                      rope_angles(1024, 2 * (d // 6))], dim=1)


def temporal_positions(f, frame_split=None, ground=None):
    """:153-191 — temporal RoPE position of each latent frame.

    plain: 0..f-1; paired (frame_split only): src 0..fs-1, tgt 0..ft-1;
    chain-of-frames (frame_split + ground): src 1..fs, ground frames all 0, tgt 1..ft.
    """
    if frame_split is None:
        return list(range(f))
    fs = frame_split
    if ground is not None:
        fg = ground[1] - ground[0]
        ft = f - fs - fg
        return list(range(1, fs + 1)) + [0] * fg + list(range(1, ft + 1))
    return list(range(fs)) + list(range(f - fs))


def rope_apply(x, grid, angles, tpos):
    """:135-205 — x [L, n, d]; rotate interleaved pairs of the first f*h*w tokens in float64."""
    f, h, w = grid
    L, n, d = x.shape
    c = d // 2
    n_t, n_hw = c - 2 * (c // 3), c // 3
    seq = f * h * w
    at = angles[:, :n_t][torch.tensor(tpos, dtype=torch.long)]          # [f, n_t]
    ah = angles[:h, n_t:n_t + n_hw]
    aw = angles[:w, n_t + n_hw:]
    ang = torch.cat([at.view(f, 1, 1, -1).expand(f, h, w, -1), ah.view(1, h, 1, -1).expand(f, h, w, -1),
                     aw.view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(seq, 1, c)
    xs = x[:seq].to(torch.float64).reshape(seq, n, c, 2)
    re = xs[..., 0] * ang.cos() - xs[..., 1] * ang.sin()
    im = xs[..., 0] * ang.sin() + xs[..., 1] * ang.cos()
    out = torch.stack([re, im], dim=-1).reshape(seq, n, d)
    return torch.cat([out.to(x.dtype), x[seq:]], dim=0)


# --------------------------------------------------------------------------------------
# block pieces
# --------------------------------------------------------------------------------------
def _linear(x, p, name):
    return torch.nn.functional.linear(x, p[name + ".weight"].float(), p[name + ".bias"].float())


def _layer_norm(x, eps, w=None, b=None):
    """:233-243 — LayerNorm over the channel dim, fp32."""
    return torch.nn.functional.layer_norm(x.float(), (x.shape[-1],), w, b, eps)


def _rms_norm(x, w, eps, emu):
    """:214-230 — RMS over the FULL channel dim (all heads); the CUDA path rounds the rsqrt factor
    and both products to bf16."""
    rs = torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    if emu:
        return _rb(_rb(x * _rb(rs, True), True) * w.float(), True)
    return x * rs * w.float()


def attention_ref(q, k, v, kv_len=None):
    """attention_utils.py:43-149 — non-causal softmax(q k^T / sqrt(d)) v; q [Lq,n,d], k,v [Lk,n,d];
    keys beyond kv_len are dropped (the flash-attn varlen path, :98-99)."""
    if kv_len is not None:
        k, v = k[:kv_len], v[:kv_len]
    d = q.shape[-1]
    s = torch.einsum("qnd,knd->nqk", q.float(), k.float()) / math.sqrt(d)
    return torch.einsum("nqk,knd->qnd", torch.softmax(s, dim=-1), v.float())


def self_attention(p, pre, x, cfg, grid, angles, tpos, kv_len, emu):
    """:271-305 — q,k,v Linear; RMSNorm(q),(k); RoPE; attention; o Linear."""
    L = x.shape[0]
    n, d = cfg.num_heads, cfg.head_dim
    q = _rms_norm(_rb(_linear(x, p, pre + "q"), emu), p[pre + "norm_q.weight"], cfg.eps, emu)
    k = _rms_norm(_rb(_linear(x, p, pre + "k"), emu), p[pre + "norm_k.weight"], cfg.eps, emu)
    v = _rb(_linear(x, p, pre + "v"), emu)
    q = _rb(rope_apply(q.view(L, n, d), grid, angles, tpos), emu)
    k = _rb(rope_apply(k.view(L, n, d), grid, angles, tpos), emu)
    o = _rb(attention_ref(q, k, v.view(L, n, d), kv_len), emu).reshape(L, n * d)
    return _rb(_linear(o, p, pre + "o"), emu)


def cross_attention(p, pre, x, ctx, cfg, emu):
    """:310-336 — q from x, k/v from the 512-row text context (unmasked), no RoPE."""
    L, S = x.shape[0], ctx.shape[0]
    n, d = cfg.num_heads, cfg.head_dim
    q = _rms_norm(_rb(_linear(x, p, pre + "q"), emu), p[pre + "norm_q.weight"], cfg.eps, emu)
    k = _rms_norm(_rb(_linear(ctx, p, pre + "k"), emu), p[pre + "norm_k.weight"], cfg.eps, emu)
    v = _rb(_linear(ctx, p, pre + "v"), emu)
    o = _rb(attention_ref(q.view(L, n, d), k.view(S, n, d), v.view(S, n, d)), emu).reshape(L, n * d)
    return _rb(_linear(o, p, pre + "o"), emu)


def block_forward(p, i, x, e0, ctx, cfg, grid, angles, tpos, kv_len, emu=False):
    """:464-515 — one WanAttentionBlock.  x [L,C] fp32, e0 [6,C] fp32, ctx [512,C]."""
    b = f"blocks.{i}."
    e = (p[b + "modulation"].float()[0] + e0)                      # [6, C]  (:491)
    t = _rb(_layer_norm(x, cfg.eps) * (1 + e[1]) + e[0], emu)       # (:495-496)
    y = self_attention(p, b + "self_attn.", t, cfg, grid, angles, tpos, kv_len, emu)
    x = x + y * e[2]                                                # (:499)
    t = _rb(_layer_norm(x, cfg.eps, p[b + "norm3.weight"].float(), p[b + "norm3.bias"].float()), emu)
    x = x + cross_attention(p, b + "cross_attn.", t, ctx, cfg, emu)  # (:504)
    t = _rb(_layer_norm(x, cfg.eps) * (1 + e[4]) + e[3], emu)       # (:507-508)
    h = _rb(_linear(t, p, b + "ffn.0"), emu)
    h = _rb(torch.nn.functional.gelu(h, approximate="tanh"), emu)   # (:458)
    y = _rb(_linear(h, p, b + "ffn.2"), emu)
    return x + y * e[5]                                             # (:511)


# --------------------------------------------------------------------------------------
# full forward
# --------------------------------------------------------------------------------------
def dit_forward(p, cfg, x, t, context, seq_len, frame_split_indices=None, ground_frame_indices=None,
                emulate_bf16=False, num_layers=None, return_tokens=False):
    """:818-1105 — x [B,16,f,h,w] (or list of [16,f,h,w]); t [B]; context: list of [len_i, text_dim].

    Returns [B,16,f,h,w] fp32.  `num_layers` truncates the block stack (bench/cpu_baseline use).
    """
    emu = emulate_bf16
    C = cfg.dim
    pt, ph, pw = cfg.patch_size
    angles = rope_table(cfg.head_dim)
    outs = []
    nl = cfg.num_layers if num_layers is None else num_layers
    for bi in range(len(x)):
        u = x[bi].float()
        cin, F_, H_, W_ = u.shape
        f, h, w = F_ // pt, H_ // ph, W_ // pw
        # patch embedding = Conv3d(k = stride = patch) (:662, :870) -> tokens (f,h,w)-major (:879)
        a = u.view(cin, f, pt, h, ph, w, pw).permute(1, 3, 5, 0, 2, 4, 6).reshape(f * h * w, -1)
        tok = _rb(a @ p["patch_embedding.weight"].float().reshape(C, -1).t()
                  + p["patch_embedding.bias"].float(), emu)
        L = tok.shape[0]
        assert L <= seq_len
        xs = torch.cat([tok, tok.new_zeros(seq_len - L, C)], dim=0)      # (:907-910)
        # time embeddings in fp32 (:913-929)
        se = sinusoidal_embedding(cfg.freq_dim, t[bi:bi + 1]).float()
        e = _linear(torch.nn.functional.silu(_linear(se, p, "time_embedding.0")), p, "time_embedding.2")
        e0 = _linear(torch.nn.functional.silu(e), p, "time_projection.1").view(6, C)
        # text context, zero padded to text_len, NOT masked (:936-942)
        c = context[bi].float()
        c = torch.cat([c, c.new_zeros(cfg.text_len - c.shape[0], c.shape[1])], dim=0)
        c = _rb(_linear(c, p, "text_embedding.0"), emu)
        c = _rb(torch.nn.functional.gelu(c, approximate="tanh"), emu)
        c = _rb(_linear(c, p, "text_embedding.2"), emu)
        fs = frame_split_indices[bi] if frame_split_indices is not None else None
        gr = ground_frame_indices[bi] if (ground_frame_indices is not None and fs is not None) else None
        tpos = temporal_positions(f, fs, gr)
        for i in range(nl):
            xs = block_forward(p, i, xs, e0, c, cfg, (f, h, w), angles, tpos, L, emu)
        if return_tokens:
            outs.append(xs)
            continue
        # head (:535-548): modulation uses e (not e0)
        eh = p["head.modulation"].float()[0] + e                       # [2, C]
        y = _rb(_layer_norm(xs, cfg.eps) * (1 + eh[1]) + eh[0], emu)
        y = _rb(_linear(y, p, "head.head"), emu)
        # unpatchify (:1108-1131)
        y = y[:L].view(f, h, w, pt, ph, pw, cfg.out_dim)
        y = torch.einsum("fhwpqrc->cfphqwr", y).reshape(cfg.out_dim, f * pt, h * ph, w * pw)
        outs.append(y)
    return torch.stack(outs)
