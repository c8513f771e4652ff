"""B200-native Wan-2.1 DiT with the reference's class API (videox_fun.models.WanTransformer3DModel).

Host-side mirror of videox_fun/models/wan_transformer3d.py (reference): the same constructor
arguments, sub-module tree and state-dict keys so `merge_lora`, checkpoint
loading and the pipeline keep working, but `forward` never calls an ATen compute kernel for
the block stack: every op is a libvcof launch (tcgen05 GEMM with fused epilogues, tcgen05
flash attention, fused LN+modulate, fused RMSNorm+RoPE …).  PyTorch provides device memory,
the stream and a few byte-sized glue ops (modulation add, sinusoid of the timestep).

There is no fallback: CPU tensors, non-bf16 weights or a missing libvcof raise.
"""
import glob
import json
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import VcofError

__all__ = ["WanTransformer3DModel", "WanAttentionBlock", "WanSelfAttention", "WanT2VCrossAttention",
           "WanRMSNorm", "WanLayerNorm", "Head", "TeaCache", "rope_params", "sinusoidal_embedding_1d"]


# ----------------------------------------------------------------------------------------------
# parameter containers (names/shape = reference; compute happens in the fused forward below)
# ----------------------------------------------------------------------------------------------
class WanRMSNorm(nn.Module):
    """reference :214-230 — RMS norm over the full channel dim; weight [dim]."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))


class WanLayerNorm(nn.LayerNorm):
    """reference :233-243."""

    def __init__(self, dim, eps=1e-6, elementwise_affine=False):
        super().__init__(dim, elementwise_affine=elementwise_affine, eps=eps)


class WanSelfAttention(nn.Module):
    """reference :246-305 (parameters: q, k, v, o, norm_q, norm_k)."""

    def __init__(self, dim, num_heads, window_size=(-1, -1), qk_norm=True, eps=1e-6):
        assert dim % num_heads == 0
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.window_size, self.qk_norm, self.eps = window_size, qk_norm, eps
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        self.o = nn.Linear(dim, dim)
        self.norm_q = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()
        self.norm_k = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()


class WanT2VCrossAttention(WanSelfAttention):
    """reference :308-336."""


class _Workspace:
    """Per-(device, L, C, F) activation buffers reused by every block of a forward."""

    def __init__(self):
        self.key = None

    def get(self, device, L, C, Fd):
        key = (str(device), L, C, Fd)
        if self.key != key:
            bf = dict(dtype=torch.bfloat16, device=device)
            self.a = torch.empty((L, C), **bf)    # LN output / attention output
            self.q = torch.empty((L, C), **bf)
            self.k = torch.empty((L, C), **bf)
            self.v = torch.empty((L, C), **bf)
            self.h = torch.empty((L, Fd), **bf)   # FFN hidden
            self.key = key
        return self


def _f32_cached(mod, name, tensor):
    """fp32 copy of a small parameter vector, rebuilt when the parameter is mutated in place."""
    cache = mod.__dict__.setdefault("_vcof_f32", {})
    key = (tensor.data_ptr(), tensor._version, tensor.device)
    hit = cache.get(name)
    if hit is None or hit[0] != key:
        cache[name] = (key, tensor.detach().to(torch.float32).contiguous())
    return cache[name][1]


class WanAttentionBlock(nn.Module):
    """reference :424-515 — AdaLN-Zero block: self-attn, text cross-attn, GELU MLP."""

    def __init__(self, cross_attn_type, dim, ffn_dim, num_heads, window_size=(-1, -1), qk_norm=True,
                 cross_attn_norm=False, eps=1e-6):
        super().__init__()
        if cross_attn_type != "t2v_cross_attn":
            raise NotImplementedError("videocof_b200 implements the T2V cross-attention only "
                                      "(the VideoCoF path); got %r" % (cross_attn_type,))
        self.dim, self.ffn_dim, self.num_heads = dim, ffn_dim, num_heads
        self.window_size, self.qk_norm, self.cross_attn_norm, self.eps = window_size, qk_norm, cross_attn_norm, eps
        self.norm1 = WanLayerNorm(dim, eps)
        self.self_attn = WanSelfAttention(dim, num_heads, window_size, qk_norm, eps)
        self.norm3 = WanLayerNorm(dim, eps, elementwise_affine=True) if cross_attn_norm else nn.Identity()
        self.cross_attn = WanT2VCrossAttention(dim, num_heads, (-1, -1), qk_norm, eps)
        self.norm2 = WanLayerNorm(dim, eps)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)

    # -- fused single-sample path -------------------------------------------------------------
    def context_kv(self, ctx):
        """Cross-attention keys and values of the embedded text, ctx bf16 [S,C] (:310-336: k = norm_k(k(context)),
        v = v(context)).  They depend on the prompt and the weights only, not on the timestep or the latents."""
        ca = self.cross_attn
        kc = ops.gemm(ctx, ca.k.weight, ca.k.bias, "bias")
        ops.rmsnorm_rope_(kc, ca.norm_k.weight, ca.eps, self.dim // self.num_heads, None)
        return kc, ops.gemm(ctx, ca.v.weight, ca.v.bias, "bias")

    def run(self, x, e, ctx, rope, kv_len, ws, sp=None, ckv=None):
        """x fp32 [L,C] (updated in place and returned); e fp32 [6,C] = modulation + e0;
        ctx bf16 [S,C]; rope: ops.RopeSpec; sp: optional sequence-parallel context (dist.py);
        ckv: this block's (K, V) of ctx when the model's context cache holds them (else computed here)."""
        sa, ca = self.self_attn, self.cross_attn
        n, hd = self.num_heads, self.dim // self.num_heads
        # self-attention (:495-499)
        ops.ln_modulate(x, None, None, e[0], e[1], self.eps, out=ws.a)
        if sp is None:
            ops.gemm(ws.a, sa.q.weight, sa.q.bias, "bias", out=ws.q)
            ops.gemm(ws.a, sa.k.weight, sa.k.bias, "bias", out=ws.k)
            ops.gemm(ws.a, sa.v.weight, sa.v.bias, "bias", out=ws.v)
            ops.rmsnorm_rope_(ws.q, sa.norm_q.weight, sa.eps, hd, rope)
            ops.rmsnorm_rope_(ws.k, sa.norm_k.weight, sa.eps, hd, rope)
            ops.attention(ws.q, ws.k, ws.v, n, kv_len=kv_len, out=ws.a)
        elif sp.use_push(n):
            # push exchange (dist.py): the producing kernels store Q, K, V straight into the other ranks' receive
            # buffers over NVLink peer memory; no collective call, two stream-ordered barriers per layer
            pb = sp.push_buffers(x.shape[0], self.dim, ws.a)
            ops.gemm(ws.a, sa.v.weight, sa.v.bias, "bias", out=ws.v)
            ops.copy_scatter(ws.v, pb.dst["v"])
            ops.gemm(ws.a, sa.q.weight, sa.q.bias, "bias", out=ws.q)
            ops.rmsnorm_rope_scatter(ws.q, sa.norm_q.weight, sa.eps, hd, rope, pb.dst["q"])
            ops.gemm(ws.a, sa.k.weight, sa.k.bias, "bias", out=ws.k)
            ops.rmsnorm_rope_scatter(ws.k, sa.norm_k.weight, sa.eps, hd, rope, pb.dst["k"])
            sp.attention_pushed(n, ws.a, ops.attention_scatter)
        elif sp.can_exchange_heads(n):
            # head exchange (dist.py): every projection lands in its all-to-all send layout — V through a pack
            # kernel, Q and K straight from the norm/RoPE kernel — and its exchange overlaps the next projection
            ops.gemm(ws.a, sa.v.weight, sa.v.bias, "bias", out=ws.v)
            sp.start_exchange("v", sp.pack(ws.v, sp.send_buffer("v", ws.v)))
            ops.gemm(ws.a, sa.q.weight, sa.q.bias, "bias", out=ws.q)
            sp.start_exchange("q", ops.rmsnorm_rope_(ws.q, sa.norm_q.weight, sa.eps, hd, rope,
                                                     out_blocked=sp.send_buffer("q", ws.q)))
            ops.gemm(ws.a, sa.k.weight, sa.k.bias, "bias", out=ws.k)
            sp.start_exchange("k", ops.rmsnorm_rope_(ws.k, sa.norm_k.weight, sa.eps, hd, rope,
                                                     out_blocked=sp.send_buffer("k", ws.k)))
            sp.attention_exchanged(n, out=ws.a)
        else:
            # K first so that its NVLink all-gather overlaps the V and Q projections, then V overlaps Q
            ops.gemm(ws.a, sa.k.weight, sa.k.bias, "bias", out=ws.k)
            ops.rmsnorm_rope_(ws.k, sa.norm_k.weight, sa.eps, hd, rope)
            sp.start_gather("k", ws.k)
            ops.gemm(ws.a, sa.v.weight, sa.v.bias, "bias", out=ws.v)
            sp.start_gather("v", ws.v)
            ops.gemm(ws.a, sa.q.weight, sa.q.bias, "bias", out=ws.q)
            ops.rmsnorm_rope_(ws.q, sa.norm_q.weight, sa.eps, hd, rope)
            sp.attention_gathered(ws.q, n, out=ws.a)
        ops.gemm(ws.a, sa.o.weight, sa.o.bias, "bias_gate_res", out=x, gate=e[2])
        # cross-attention (:504)
        if self.cross_attn_norm:
            ops.ln_modulate(x, _f32_cached(self, "n3w", self.norm3.weight),
                            _f32_cached(self, "n3b", self.norm3.bias), None, None, self.eps, out=ws.a)
        else:
            raise NotImplementedError("cross_attn_norm=False is not on the VideoCoF path")
        ops.gemm(ws.a, ca.q.weight, ca.q.bias, "bias", out=ws.q)
        ops.rmsnorm_rope_(ws.q, ca.norm_q.weight, ca.eps, hd, None)
        kc, vc = ckv if ckv is not None else self.context_kv(ctx)
        ops.attention(ws.q, kc, vc, n, out=ws.a)
        ops.gemm(ws.a, ca.o.weight, ca.o.bias, "bias_gate_res", out=x, gate=None)
        # MLP (:507-511)
        ops.ln_modulate(x, None, None, e[3], e[4], self.eps, out=ws.a)
        ops.gemm(ws.a, self.ffn[0].weight, self.ffn[0].bias, "bias_gelu", out=ws.h)
        ops.gemm(ws.h, self.ffn[2].weight, self.ffn[2].bias, "bias_gate_res", out=x, gate=e[5])
        return x

    def run_batched(self, X, es, ctx, ropes, kv_lens, ws, rows, ckv=None):
        """Batch-aware block (classifier-free guidance: the uncond + cond samples of one step, pipeline_wan.py:700).
        X fp32 [B*rows, C]: the samples' residual streams stacked along M (updated in place); es[b] fp32 [6, C];
        ctx bf16 [B*S, C] the samples' text contexts stacked; ropes / kv_lens per sample.

        Every Linear whose epilogue is sample-independent runs as ONE GEMM over all B*rows tokens — q, k, v, the
        cross-attention q / k / v / o and the first FFN Linear: the weights stream once per step instead of once per
        sample — while what depends on the sample runs per row range: AdaLN modulation, RoPE (positions restart),
        attention (keys of one sample), and the two gated residual GEMMs (per-sample gate vector).  Row-wise the
        arithmetic is that of `run`, so the result is bit-identical to the per-sample loop.
        ckv: per sample, this block's cached (K, V) of that sample's context (see `run`)."""
        sa, ca = self.self_attn, self.cross_attn
        n, hd = self.num_heads, self.dim // self.num_heads
        B = len(es)
        rng = [slice(b * rows, (b + 1) * rows) for b in range(B)]
        for b in range(B):
            ops.ln_modulate(X[rng[b]], None, None, es[b][0], es[b][1], self.eps, out=ws.a[rng[b]])
        ops.gemm(ws.a, sa.q.weight, sa.q.bias, "bias", out=ws.q)
        ops.gemm(ws.a, sa.k.weight, sa.k.bias, "bias", out=ws.k)
        ops.gemm(ws.a, sa.v.weight, sa.v.bias, "bias", out=ws.v)
        for b in range(B):
            ops.rmsnorm_rope_(ws.q[rng[b]], sa.norm_q.weight, sa.eps, hd, ropes[b])
            ops.rmsnorm_rope_(ws.k[rng[b]], sa.norm_k.weight, sa.eps, hd, ropes[b])
            ops.attention(ws.q[rng[b]], ws.k[rng[b]], ws.v[rng[b]], n, kv_len=kv_lens[b], out=ws.a[rng[b]])
        for b in range(B):
            ops.gemm(ws.a[rng[b]], sa.o.weight, sa.o.bias, "bias_gate_res", out=X[rng[b]], gate=es[b][2])
        if not self.cross_attn_norm:
            raise NotImplementedError("cross_attn_norm=False is not on the VideoCoF path")
        ops.ln_modulate(X, _f32_cached(self, "n3w", self.norm3.weight), _f32_cached(self, "n3b", self.norm3.bias),
                        None, None, self.eps, out=ws.a)
        ops.gemm(ws.a, ca.q.weight, ca.q.bias, "bias", out=ws.q)
        ops.rmsnorm_rope_(ws.q, ca.norm_q.weight, ca.eps, hd, None)
        if ckv is None:
            S = ctx.shape[0] // B
            kc, vc = self.context_kv(ctx)
            ckv = [(kc[b * S:(b + 1) * S], vc[b * S:(b + 1) * S]) for b in range(B)]
        for b in range(B):
            ops.attention(ws.q[rng[b]], ckv[b][0], ckv[b][1], n, out=ws.a[rng[b]])
        ops.gemm(ws.a, ca.o.weight, ca.o.bias, "bias_gate_res", out=X, gate=None)
        for b in range(B):
            ops.ln_modulate(X[rng[b]], None, None, es[b][3], es[b][4], self.eps, out=ws.a[rng[b]])
        ops.gemm(ws.a, self.ffn[0].weight, self.ffn[0].bias, "bias_gelu", out=ws.h)
        for b in range(B):
            ops.gemm(ws.h[rng[b]], self.ffn[2].weight, self.ffn[2].bias, "bias_gate_res", out=X[rng[b]], gate=es[b][5])
        return X

    def forward(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens=None, dtype=torch.bfloat16,
                t=0, frame_split_indices=None, ground_frame_indices=None):
        """Reference signature (:464-477): x [B,L,C], e [B,6,C] fp32, context [B,S,C]."""
        if context_lens is not None:
            raise NotImplementedError("context_lens masking is never used by the reference T2V path")
        outs = []
        ws = _Workspace().get(x.device, x.shape[1], self.dim, self.ffn_dim)
        for b in range(x.shape[0]):
            f, h, w = [int(v) for v in grid_sizes[b].tolist()]
            fs = frame_split_indices[b] if frame_split_indices is not None and b < len(frame_split_indices) else None
            gr = ground_frame_indices[b] if (fs is not None and ground_frame_indices is not None
                                             and b < len(ground_frame_indices)) else None
            rope = make_rope_spec(freqs, x.device, f, h, w, fs, gr)
            xb = x[b].to(torch.float32).contiguous().clone()
            eb = (self.modulation.to(torch.float32)[0] + e[b].to(torch.float32)).contiguous()
            outs.append(self.run(xb, eb, context[b].to(torch.bfloat16).contiguous(), rope,
                                 int(seq_lens[b]), ws))
        return torch.stack(outs)


class Head(nn.Module):
    """reference :518-548."""

    def __init__(self, dim, out_dim, patch_size, eps=1e-6):
        super().__init__()
        self.dim, self.out_dim, self.patch_size, self.eps = dim, out_dim, patch_size, eps
        self.norm = WanLayerNorm(dim, eps)
        self.head = nn.Linear(dim, math.prod(patch_size) * out_dim)
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim, position):
    """reference :31-41 (float64)."""
    half = dim // 2
    position = position.type(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def rope_params(max_seq_len, dim, theta=10000):
    """reference :44-52 — complex128 exp(i * pos * theta^(-2k/dim))."""
    freqs = torch.outer(torch.arange(max_seq_len),
                        1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


def temporal_positions(f, frame_split=None, ground=None):
    """Temporal RoPE position per latent frame for the three reference modes (:153-191)."""
    if frame_split is None:
        return list(range(f))
    if ground is not None:
        fg = ground[1] - ground[0]
        return list(range(1, frame_split + 1)) + [0] * fg + list(range(1, f - frame_split - fg + 1))
    return list(range(frame_split)) + list(range(f - frame_split))


_rope_cache = {}      # id(freqs) -> (freqs, _version, device, table); the entry HOLDS the tensor so its id stays unique


def make_rope_spec(freqs, device, f, h, w, frame_split=None, ground=None, row_offset=0):
    """Device tables for vcof_rmsnorm_rope from the model's complex128 `freqs` [1024, d/2].  The cos/sin table is
    cached per `freqs` tensor object (enable_riflex / disable_riflex install a new tensor, an in-place edit bumps
    `_version`); a few entries are kept so that several models do not evict each other."""
    hit = _rope_cache.get(id(freqs))
    if hit is not None and hit[0] is freqs and hit[1] == freqs._version and hit[2] == str(device):
        table = hit[3]
    else:
        fr = freqs.detach().to("cpu")
        table = torch.stack([fr.real, fr.imag], dim=-1).to(torch.float32).contiguous().to(device)
        if len(_rope_cache) >= 4:
            _rope_cache.pop(next(iter(_rope_cache)))
        _rope_cache[id(freqs)] = (freqs, freqs._version, str(device), table)
    c = freqs.shape[1]
    n_t, n_h = c - 2 * (c // 3), c // 3
    tpos = torch.tensor(temporal_positions(f, frame_split, ground), dtype=torch.int32, device=device)
    return ops.RopeSpec(table, tpos, f, h, w, n_t, n_h, row_offset)


def _reference_init(name, shape, dim):
    """Value the reference constructor + init_weights (:462, :1133-1155) leave in a parameter the checkpoint does not
    hold (the reference loads with strict=False, so such parameters keep their initial values, :1282)."""
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "modulation":
        return torch.randn(shape) / dim ** 0.5
    if leaf == "bias":
        return torch.zeros(shape)
    if len(shape) == 1:                                   # RMSNorm / LayerNorm gains
        return torch.ones(shape)
    w = torch.empty(shape)
    if name == "head.head.weight":
        return w.zero_()
    if name.startswith(("text_embedding", "time_embedding")):
        return nn.init.normal_(w, std=0.02)
    if name == "patch_embedding.weight":
        nn.init.xavier_uniform_(w.flatten(1))
        return w
    return nn.init.xavier_uniform_(w)


class _Config(dict):
    """Attribute-style config, enough of diffusers' FrozenDict for the pipeline (:634, :689)."""
    __getattr__ = dict.get


class TeaCache:
    """reference videox_fun/models/cache_utils.py:21-76 (timestep-embedding aware step skipping)."""

    def __init__(self, coefficients, num_steps, rel_l1_thresh=0.0, num_skip_start_steps=0, offload=True):
        if num_steps < 1:
            raise ValueError(f"`num_steps` must be greater than 0 but is {num_steps}.")
        if rel_l1_thresh < 0:
            raise ValueError(f"`rel_l1_thresh` must be greater than or equal to 0 but is {rel_l1_thresh}.")
        if num_skip_start_steps < 0 or num_skip_start_steps > num_steps:
            raise ValueError("`num_skip_start_steps` must be in [0, num_steps]")
        self.coefficients, self.num_steps = coefficients, num_steps
        self.rel_l1_thresh, self.num_skip_start_steps, self.offload = rel_l1_thresh, num_skip_start_steps, offload
        self.rescale_func = np.poly1d(self.coefficients)
        self.reset()

    @staticmethod
    def compute_rel_l1_distance(prev, cur):
        return ((cur - prev).abs().mean() / prev.abs().mean()).cpu().item()

    def reset(self):
        self.cnt = 0
        self.should_calc = True
        self.accumulated_rel_l1_distance = 0
        self.previous_modulated_input = None
        self.previous_residual = None
        self.previous_residual_cond = None
        self.previous_residual_uncond = None


# ----------------------------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------------------------
class WanTransformer3DModel(nn.Module):
    """Drop-in for the reference class (:567-1105).  Same __init__ kwargs and state-dict keys."""

    _supports_gradient_checkpointing = False

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048, ffn_dim=8192,
                 freq_dim=256, text_dim=4096, out_dim=16, num_heads=16, num_layers=32, window_size=(-1, -1),
                 qk_norm=True, cross_attn_norm=True, eps=1e-6, in_channels=16, hidden_size=2048,
                 add_control_adapter=False, in_dim_control_adapter=24, downscale_factor_control_adapter=8,
                 add_ref_conv=False, in_dim_ref_conv=16, cross_attn_type=None):
        super().__init__()
        if add_control_adapter or add_ref_conv or model_type != "t2v":
            raise NotImplementedError("videocof_b200 covers the Wan-2.1 T2V DiT used by VideoCoF "
                                      "(no control adapter / ref conv / i2v)")
        patch_size = tuple(patch_size)
        self.config = _Config(model_type=model_type, patch_size=patch_size, text_len=text_len, in_dim=in_dim,
                              dim=dim, ffn_dim=ffn_dim, freq_dim=freq_dim, text_dim=text_dim, out_dim=out_dim,
                              num_heads=num_heads, num_layers=num_layers, window_size=tuple(window_size),
                              qk_norm=qk_norm, cross_attn_norm=cross_attn_norm, eps=eps, in_channels=in_channels,
                              hidden_size=hidden_size, add_control_adapter=False, add_ref_conv=False,
                              cross_attn_type=cross_attn_type)
        self.model_type, self.patch_size, self.text_len = model_type, patch_size, text_len
        self.in_dim, self.dim, self.ffn_dim, self.freq_dim = in_dim, dim, ffn_dim, freq_dim
        self.text_dim, self.out_dim, self.num_heads, self.num_layers = text_dim, out_dim, num_heads, num_layers
        self.window_size, self.qk_norm, self.cross_attn_norm, self.eps = window_size, qk_norm, cross_attn_norm, eps
        if patch_size != (1, 2, 2):
            raise NotImplementedError("patch_size must be (1,2,2)")
        if dim // num_heads != 128:
            raise NotImplementedError("libvcof attention is specialised for head_dim 128 (Wan 1.3B / 14B)")

        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=patch_size, stride=patch_size)
        self.text_embedding = nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate="tanh"),
                                            nn.Linear(dim, dim))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
        cross_attn_type = cross_attn_type or "t2v_cross_attn"
        self.blocks = nn.ModuleList([
            WanAttentionBlock(cross_attn_type, dim, ffn_dim, num_heads, window_size, qk_norm, cross_attn_norm, eps)
            for _ in range(num_layers)])
        for i, blk in enumerate(self.blocks):
            blk.self_attn.layer_idx, blk.self_attn.num_layers = i, num_layers
        self.head = Head(dim, out_dim, patch_size, eps)

        d = dim // num_heads
        self.d = d
        # plain attribute (not a buffer) exactly as the reference (:690-699)
        self.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                                rope_params(1024, 2 * (d // 6))], dim=1)
        self.control_adapter = None
        self.ref_conv = None
        self.teacache = None
        self.cfg_skip_ratio = None
        self.current_steps = 0
        self.num_inference_steps = None
        self.gradient_checkpointing = False
        self.sp_world_size = 1
        self.sp_world_rank = 0
        self._sp = None
        self._ws = _Workspace()
        self._ctx_cache = None
        self.init_weights()

    # ---- diffusers-ModelMixin surface the callers touch -----------------------------------------
    @property
    def dtype(self):
        return self.patch_embedding.weight.dtype

    @property
    def device(self):
        return self.patch_embedding.weight.device

    def init_weights(self):
        """reference :1133-1155."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.xavier_uniform_(self.patch_embedding.weight.flatten(1))
        for m in self.text_embedding.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=.02)
        for m in self.time_embedding.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=.02)
        nn.init.zeros_(self.head.head.weight)

    # ---- TeaCache / cfg-skip / RIFLEx switches (:731-800) ----------------------------------------
    def enable_teacache(self, coefficients, num_steps, rel_l1_thresh, num_skip_start_steps=0, offload=True):
        self.teacache = TeaCache(coefficients, num_steps, rel_l1_thresh=rel_l1_thresh,
                                 num_skip_start_steps=num_skip_start_steps, offload=offload)

    def share_teacache(self, transformer=None):
        self.teacache = transformer.teacache

    def disable_teacache(self):
        self.teacache = None

    def enable_cfg_skip(self, cfg_skip_ratio, num_steps):
        if cfg_skip_ratio != 0:
            self.cfg_skip_ratio, self.current_steps, self.num_inference_steps = cfg_skip_ratio, 0, num_steps
        else:
            self.disable_cfg_skip()

    def share_cfg_skip(self, transformer=None):
        self.cfg_skip_ratio = transformer.cfg_skip_ratio
        self.current_steps = transformer.current_steps
        self.num_inference_steps = transformer.num_inference_steps

    def disable_cfg_skip(self):
        self.cfg_skip_ratio, self.current_steps, self.num_inference_steps = None, 0, None

    # ---- step-invariant context work ("K3 / K10") ---------------------------------
    def enable_context_cache(self, max_entries=4):
        """Keep, per distinct prompt embedding, the embedded text (:936-942) and every block's cross-attention K / V
        of it (:310-336) across forwards: they depend on the prompt and the weights only, and the reference recomputes
        all 2 + 3 x layers launches at every timestep.  Opt-in and scoped: `WanPipeline.__call__` turns it on around
        its denoising loop — where the embeddings are fixed before the first step (pipeline_wan.py:606) — and off
        (freed) after it; a bare `forward` recomputes, like the reference.  The kernels and their inputs are the same
        ones, so a cached forward is bit-identical to an uncached one.

        An entry is keyed by the embedding tensor's storage address, shape and version counter and HOLDS the tensor (an
        address cannot be recycled while its entry lives), plus the address / version of every weight it was computed
        from: in-place edits (LoRA merge / unmerge, load_state_dict) and moved weights miss."""
        self._ctx_cache = {"max": int(max_entries), "entries": []}

    def disable_context_cache(self):
        self._ctx_cache = None

    def _context_weights_key(self):
        ps = [p for p in self.text_embedding.parameters()]
        for blk in self.blocks:
            ca = blk.cross_attn
            ps += [ca.k.weight, ca.k.bias, ca.v.weight, ca.v.bias, ca.norm_k.weight]
        return tuple((p.data_ptr(), p._version) for p in ps)

    def _contexts(self, context, n):
        """Per sample: (embedded text bf16 [text_len, C], per-block (K, V) list or None when the cache is off)."""
        cache = getattr(self, "_ctx_cache", None)
        if cache is None:
            return [(self._text_embed(context[b]), None) for b in range(n)]
        wkey = self._context_weights_key()
        out = []
        for b in range(n):
            c = context[b]
            key = (c.data_ptr(), c._version, tuple(c.shape), tuple(c.stride()), c.dtype, str(c.device), wkey)
            hit = next((e for e in cache["entries"] if e[0] == key), None)
            if hit is None:
                ctx = self._text_embed(c)
                hit = (key, c, ctx, [blk.context_kv(ctx) for blk in self.blocks])
                cache["entries"].append(hit)
                del cache["entries"][:-cache["max"]]
            out.append((hit[2], hit[3]))
        return out

    def enable_riflex(self, k=6, L_test=66, L_test_scale=4.886):
        d = self.d
        dim_t = d - 4 * (d // 6)
        fr = 1.0 / torch.pow(10000.0, torch.arange(0, dim_t, 2).to(torch.float64).div(dim_t))
        fr[k - 1] = 0.9 * 2 * torch.pi / L_test
        if L_test_scale is not None:
            fr[k - 1] = fr[k - 1] / L_test_scale
        ang = torch.outer(torch.arange(1024), fr)
        device = self.freqs.device
        self.freqs = torch.cat([torch.polar(torch.ones_like(ang), ang), rope_params(1024, 2 * (d // 6)),
                                rope_params(1024, 2 * (d // 6))], dim=1).to(device)

    def disable_riflex(self):
        d = self.d
        device = self.freqs.device
        self.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                                rope_params(1024, 2 * (d // 6))], dim=1).to(device)

    def enable_multi_gpus_inference(self, group=None):
        """Sequence-parallel forward over `group` (default: WORLD): tokens sharded contiguously,
        one K/V all-gather per self-attention layer, head output gathered (replaces the reference's
        xfuser USP path :802-816, which cannot run VideoCoF's chain-of-frames kwargs)."""
        from .dist import SequenceParallel
        self._sp = SequenceParallel(group)
        self.sp_world_size, self.sp_world_rank = self._sp.world, self._sp.rank

    # ---- forward ------------------------------------------------------------------------------------
    def _check_ready(self, x):
        if not x.is_cuda:
            raise VcofError("WanTransformer3DModel.forward needs CUDA tensors: libvcof has no CPU path")
        w = self.patch_embedding.weight
        if w.device != x.device or w.dtype != torch.bfloat16:
            raise VcofError(f"weights must be bf16 on {x.device} (got {w.dtype} on {w.device}); "
                            "call .to(device, torch.bfloat16) as the reference CLIs do")

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None, y_camera=None, full_ref=None,
                subject_ref=None, cond_flag=True, frame_split_indices=None, ground_frame_indices=None):
        """reference :818-1105.  x: [B,16,f,h,w] tensor or list of [16,f,h,w]; returns a fresh
        [B,16,f,h,w] bf16 tensor (the pipeline mutates it in place, pipeline_wan.py:736)."""
        if any(v is not None for v in (clip_fea, y, y_camera, full_ref, subject_ref)):
            raise NotImplementedError("clip_fea / y / y_camera / full_ref / subject_ref are not on the "
                                      "VideoCoF T2V path")
        # cfg_skip (reference utils/cfg_optimization.py:5-38): drop the uncond half late in sampling
        bs = len(x)
        skip = (bs >= 2 and self.cfg_skip_ratio is not None and
                self.current_steps >= self.num_inference_steps * (1 - self.cfg_skip_ratio))
        if skip:
            h = bs // 2
            x, t, context = x[h:], t[h:], context[h:]
            frame_split_indices = frame_split_indices[h:] if frame_split_indices is not None else None
            ground_frame_indices = ground_frame_indices[h:] if ground_frame_indices is not None else None
        out = self._forward(x, t, context, seq_len, cond_flag, frame_split_indices, ground_frame_indices)
        if skip:
            out = torch.cat([out, out], dim=0)
        return out

    def _time_embed(self, t_b):
        """:913-929 — fp32 throughout.  t_b: 1-element tensor."""
        se = sinusoidal_embedding_1d(self.freq_dim, t_b.reshape(1).to(self.device)).float().contiguous()
        te0, te2, tp = self.time_embedding[0], self.time_embedding[2], self.time_projection[1]
        e = ops.linear_f32(ops.linear_f32(se, te0.weight, te0.bias, act_out=True), te2.weight, te2.bias)
        e0 = ops.linear_f32(e, tp.weight, tp.bias, act_in=True)
        return e, e0.view(6, self.dim)

    def _text_embed(self, c):
        """:936-942 — zero-pad to text_len, Linear -> GELU(tanh) -> Linear; unmasked."""
        cin = torch.zeros((self.text_len, self.text_dim), dtype=torch.bfloat16, device=self.device)
        cin[:c.shape[0]] = c.to(torch.bfloat16)
        t0, t2 = self.text_embedding[0], self.text_embedding[2]
        return ops.gemm(ops.gemm(cin, t0.weight, t0.bias, "bias_gelu"), t2.weight, t2.bias, "bias")

    def _mod_stack(self):
        """[layers, 6, C] fp32 stack of the block modulation parameters (rebuilt on mutation)."""
        key = tuple((b.modulation.data_ptr(), b.modulation._version) for b in self.blocks)
        if getattr(self, "_mod_key", None) != key:
            self._mod = torch.stack([b.modulation.detach().to(torch.float32)[0] for b in self.blocks]).contiguous()
            self._mod_key = key
        return self._mod

    def _forward(self, x, t, context, seq_len, cond_flag, frame_split_indices, ground_frame_indices):
        xs = list(x) if not isinstance(x, torch.Tensor) else [x[i] for i in range(x.shape[0])]
        self._check_ready(xs[0])
        dev = xs[0].device
        if self.freqs.device != dev:
            self.freqs = self.freqs.to(dev)
        if t.dim() != 1:
            raise NotImplementedError("per-token timesteps (t.dim() != 1) are not on the VideoCoF path")
        C = self.dim
        sp = self._sp
        P = sp.world if sp is not None else 1
        if P > 1:
            seq_len = int(math.ceil(seq_len / P)) * P                    # (:904-905)
        outs = []
        tc = self.teacache
        embeds = [self._time_embed(t[b]) for b in range(len(xs))]
        ctxs = self._contexts(context, len(xs))

        # TeaCache gate (:956-1031): ONE decision per forward, from the modulated timestep embedding of the batch
        should_calc = True
        if tc is not None:
            if cond_flag:
                mod_inp = torch.stack([e0 for _, e0 in embeds])
                if tc.cnt < tc.num_skip_start_steps:
                    should_calc, tc.accumulated_rel_l1_distance = True, 0
                else:
                    d = tc.compute_rel_l1_distance(tc.previous_modulated_input, mod_inp)
                    tc.accumulated_rel_l1_distance += tc.rescale_func(d)
                    if tc.accumulated_rel_l1_distance < tc.rel_l1_thresh:
                        should_calc = False
                    else:
                        should_calc, tc.accumulated_rel_l1_distance = True, 0
                tc.previous_modulated_input = mod_inp
                tc.should_calc = should_calc
            else:
                should_calc = tc.should_calc
            self.should_calc = should_calc          # attribute the reference also exposes (:968-981)
        residuals = []

        # Batch-aware path (single GPU, B >= 2 samples of one shape — the CFG pair of inference.py): the samples'
        # tokens are stacked along M so that the block stack streams its weights once per step (run_batched).
        shapes = {tuple(u.shape) for u in xs}
        if (len(xs) >= 2 and P == 1 and len(shapes) == 1 and os.environ.get("VCOF_DIT_BATCHED", "1") != "0"
                and not (tc is not None and not should_calc)):
            return self._forward_batched(xs, t, ctxs, seq_len, cond_flag, frame_split_indices,
                                         ground_frame_indices, embeds, tc)

        for b, u in enumerate(xs):
            u = u.to(torch.bfloat16).contiguous()
            cin, F_, H_, W_ = u.shape
            f, h, w = F_, H_ // 2, W_ // 2
            L = f * h * w
            assert L <= seq_len, "seq_len shorter than the token count"      # (:906)
            # patch embedding as a K=64 GEMM; the residual stream is fp32 from here on
            a = ops.patchify(u)
            xb = torch.zeros((seq_len, C), dtype=torch.float32, device=dev) if seq_len > L else \
                torch.empty((L, C), dtype=torch.float32, device=dev)
            ops.gemm(a, self.patch_embedding.weight.view(C, -1), self.patch_embedding.bias, "bias_f32",
                     out=xb[:L])
            e, e0 = embeds[b]
            ctx, ckv = ctxs[b]
            fs = frame_split_indices[b] if frame_split_indices is not None and b < len(frame_split_indices) else None
            gr = ground_frame_indices[b] if (fs is not None and ground_frame_indices is not None
                                             and b < len(ground_frame_indices)) else None
            row0, rows = 0, seq_len
            if P > 1:                                                        # token-chunk SP (:949-953)
                rows = seq_len // P
                row0 = sp.rank * rows
                xb = xb[row0:row0 + rows].contiguous()
                sp.configure(kv_len=L, rows=rows)
            rope = make_rope_spec(self.freqs, dev, f, h, w, fs, gr, row_offset=row0)
            mod_all = self._mod_stack() + e0                                 # [layers, 6, C] (:491)

            if tc is not None and not should_calc:
                prev = tc.previous_residual_cond if cond_flag else tc.previous_residual_uncond
                xb = xb + prev[b - len(xs)].to(dev)                          # `[-x.size(0):]` of the stored batch
            else:
                ori = xb.clone() if tc is not None else None
                ws = self._ws.get(dev, rows, C, self.ffn_dim)
                for i, blk in enumerate(self.blocks):
                    blk.run(xb, mod_all[i], ctx, rope, L, ws, sp if P > 1 else None,
                            ckv[i] if ckv is not None else None)
                if tc is not None:
                    res = xb - ori
                    residuals.append(res.cpu() if tc.offload else res)

            # head (:535-548) — modulation uses e, not e0
            eh = (self.head.modulation.detach().to(torch.float32)[0] + e).contiguous()   # [2, C]
            yb = ops.ln_modulate(xb, None, None, eh[0], eh[1], self.eps)
            yo = ops.gemm(yb, self.head.head.weight, self.head.head.bias, "bias")            # [rows, 64]
            if P > 1:
                yo = sp.all_gather_rows(yo)                                                   # (:1085-1086)
            outs.append(ops.unpatchify(yo[:L], self.out_dim, f, H_, W_))                     # (:1108-1131)
        if tc is not None and residuals:
            stacked = torch.stack(residuals)
            if cond_flag:
                tc.previous_residual_cond = stacked
            else:
                tc.previous_residual_uncond = stacked
        if tc is not None and cond_flag:
            tc.cnt += 1
            if tc.cnt == tc.num_steps:
                tc.reset()
        return torch.stack(outs)

    def _forward_batched(self, xs, t, ctxs, seq_len, cond_flag, frame_split_indices, ground_frame_indices, embeds, tc):
        """`_forward` for B >= 2 same-shape samples on one GPU: tokens stacked [B*seq_len, C] (see run_batched).
        ctxs: `_contexts` of the batch."""
        dev = xs[0].device
        C, B = self.dim, len(xs)
        cin, F_, H_, W_ = xs[0].shape
        f, h, w = F_, H_ // 2, W_ // 2
        L = f * h * w
        assert L <= seq_len, "seq_len shorter than the token count"          # (:906)
        X = torch.zeros((B * seq_len, C), dtype=torch.float32, device=dev) if seq_len > L else \
            torch.empty((B * seq_len, C), dtype=torch.float32, device=dev)
        ropes, mods = [], []
        for b, u in enumerate(xs):
            a = ops.patchify(u.to(torch.bfloat16).contiguous())
            ops.gemm(a, self.patch_embedding.weight.view(C, -1), self.patch_embedding.bias, "bias_f32",
                     out=X[b * seq_len:b * seq_len + L])
            fs = frame_split_indices[b] if frame_split_indices is not None and b < len(frame_split_indices) else None
            gr = ground_frame_indices[b] if (fs is not None and ground_frame_indices is not None
                                             and b < len(ground_frame_indices)) else None
            ropes.append(make_rope_spec(self.freqs, dev, f, h, w, fs, gr))
            mods.append(self._mod_stack() + embeds[b][1])                    # [layers, 6, C] (:491)
        cached = ctxs[0][1] is not None
        ctx = None if cached else torch.cat([c for c, _ in ctxs], dim=0)
        ori = X.clone() if tc is not None else None
        ws = self._ws.get(dev, B * seq_len, C, self.ffn_dim)
        for i, blk in enumerate(self.blocks):
            blk.run_batched(X, [m[i] for m in mods], ctx, ropes, [L] * B, ws, seq_len,
                            [kv[i] for _, kv in ctxs] if cached else None)
        if tc is not None:
            res = (X - ori).view(B, seq_len, C)
            res = res.cpu() if tc.offload else res
            if cond_flag:
                tc.previous_residual_cond = res
            else:
                tc.previous_residual_uncond = res
        outs = []
        for b in range(B):
            e = embeds[b][0]
            eh = (self.head.modulation.detach().to(torch.float32)[0] + e).contiguous()   # [2, C]
            yb = ops.ln_modulate(X[b * seq_len:(b + 1) * seq_len], None, None, eh[0], eh[1], self.eps)
            yo = ops.gemm(yb, self.head.head.weight, self.head.head.bias, "bias")
            outs.append(ops.unpatchify(yo[:L], self.out_dim, f, H_, W_))
        if tc is not None and cond_flag:
            tc.cnt += 1
            if tc.cnt == tc.num_steps:
                tc.reset()
        return torch.stack(outs)

    # ---- synthetic weights (bench / smoke: no checkpoints are reachable offline) ------------------------
    @classmethod
    def random_init(cls, device="cuda", dtype=torch.bfloat16, seed=0, **config):
        """Materialise the architecture directly on `device` with random weights of realistic scale
        (xavier-like Linear weights, N(0,0.02) embeddings/biases, modulation ~ N(0,1/C) as :462;
        head.head.weight ~ N(0,0.02) instead of the reference's zeros so the output is not 0)."""
        with torch.device("meta"):
            model = cls(**config)
        model.to_empty(device=device)
        d = model.d
        model.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                                 rope_params(1024, 2 * (d // 6))], dim=1).to(device)
        g = torch.Generator(device=device).manual_seed(seed)
        with torch.no_grad():
            for name, prm in model.named_parameters():
                if name.endswith("modulation"):
                    std = 1.0 / math.sqrt(prm.shape[-1])
                elif name.endswith(".bias"):
                    std = 0.02
                elif name.endswith("norm_q.weight") or name.endswith("norm_k.weight") or name.endswith("norm3.weight"):
                    prm.data = (1.0 + 0.05 * torch.randn(prm.shape, generator=g, device=device)).to(dtype)
                    continue
                elif name.startswith(("text_embedding", "time_embedding", "head.head")):
                    std = 0.02
                elif prm.dim() >= 2:
                    fan_out, fan_in = prm.shape[0], prm[0].numel()
                    std = math.sqrt(2.0 / (fan_in + fan_out))
                else:
                    std = 0.02
                prm.data = (torch.randn(prm.shape, generator=g, device=device, dtype=torch.float32) * std).to(dtype)
        for prm in model.parameters():
            prm.requires_grad_(False)
        return model.eval()

    # ---- loading ---------------------------------------------------------------------------------------
    @classmethod
    def from_config(cls, config, **kwargs):
        import inspect
        valid = set(inspect.signature(cls.__init__).parameters) - {"self"}
        merged = {k: v for k, v in dict(config, **kwargs).items() if k in valid}
        return cls(**merged)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, subfolder=None, transformer_additional_kwargs={},
                        low_cpu_mem_usage=False, torch_dtype=torch.bfloat16):
        """reference :1157-1299: config.json + (sharded) safetensors / .bin in a directory."""
        if subfolder is not None:
            pretrained_model_path = os.path.join(pretrained_model_path, subfolder)
        config_file = os.path.join(pretrained_model_path, "config.json")
        if not os.path.isfile(config_file):
            raise RuntimeError(f"{config_file} does not exist")
        with open(config_file) as fh:
            config = json.load(fh)
        kw = dict(transformer_additional_kwargs)
        for key, dst in kw.pop("dict_mapping", {}).items():
            kw[dst] = config[key]
        with torch.device("meta"):
            model = cls.from_config(config, **kw)
        freqs = None
        with torch.device("cpu"):
            d = model.d
            freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                               rope_params(1024, 2 * (d // 6))], dim=1)
        model.freqs = freqs
        state = {}
        bin_file = os.path.join(pretrained_model_path, "diffusion_pytorch_model.bin")
        if os.path.exists(bin_file):
            state = torch.load(bin_file, map_location="cpu")
        else:
            from safetensors.torch import load_file
            files = sorted(glob.glob(os.path.join(pretrained_model_path, "*.safetensors")))
            if not files:
                raise RuntimeError(f"no weights found under {pretrained_model_path}")
            for fpath in files:
                state.update(load_file(fpath))
        own = dict(model.named_parameters())
        # patch_embedding with a different number of input channels: overlapping channels copied, the rest zero (:1270-1273)
        pe = state.get("patch_embedding.weight")
        if pe is not None and tuple(pe.shape) != tuple(own["patch_embedding.weight"].shape) and pe.dim() == 5:
            tgt = torch.zeros(tuple(own["patch_embedding.weight"].shape), dtype=pe.dtype)
            n = min(tgt.shape[1], pe.shape[1])
            if tuple(tgt[:, :n].shape) == tuple(pe[:, :n].shape):
                tgt[:, :n] = pe[:, :n]
                state["patch_embedding.weight"] = tgt
        for key in list(state):
            if key in own and tuple(state[key].shape) != tuple(own[key].shape):
                print(key, "Size don't match, skip")                                       # (:1276-1280)
                del state[key]
        missing = [name for name in own if name not in state]
        unexpected = sorted(set(state) - set(own))
        if missing and low_cpu_mem_usage:
            # the reference's meta-device loader refuses an incomplete checkpoint (:1222-1229)
            raise ValueError(f"Cannot load {cls} from {pretrained_model_path} because the following keys are missing: \n "
                             f"{', '.join(missing)}. \n Please make sure to pass `low_cpu_mem_usage=False` if you want "
                             "to randomly initialize those weights or else make sure your checkpoint file is correct.")
        for name, prm in own.items():
            src = state.get(name)
            if src is None:
                src = _reference_init(name, tuple(prm.shape), model.dim)     # what init_weights left there (:1133-1155)
            mod, _, leaf = name.rpartition(".")
            target = model.get_submodule(mod) if mod else model
            setattr(target, leaf, nn.Parameter(src.to(torch_dtype), requires_grad=False))
        print(f"### missing keys: {len(missing)}; \n### unexpected keys: {len(unexpected)};")
        print(missing)                                                                     # (:1283-1284)
        return model
