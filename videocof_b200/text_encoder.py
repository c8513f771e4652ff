"""B200-native umT5 text encoder with the reference's class API (videox_fun.models.WanT5EncoderModel).

Host-side mirror of videox_fun/models/wan_text_encoder.py (reference; SURVEY.md §8f rank 3 — the step right before
the denoising path): same constructor arguments, sub-module tree and state-dict keys
(`token_embedding.weight`, `blocks.N.{norm1,norm2}.weight`, `blocks.N.attn.{q,k,v,o}.weight`,
`blocks.N.ffn.{gate.0,fc1,fc2}.weight`, `blocks.N.pos_embedding.embedding.weight`, `norm.weight`), same
`forward(input_ids, attention_mask) -> (hidden,)`, but the forward is libvcof launches only:

    embed_rows -> per block [ t5_rmsnorm -> 3 x GEMM -> t5_attn -> GEMM(+= residual)
                               t5_rmsnorm -> GEMM(gelu) -> GEMM(*= gate) -> GEMM(+= residual) ] -> t5_rmsnorm

The bias-free Linears run on the tcgen05 GEMM; its "mul" / "add" epilogues keep the gated product and the bf16 residual
stream inside the GEMM (no separate elementwise passes).  The [heads, L, L] position-bias tensor of the reference is
never built: it depends on (key - query) only, so each layer hands the attention kernel a [heads, 2L-1] table.

There is no fallback: CPU tensors, non-bf16 weights or a missing libvcof raise.
"""
import inspect
import math

import torch
import torch.nn as nn

from . import ops
from ._lib import VcofError

__all__ = ["WanT5EncoderModel", "T5SelfAttention", "T5Attention", "T5FeedForward", "T5LayerNorm",
           "T5RelativeEmbedding", "GELU"]


class GELU(nn.Module):
    """reference :39-42 (parameter-free; fused into the gate GEMM's epilogue here)."""


class T5LayerNorm(nn.Module):
    """reference :45-57."""

    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))


class T5Attention(nn.Module):
    """reference :60-112 (parameters: bias-free q, k, v, o)."""

    def __init__(self, dim, dim_attn, num_heads, dropout=0.1):
        assert dim_attn % num_heads == 0
        super().__init__()
        self.dim, self.dim_attn, self.num_heads, self.head_dim = dim, dim_attn, num_heads, dim_attn // num_heads
        self.q = nn.Linear(dim, dim_attn, bias=False)
        self.k = nn.Linear(dim, dim_attn, bias=False)
        self.v = nn.Linear(dim, dim_attn, bias=False)
        self.o = nn.Linear(dim_attn, dim, bias=False)
        self.dropout = nn.Dropout(dropout)


class T5FeedForward(nn.Module):
    """reference :115-133: fc2(fc1(x) * gelu(gate(x)))."""

    def __init__(self, dim, dim_ffn, dropout=0.1):
        super().__init__()
        self.dim, self.dim_ffn = dim, dim_ffn
        self.gate = nn.Sequential(nn.Linear(dim, dim_ffn, bias=False), GELU())
        self.fc1 = nn.Linear(dim, dim_ffn, bias=False)
        self.fc2 = nn.Linear(dim_ffn, dim, bias=False)
        self.dropout = nn.Dropout(dropout)


class T5RelativeEmbedding(nn.Module):
    """reference :199-247."""

    def __init__(self, num_buckets, num_heads, bidirectional, max_dist=128):
        super().__init__()
        self.num_buckets, self.num_heads, self.bidirectional, self.max_dist = num_buckets, num_heads, bidirectional, max_dist
        self.embedding = nn.Embedding(num_buckets, num_heads)

    def _relative_position_bucket(self, rel_pos):
        """:224-247, verbatim arithmetic order (fp32 log, truncation) so boundaries fall where the reference's do."""
        if self.bidirectional:
            num_buckets = self.num_buckets // 2
            rel_buckets = (rel_pos > 0).long() * num_buckets
            rel_pos = torch.abs(rel_pos)
        else:
            num_buckets = self.num_buckets
            rel_buckets = 0
            rel_pos = -torch.min(rel_pos, torch.zeros_like(rel_pos))
        max_exact = num_buckets // 2
        large = max_exact + (torch.log(rel_pos.float() / max_exact) / math.log(self.max_dist / max_exact)
                             * (num_buckets - max_exact)).long()
        large = torch.min(large, torch.full_like(large, num_buckets - 1))
        return rel_buckets + torch.where(rel_pos < max_exact, rel_pos, large)

    def table(self, L):
        """fp32 [heads, 2L-1] on the weight's device: entry (key - query) + L - 1 is the bias of that offset.  The
        2L-1 bucket ids are integer work done on the host, so they do not depend on a device's logf rounding and
        equal the reference's CPU result bit for bit (tests/golden/t5_tiny.npz)."""
        w = self.embedding.weight
        key = (L, w.data_ptr(), w._version, w.device)
        hit = self.__dict__.get("_vcof_table")
        if hit is None or hit[0] != key:
            buckets = self._relative_position_bucket(torch.arange(-(L - 1), L))
            tab = w.detach()[buckets.to(w.device)].t().to(torch.float32).contiguous()
            self.__dict__["_vcof_table"] = (key, tab)
        return self.__dict__["_vcof_table"][1]

    def forward(self, lq, lk):
        """[1, heads, lq, lk] as the reference builds it (:207-222); the fused path uses table() instead."""
        rel = torch.arange(lk).unsqueeze(0) - torch.arange(lq).unsqueeze(1)
        buckets = self._relative_position_bucket(rel).to(self.embedding.weight.device)
        return self.embedding.weight[buckets].permute(2, 0, 1).unsqueeze(0).contiguous()


class T5SelfAttention(nn.Module):
    """reference :136-158 — one encoder block."""

    def __init__(self, dim, dim_attn, dim_ffn, num_heads, num_buckets, shared_pos=True, dropout=0.1):
        super().__init__()
        self.dim, self.dim_attn, self.dim_ffn = dim, dim_attn, dim_ffn
        self.num_heads, self.num_buckets, self.shared_pos = num_heads, num_buckets, shared_pos
        self.norm1 = T5LayerNorm(dim)
        self.attn = T5Attention(dim, dim_attn, num_heads, dropout)
        self.norm2 = T5LayerNorm(dim)
        self.ffn = T5FeedForward(dim, dim_ffn, dropout)
        self.pos_embedding = None if shared_pos else T5RelativeEmbedding(num_buckets, num_heads, bidirectional=True)

    def run(self, x, ws, bias_rel, mask, B, L):
        """x bf16 [B*L, dim], updated in place (the bf16 residual stream of the reference's eager path)."""
        at, ff = self.attn, self.ffn
        n = x.shape[0]                                                        # <= a few 128-row tiles: pick the
        na, nd, nf = (ops.narrow_tiles_pay(n, w, ws.sms) for w in (self.dim_attn, self.dim, self.dim_ffn))  # tile width
        ops.t5_rmsnorm(x, self.norm1.weight, self.norm1.eps, out=ws.h)
        ops.gemm(ws.h, at.q.weight, None, "bias", out=ws.q, narrow=na)
        ops.gemm(ws.h, at.k.weight, None, "bias", out=ws.k, narrow=na)
        ops.gemm(ws.h, at.v.weight, None, "bias", out=ws.v, narrow=na)
        ops.t5_attention(ws.q, ws.k, ws.v, bias_rel, B, L, self.num_heads, key_mask=mask, out=ws.a)
        ops.gemm(ws.a, at.o.weight, None, "add", out=x, narrow=nd)            # x = x + o(attn)          (:156)
        ops.t5_rmsnorm(x, self.norm2.weight, self.norm2.eps, out=ws.h)
        ops.gemm(ws.h, ff.gate[0].weight, None, "bias_gelu", out=ws.g, narrow=nf)   # gelu(gate(x))      (:120, :129)
        ops.gemm(ws.h, ff.fc1.weight, None, "mul", out=ws.g, narrow=nf)       # fc1(x) * gelu(gate(x))   (:129)
        ops.gemm(ws.g, ff.fc2.weight, None, "add", out=x, narrow=nd)          # x = x + fc2(.)           (:131, :157)
        return x


class _Workspace:
    def __init__(self):
        self.key = None

    def get(self, device, n, dim, dim_attn, dim_ffn):
        key = (str(device), n, dim, dim_attn, dim_ffn)
        if self.key != key:
            bf = dict(dtype=torch.bfloat16, device=device)
            self.h = torch.empty((n, dim), **bf)
            self.q = torch.empty((n, dim_attn), **bf)
            self.k = torch.empty((n, dim_attn), **bf)
            self.v = torch.empty((n, dim_attn), **bf)
            self.a = torch.empty((n, dim_attn), **bf)
            self.g = torch.empty((n, dim_ffn), **bf)
            self.sms = torch.cuda.get_device_properties(device).multi_processor_count if torch.cuda.is_available() \
                else 148
            self.key = key
        return self


class WanT5EncoderModel(nn.Module):
    """reference :249-394."""

    def __init__(self, vocab, dim, dim_attn, dim_ffn, num_heads, num_layers, num_buckets, shared_pos=True,
                 dropout=0.1):
        super().__init__()
        self.dim, self.dim_attn, self.dim_ffn = dim, dim_attn, dim_ffn
        self.num_heads, self.num_layers, self.num_buckets, self.shared_pos = num_heads, num_layers, num_buckets, shared_pos
        self.token_embedding = vocab if isinstance(vocab, nn.Embedding) else nn.Embedding(vocab, dim)
        self.pos_embedding = T5RelativeEmbedding(num_buckets, num_heads, bidirectional=True) if shared_pos else None
        self.dropout = nn.Dropout(dropout)
        self.blocks = nn.ModuleList([T5SelfAttention(dim, dim_attn, dim_ffn, num_heads, num_buckets, shared_pos,
                                                     dropout) for _ in range(num_layers)])
        self.norm = T5LayerNorm(dim)
        self.__dict__["_ws"] = _Workspace()

    # ---- diffusers ModelMixin surface the CLIs / pipeline touch -------------------------------------------
    @property
    def dtype(self):
        return self.token_embedding.weight.dtype

    @property
    def device(self):
        return self.token_embedding.weight.device

    def _check(self):
        w = self.token_embedding.weight
        if not w.is_cuda:
            raise VcofError("WanT5EncoderModel: weights are on %s — libvcof has no CPU path; call .to('cuda')" % w.device)
        if w.dtype != torch.bfloat16:
            raise VcofError(f"WanT5EncoderModel: weights are {w.dtype}; the fused path computes in bf16 "
                            "(the CLIs load the text encoder with torch_dtype=bfloat16, fast_infer.py:308-312)")
        if self.training and self.dropout.p > 0:
            raise VcofError("WanT5EncoderModel: inference only (dropout is not implemented); call .eval() "
                            "(the shipped config sets dropout 0.0, config/wan2.1/wan_civitai.yaml:26)")

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None):
        """input_ids int64 [B, L <= 512]; attention_mask [B, L] (0 = padding) or None -> (bf16 [B, L, dim],).
        Masked KEYS are excluded; padded query rows are still computed, exactly as the reference does (:94-98) —
        the pipeline trims them afterwards (pipeline_wan.py:183)."""
        self._check()
        if input_ids.dim() != 2:
            raise VcofError(f"input_ids must be [B, L], got {tuple(input_ids.shape)}")
        vocab = self.token_embedding.weight.shape[0]
        if input_ids.numel() and (int(input_ids.min()) < 0 or int(input_ids.max()) >= vocab):
            raise IndexError(f"token id out of range [0, {vocab})")                 # nn.Embedding's behaviour
        dev = self.device
        B, L = input_ids.shape
        ids = input_ids.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        mask = None
        if attention_mask is not None:
            if attention_mask.dim() != 2 or tuple(attention_mask.shape) != (B, L):
                raise NotImplementedError("only [B, L] key masks are implemented (what the pipeline passes, "
                                          f"pipeline_wan.py:175); got {tuple(attention_mask.shape)}")
            mask = (attention_mask.to(dev) != 0).to(torch.int32).contiguous()
        x = ops.embed_rows(ids, self.token_embedding.weight)
        ws = self._ws.get(dev, B * L, self.dim, self.dim_attn, self.dim_ffn)
        shared = self.pos_embedding.table(L) if self.shared_pos else None
        for blk in self.blocks:
            blk.run(x, ws, shared if self.shared_pos else blk.pos_embedding.table(L), mask, B, L)
        out = ops.t5_rmsnorm(x, self.norm.weight, self.norm.eps)
        return (out.view(B, L, self.dim),)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, additional_kwargs={}, low_cpu_mem_usage=False,
                        torch_dtype=torch.bfloat16):
        """reference :296-394: a single .safetensors / .pth state dict + constructor kwargs from the YAML config
        (`text_encoder_kwargs`); unknown kwargs are filtered like the reference's filter_kwargs."""
        valid = set(inspect.signature(cls.__init__).parameters) - {"self"}
        kw = {k: v for k, v in dict(additional_kwargs).items() if k in valid}
        with torch.device("meta"):
            model = cls(**kw)
        if pretrained_model_path.endswith(".safetensors"):
            from safetensors.torch import load_file
            state = load_file(pretrained_model_path)
        else:
            state = torch.load(pretrained_model_path, map_location="cpu")
        own = dict(model.named_parameters())
        missing = []
        for name, prm in own.items():
            src = state.get(name)
            if src is None or tuple(src.shape) != tuple(prm.shape):
                missing.append(name)
                src = torch.zeros(prm.shape)
            mod, _, leaf = name.rpartition(".")
            setattr(model.get_submodule(mod) if mod else model, leaf,
                    nn.Parameter(src.to(torch_dtype), requires_grad=False))
        print(f"### missing keys: {len(missing)}; \n### unexpected keys: {len(set(state) - set(own))};")
        print(missing, sorted(set(state) - set(own)))
        return model.eval()
