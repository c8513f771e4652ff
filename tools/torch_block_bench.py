"""Library baseline on the SAME GPU: one Wan DiT block of the workload's widths written in plain PyTorch the way the
reference's own CUDA path runs it — cuBLAS Linears on bf16 weights, flash-attn 2 for self- and cross-attention (the
reference's default backend, videox_fun/models/attention_utils.py:115-149), composed elementwise ops for LayerNorm /
modulate / RMSNorm / RoPE / gates — timed with CUDA events.  It answers "how fast would the reference's kernels be on
this B200", which bench.py's CPU reference arm cannot: the reference package itself cannot be installed or travel to the
GPU box, so this is a restatement of its op sequence (wan_transformer3d.py:271-305, 310-336, 464-515),
not its code.  A measurement tool only — nothing in the product or in bench.py imports it.

    python tools/torch_block_bench.py                      # c2 widths, 75 600 tokens, 3 warm-up + 5 timed blocks
    python tools/torch_block_bench.py --tokens 1280 --dim 1536 --ffn 8960 --heads 12
Prints one JSON line: ms per block, the x40-layer step estimate, TFLOP/s over the algorithmic FLOPs.
On a CPU-only box it runs tiny shapes through SDPA (tests/test_tools_cpu.py) to keep the script honest."""
import argparse
import json
import math
import sys
import time

import torch
import torch.nn.functional as F


def attention(q, k, v):
    """q [L, n, d], k / v [S, n, d] bf16 -> [L, n, d]; flash-attn 2 on CUDA (the reference's backend), SDPA elsewhere."""
    if q.is_cuda:
        try:
            from flash_attn import flash_attn_func
            return flash_attn_func(q[None], k[None], v[None])[0], "flash_attn2"
        except ImportError:
            pass
    o = F.scaled_dot_product_attention(q.transpose(0, 1)[None], k.transpose(0, 1)[None], v.transpose(0, 1)[None])
    return o[0].transpose(0, 1), "sdpa"


def rms_norm(x, w, eps=1e-6):
    # WanRMSNorm (:214-230): fp32 mean, factor cast back to the activation dtype, two bf16 products
    return x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps).to(x.dtype) * w


def rope(x, freqs_cis):
    # rope_apply (:135-211): interleaved pairs as complex numbers in float64, back to the activation dtype
    L, n, d = x.shape
    xc = torch.view_as_complex(x.to(torch.float64).reshape(L, n, d // 2, 2))
    return torch.view_as_real(xc * freqs_cis[:, None, :]).flatten(2).to(x.dtype)


class Block:
    def __init__(self, C, Fd, n, dev, g):
        r = lambda *s: (torch.randn(*s, generator=g) * 0.02).to(dev, torch.bfloat16)
        self.n, self.C = n, C
        self.w = {k: r(C, C) for k in ("sq", "sk", "sv", "so", "cq", "ck", "cv", "co")}
        self.b = {k: r(C) for k in self.w}
        self.w1, self.b1, self.w2, self.b2 = r(Fd, C), r(Fd), r(C, Fd), r(C)
        self.nq, self.nk, self.cnq, self.cnk = (torch.ones(C, device=dev, dtype=torch.bfloat16) for _ in range(4))
        self.n3w, self.n3b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        self.modulation = (torch.randn(6, C, generator=g) / math.sqrt(C)).to(dev)

    def forward(self, x, e0, ctx, freqs_cis):
        """x fp32 [L, C]; e0 fp32 [6, C]; ctx bf16 [S, C] -> fp32 [L, C]  (:464-515)."""
        L, C, n = x.shape[0], self.C, self.n
        d = C // n
        e = (self.modulation + e0).chunk(6, dim=0)
        t = (F.layer_norm(x, (C,), eps=1e-6) * (1 + e[1]) + e[0]).to(torch.bfloat16)
        q = rms_norm(F.linear(t, self.w["sq"], self.b["sq"]), self.nq).view(L, n, d)
        k = rms_norm(F.linear(t, self.w["sk"], self.b["sk"]), self.nk).view(L, n, d)
        v = F.linear(t, self.w["sv"], self.b["sv"]).view(L, n, d)
        o, self.backend = attention(rope(q, freqs_cis), rope(k, freqs_cis), v)
        x = x + F.linear(o.reshape(L, C), self.w["so"], self.b["so"]) * e[2]
        t = F.layer_norm(x, (C,), self.n3w, self.n3b, 1e-6).to(torch.bfloat16)
        S = ctx.shape[0]
        q = rms_norm(F.linear(t, self.w["cq"], self.b["cq"]), self.cnq).view(L, n, d)
        k = rms_norm(F.linear(ctx, self.w["ck"], self.b["ck"]), self.cnk).view(S, n, d)
        v = F.linear(ctx, self.w["cv"], self.b["cv"]).view(S, n, d)
        o, _ = attention(q, k, v)
        x = x + F.linear(o.reshape(L, C), self.w["co"], self.b["co"])
        t = (F.layer_norm(x, (C,), eps=1e-6) * (1 + e[4]) + e[3]).to(torch.bfloat16)
        y = F.linear(F.gelu(F.linear(t, self.w1, self.b1), approximate="tanh"), self.w2, self.b2)
        return x + y * e[5]


def block_flops(L, C, Fd, S=512):
    """8LC^2 (self q,k,v,o) + 4L^2C (self attention) + 4LC^2 + 4SC^2 (cross) + 4LSC + 4LCF (FFN)."""
    return 8 * L * C * C + 4 * L * L * C + 4 * L * C * C + 4 * S * C * C + 4 * L * S * C + 4 * L * C * Fd


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=75600)
    ap.add_argument("--dim", type=int, default=5120)
    ap.add_argument("--ffn", type=int, default=13824)
    ap.add_argument("--heads", type=int, default=40)
    ap.add_argument("--layers", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    a = ap.parse_args(argv)
    dev = torch.device(a.device)
    g = torch.Generator().manual_seed(0)
    L, C = a.tokens, a.dim
    blk = Block(C, a.ffn, a.heads, dev, g)
    x = torch.randn(L, C, generator=g).to(dev)
    e0 = (torch.randn(6, C, generator=g) * 0.1).to(dev)
    ctx = torch.randn(512, C, generator=g).to(dev, torch.bfloat16)
    ang = torch.rand(L, C // a.heads // 2, generator=g, dtype=torch.float64) * 6.28
    freqs_cis = torch.polar(torch.ones_like(ang), ang).to(dev)
    with torch.no_grad():
        for _ in range(a.warmup):
            blk.forward(x, e0, ctx, freqs_cis)
        if dev.type == "cuda":
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(a.iters):
                y = blk.forward(x, e0, ctx, freqs_cis)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / a.iters
        else:
            t0 = time.perf_counter()
            for _ in range(a.iters):
                y = blk.forward(x, e0, ctx, freqs_cis)
            ms = (time.perf_counter() - t0) * 1e3 / a.iters
    fl = block_flops(L, C, a.ffn)
    line = {"tool": "torch_block_bench", "what": "one DiT block, plain PyTorch ops + " + blk.backend + " (library baseline, "
            "restated op sequence of the reference's CUDA path)", "device": torch.cuda.get_device_name(0) if dev.type == "cuda"
            else "cpu", "tokens": L, "dim": C, "ffn": a.ffn, "heads": a.heads, "ms_per_block": ms,
            "tflops": fl / ms / 1e9, "est_ms_per_step": ms * a.layers, "est_steps_per_sec": 1e3 / (ms * a.layers),
            "finite": bool(torch.isfinite(y).all()), "iters": a.iters, "warmup": a.warmup}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
