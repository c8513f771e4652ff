"""Stand-alone timing of single libvcof kernels at workload shapes (CUDA events, L2-cold inputs).

    python tools/kbench.py attn --L 75600 --heads 8 [--Lk 512] [--iters 5]
    python tools/kbench.py gemm --M 75600 --N 5120 --K 5120 [--epi bias]
    python tools/kbench.py ln|rms --L 75600 --C 5120

Prints one JSON line per run (achieved TFLOP/s or GB/s against MEASURED_PEAKS.json).  Also the
command ncu captures are taken from (profiles/README.md).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videocof_b200 import ops  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    return d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0)


def timeit(fn, iters, flush):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()          # > L2-sized write: evicts inputs between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return min(ms), sum(ms) / len(ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kind", choices=["attn", "gemm", "ln", "rms"])
    ap.add_argument("--L", type=int, default=75600)
    ap.add_argument("--Lk", type=int, default=0)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--M", type=int, default=75600)
    ap.add_argument("--N", type=int, default=5120)
    ap.add_argument("--K", type=int, default=5120)
    ap.add_argument("--C", type=int, default=5120)
    ap.add_argument("--epi", default="bias")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--vt", type=int, default=0)
    a = ap.parse_args()
    dev = "cuda"
    tf_peak, hbm_peak = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = dict(kind=a.kind)
    if a.kind == "attn":
        Lk = a.Lk or a.L
        C = a.heads * 128
        q = torch.randn(a.L, C, device=dev).bfloat16()
        k = torch.randn(Lk, C, device=dev).bfloat16()
        v = torch.randn(Lk, C, device=dev).bfloat16()
        o = torch.empty_like(q)
        if a.vt:
            v = v.t().contiguous()
        best, avg = timeit(lambda: ops.attention(q, k, v, a.heads, out=o, v_transposed=bool(a.vt)), a.iters, flush)
        fl = 4.0 * a.L * Lk * C
        out.update(L=a.L, Lk=Lk, heads=a.heads, ms_best=best, ms_avg=avg, tflops=fl / best / 1e9,
                   frac_of_measured_burst=fl / best / 1e9 / tf_peak)
    elif a.kind == "gemm":
        x = torch.randn(a.M, a.K, device=dev).bfloat16()
        w = (torch.randn(a.N, a.K, device=dev) / a.K ** 0.5).bfloat16()
        b = torch.randn(a.N, device=dev).bfloat16()
        kw = {}
        if a.epi == "bias_gate_res":
            kw = dict(out=torch.zeros(a.M, a.N, device=dev), gate=torch.randn(a.N, device=dev))
        best, avg = timeit(lambda: ops.gemm(x, w, b, a.epi, **kw), a.iters, flush)
        fl = 2.0 * a.M * a.N * a.K
        out.update(M=a.M, N=a.N, K=a.K, epi=a.epi, ms_best=best, ms_avg=avg, tflops=fl / best / 1e9,
                   frac_of_measured_burst=fl / best / 1e9 / tf_peak)
    elif a.kind == "ln":
        x = torch.randn(a.L, a.C, device=dev)
        sh, sc = torch.randn(a.C, device=dev), torch.randn(a.C, device=dev)
        o = torch.empty(a.L, a.C, device=dev, dtype=torch.bfloat16)
        best, avg = timeit(lambda: ops.ln_modulate(x, None, None, sh, sc, 1e-6, out=o), a.iters, flush)
        by = a.L * a.C * 6.0
        out.update(L=a.L, C=a.C, ms_best=best, gbs=by / best / 1e6, frac_of_measured_hbm=by / best / 1e6 / hbm_peak)
    else:
        x = torch.randn(a.L, a.C, device=dev).bfloat16()
        w = torch.ones(a.C, device=dev).bfloat16()
        best, avg = timeit(lambda: ops.rmsnorm_rope_(x, w, 1e-6, 128, None), a.iters, flush)
        by = a.L * a.C * 4.0
        out.update(L=a.L, C=a.C, ms_best=best, gbs=by / best / 1e6, frac_of_measured_hbm=by / best / 1e6 / hbm_peak)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
