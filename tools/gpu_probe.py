"""Kernel-by-kernel parity probe against plain torch fp32 math ON THE GPU.

Each case runs in its own subprocess with a timeout: a trapping / hanging kernel poisons
only its own CUDA context and the rest of the probe still reports.  Results go to
gpurun_out/probe.jsonl (one JSON object per case) and a summary is printed.

    python tools/gpu_probe.py            # run everything
    python tools/gpu_probe.py gemm attn  # run the groups named
    python tools/gpu_probe.py --case <json>   (internal)
"""
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cases():
    cs = [dict(kind="umma", mode=m) for m in range(4)]
    for (M, N, K) in [(128, 256, 64), (256, 256, 256), (300, 512, 256), (1000, 1536, 1536),
                      (512, 64, 5120), (777, 8960, 1536), (640, 128, 512), (4096, 5120, 5120)]:
        for epi in ("bias", "bias_gelu", "bias_gate_res", "bias_f32"):
            if M * N * K > 1e9 and epi not in ("bias", "bias_gate_res"):
                continue
            cs.append(dict(kind="gemm", M=M, N=N, K=K, epi=epi))
    cs.append(dict(kind="gemm", M=200, N=104, K=64, epi="bias"))      # ragged N, BN=128
    cs.append(dict(kind="gemm", M=200, N=300, K=192, epi="bias_gate_res"))  # ragged N, BN=256
    for vt in (0, 1):
        for (Lq, Lk, kv, h) in [(128, 128, 128, 1), (256, 128, 128, 1), (256, 256, 256, 2),
                                (300, 333, 333, 2), (512, 640, 600, 3), (1280, 1280, 1280, 12),
                                (2048, 512, 512, 4), (5000, 5000, 5000, 2)]:
            if vt and kv % 8:
                continue
            cs.append(dict(kind="attn", Lq=Lq, Lk=Lk, kv=kv, heads=h, vt=vt))
    cs.append(dict(kind="attn", Lq=1024, Lk=4096, kv=4096, heads=2, vt=0, big_scores=1))
    # > 148 work items: persistent CTAs walk several (head, q-block) items (barrier phases across items)
    cs.append(dict(kind="attn", Lq=4096, Lk=4000, kv=3999, heads=12, vt=0))
    cs.append(dict(kind="attn", Lq=9000, Lk=1111, kv=1111, heads=10, vt=0))
    for (L, C) in [(64, 1536), (1000, 5120), (333, 2048)]:
        for mode in ("plain", "mod", "affine"):
            cs.append(dict(kind="ln", L=L, C=C, mode=mode))
    for (F, H, W, heads) in [(5, 16, 16, 12), (3, 21, 37, 4)]:
        for mode in ("norope", "plain", "cot"):
            cs.append(dict(kind="rms", F=F, H=H, W=W, heads=heads, mode=mode))
    cs.append(dict(kind="patch", F=5, H=32, W=32))
    cs.append(dict(kind="patch", F=3, H=42, W=74))
    cs.append(dict(kind="lin", B=2, N=1536, K=256))
    cs.append(dict(kind="lin", B=1, N=9216, K=1536))
    return cs


def run_case(c):
    import torch
    from videocof_b200 import ops
    torch.manual_seed(1234)
    dev = "cuda"
    kind = c["kind"]
    res = dict(c)

    def stats(got, ref):
        got = got.float()
        ref = ref.float()
        d = (got - ref).abs()
        res["max_abs"] = float(d.max())
        res["ref_absmax"] = float(ref.abs().max())
        res["rel_fro"] = float(d.norm() / (ref.norm() + 1e-30))
        res["nan"] = bool(torch.isnan(got).any())

    if kind == "umma":
        import probe_lib as _lib            # tests/native/libvcof_probes.so: the probes are not in the product library
        a = torch.randn(128, 64, device=dev).bfloat16()
        b = torch.randn(128, 64, device=dev).bfloat16()
        d = torch.zeros(128, 128, device=dev)
        _lib.call("vcof_debug_umma_probe", a.data_ptr(), b.data_ptr(), d.data_ptr(), c["mode"],
                  torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        stats(d, a.float() @ b.float().t())
        res["ok"] = (not res["nan"]) and res["rel_fro"] < 1e-5
    elif kind == "gemm":
        M, N, K = c["M"], c["N"], c["K"]
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        acc = a.float() @ w.float().t() + b.float()
        epi = c["epi"]
        if epi == "bias":
            got = ops.gemm(a, w, b, "bias")
            ref = acc
        elif epi == "bias_gelu":
            got = ops.gemm(a, w, b, "bias_gelu")
            ref = torch.nn.functional.gelu(acc.bfloat16().float(), approximate="tanh")
        elif epi == "bias_f32":
            got = ops.gemm(a, w, b, "bias_f32")
            ref = acc.bfloat16().float()
        else:
            x0 = torch.randn(M, N, device=dev)
            g = torch.randn(N, device=dev)
            got = x0.clone()
            ops.gemm(a, w, b, "bias_gate_res", out=got, gate=g)
            ref = x0 + g * acc.bfloat16().float()
        torch.cuda.synchronize()
        stats(got, ref)
        res["ok"] = (not res["nan"]) and res["rel_fro"] < 6e-3
    elif kind == "attn":
        Lq, Lk, kv, h = c["Lq"], c["Lk"], c["kv"], c["heads"]
        C = h * 128
        sc = 4.0 if c.get("big_scores") else 1.0
        q = (torch.randn(Lq, C, device=dev) * sc).bfloat16()
        k = (torch.randn(Lk, C, device=dev) * sc).bfloat16()
        v = torch.randn(Lk, C, device=dev).bfloat16()
        if c["vt"]:
            ldv = (Lk + 7) // 8 * 8
            vt = torch.zeros(C, ldv, device=dev, dtype=torch.bfloat16)
            vt[:, :Lk] = v.t()
            got = ops.attention(q, k, vt, h, kv_len=kv, v_transposed=True)
        else:
            got = ops.attention(q, k, v, h, kv_len=kv)
        torch.cuda.synchronize()
        qf = q.float().view(Lq, h, 128).transpose(0, 1)
        kf = k.float()[:kv].view(kv, h, 128).transpose(0, 1)
        vf = v.float()[:kv].view(kv, h, 128).transpose(0, 1)
        s = (qf @ kf.transpose(1, 2)) / math.sqrt(128)
        ref = (torch.softmax(s, dim=-1) @ vf).transpose(0, 1).reshape(Lq, C)
        stats(got, ref)
        res["ok"] = (not res["nan"]) and res["rel_fro"] < 1.5e-2
    elif kind == "ln":
        L, C = c["L"], c["C"]
        x = torch.randn(L, C, device=dev) * 3 + 0.5
        w = b = sh = scl = None
        if c["mode"] == "mod":
            sh = torch.randn(C, device=dev)
            scl = torch.randn(C, device=dev) * 0.1
        if c["mode"] == "affine":
            w = torch.randn(C, device=dev)
            b = torch.randn(C, device=dev)
        got = ops.ln_modulate(x, w, b, sh, scl, 1e-6)
        ref = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-6)
        if sh is not None:
            ref = ref * (1 + scl) + sh
        torch.cuda.synchronize()
        stats(got, ref)
        res["ok"] = (not res["nan"]) and res["rel_fro"] < 4e-3
    elif kind == "rms":
        F, H, W, h = c["F"], c["H"], c["W"], c["heads"]
        d = 128
        C = h * d
        L = F * H * W + 5  # 5 padding rows (normalised, not rotated)
        x = torch.randn(L, C, device=dev).bfloat16()
        wgt = (1 + 0.1 * torch.randn(C, device=dev)).bfloat16()
        xf = x.float()
        rs = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6).bfloat16()
        y = ((x * rs) * wgt)  # bf16 double rounding as the reference
        ref = y.float()
        rope = None
        if c["mode"] != "norope":
            n_t, n_h = 22, 21
            def params(n, dim):
                fr = torch.outer(torch.arange(n, dtype=torch.float64),
                                 1.0 / torch.pow(10000, torch.arange(0, dim, 2, dtype=torch.float64) / dim))
                return fr
            ang = torch.cat([params(1024, 44), params(1024, 42), params(1024, 42)], dim=1)  # [1024,64]
            table = torch.stack([ang.cos(), ang.sin()], dim=-1).float().to(dev).contiguous()
            if c["mode"] == "plain":
                tpos = torch.arange(F, dtype=torch.int32)
            else:
                fs = F // 2
                tpos = torch.tensor(list(range(1, fs + 1)) + [0] + list(range(1, F - fs)), dtype=torch.int32)
            tpos_d = tpos.to(dev)
            rope = ops.RopeSpec(table, tpos_d, F, H, W, n_t, n_h, 0)
            # reference in float64
            n = F * H * W
            l = torch.arange(n)
            f = l // (H * W)
            hh = (l // W) % H
            ww = l % W
            pos = torch.empty(n, 64, dtype=torch.long)
            pos[:, :22] = tpos[f].long()[:, None]
            pos[:, 22:43] = hh[:, None]
            pos[:, 43:] = ww[:, None]
            a = ang[pos, torch.arange(64)[None, :]].to(dev)  # [n,64]
            yy = y[:n].double().view(n, h, 64, 2)
            re = yy[..., 0] * a.cos()[:, None] - yy[..., 1] * a.sin()[:, None]
            im = yy[..., 0] * a.sin()[:, None] + yy[..., 1] * a.cos()[:, None]
            ref = ref.clone()
            ref[:n] = torch.stack([re, im], dim=-1).reshape(n, C).float()
        got = ops.rmsnorm_rope_(x.clone(), wgt, 1e-6, d, rope)
        torch.cuda.synchronize()
        stats(got, ref.bfloat16())
        res["ok"] = (not res["nan"]) and res["rel_fro"] < 3e-3
        # column-blocked output (head-exchange send layout) and the pack / unpack kernel: bit-exact rearrangements
        for P in (p_ for p_ in (2, 4) if h % p_ == 0):
            rows = x.shape[0]
            blk = torch.full((P, rows, C // P), 7.0, dtype=torch.bfloat16, device=dev)
            x0 = x.clone()
            ops.rmsnorm_rope_(x0, wgt, 1e-6, d, rope, out_blocked=blk)
            want = got.view(rows, P, C // P).transpose(0, 1)
            packed = ops.copy_blocked(got, torch.empty_like(blk), True)
            back = ops.copy_blocked(torch.empty_like(got), packed, False)
            torch.cuda.synchronize()
            exact = bool((blk == want).all()) and bool((x0 == x).all()) and bool((packed == want).all()) \
                and bool((back == got).all())
            res[f"blocked_P{P}_exact"] = exact
            res["ok"] = res["ok"] and exact
    elif kind == "patch":
        F, H, W = c["F"], c["H"], c["W"]
        x = torch.randn(16, F, H, W, device=dev).bfloat16()
        a = ops.patchify(x)
        ref = x.view(16, F, H // 2, 2, W // 2, 2).permute(1, 2, 4, 0, 3, 5).reshape(-1, 64)
        ok1 = bool((a == ref).all())
        y = torch.randn(F * (H // 2) * (W // 2), 64, device=dev).bfloat16()
        u = ops.unpatchify(y, 16, F, H, W)
        r2 = y.view(F, H // 2, W // 2, 1, 2, 2, 16)
        r2 = torch.einsum("fhwpqrc->cfphqwr", r2).reshape(16, F, H, W)
        ok2 = bool((u == r2).all())
        torch.cuda.synchronize()
        res["ok"] = ok1 and ok2
        res["patchify_exact"], res["unpatchify_exact"] = ok1, ok2
    elif kind == "lin":
        B, N, K = c["B"], c["N"], c["K"]
        x = torch.randn(B, K, device=dev)
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        got = ops.linear_f32(x, w, b, act_in=True, act_out=True)
        ref = torch.nn.functional.silu(torch.nn.functional.silu(x) @ w.float().t() + b.float())
        torch.cuda.synchronize()
        stats(got, ref)
        res["ok"] = (not res["nan"]) and res["rel_fro"] < 1e-5
    return res


def batch_key(c):
    return (c["kind"], c.get("epi"), c.get("vt"), c.get("mode") if c["kind"] == "umma" else None)


def run_batch(cs):
    """Run several cases in this process; stop once the CUDA context is dead."""
    import torch
    dead = False
    for c in cs:
        if dead:
            r = dict(c, ok=False, error="skipped: CUDA context dead after an earlier failure")
        else:
            try:
                r = run_case(c)
            except Exception as e:  # noqa: BLE001 - report everything
                r = dict(c, ok=False, error=f"{type(e).__name__}: {e}"[:400])
                try:
                    torch.cuda.synchronize()
                except Exception:  # noqa: BLE001
                    dead = True
        print("PROBE_RESULT " + json.dumps(r), flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--cases":
        run_batch(json.loads(sys.argv[2]))
        return 0
    groups = set(sys.argv[1:])
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    batches = {}
    for c in cases():
        if groups and c["kind"] not in groups:
            continue
        batches.setdefault(batch_key(c), []).append(c)
    results = []
    t0 = time.time()
    with open(os.path.join(out_dir, "probe.jsonl"), "w") as f:
        for key, cs in batches.items():
            got = []
            try:
                p = subprocess.run([sys.executable, __file__, "--cases", json.dumps(cs)],
                                   capture_output=True, text=True, timeout=240, cwd=ROOT)
                out, err, rc = p.stdout, p.stderr, p.returncode
            except subprocess.TimeoutExpired as e:
                out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                err, rc = "timeout", -9
            for l in out.splitlines():
                if l.startswith("PROBE_RESULT "):
                    got.append(json.loads(l[len("PROBE_RESULT "):]))
            for c in cs[len(got):]:
                got.append(dict(c, ok=False, error="no result; rc=%s" % rc,
                                tail=(out[-300:] + err[-800:])))
            for r in got:
                results.append(r)
                f.write(json.dumps(r) + "\n")
                brief = {k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items()}
                print(("PASS " if r.get("ok") else "FAIL ") + json.dumps(brief), flush=True)
            f.flush()
    bad = [r for r in results if not r.get("ok")]
    print(f"probe: {len(results) - len(bad)}/{len(results)} passed in {time.time() - t0:.0f}s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
