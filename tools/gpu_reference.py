"""The reference's OWN CUDA path on this GPU: the unmodified files of knightyxp/VideoCoF (staged under the git-ignored
baseline/_ref by tools/stage_reference.py, loaded through tools/ref_loader.py's diffusers shim) run the way the CLIs
run them — bf16 weights, `torch.cuda.amp.autocast(bf16)` around the DiT call (pipeline_wan.py:707), flash-attn 2 as the
attention backend (attention_utils.py:113-146: flash_attn_varlen_func; FA3 / SageAttention are not installed), cuBLAS
Linears (wan_transformer3d.py:264-267, 457-459), cuDNN convolutions in the VAE (wan_vae.py:21-40).

This is "the kernel to beat" (SURVEY §8d last row): bench.py calls `measure()` for its
`gpu_reference` object and tests/test_widen_y_gpu_reference.py uses the same modules as the CUDA parity reference.
It is measurement / test infrastructure: nothing in the product imports it.

    python tools/gpu_reference.py                 # c2 widths: one WanAttentionBlock, FA2 self-attention, one Linear
    python tools/gpu_reference.py --full-step     # + the whole 40-layer reference DiT forward (needs ~90 GB of HBM)
    python tools/gpu_reference.py --vae           # + reference VAE decode / encode at 720p (cuDNN)
"""
import argparse
import contextlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG = "_vcof_gpu_ref"


def available():
    import ref_loader
    return os.path.isdir(os.path.join(ref_loader.REF, "videox_fun", "models"))


def load():
    """-> namespace (dit, vae, unipc, attention_utils) of the unmodified reference modules."""
    import ref_loader
    return ref_loader.load_reference(pkg=PKG)


@contextlib.contextmanager
def backend(name):
    """attention() reads VIDEOX_ATTENTION_TYPE at call time (attention_utils.py:173): FLASH_ATTENTION is the
    reference's default, anything unknown falls through to torch SDPA."""
    old = os.environ.get("VIDEOX_ATTENTION_TYPE")
    os.environ["VIDEOX_ATTENTION_TYPE"] = name
    try:
        yield
    finally:
        if old is None:
            os.environ.pop("VIDEOX_ATTENTION_TYPE", None)
        else:
            os.environ["VIDEOX_ATTENTION_TYPE"] = old


def autocast():
    """torch.cuda.amp.autocast(dtype=bf16) of pipeline_wan.py:707 (a CPU autocast on a CPU-only box, where only
    tests/test_tools_cpu.py exercises this file's plumbing with the SDPA backend)."""
    return torch.autocast("cuda" if torch.cuda.is_available() else "cpu", dtype=torch.bfloat16)


def cuda_time(fn, warmup=2, iters=3):
    """mean ms per call, CUDA events on the current stream, synchronised on both sides."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def make_block(cfg_kw, device, seed=0):
    """One reference WanAttentionBlock of the workload's widths, weights as the CLIs hold them (bf16, incl. the
    modulation parameter; fast_infer.py:281-286)."""
    ns = load()
    torch.manual_seed(seed)
    with torch.device(device):
        blk = ns.dit.WanAttentionBlock("t2v_cross_attn", cfg_kw["dim"], cfg_kw["ffn_dim"], cfg_kw["num_heads"],
                                       (-1, -1), True, True, 1e-6)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for name, prm in blk.named_parameters():
            if name.endswith("modulation"):
                continue                      # randn / sqrt(C) by the constructor (:462)
            if name.endswith("norm_q.weight") or name.endswith("norm_k.weight") or name == "norm3.weight":
                prm.copy_(1.0 + 0.05 * torch.randn(prm.shape, generator=g, device=device))
            elif name.endswith(".bias"):
                prm.copy_(0.02 * torch.randn(prm.shape, generator=g, device=device))
            else:
                std = (2.0 / (prm.shape[0] + prm.shape[1])) ** 0.5
                prm.copy_(std * torch.randn(prm.shape, generator=g, device=device))
    # (CPU-only plumbing check: fp32 weights — CPU autocast has no fp32 policy for layer_norm with bf16 parameters)
    return blk.to(torch.bfloat16 if torch.device(device).type == "cuda" else torch.float32).eval().requires_grad_(False)


def block_inputs(cfg_kw, f, h, w, device, seed=1, x_dtype=torch.float32, text_len=512):
    """Synthetic block inputs of the shapes WanTransformer3DModel.forward hands to a block (:1057-1069)."""
    C = cfg_kw["dim"]
    L = f * h * w
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(1, L, C, generator=g, device=device).to(x_dtype)
    e = 0.1 * torch.randn(1, 6, C, generator=g, device=device)
    ctx = torch.randn(1, text_len, C, generator=g, device=device).to(torch.bfloat16)
    return dict(x=x, e=e, context=ctx, seq_lens=torch.tensor([L], dtype=torch.long),
                grid_sizes=torch.tensor([[f, h, w]], dtype=torch.long))


def rope_freqs(head_dim, device):
    ns = load()
    d = head_dim
    return torch.cat([ns.dit.rope_params(1024, d - 4 * (d // 6)), ns.dit.rope_params(1024, 2 * (d // 6)),
                      ns.dit.rope_params(1024, 2 * (d // 6))], dim=1).to(device)


def run_block(blk, inp, freqs, fs=None, ground=None, attention_type="FLASH_ATTENTION"):
    """The call of wan_transformer3d.py:1057-1069 under the pipeline's autocast (pipeline_wan.py:707)."""
    with torch.no_grad(), autocast(), backend(attention_type):
        return blk(inp["x"], e=inp["e"], seq_lens=inp["seq_lens"], grid_sizes=inp["grid_sizes"], freqs=freqs,
                   context=inp["context"], context_lens=None, dtype=torch.bfloat16, t=0,
                   frame_split_indices=None if fs is None else [fs],
                   ground_frame_indices=None if ground is None else [ground])


def block_flops(L, C, Fd, S=512):
    return 8 * L * C * C + 4 * L * L * C + 4 * L * C * C + 4 * S * C * C + 4 * L * S * C + 4 * L * C * Fd


def measure(cfg_kw, lat, fs, device="cuda", iters=3, sdpa=True):
    """-> dict for bench.py's `gpu_reference`: the reference's block, its FA2 self-attention call and one cuBLAS
    Linear at the workload's shapes, CUDA-event timed on this GPU."""
    ns = load()
    C, Fd, n, layers = cfg_kw["dim"], cfg_kw["ffn_dim"], cfg_kw["num_heads"], cfg_kw["num_layers"]
    f, h, w = lat[1], lat[2] // 2, lat[3] // 2
    L = f * h * w
    d = C // n
    dev = torch.device(device)
    blk = make_block(cfg_kw, dev)
    inp = block_inputs(cfg_kw, f, h, w, dev)
    freqs = rope_freqs(d, dev)
    gr = (fs, fs + 1)
    out = {"what": "UNMODIFIED reference modules (baseline/_ref) on this GPU: bf16 weights, autocast(bf16), "
                   "flash-attn 2 varlen (attention_utils.py:113-146), cuBLAS Linears, eager PyTorch elementwise ops",
           "device": torch.cuda.get_device_name(dev), "tokens": L,
           "flash_attn": getattr(sys.modules.get("flash_attn"), "__version__", None), "torch": torch.__version__}
    y = run_block(blk, inp, freqs, fs, gr)
    out["finite"] = bool(torch.isfinite(y).all())
    del y
    ms_block = cuda_time(lambda: run_block(blk, inp, freqs, fs, gr), warmup=1, iters=iters)
    out["block_ms"] = ms_block
    out["block_tflops"] = block_flops(L, C, Fd) / ms_block / 1e9
    out["est_ms_per_step"] = ms_block * layers
    out["est_steps_per_sec"] = 1e3 / (ms_block * layers)
    out["est_note"] = f"one WanAttentionBlock x {layers} layers (embeddings / head / scheduler excluded: favours the reference)"

    # the attention call alone, exactly as WanSelfAttention.forward makes it (:294-299)
    g = torch.Generator(device=dev).manual_seed(3)
    q, k, v = (torch.randn(1, L, n, d, generator=g, device=dev).to(torch.bfloat16) for _ in range(3))
    k_lens = torch.tensor([L], dtype=torch.long)

    def attn(kind):
        with torch.no_grad(), backend(kind):
            return ns.attention_utils.attention(q, k, v, k_lens=k_lens, window_size=(-1, -1))

    ms = cuda_time(lambda: attn("FLASH_ATTENTION"), warmup=1, iters=iters)
    out["self_attention_fa2_ms"] = ms
    out["self_attention_fa2_tflops"] = 4.0 * L * L * C / ms / 1e9
    if sdpa:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ms = cuda_time(lambda: attn("SDPA"), warmup=1, iters=iters)
        out["self_attention_sdpa_ms"] = ms
        out["self_attention_sdpa_tflops"] = 4.0 * L * L * C / ms / 1e9
    del q, k, v

    # one C x C Linear with bias as self_attn.q runs it (:264, cuBLAS under autocast) and the two FFN Linears (:457-459)
    a = torch.randn(L, C, generator=g, device=dev).to(torch.bfloat16)

    def lin(mod):
        with torch.no_grad(), autocast():
            return mod(a)

    ms = cuda_time(lambda: lin(blk.self_attn.q), warmup=2, iters=10)
    out["linear_cxc_ms"] = ms
    out["linear_cxc_tflops"] = 2.0 * L * C * C / ms / 1e9
    ms = cuda_time(lambda: lin(blk.ffn), warmup=2, iters=5)
    out["ffn_ms"] = ms
    out["ffn_tflops"] = 4.0 * L * C * Fd / ms / 1e9
    del a, blk, inp
    torch.cuda.empty_cache()
    return out


def measure_full_step(cfg_kw, lat, fs, device="cuda", iters=1, n_ctx=77):
    """The whole reference WanTransformer3DModel.forward + scheduler step at the workload's size (random-init)."""
    ns = load()
    dev = torch.device(device)
    t0 = time.time()
    with torch.device(dev):
        model = ns.dit.WanTransformer3DModel(model_type="t2v", in_dim=16, out_dim=16, text_len=512, **cfg_kw)
    torch.nn.init.normal_(model.head.head.weight, std=0.02)
    model = model.to(torch.bfloat16).eval().requires_grad_(False)
    model.freqs = model.freqs.to(dev)
    sched = ns.unipc.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, solver_order=2)
    sched.set_timesteps(4, device=dev, shift=3.0)
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(1, *lat, generator=g).to(torch.bfloat16).to(dev)
    ctx = [torch.randn(n_ctx, 4096, generator=g).to(torch.bfloat16).to(dev)]
    L = lat[1] * (lat[2] // 2) * (lat[3] // 2)
    state = {"x": x, "i": 0}

    def step():
        i = state["i"] % 4
        if i == 0:
            sched.set_timesteps(4, device=dev, shift=3.0)
        t = sched.timesteps[i]
        with torch.no_grad(), autocast(), backend("FLASH_ATTENTION"):
            vel = model(x=state["x"], t=t.expand(1), context=ctx, seq_len=L, frame_split_indices=[fs],
                        ground_frame_indices=[(fs, fs + 1)])
        vel[:, :, :fs] = 0
        state["x"] = sched.step(vel, t, state["x"], return_dict=False)[0]
        state["i"] += 1

    init_s = time.time() - t0
    ms = cuda_time(step, warmup=1, iters=iters)
    out = {"full_step_ms": ms, "full_steps_per_sec": 1e3 / ms, "init_s": init_s,
           "what": "reference WanTransformer3DModel.forward (40 layers, FA2) + FlowUniPCMultistepScheduler.step, "
                   "inputs resident", "finite": bool(torch.isfinite(state["x"].float()).all())}
    del model
    torch.cuda.empty_cache()
    return out


def measure_vae(device="cuda", latent_frames=3, H=720, W=1280):
    """Reference AutoencoderKLWan (cuDNN Conv3d, chunked loop with feature caches) decode / encode at 720p."""
    ns = load()
    dev = torch.device(device)
    torch.manual_seed(2)
    vae = ns.vae.AutoencoderKLWan().to(dev, torch.bfloat16).eval().requires_grad_(False)
    g = torch.Generator(device="cpu").manual_seed(4)
    z = torch.randn(1, 16, latent_frames, H // 8, W // 8, generator=g).to(torch.bfloat16).to(dev)
    T = 4 * (latent_frames - 1) + 1
    video = (torch.rand(1, 3, T, H, W, generator=g) * 2 - 1).to(torch.bfloat16).to(dev)
    out = {"latent_frames": latent_frames, "frames": T, "H": H, "W": W}
    with torch.no_grad():
        out["decode_ms"] = cuda_time(lambda: vae.decode(z).sample, warmup=1, iters=2)
        out["encode_ms"] = cuda_time(lambda: vae.encode(video)[0].mode(), warmup=1, iters=2)
    del vae
    torch.cuda.empty_cache()
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-step", action="store_true")
    ap.add_argument("--vae", action="store_true")
    ap.add_argument("--vae-latent-frames", type=int, default=3)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args(argv)
    if not available():
        print(json.dumps({"tool": "gpu_reference", "unavailable": "baseline/_ref not staged (tools/stage_reference.py)"}))
        return 0
    cfg = dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40)
    lat = (16, 21, 90, 160)
    line = {"tool": "gpu_reference", **measure(cfg, lat, 10, iters=a.iters)}
    if a.full_step:
        line["full"] = measure_full_step(cfg, lat, 10)
    if a.vae:
        line["vae"] = measure_vae(latent_frames=a.vae_latent_frames)
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
