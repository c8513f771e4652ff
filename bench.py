#!/usr/bin/env python
"""bench.py — denoising-steps/s of the VideoCoF hot path (Wan-2.1 14B DiT, 81 f x 720p latents).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl vcof|reference] [--workload c2|c3|c1|c5]

One "step" = the pipeline's per-timestep work (reference videox_fun/pipeline/pipeline_wan.py:694-740):
DiT forward on the [src | ground | target] latents with chain-of-frames RoPE, zero the source-frame
velocity, UniPC scheduler step.  Workload C2 (BASELINE.json configs[1]): 14B widths, 40 layers,
latents [16,21,90,160] -> 75,600 tokens, 4-step fast_infer.py schedule, bf16, random-init weights,
synthetic latents/prompt embeddings.

Prints ONE JSON line (rank 0).  `value` = steps/s with inputs resident in HBM; `e2e` = the same
through the public API with HOST (pinned) buffers copied in and the new latents copied out every
step; `roofline` = the dominant kernel (self-attention) timed live with CUDA events inside the
timed region; `cpu_baseline` = the CPU oracle timed on this box's host cores on a bounded sample.
N > 1 (torchrun): sequence-parallel over the ranks, strong scaling.
`--impl reference` times the reference's CPU implementation (oracle port; /root/reference cannot
travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dit config, latent [C,f,h,w], frame_split, n_ctx tokens)
    "c2": (dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40), (16, 21, 90, 160), 10, 77),
    "c1": (dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30), (16, 5, 32, 32), 2, 77),
    # BASELINE.json configs[4]: 321 frames (4x length extrapolation) x 720p, meant for 8 GPUs
    "c5": (dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40), (16, 81, 90, 160), 40, 77),
    # BASELINE.json configs[2]: the inference.py path — guidance 5.0 -> batch 2 per step, 50-step schedule
    "c3": (dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40), (16, 21, 90, 160), 10, 77),
}
EXTRA = {"c3": dict(guidance=5.0, sched_steps=50)}
METRIC = "denoising_steps_per_sec"


def flops_per_forward(cfg, L, S=512):
    C, F, n = cfg["dim"], cfg["ffn_dim"], cfg["num_layers"]
    per_block = 8 * L * C * C + 4 * L * L * C + (4 * L * C * C + 4 * S * C * C) + 4 * L * S * C + 4 * L * C * F
    return per_block * n


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", 1400.0), hbm=d.get("hbm_gbs", 6650.0), src="measured")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port) — bounded sample, extrapolated by token count
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(cfg_kw, L, Ls=1024, Lq=256):
    """Time ONE oracle block of the workload's widths on this box's cores: token-local ops on Ls tokens,
    self-attention of Lq queries against all L keys; scale both linearly to L tokens, x layers."""
    import torch
    from oracle.dit_oracle import DiTConfig, attention_ref, block_forward, make_block_params, rope_table, \
        temporal_positions
    # all the host cores this process may use, whatever OMP_NUM_THREADS says: torch.distributed.run exports
    # OMP_NUM_THREADS=1 to its workers, which made the N > 1 reference arm ten times slower than the N = 1 one
    torch.set_num_threads(host_cores())
    cfg = DiTConfig(**cfg_kw)
    torch.manual_seed(0)
    p = make_block_params(cfg, 0, seed=1)
    C, n, d = cfg.dim, cfg.num_heads, cfg.head_dim
    x = torch.randn(Ls, C)
    e0 = torch.randn(6, C) * 0.1
    ctx = torch.randn(cfg.text_len, C)
    f, h, w = 1, 1, Ls
    t0 = time.perf_counter()
    with torch.no_grad():
        # token-local part (+ an Ls x Ls attention that is subtracted analytically below)
        block_forward(p, 0, x, e0, ctx, cfg, (f, h, w), rope_table(d), temporal_positions(f), Ls)
    t_block = time.perf_counter() - t0
    q = torch.randn(Lq, n, d)
    k = torch.randn(L, n, d)
    v = torch.randn(L, n, d)
    t0 = time.perf_counter()
    with torch.no_grad():
        for h0 in range(0, n, 8):   # 8 heads at a time bounds the score matrix
            attention_ref(q[:, h0:h0 + 8], k[:, h0:h0 + 8], v[:, h0:h0 + 8])
    t_attn = time.perf_counter() - t0
    # attention inside the block sample cost ~ t_attn * (Ls*Ls)/(Lq*L); remove it from the linear part
    t_lin = max(t_block - t_attn * (Ls * Ls) / (Lq * L), 0.0)
    per_block = t_lin * (L / Ls) + t_attn * (L / Lq)
    step_s = per_block * cfg.num_layers
    return dict(step_s=step_s, t_block=t_block, t_attn=t_attn, cores=torch.get_num_threads(),
                sample=f"oracle block ({C}/{cfg.ffn_dim}/{n} heads): token-local ops on {Ls} tokens + "
                       f"self-attention of {Lq} queries x {L} keys, scaled linearly to {L} tokens x "
                       f"{cfg.num_layers} layers (extrapolated)")


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port) on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg_kw, lat, fs, n_ctx = WORKLOADS[args.workload]
    L = lat[1] * (lat[2] // 2) * (lat[3] // 2)
    for _ in range(args.warmup):
        cpu_reference_sample(cfg_kw, L, Ls=256, Lq=64)
    vals, last = [], None
    for _ in range(args.steps):
        last = cpu_reference_sample(cfg_kw, L)
        vals.append(last["step_s"])
    step_s = sum(vals) / len(vals)
    v = 1.0 / step_s
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, args.gpus),     # the libvcof arm's config, verbatim (tier rule 4)
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": last["cores"], "kind": "port",
                             "sample": last["sample"]},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def l2_note(cfg_kw, rows):
    """Timing rule: inputs larger than L2 or a flush between iterations — say which.  Per layer and GPU the step streams
    the fp32 residual, four bf16 [rows, C] activations, the bf16 FFN hidden and the layer's weights."""
    C, F = cfg_kw["dim"], cfg_kw["ffn_dim"]
    act = rows * (C * 4 + 4 * C * 2 + F * 2)
    wts = (8 * C * C + 2 * C * F) * 2
    if act + wts > 4 * 126e6:
        return (f"working set per layer ({act / 1e9:.2f} GB activations + {wts / 1e9:.2f} GB weights) >> 126 MB L2; "
                "no flush needed")
    return (f"working set per layer ({(act + wts) / 1e6:.0f} MB) is not >> the 126 MB L2 and no flush is done: "
            "debug workload, not a valid bench configuration")


def workload_config(name, n_gpus):
    cfg_kw, lat, fs, n_ctx = WORKLOADS[name]
    ex = EXTRA.get(name, {})
    L = lat[1] * (lat[2] // 2) * (lat[3] // 2)
    batch = 2 if ex.get("guidance", 1.0) > 1.0 else 1
    return {"workload": f"{name}: Wan-2.1 DiT dim={cfg_kw['dim']} ffn={cfg_kw['ffn_dim']} heads={cfg_kw['num_heads']} "
                        f"layers={cfg_kw['num_layers']}, latents {list(lat)} -> {L} tokens, chain-of-frames "
                        f"{fs}|1|{lat[1] - fs - 1}, {ex.get('sched_steps', 4)}-step UniPC schedule (shift 3), "
                        + (f"guidance {ex['guidance']} -> batch 2 (uncond + cond) per step" if batch == 2 else "batch 1"),
            "tokens": L, "parallelism": f"sp{n_gpus}" if n_gpus > 1 else "single",
            "l2": l2_note(cfg_kw, L // n_gpus)}


def sha256_of(t):
    """Checksum of a tensor's bytes (bf16 viewed as int16): equal across N when the sharded forward reproduces the
    single-GPU bits, so every scaling record carries its own parity signal."""
    import hashlib
    import torch
    return hashlib.sha256(t.detach().contiguous().view(torch.int16).cpu().numpy().tobytes()).hexdigest()


class StepRunner:
    """The per-timestep work of pipeline_wan.py:694-740 on one workload: DiT forward (batch 2 with classifier-free
    guidance), source-frame velocity zeroed, UniPC step.  Host copies of the synthetic inputs are pinned."""

    def __init__(self, torch, model, name, dev):
        from videocof_b200.scheduler import FlowUniPCMultistepScheduler
        self.torch, self.model, self.dev = torch, model, dev
        cfg_kw, lat, fs, n_ctx = WORKLOADS[name]
        ex = EXTRA.get(name, {})
        self.fs, self.lat = fs, lat
        self.guidance = ex.get("guidance", 1.0)
        self.n_sched = ex.get("sched_steps", 4)
        self.L = lat[1] * (lat[2] // 2) * (lat[3] // 2)
        self.sched = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, solver_order=2)
        g = torch.Generator(device="cpu").manual_seed(1)
        self.lat_host = torch.randn(1, *lat, generator=g).to(torch.bfloat16).pin_memory()
        self.ctx_host = torch.randn(n_ctx, 4096, generator=g).to(torch.bfloat16).pin_memory()
        self.neg_host = torch.randn(n_ctx // 2, 4096, generator=g).to(torch.bfloat16).pin_memory()
        self.B = 2 if self.guidance > 1.0 else 1
        self.kw = dict(seq_len=self.L, frame_split_indices=[fs] * self.B,
                       ground_frame_indices=[(fs, fs + 1)] * self.B)
        self.reset()

    def contexts(self, non_blocking=False):
        c = [self.ctx_host.to(self.dev, non_blocking=non_blocking)]
        if self.B == 2:      # pipeline_wan.py:606: in_prompt_embeds = negative + positive
            c = [self.neg_host.to(self.dev, non_blocking=non_blocking)] + c
        return c

    def reset(self):
        self.i = 0
        self.latents = self.lat_host.to(self.dev)
        self.ctx_dev = self.contexts()
        self.sched.set_timesteps(self.n_sched, device=self.dev, shift=3.0)

    def one_step(self, latents, ctx):
        torch = self.torch
        if self.i % self.n_sched == 0:
            self.sched.set_timesteps(self.n_sched, device=self.dev, shift=3.0)
        t = self.sched.timesteps[self.i % self.n_sched]
        x = torch.cat([latents] * 2) if self.B == 2 else latents                 # (:700)
        with torch.no_grad():
            v = self.model(x=x, t=t.expand(self.B), context=ctx, **self.kw)
        if self.B == 2:                                                            # (:731-733)
            vu, vt = v.chunk(2)
            v = vu + self.guidance * (vt - vu)
        v[:, :, :self.fs] = 0                                                      # (:736)
        self.i += 1
        return self.sched.step(v, t, latents, return_dict=False)[0]

    def resident_step(self):
        self.latents = self.one_step(self.latents, self.ctx_dev)


# ------------------------------------------------------------------------------------------------
# the libvcof arm
# ------------------------------------------------------------------------------------------------
def run_vcof(args):
    import torch
    import torch.distributed as dist
    from videocof_b200 import ops
    from videocof_b200.dit import WanTransformer3DModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg_kw, lat, fs, n_ctx = WORKLOADS[args.workload]
    model = WanTransformer3DModel.random_init(device=dev, seed=0, **cfg_kw)
    if world > 1:
        model.enable_multi_gpus_inference()
    run = StepRunner(torch, model, args.workload, dev)
    L, sched = run.L, run.sched
    lat_host, ctx_host = run.lat_host, run.ctx_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        run.resident_step()
    run.reset()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ops.reset_launches()
    ops.enable_timing()
    if args.profile_range:
        torch.cuda.profiler.start()      # ncu --profile-from-start off captures only the timed steps
    total_ms = timed(args.steps, run.resident_step)
    if args.profile_range:
        torch.cuda.profiler.stop()
    timing = ops.collect_timing()
    launches = ops.launches()
    clk = clocks.stop() if rank == 0 else None
    latents_sha = sha256_of(run.latents)        # after exactly `steps` steps from the seeded start: same at every N

    # ---- end to end through the public API with host buffers
    out_host = torch.empty_like(lat_host)

    def e2e_step():
        x = lat_host.to(dev, non_blocking=True)
        c = run.contexts(non_blocking=True)
        y = run.one_step(x, c)
        out_host.copy_(y, non_blocking=True)

    run.reset()
    e2e_ms = timed(args.steps, e2e_step)

    # ---- whole pipeline through WanPipeline.__call__ (VAE encode -> 4 steps -> VAE decode x2), host in/out: the
    #      frames/s half of BASELINE.json's metric, at every N (DiT token-sharded, VAE frame-sharded with halos)
    pipe_stats = None
    if not args.no_pipeline and args.workload == "c2":
        from videocof_b200.pipeline import WanPipeline
        from videocof_b200.vae import AutoencoderKLWan
        torch.manual_seed(2)
        vae = AutoencoderKLWan().to(dev, torch.bfloat16).eval()
        pipe = WanPipeline(None, None, vae, model, sched)
        if world > 1:
            vae.enable_temporal_sharding()          # the DiT is already sequence-parallel (see above)
        src_frames = 4 * fs - 3                                   # fs latent frames of source video
        H, W = lat[2] * 8, lat[3] * 8
        g = torch.Generator(device="cpu").manual_seed(3)
        if args.pipeline_bytes:     # byte frames in and out (videocof_b200/video_io.py)
            video_host = torch.randint(0, 256, (1, src_frames, H, W, 3), generator=g, dtype=torch.uint8).pin_memory()
        else:
            video_host = (torch.rand(1, 3, src_frames, H, W, generator=g) * 2 - 1).to(torch.bfloat16).pin_memory()
        t_axis = 1 if args.pipeline_bytes else 2

        def run_pipe():
            return pipe(video=video_host, prompt_embeds=[ctx_host.to(dev)], height=H, width=W,
                        source_frames=src_frames, reasoning_frames=4, num_inference_steps=4, guidance_scale=1.0,
                        shift=3, repeat_rope=True, cot=True, generator=torch.Generator(device="cpu").manual_seed(4),
                        output_type="uint8" if args.pipeline_bytes else "numpy")

        run_pipe()                                                # warm-up (allocator, weight packs)
        barrier()
        t0 = time.perf_counter()
        out = run_pipe()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        n_edit = int(out.edit_videos.shape[t_axis])
        import hashlib
        import numpy as np
        frames = np.ascontiguousarray(np.asarray(out.edit_videos))
        pipe_stats = {"seconds": dt, "frames_per_sec": n_edit / dt, "edit_frames": n_edit,
                      "ground_frames": int(out.ground_videos.shape[t_axis]), "source_frames": src_frames,
                      "edit_frames_sha256": hashlib.sha256(frames.tobytes()).hexdigest(),
                      "what": ("WanPipeline.__call__: VAE encode(source, host uint8 frames) + 4 DiT steps + VAE "
                               "decode(ground) + VAE decode(edit) -> uint8 frames on the host; random-init weights"
                               if args.pipeline_bytes else
                               "WanPipeline.__call__: VAE encode(source, host bf16) + 4 DiT steps + VAE decode(ground) + "
                               "VAE decode(edit) -> fp32 numpy frames on the host; random-init weights")
                              + ("; DiT token-sharded, VAE frame-sharded with halos" if world > 1 else "")}
        del vae, pipe, out, frames
        torch.cuda.empty_cache()

    # ---- BASELINE.json configs[4] (321 frames x 720p, 8 GPUs): one timed step of the same model on the long clip
    c5_stats = None
    if args.workload == "c2" and (args.c5 or (world == 8 and not args.no_c5)):
        run5 = StepRunner(torch, model, "c5", dev)
        run5.resident_step()
        run5.reset()
        ms5 = timed(2, run5.resident_step)
        c5_stats = {"ms_per_step": ms5 / 2, "steps_per_sec": 2e3 / ms5, "steps": 2, "warmup": 1,
                    "latents_sha256": sha256_of(run5.latents), "config": workload_config("c5", world)}
        del run5
        torch.cuda.empty_cache()

    # ---- the reference's own CUDA path on this GPU (tools/gpu_reference.py; N = 1 only)
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_gpu_reference and args.workload in ("c2", "c3"):
        gpu_ref = gpu_reference_leg(torch, model, cfg_kw, lat, fs, timing, total_ms / args.steps / run.B)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ms_per_step = total_ms / args.steps
    value = 1000.0 / ms_per_step
    peaks = measured_peaks()
    # dominant kernel: self-attention (72 % of the algorithmic FLOPs at C2)
    rows = L // world
    attn_key = [k for k in timing if k.startswith("attn") and f"Lk={L} " in k]
    roof = None
    if attn_key:
        # every self-attention launch of the timed region (the head exchange splits a layer's heads over two launches):
        # algorithmic FLOPs 4 * Lq * Lk * heads * 128 per launch, summed, over the summed launch time
        import re
        n = ms = flops_total = 0.0
        for k in attn_key:
            m = re.match(r"attn Lq=(\d+) Lk=(\d+) heads=(\d+)", k)
            kn, kms = timing[k]
            n, ms = n + kn, ms + kms
            flops_total += kn * 4.0 * int(m.group(1)) * int(m.group(2)) * int(m.group(3)) * 128
        ach = flops_total / (ms * 1e-3) / 1e12
        roof = {"kernel": "attn_fwd_kernel (self-attention)", "bound": "tensor", "achieved": ach,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"],
                "peak_source": f"{peaks['src']} bf16_tflops_sustained", "traffic": None,
                "flops_per_launch": flops_total / n, "avg_launch_ms": ms / n, "launches_timed": int(n)}
        if len(attn_key) == 1:
            roof.update(ncu_traffic(attn_key[0]))
    gpu_ms = sum(ms for _, ms in timing.values())
    breakdown = sorted(((k, n, ms) for k, (n, ms) in timing.items()), key=lambda r: -r[2])[:8]
    fl = flops_per_forward(cfg_kw, L) * run.B
    line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args.workload, world),
            "clocks": clk, "gpu_launches": launches, "latents_sha256": latents_sha,
            "e2e": {"value": 1000.0 / (e2e_ms / args.steps), "unit": "steps/s",
                    "h2d_bytes_per_step": lat_host.numel() * 2 + sum(c.numel() * 2 for c in run.ctx_dev),
                    "d2h_bytes_per_step": out_host.numel() * 2},
            "roofline": roof,
            "model_tflops": fl / (ms_per_step * 1e-3) / 1e12 / world,
            "model_frac_of_peak": fl / (ms_per_step * 1e-3) / 1e12 / world / peaks["tflops"],
            "pipeline": pipe_stats,
            "frames_per_sec": pipe_stats["frames_per_sec"] if pipe_stats else None,
            "kernel_ms_per_step": gpu_ms / args.steps,
            "top_kernels": [{"key": k, "launches": n, "ms": round(ms, 3)} for k, n, ms in breakdown]}
    if c5_stats:
        line["c5"] = c5_stats
    if gpu_ref:
        line["gpu_reference"] = gpu_ref
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_sample(cfg_kw, L)
        line["cpu_baseline"] = {"value": 1.0 / r["step_s"] / run.B, "unit": "steps/s", "cores": r["cores"],
                                "kind": "port", "sample": r["sample"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def gpu_reference_leg(torch, model, cfg_kw, lat, fs, timing, ours_forward_ms):
    """`gpu_reference`: the UNMODIFIED reference modules (baseline/_ref) on this same GPU — one WanAttentionBlock of the
    workload's widths at the workload's token count, its flash-attn 2 self-attention call and its cuBLAS Linears,
    CUDA-event timed right after the libvcof arm — and libvcof's speed-up per op and per block."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import gpu_reference as gr
    except Exception as exc:       # noqa: BLE001
        return {"unavailable": f"tools/gpu_reference.py failed to import: {exc!r}"}
    if not gr.available():
        return {"unavailable": "baseline/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)"}
    torch.cuda.empty_cache()
    clocks = ClockSampler(torch.cuda.current_device())
    clocks.start()
    try:
        ref = gr.measure(cfg_kw, lat, fs, device=f"cuda:{torch.cuda.current_device()}")
    except Exception as exc:       # noqa: BLE001
        clocks.stop()
        return {"unavailable": f"reference CUDA path failed: {exc!r}"}
    ref["clocks"] = clocks.stop()
    L = ref["tokens"]
    layers = cfg_kw["num_layers"]

    def ours(prefix):
        hit = [(n, ms) for k, (n, ms) in timing.items() if k.startswith(prefix)]
        return hit[0][1] / hit[0][0] if hit else None
    o_attn = ours(f"attn Lq={L} Lk={L} ")
    o_lin = ours(f"gemm[bias] M={L} N={cfg_kw['dim']} K={cfg_kw['dim']}")
    o_f1 = ours(f"gemm[bias_gelu] M={L} N={cfg_kw['ffn_dim']}")
    o_f2 = ours(f"gemm[bias_gate_res] M={L} N={cfg_kw['dim']} K={cfg_kw['ffn_dim']}")
    ref["libvcof"] = {"block_ms": ours_forward_ms / layers,
                      "block_note": "whole forward + scheduler step / layers (embeddings, head and scheduler included: "
                                    "favours the reference)",
                      "self_attention_ms": o_attn, "linear_cxc_ms": o_lin,
                      "ffn_ms": (o_f1 + o_f2) if o_f1 and o_f2 else None}
    ref["speedup"] = {"block": ref["block_ms"] / (ours_forward_ms / layers),
                      "self_attention_vs_fa2": ref["self_attention_fa2_ms"] / o_attn if o_attn else None,
                      "linear_cxc_vs_cublas": ref["linear_cxc_ms"] / o_lin if o_lin else None,
                      "ffn_vs_cublas": ref["ffn_ms"] / (o_f1 + o_f2) if o_f1 and o_f2 else None}
    return ref


def ncu_traffic(attn_key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the same launch
    shape (profiles/r2_attn_c2_ncu.json: dram__bytes_read.sum + dram__bytes_write.sum); null for other shapes."""
    import re
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_attn_c2_ncu.json")
    m = re.match(r"attn Lq=(\d+) Lk=(\d+) heads=(\d+)", attn_key)
    if not m or not os.path.exists(path):
        return {}
    d = json.load(open(path))
    sh = d["shape"]
    if (int(m.group(1)), int(m.group(2)), int(m.group(3))) != (sh["Lq"], sh["Lk"], sh["heads"]):
        return {}
    return {"traffic": d["dram_bytes_per_launch"], "traffic_unit": "bytes/launch",
            "traffic_algorithmic": d["algorithmic_bytes"], "traffic_source": "profiles/r2_attn_c2_ncu.json"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vcof", choices=["vcof", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (use with ncu --profile-from-start off)")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the WanPipeline (VAE + 4 steps) end-to-end leg")
    ap.add_argument("--pipeline", action="store_true", help="(kept for old command lines: the leg now runs at every N)")
    ap.add_argument("--c5", action="store_true", help="also time one step of the 321-frame workload (default at N = 8)")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip the reference's own CUDA path (baseline/_ref: flash-attn 2 + cuBLAS) timed beside libvcof")
    ap.add_argument("--pipeline-bytes", action="store_true",
                    help="WanPipeline leg with uint8 frames in and out (conversions on the device) instead of bf16 in / "
                         "fp32 out")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_vcof(args)


if __name__ == "__main__":
    sys.exit(main())
