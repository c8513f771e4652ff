"""Host-side plumbing of videocof_b200.pipeline.WanPipeline on CPU with stub DiT / VAE modules: chain-of-frames
latent assembly, frame-split kwargs, CFG batching, frozen source frames, split ground/edit decode
(reference videox_fun/pipeline/pipeline_wan.py:381-428, 592-799)."""
import pytest
import torch

from videocof_b200.pipeline import WanPipeline, randn_tensor
from videocof_b200.scheduler import FlowUniPCMultistepScheduler


class _Cfg(dict):
    __getattr__ = dict.get


class StubVAE:
    temporal_compression_ratio = 4
    spatial_compression_ratio = 8
    latent_channels = 16
    dtype = torch.float32

    class _D:
        def __init__(self, m):
            self.m = m

        def mode(self):
            return self.m

    def encode(self, x):
        b, _, t, h, w = x.shape
        f = (t - 1) // 4 + 1
        return (self._D(torch.full((b, 16, f, h // 8, w // 8), 0.5)),)

    def decode(self, z):
        b, _, f, h, w = z.shape
        out = z.mean(dim=1, keepdim=True).repeat_interleave(4, dim=2)[:, :, :4 * (f - 1) + 1]
        out = out.repeat(1, 3, 1, 1, 1).repeat_interleave(8, 3).repeat_interleave(8, 4)

        class O:
            sample = out
        return O


class StubDiT:
    config = _Cfg(in_channels=16, patch_size=(1, 2, 2))
    dtype = torch.float32
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []
        self.num_inference_steps = None
        self.current_steps = 0

    def __call__(self, x, t, context, seq_len, frame_split_indices=None, ground_frame_indices=None):
        self.calls.append(dict(shape=tuple(x.shape), t=t.clone(), n_ctx=len(context), seq_len=seq_len,
                               fsi=frame_split_indices, gfi=ground_frame_indices, step=self.current_steps))
        return torch.ones_like(x) * (1 + torch.arange(x.shape[0]).view(-1, 1, 1, 1, 1))


def make(guidance):
    dit, vae = StubDiT(), StubVAE()
    pipe = WanPipeline(None, None, vae, dit, FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1))
    video = torch.zeros(1, 3, 9, 32, 48)                       # 9 source frames -> 3 latent frames
    emb = [torch.randn(5, 8)]
    out = pipe(video=video, prompt_embeds=emb, negative_prompt_embeds=[torch.randn(3, 8)] if guidance > 1 else None,
               height=32, width=48, source_frames=9, reasoning_frames=4, num_inference_steps=4,
               guidance_scale=guidance, shift=3, repeat_rope=True, cot=True, generator=torch.Generator().manual_seed(0))
    return dit, out


def test_cot_layout_and_kwargs():
    dit, out = make(1.0)
    assert len(dit.calls) == 4
    c = dit.calls[0]
    assert c["shape"] == (1, 16, 3 + 1 + 3, 4, 6)             # [src 3 | ground 1 | target 3]
    assert c["fsi"] == [3] and c["gfi"] == [(3, 4)]
    assert c["seq_len"] == 7 * 2 * 3
    assert [int(k["t"][0]) for k in dit.calls] == [999, 899, 749, 499]
    assert [k["step"] for k in dit.calls] == [0, 1, 2, 3]
    # ground segment: 1 latent -> 1 frame; edit segment: 3 latents -> 9 frames; concatenated along time
    assert tuple(out.ground_videos.shape) == (1, 3, 1, 32, 48)
    assert tuple(out.edit_videos.shape) == (1, 3, 9, 32, 48)
    assert tuple(out.videos.shape) == (1, 3, 10, 32, 48)
    assert float(out.videos.min()) >= 0.0 and float(out.videos.max()) <= 1.0
    # one host copy of the clip: the two segments are views of `videos`, in the reference's order (ground first)
    assert torch.equal(out.videos[:, :, :1], out.ground_videos) and torch.equal(out.videos[:, :, 1:], out.edit_videos)
    import numpy as np
    assert np.shares_memory(out.videos.numpy(), out.ground_videos.numpy())
    assert np.shares_memory(out.videos.numpy(), out.edit_videos.numpy())


def test_cfg_batches_two_and_combines():
    dit, _ = make(5.0)
    assert dit.calls[0]["shape"][0] == 2 and dit.calls[0]["n_ctx"] == 2
    assert dit.calls[0]["fsi"] == [3, 3]


def test_source_frames_keep_their_latents():
    """noise_pred[:, :, :condition_count] = 0 (:736): UniPC with zero velocity leaves x unchanged."""
    dit, vae = StubDiT(), StubVAE()
    sched = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1)
    pipe = WanPipeline(None, None, vae, dit, sched)
    seen = {}

    def cb(p, i, t, kw):
        seen[i] = kw["latents"].clone()
        return {}
    pipe(video=torch.zeros(1, 3, 5, 16, 16), prompt_embeds=[torch.randn(2, 8)], height=16, width=16, source_frames=5,
         reasoning_frames=4, num_inference_steps=4, guidance_scale=1.0, shift=3, cot=True, callback_on_step_end=cb,
         generator=torch.Generator().manual_seed(1))
    for i in range(4):
        assert torch.allclose(seen[i][:, :, :2], torch.full_like(seen[i][:, :, :2], 0.5))   # encoded source = 0.5
    assert not torch.allclose(seen[3][:, :, 2:], seen[0][:, :, 2:])


def test_randn_tensor_cpu_generator_is_device_independent():
    a = randn_tensor((2, 3), generator=torch.Generator().manual_seed(3), device="cpu", dtype=torch.float32)
    b = torch.randn((2, 3), generator=torch.Generator().manual_seed(3))
    assert torch.equal(a, b)


def test_context_cache_is_scoped_to_the_denoising_loop(monkeypatch):
    """`__call__` turns the DiT's step-invariant context cache on for its loop and frees it afterwards — also when a
    step raises — leaves a cache the caller enabled alone, and VCOF_CONTEXT_CACHE=0 keeps it off."""
    class CachingDiT(StubDiT):
        def __init__(self, fail_at=None):
            super().__init__()
            self._ctx_cache, self.seen, self.events, self.fail_at = None, [], [], fail_at

        def enable_context_cache(self, max_entries=4):
            self.events.append("on")
            self._ctx_cache = {"max": max_entries, "entries": []}

        def disable_context_cache(self):
            self.events.append("off")
            self._ctx_cache = None

        def __call__(self, x, **kw):
            self.seen.append(self._ctx_cache is not None)
            if self.fail_at == len(self.seen):
                raise RuntimeError("step failed")
            return super().__call__(x, **kw)

    def run(dit):
        pipe = WanPipeline(None, None, StubVAE(), dit, FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1))
        return pipe(video=torch.zeros(1, 3, 5, 16, 16), prompt_embeds=[torch.randn(2, 8)], height=16, width=16,
                    source_frames=5, reasoning_frames=4, num_inference_steps=4, guidance_scale=1.0, shift=3, cot=True,
                    generator=torch.Generator().manual_seed(1))

    dit = CachingDiT()
    run(dit)
    assert dit.seen == [True] * 4 and dit.events == ["on", "off"] and dit._ctx_cache is None
    dit = CachingDiT(fail_at=2)
    with pytest.raises(RuntimeError, match="step failed"):
        run(dit)
    assert dit.events == ["on", "off"] and dit._ctx_cache is None
    dit = CachingDiT()
    dit.enable_context_cache()
    run(dit)
    assert dit.events == ["on"] and dit._ctx_cache is not None          # the caller's cache is the caller's to free
    monkeypatch.setenv("VCOF_CONTEXT_CACHE", "0")
    dit = CachingDiT()
    run(dit)
    assert dit.seen == [False] * 4 and dit.events == []
    run(StubDiT())                                                       # a DiT without the cache surface: untouched


@pytest.mark.parametrize("guidance", [1.0, 5.0])
def test_pipeline_with_the_real_dit_is_identical_with_and_without_the_context_cache(guidance, monkeypatch):
    """The whole denoising loop through videocof_b200.dit (libvcof entry points replaced by their contract statements,
    tests/vcof_emulator.py) with the pipeline's scoped context cache on (default) and off: same videos, bit for bit,
    with and without classifier-free guidance (batched CFG pair)."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import vcof_emulator
    from gen_golden import DIT_CASES
    from oracle.dit_oracle import DiTConfig, make_dit_params
    from videocof_b200.dit import WanTransformer3DModel
    vcof_emulator.install_dit(monkeypatch)
    ckw, _, _, _ = DIT_CASES["dit_tiny"]
    cfg = DiTConfig(**ckw)
    dit = WanTransformer3DModel(**cfg.to_kwargs())
    dit.load_state_dict(make_dit_params(cfg, seed=11), strict=True)
    dit = dit.to(torch.bfloat16).eval()
    vae = StubVAE()
    vae.dtype = torch.bfloat16
    g = torch.Generator().manual_seed(5)
    emb, neg = [torch.randn(6, cfg.text_dim, generator=g).bfloat16()], [torch.randn(4, cfg.text_dim, generator=g).bfloat16()]

    def run():
        pipe = WanPipeline(None, None, vae, dit, FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1))
        return pipe(video=torch.zeros(1, 3, 5, 16, 16), prompt_embeds=emb, negative_prompt_embeds=neg if guidance > 1 else None,
                    height=16, width=16, source_frames=5, reasoning_frames=4, num_inference_steps=3,
                    guidance_scale=guidance, shift=3, repeat_rope=True, cot=True,
                    generator=torch.Generator().manual_seed(1)).videos

    cached = run()
    assert dit._ctx_cache is None
    monkeypatch.setenv("VCOF_CONTEXT_CACHE", "0")
    plain = run()
    assert torch.equal(cached, plain) and float(cached.std()) > 0
