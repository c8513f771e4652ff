"""WanPipeline — the caller of the hot path, with the reference's constructor and __call__ contract.

Mirrors videox_fun/pipeline/pipeline_wan.py (reference) without diffusers: prompt encoding
(:139-260), chain-of-frames latent assembly (:343-419), the denoise loop with CFG, frozen
source frames and the scheduler step (:592-755), and the split ground/edit decode (:757-799).
The two heavy callees are the libvcof-backed models (videocof_b200.dit / videocof_b200.vae);
everything in this file is host-side plumbing on small tensors.
"""
import math
import os
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np
import torch

from .scheduler import FlowUniPCMultistepScheduler


@dataclass
class WanPipelineOutput:
    """reference :92-105."""
    videos: Any
    ground_videos: Optional[Any] = None
    edit_videos: Optional[Any] = None


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor semantics: draw on the generator's device (CPU generators
    give device-independent noise), then move."""
    gen = generator[0] if isinstance(generator, (list, tuple)) else generator
    rand_device = device
    if gen is not None and gen.device.type != torch.device(device).type and gen.device.type == "cpu":
        rand_device = "cpu"
    return torch.randn(shape, generator=gen, device=rand_device, dtype=dtype).to(device)


class WanPipeline:
    def __init__(self, tokenizer, text_encoder, vae, transformer, scheduler):
        self.tokenizer, self.text_encoder, self.vae = tokenizer, text_encoder, vae
        self.transformer, self.scheduler = transformer, scheduler
        self._guidance_scale = 1.0
        self._interrupt = False
        self._num_timesteps = 0
        self._device = None

    # ---- diffusers.DiffusionPipeline surface used by the CLIs -------------------------------------------
    def to(self, device=None, dtype=None):
        for m in (self.text_encoder, self.vae, self.transformer):
            if m is not None and hasattr(m, "to"):
                m.to(device) if dtype is None else m.to(device, dtype)
        self._device = torch.device(device) if device is not None else self._device
        return self

    @property
    def _execution_device(self):
        return self._device if self._device is not None else self.transformer.device

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def interrupt(self):
        return self._interrupt

    def maybe_free_model_hooks(self):
        pass

    def enable_multi_gpus(self, group=None):
        """One call for a multi-GPU box (one process per GPU, torch.distributed initialised): token-sharded DiT with a
        K/V all-gather per layer (videocof_b200.dist) and frame-sharded VAE with 2-frame halos (vae.TimeShard).  Every
        rank must call the pipeline with the same inputs / seeds and receives the full result."""
        self.transformer.enable_multi_gpus_inference(group)
        self.vae.enable_temporal_sharding(group)
        return self

    def enable_model_cpu_offload(self, gpu_id=None, device="cuda", **_):
        """The CLIs default to offload modes sized for 24-80 GB GPUs (fast_infer.py:136, :348-361).  A fused path
        that reads weight pointers directly cannot be paged by accelerate hooks, and a B200 holds the whole
        pipeline (28 GB DiT + 11 GB T5 + 0.25 GB VAE of 180 GB): keep everything resident instead."""
        print("[videocof_b200] offload request ignored: keeping DiT / VAE / text encoder resident on", device)
        return self.to(device)

    enable_sequential_cpu_offload = enable_model_cpu_offload

    # ---- prompt encoding (:139-260) -----------------------------------------------------------------------
    def _get_t5_prompt_embeds(self, prompt, num_videos_per_prompt=1, max_sequence_length=512, device=None, dtype=None):
        device = device or self._execution_device
        dtype = dtype or self.text_encoder.dtype
        prompt = [prompt] if isinstance(prompt, str) else prompt
        ti = self.tokenizer(prompt, padding="max_length", max_length=max_sequence_length, truncation=True,
                            add_special_tokens=True, return_tensors="pt")
        ids, mask = ti.input_ids, ti.attention_mask
        seq_lens = mask.gt(0).sum(dim=1).long()
        emb = self.text_encoder(ids.to(device), attention_mask=mask.to(device))[0].to(dtype=dtype, device=device)
        _, seq_len, _ = emb.shape
        emb = emb.repeat(1, num_videos_per_prompt, 1).view(len(prompt) * num_videos_per_prompt, seq_len, -1)
        return [u[:v] for u, v in zip(emb, seq_lens)]

    def encode_prompt(self, prompt, negative_prompt=None, do_classifier_free_guidance=True, num_videos_per_prompt=1,
                      prompt_embeds=None, negative_prompt_embeds=None, max_sequence_length=512, device=None, dtype=None):
        device = device or self._execution_device
        prompt = [prompt] if isinstance(prompt, str) else prompt
        batch_size = len(prompt) if prompt is not None else len(prompt_embeds)
        if prompt_embeds is None:
            prompt_embeds = self._get_t5_prompt_embeds(prompt, num_videos_per_prompt, max_sequence_length, device, dtype)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            negative_prompt = negative_prompt or ""
            negative_prompt = batch_size * [negative_prompt] if isinstance(negative_prompt, str) else negative_prompt
            if prompt is not None and type(prompt) is not type(negative_prompt):
                raise TypeError("`negative_prompt` should be the same type to `prompt`")
            if batch_size != len(negative_prompt):
                raise ValueError("`negative_prompt` batch size does not match `prompt`")
            negative_prompt_embeds = self._get_t5_prompt_embeds(negative_prompt, num_videos_per_prompt,
                                                                max_sequence_length, device, dtype)
        return prompt_embeds, negative_prompt_embeds

    _callback_tensor_inputs = ["latents", "prompt_embeds", "negative_prompt_embeds"]      # (:119-123)

    def check_inputs(self, prompt, height, width, negative_prompt, callback_on_step_end_tensor_inputs,
                     prompt_embeds=None, negative_prompt_embeds=None):
        """:449-498 — same positional order, same conditions, ValueError as there."""
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_on_step_end_tensor_inputs is not None and not all(
                k in self._callback_tensor_inputs for k in callback_on_step_end_tensor_inputs):
            bad = [k for k in callback_on_step_end_tensor_inputs if k not in self._callback_tensor_inputs]
            raise ValueError(f"`callback_on_step_end_tensor_inputs` has to be in {self._callback_tensor_inputs}, "
                             f"but found {bad}")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError("Cannot forward both `prompt` and `prompt_embeds`. Please make sure to only forward one "
                             "of the two.")
        elif prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and "
                             "`prompt_embeds` undefined.")
        elif prompt is not None and not isinstance(prompt, (str, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if prompt is not None and negative_prompt_embeds is not None:
            raise ValueError("Cannot forward both `prompt` and `negative_prompt_embeds`. Please make sure to only "
                             "forward one of the two.")
        if negative_prompt is not None and negative_prompt_embeds is not None:
            raise ValueError("Cannot forward both `negative_prompt` and `negative_prompt_embeds`. Please make sure to "
                             "only forward one of the two.")
        if hasattr(prompt_embeds, "shape") and hasattr(negative_prompt_embeds, "shape"):
            # (:492-498) tensors only: the pipeline's own form is a list of per-sample [tokens, 4096] tensors, whose
            # token counts may differ between the prompt and the negative prompt
            if prompt_embeds.shape != negative_prompt_embeds.shape:
                raise ValueError("`prompt_embeds` and `negative_prompt_embeds` must have the same shape when passed "
                                 f"directly, but got: `prompt_embeds` {prompt_embeds.shape} != "
                                 f"`negative_prompt_embeds` {negative_prompt_embeds.shape}.")

    # ---- latents (:343-419) ------------------------------------------------------------------------------------
    def _encode_source(self, video, dtype, device):
        if video.dtype == torch.uint8:
            # byte frames [B, T, H, W, 3] (videocof_b200.video_io.load_video_frames(..., as_uint8=True)): one quarter of
            # the fp32 bytes go over PCIe; the [-1, 1] scaling and the cast to bf16 run inside the VAE's layout kernel
            video = video.to(device=device, non_blocking=True)
            lat = [self.vae.encode_frames(video[i:i + 1])[0].mode() for i in range(video.shape[0])]
            return torch.cat(lat, dim=0).to(dtype)
        video = video.to(device=device, dtype=dtype)
        lat = [self.vae.encode(video[i:i + 1])[0].mode() for i in range(video.shape[0])]
        return torch.cat(lat, dim=0)

    def prepare_cot_video_latents(self, video, reasoning_latent_count=1, batch_size=1, num_channels_latents=16,
                                  height=480, width=832, dtype=torch.float32, device=None, generator=None,
                                  condition_count=None, latents=None, timestep=None):
        """[ src | noise(ground + target) ] along latent time."""
        if latents is not None:
            return latents.to(device=device, dtype=dtype)
        org = self._encode_source(video, dtype, device)
        b, c, f, h, w = org.shape
        noise = randn_tensor((b, c, f + reasoning_latent_count, h, w), generator=generator, device=device, dtype=dtype)
        return torch.cat([org, noise], dim=2)

    def prepare_video_latents_new(self, video, batch_size=1, num_channels_latents=16, height=480, width=832,
                                  dtype=torch.float32, device=None, generator=None, condition_count=None,
                                  latents=None, timestep=None):
        """[ src | noise(target) ] (non-CoT paired mode)."""
        if latents is not None:
            return latents.to(device=device, dtype=dtype)
        org = self._encode_source(video, dtype, device)
        noise = randn_tensor(tuple(org.shape), generator=generator, device=device, dtype=dtype)
        return torch.cat([org, noise], dim=2)

    def _decode_device(self, latents):
        """:423-426 on the device: decode, map [-1,1] -> [0,1] in the VAE's dtype as the reference does, widen to fp32
        (bf16 -> fp32 is exact wherever it happens; on the device it is not an element-by-element host loop)."""
        frames = self.vae.decode(latents.to(self.vae.dtype)).sample
        return (frames / 2 + 0.5).clamp(0, 1).float()

    def decode_latents(self, latents):
        """:423-428 — fp32 numpy [B, 3, T, H, W] in [0,1] on the host; same values as the reference's
        `frames.cpu().float().numpy()`."""
        return self._decode_device(latents).cpu().numpy()

    def _decode_frames_device(self, latents):
        return self.vae.decode_frames(latents.to(self.vae.dtype))

    def decode_frames(self, latents):
        """decode_latents + the uint8 conversion of save_videos_grid (utils/utils.py:66) on the device: uint8
        [B, T, H, W, 3] numpy on the host, bit-identical to what the reference writes to the video file."""
        return self._decode_frames_device(latents).cpu().numpy()

    # ---- __call__ (:518-799) -----------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, video=None, prompt=None, negative_prompt=None, height=480, width=720, num_frames=49,
                 source_frames=33, reasoning_frames=4, num_inference_steps=50, timesteps=None, guidance_scale=6,
                 num_videos_per_prompt=1, eta=0.0, generator=None, latents=None, prompt_embeds=None,
                 negative_prompt_embeds=None, output_type="numpy", return_dict=False, callback_on_step_end=None,
                 attention_kwargs=None, callback_on_step_end_tensor_inputs=("latents",), max_sequence_length=512,
                 comfyui_progressbar=False, shift=5, repeat_rope=True, cot=False):
        # a diffusers PipelineCallback / MultiPipelineCallbacks object names its own tensor inputs (:559-560)
        if hasattr(callback_on_step_end, "tensor_inputs"):
            callback_on_step_end_tensor_inputs = callback_on_step_end.tensor_inputs
        num_videos_per_prompt = 1
        self.check_inputs(prompt, height, width, negative_prompt, callback_on_step_end_tensor_inputs, prompt_embeds,
                          negative_prompt_embeds)
        self._guidance_scale = guidance_scale
        self._interrupt = False
        device = self._execution_device
        weight_dtype = self.text_encoder.dtype if self.text_encoder is not None else self.transformer.dtype
        do_cfg = guidance_scale > 1.0
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None:
            batch_size = len(prompt)
        else:
            batch_size = len(prompt_embeds)

        prompt_embeds, negative_prompt_embeds = self.encode_prompt(
            prompt, negative_prompt, do_cfg, num_videos_per_prompt, prompt_embeds, negative_prompt_embeds,
            max_sequence_length, device)
        in_prompt_embeds = (list(negative_prompt_embeds) + list(prompt_embeds)) if do_cfg else list(prompt_embeds)

        if isinstance(self.scheduler, FlowUniPCMultistepScheduler):
            self.scheduler.set_timesteps(num_inference_steps, device=device, shift=shift)
        else:  # any diffusers-style scheduler object
            self.scheduler.set_timesteps(num_inference_steps, device=device)
        timesteps = self.scheduler.timesteps
        self._num_timesteps = len(timesteps)

        ratio = getattr(self.vae, "temporal_compression_ratio", 4)
        condition_count = 1 if source_frames == 1 else (source_frames - 1) // ratio + 1
        latent_channels = self.transformer.config.in_channels
        ground_latent_count = 0
        if cot:
            ground_latent_count = 1 if reasoning_frames <= 1 else (reasoning_frames - 1) // ratio + 1
            latents = self.prepare_cot_video_latents(video, ground_latent_count, batch_size, latent_channels, height,
                                                     width, weight_dtype, device, generator, condition_count, latents)
        else:
            latents = self.prepare_video_latents_new(video, batch_size, latent_channels, height, width, weight_dtype,
                                                     device, generator, condition_count, latents)

        _, _, f_lat, h_lat, w_lat = latents.shape
        ps = self.transformer.config.patch_size
        seq_len = math.ceil((h_lat * w_lat) / (ps[1] * ps[2]) * f_lat)
        self.transformer.num_inference_steps = num_inference_steps
        # The prompt embeddings are fixed from here to the last step (:606; the callback's replacements never reach the
        # DiT, see below), so the text embedding and the blocks' cross-attention K / V are computed at the first step
        # only (dit.enable_context_cache); the cache lives for this call alone.
        scoped_cache = (getattr(self.transformer, "_ctx_cache", 0) is None
                        and os.environ.get("VCOF_CONTEXT_CACHE", "1") != "0")
        if scoped_cache:
            self.transformer.enable_context_cache()
        try:
            for i, t in enumerate(timesteps):
                self.transformer.current_steps = i
                if self._interrupt:
                    continue
                x = torch.cat([latents] * 2) if do_cfg else latents
                if hasattr(self.scheduler, "scale_model_input"):
                    x = self.scheduler.scale_model_input(x, t)
                timestep = t.expand(x.shape[0])
                fsi = gfi = None
                if repeat_rope and video is not None:
                    fsi = [condition_count] * x.shape[0]
                    if cot:
                        gfi = [(condition_count, condition_count + ground_latent_count)] * x.shape[0]
                noise_pred = self.transformer(x=x, context=in_prompt_embeds, t=timestep, seq_len=seq_len,
                                              frame_split_indices=fsi, ground_frame_indices=gfi)
                if do_cfg:
                    uncond, text = noise_pred.chunk(2)
                    noise_pred = uncond + self.guidance_scale * (text - uncond)
                noise_pred[:, :, :condition_count] = 0          # source frames get zero velocity (:736)
                latents = self.scheduler.step(noise_pred, t, latents, return_dict=False)[0]
                if callback_on_step_end is not None:
                    kw = {"latents": latents, "prompt_embeds": prompt_embeds,
                          "negative_prompt_embeds": negative_prompt_embeds}
                    outs = callback_on_step_end(self, i, t, {k: kw[k] for k in callback_on_step_end_tensor_inputs})
                    latents = outs.pop("latents", latents)
                    # (:747-748) popped like the reference does; like there, the DiT keeps the embeddings of step 0
                    prompt_embeds = outs.pop("prompt_embeds", prompt_embeds)
                    negative_prompt_embeds = outs.pop("negative_prompt_embeds", negative_prompt_embeds)
        finally:
            if scoped_cache:
                self.transformer.disable_context_cache()

        ground_video = edit_video = None
        # any output_type other than "numpy" (or this repo's "uint8") decodes nothing: the reference then hands the INPUT
        # clip back as `.videos` (:757-799, `video` is never reassigned); the final latents reach the caller through
        # callback_on_step_end
        out_video = video
        if output_type in ("numpy", "uint8"):
            # "uint8" (not in the reference): byte frames [B, T, H, W, 3] converted on the device; time is axis 1 there
            decode, t_axis = (self._decode_device, 2) if output_type == "numpy" else (self._decode_frames_device, 1)
            if cot:
                # (:770-777) the ground and the edit segment are decoded separately and concatenated along time.  The
                # concatenation happens on the device and ONE copy brings the clip to the host (0.4 GB of fp32 for 38
                # frames of 720p: the reference's two host arrays + np.concatenate move it three times);
                # `ground_videos` / `edit_videos` are views of `videos`.
                g0, g1 = condition_count, condition_count + ground_latent_count
                parts = []
                if g1 > g0 and g0 < latents.shape[2]:
                    parts.append(decode(latents[:, :, g0:g1]))
                n_ground = parts[0].shape[t_axis] if parts else 0
                if g1 < latents.shape[2]:
                    parts.append(decode(latents[:, :, g1:]))
                if not parts:
                    raise ValueError("chain-of-frames latents hold neither a ground nor an edit segment to decode")
                out_video = (torch.cat(parts, dim=t_axis) if len(parts) > 1 else parts[0]).cpu().numpy()
                cut = [slice(None)] * t_axis
                if n_ground:
                    ground_video = out_video[tuple(cut + [slice(0, n_ground)])]
                if out_video.shape[t_axis] > n_ground:
                    edit_video = out_video[tuple(cut + [slice(n_ground, None)])]
            else:
                if condition_count < latents.shape[2]:
                    edit_video = decode(latents[:, :, condition_count:]).cpu().numpy()
                out_video = edit_video
        self.maybe_free_model_hooks()
        if not return_dict:
            conv = lambda v: torch.from_numpy(v) if isinstance(v, np.ndarray) else v  # noqa: E731
            out_video, ground_video, edit_video = conv(out_video), conv(ground_video), conv(edit_video)
        return WanPipelineOutput(videos=out_video, ground_videos=ground_video, edit_videos=edit_video)
