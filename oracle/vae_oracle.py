"""CPU oracle for the Wan-2.1 3D causal VAE (encode / decode).  TEST INFRASTRUCTURE ONLY.

A from-scratch, UN-CHUNKED restatement (functional torch, fp32 on CPU) of the reference's
videox_fun/models/wan_vae.py.  The reference streams the clip through the network in chunks
(frame 0 alone, then 4 frames at a time for encode; one latent frame at a time for decode)
with a 2-frame feature cache per causal convolution (:520-575).  That chunking is not
semantic: every stride-1 CausalConv3d is a causal convolution over the whole sequence with
two zero frames of left padding, and only frame 0 is special in the temporal resamplers
(probe-verified).  This file states that closed form directly; the goldens in
tests/golden/vae_*.npz come from the reference's own chunked loop (tools/gen_golden.py), so
tests/test_oracle_golden.py also proves the closed form.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.
Citations are to videox_fun/models/wan_vae.py in the reference repo.
"""
import math

import torch
import torch.nn.functional as F

__all__ = ["VAEConfig", "make_vae_params", "vae_encode", "vae_decode", "LATENT_MEAN", "LATENT_STD"]

# :630-640
LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
               0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
              3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


class VAEConfig:
    """_video_vae defaults (:599-617): dim 96, mult [1,2,4,4], z 16, temporal downsample on the last two levels."""

    def __init__(self, dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2,
                 temperal_downsample=(False, True, True)):
        self.dim, self.z_dim, self.dim_mult = dim, z_dim, tuple(dim_mult)
        self.num_res_blocks, self.temperal_downsample = num_res_blocks, tuple(temperal_downsample)


# --------------------------------------------------------------------------------------
# layer plan: the module tree of Encoder3d (:269-320) / Decoder3d (:373-425) as a flat list
# --------------------------------------------------------------------------------------
def encoder_plan(cfg):
    dims = [cfg.dim * u for u in (1,) + cfg.dim_mult]
    plan, idx = [], 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(cfg.num_res_blocks):
            plan.append(("res", f"encoder.downsamples.{idx}", cin, cout))
            cin = cout
            idx += 1
        if i != len(cfg.dim_mult) - 1:
            mode = "downsample3d" if cfg.temperal_downsample[i] else "downsample2d"
            plan.append((mode, f"encoder.downsamples.{idx}", cout, cout))
            idx += 1
    return dims, plan


def decoder_plan(cfg):
    dims = [cfg.dim * u for u in (cfg.dim_mult[-1],) + cfg.dim_mult[::-1]]
    up = cfg.temperal_downsample[::-1]
    plan, idx = [], 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(cfg.num_res_blocks + 1):
            plan.append(("res", f"decoder.upsamples.{idx}", cin, cout))
            cin = cout
            idx += 1
        if i != len(cfg.dim_mult) - 1:
            mode = "upsample3d" if up[i] else "upsample2d"
            plan.append((mode, f"decoder.upsamples.{idx}", cout, cout // 2))
            idx += 1
    return dims, plan


def make_vae_params(cfg, seed=0, bf16_exact=True):
    """Deterministic random parameters keyed by the reference's state-dict names (`model.` prefix as in
    AutoencoderKLWan, :620-645).  He-style conv scales keep activations O(1) through the stack; the
    attention `proj` is NOT zero (the reference zero-inits it, :241, which would hide the attention)."""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def rnd(*shape, std=1.0, mean=0.0):
        t = mean + torch.randn(*shape, generator=g) * std
        return t.to(torch.bfloat16).float() if bf16_exact else t

    def conv(name, cout, cin, *k):
        fan_in = cin * math.prod(k)
        p[f"model.{name}.weight"] = rnd(cout, cin, *k, std=1.0 / math.sqrt(fan_in))
        p[f"model.{name}.bias"] = rnd(cout, std=0.02)

    def res(name, cin, cout):
        p[f"model.{name}.residual.0.gamma"] = rnd(cin, 1, 1, 1, std=0.05, mean=1.0)
        conv(f"{name}.residual.2", cout, cin, 3, 3, 3)
        p[f"model.{name}.residual.3.gamma"] = rnd(cout, 1, 1, 1, std=0.05, mean=1.0)
        conv(f"{name}.residual.6", cout, cout, 3, 3, 3)
        if cin != cout:
            conv(f"{name}.shortcut", cout, cin, 1, 1, 1)

    def attn(name, c):
        p[f"model.{name}.norm.gamma"] = rnd(c, 1, 1, std=0.05, mean=1.0)
        conv(f"{name}.to_qkv", 3 * c, c, 1, 1)
        conv(f"{name}.proj", c, c, 1, 1)

    def middle(prefix, c):
        res(f"{prefix}.middle.0", c, c)
        attn(f"{prefix}.middle.1", c)
        res(f"{prefix}.middle.2", c, c)

    dims, plan = encoder_plan(cfg)
    conv("encoder.conv1", dims[0], 3, 3, 3, 3)
    for kind, name, cin, cout in plan:
        if kind == "res":
            res(name, cin, cout)
        else:
            conv(f"{name}.resample.1", cin, cin, 3, 3)
            if kind == "downsample3d":
                conv(f"{name}.time_conv", cin, cin, 3, 1, 1)
    middle("encoder", dims[-1])
    p["model.encoder.head.0.gamma"] = rnd(dims[-1], 1, 1, 1, std=0.05, mean=1.0)
    conv("encoder.head.2", cfg.z_dim * 2, dims[-1], 3, 3, 3)
    conv("conv1", cfg.z_dim * 2, cfg.z_dim * 2, 1, 1, 1)
    conv("conv2", cfg.z_dim, cfg.z_dim, 1, 1, 1)
    ddims, dplan = decoder_plan(cfg)
    conv("decoder.conv1", ddims[0], cfg.z_dim, 3, 3, 3)
    middle("decoder", ddims[0])
    for kind, name, cin, cout in dplan:
        if kind == "res":
            res(name, cin, cout)
        else:
            conv(f"{name}.resample.1", cout, cin, 3, 3)
            if kind == "upsample3d":
                conv(f"{name}.time_conv", cin * 2, cin, 3, 1, 1)
    p["model.decoder.head.0.gamma"] = rnd(ddims[-1], 1, 1, 1, std=0.05, mean=1.0)
    conv("decoder.head.2", 3, ddims[-1], 3, 3, 3)
    return p


# --------------------------------------------------------------------------------------
# ops (x is [C, T, H, W] fp32, batch handled by the caller like the reference :647-653)
# --------------------------------------------------------------------------------------
def causal_conv3d(x, w, b, stride=(1, 1, 1)):
    """:21-40 over the WHOLE sequence: zero left-pad time by 2*pad_t (pad_t = (kt-1)/2 for the 'same'
    convs), symmetric zero pad in H/W, then a plain conv."""
    kt, kh, kw = w.shape[2:]
    x = F.pad(x[None], (kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0))
    return F.conv3d(x, w, b, stride=stride)[0]


def rms_norm(x, gamma):
    """:43-58 — L2-normalise over channels per (t,h,w) position, times sqrt(C) * gamma."""
    return F.normalize(x, dim=0) * math.sqrt(x.shape[0]) * gamma.reshape(-1, *([1] * (x.dim() - 1)))


def res_block(p, name, x):
    """:190-224 — x + conv(silu(norm(conv(silu(norm(x)))))), 1x1x1 shortcut if channels change."""
    pre = f"model.{name}."
    h = x
    if pre + "shortcut.weight" in p:
        h = causal_conv3d(x, p[pre + "shortcut.weight"], p[pre + "shortcut.bias"])
    y = F.silu(rms_norm(x, p[pre + "residual.0.gamma"]))
    y = causal_conv3d(y, p[pre + "residual.2.weight"], p[pre + "residual.2.bias"])
    y = F.silu(rms_norm(y, p[pre + "residual.3.gamma"]))
    y = causal_conv3d(y, p[pre + "residual.6.weight"], p[pre + "residual.6.bias"])
    return y + h


def attn_block(p, name, x):
    """:227-266 — per frame, single head with d = C over the h*w positions."""
    pre = f"model.{name}."
    C, T, H, W = x.shape
    xf = x.permute(1, 0, 2, 3)                                                # [T, C, H, W]
    y = F.normalize(xf, dim=1) * math.sqrt(C) * p[pre + "norm.gamma"].reshape(1, C, 1, 1)
    qkv = F.conv2d(y, p[pre + "to_qkv.weight"], p[pre + "to_qkv.bias"])       # [T, 3C, H, W]
    q, k, v = qkv.reshape(T, 3 * C, H * W).permute(0, 2, 1).chunk(3, dim=-1)  # [T, HW, C] each
    a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(C), dim=-1) @ v
    a = a.permute(0, 2, 1).reshape(T, C, H, W)
    a = F.conv2d(a, p[pre + "proj.weight"], p[pre + "proj.bias"])
    return x + a.permute(1, 0, 2, 3)


def conv2d_frames(x, w, b, stride=1, pad=(1, 1, 1, 1)):
    """Spatial conv applied to every frame ('b c t h w -> (b t) c h w', :142-145)."""
    y = F.conv2d(F.pad(x.permute(1, 0, 2, 3), pad), w, b, stride=stride)
    return y.permute(1, 0, 2, 3)


def downsample(p, name, x, temporal):
    """:91-100, :147-163 — ZeroPad2d(0,1,0,1) + Conv2d 3x3 stride 2; then (3d) frame 0 passes through
    and out[1+m] = time_conv(y[2m], y[2m+1], y[2m+2]) (kernel 3, stride 2, no padding)."""
    pre = f"model.{name}."
    y = conv2d_frames(x, p[pre + "resample.1.weight"], p[pre + "resample.1.bias"], stride=2, pad=(0, 1, 0, 1))
    if temporal and y.shape[1] > 1:
        z = F.conv3d(y[None], p[pre + "time_conv.weight"], p[pre + "time_conv.bias"], stride=(2, 1, 1))[0]
        y = torch.cat([y[:, :1], z], dim=1)
    return y


def upsample(p, name, x, temporal):
    """:80-89, :107-145 — (3d) frame 0 passes through; frames 1.. go through the causal (3,1,1) time_conv
    C->2C and each yields two frames (channels [0:C] then [C:2C]); then nearest x2 + Conv2d 3x3."""
    pre = f"model.{name}."
    C = x.shape[0]
    if temporal and x.shape[1] > 1:
        z = causal_conv3d(x[:, 1:], p[pre + "time_conv.weight"], p[pre + "time_conv.bias"])   # [2C, T-1, H, W]
        z = torch.stack([z[:C], z[C:]], dim=2).reshape(C, -1, *x.shape[2:])                   # interleave
        x = torch.cat([x[:, :1], z], dim=1)
    y = x.permute(1, 0, 2, 3)
    y = F.interpolate(y, scale_factor=(2.0, 2.0), mode="nearest-exact")
    y = F.conv2d(y, p[pre + "resample.1.weight"], p[pre + "resample.1.bias"], padding=1)
    return y.permute(1, 0, 2, 3)


def middle(p, prefix, x):
    x = res_block(p, f"{prefix}.middle.0", x)
    x = attn_block(p, f"{prefix}.middle.1", x)
    return res_block(p, f"{prefix}.middle.2", x)


# --------------------------------------------------------------------------------------
# encode / decode
# --------------------------------------------------------------------------------------
def vae_encode(p, cfg, video):
    """:520-548 + :655-665 — video [3, T, H, W] in [-1,1], T = 1 + 4k -> (mu [16, f, H/8, W/8] normalised with
    the latent mean/std, logvar).  The pipeline uses latent_dist.mode() = mu (pipeline_wan.py:406-407)."""
    _, plan = encoder_plan(cfg)
    x = causal_conv3d(video.float(), p["model.encoder.conv1.weight"], p["model.encoder.conv1.bias"])
    for kind, name, cin, cout in plan:
        x = res_block(p, name, x) if kind == "res" else downsample(p, name, x, kind == "downsample3d")
    x = middle(p, "encoder", x)
    x = F.silu(rms_norm(x, p["model.encoder.head.0.gamma"]))
    x = causal_conv3d(x, p["model.encoder.head.2.weight"], p["model.encoder.head.2.bias"])
    x = causal_conv3d(x, p["model.conv1.weight"], p["model.conv1.bias"])
    mu, logvar = x.chunk(2, dim=0)
    mean = torch.tensor(LATENT_MEAN[:cfg.z_dim]).reshape(-1, 1, 1, 1)
    inv_std = 1.0 / torch.tensor(LATENT_STD[:cfg.z_dim]).reshape(-1, 1, 1, 1)
    return (mu - mean) * inv_std, logvar


def vae_decode(p, cfg, z):
    """:550-575 + :667-682 — z [16, f, h, w] (normalised) -> video [3, 4(f-1)+1, 8h, 8w] clamped to [-1,1]."""
    _, plan = decoder_plan(cfg)
    mean = torch.tensor(LATENT_MEAN[:cfg.z_dim]).reshape(-1, 1, 1, 1)
    inv_std = 1.0 / torch.tensor(LATENT_STD[:cfg.z_dim]).reshape(-1, 1, 1, 1)
    z = z.float() / inv_std + mean
    x = causal_conv3d(z, p["model.conv2.weight"], p["model.conv2.bias"])
    x = causal_conv3d(x, p["model.decoder.conv1.weight"], p["model.decoder.conv1.bias"])
    x = middle(p, "decoder", x)
    for kind, name, cin, cout in plan:
        x = res_block(p, name, x) if kind == "res" else upsample(p, name, x, kind == "upsample3d")
    x = F.silu(rms_norm(x, p["model.decoder.head.0.gamma"]))
    x = causal_conv3d(x, p["model.decoder.head.2.weight"], p["model.decoder.head.2.bias"])
    return x.clamp(-1, 1)
