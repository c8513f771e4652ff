"""Time the CUDA VAE (random-init reference-shaped weights) on a 720p clip: encode T frames, decode the latents.

    python tools/vae_bench.py [--frames 9] [--height 720] [--width 1280]
Prints a JSON line with ms, algorithmic conv TFLOP/s and the per-kernel-class breakdown."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videocof_b200 import ops  # noqa: E402
from videocof_b200.vae import AutoencoderKLWan  # noqa: E402


def conv_flops(model, T, H, W):
    """Sum 2*Cin*Cout*k*positions over the convs, following the layer plan (positions per stage)."""
    import torch.nn as nn
    from videocof_b200 import vae as V
    fl = {"enc": 0.0, "dec": 0.0}

    def walk(layers, t, h, w, key, up):
        for m in layers:
            if isinstance(m, V.ResidualBlock):
                for c in (m.residual[2], m.residual[6]):
                    fl[key] += 2.0 * c.weight.numel() * t * h * w
                if not isinstance(m.shortcut, nn.Identity):
                    fl[key] += 2.0 * m.shortcut.weight.numel() * t * h * w
            elif isinstance(m, V.Resample):
                if m.mode == "upsample3d":
                    fl[key] += 2.0 * m.time_conv.weight.numel() * max(t - 1, 0) * h * w
                    t = 1 + 2 * (t - 1)
                if m.mode.startswith("up"):
                    h, w = 2 * h, 2 * w
                    fl[key] += 2.0 * m.resample[1].weight.numel() * t * h * w
                else:
                    h, w = h // 2, w // 2
                    fl[key] += 2.0 * m.resample[1].weight.numel() * t * h * w
                    if m.mode == "downsample3d":
                        t = 1 + (t - 1) // 2
                        fl[key] += 2.0 * m.time_conv.weight.numel() * max(t - 1, 0) * h * w
        return t, h, w

    e, d = model.model.encoder, model.model.decoder
    fl["enc"] += 2.0 * e.conv1.weight.numel() * T * H * W
    t, h, w = walk(e.downsamples, T, H, W, "enc", False)
    walk(e.middle, t, h, w, "enc", False)
    fl["enc"] += 2.0 * e.head[2].weight.numel() * t * h * w
    f = t
    fl["dec"] += 2.0 * d.conv1.weight.numel() * f * h * w
    walk(d.middle, f, h, w, "dec", True)
    t2, h2, w2 = walk(d.upsamples, f, h, w, "dec", True)
    fl["dec"] += 2.0 * d.head[2].weight.numel() * t2 * h2 * w2
    return fl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=9)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    a = ap.parse_args()
    torch.manual_seed(0)
    m = AutoencoderKLWan()
    for p in m.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=1.0 / (p[0].numel() ** 0.5))
    m = m.to("cuda", torch.bfloat16).eval()
    video = (torch.rand(1, 3, a.frames, a.height, a.width, device="cuda") * 2 - 1).bfloat16()
    fl = conv_flops(m, a.frames, a.height, a.width)
    out = {}
    with torch.no_grad():
        for _ in range(2):
            z = m.encode(video)[0].mode()
            y = m.decode(z).sample
        torch.cuda.synchronize()
        for name, fn in (("enc", lambda: m.encode(video)[0].mode()), ("dec", lambda: m.decode(z).sample)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ops.enable_timing()
            e0.record()
            fn()
            e1.record()
            tm = ops.collect_timing()
            ms = e0.elapsed_time(e1)
            top = sorted(((k, n, t) for k, (n, t) in tm.items()), key=lambda r: -r[2])[:6]
            out[name] = dict(ms=ms, conv_tflops=fl[name] / ms / 1e9, kernel_ms=sum(t for _, t in tm.values()),
                             top=[(k, n, round(t, 2)) for k, n, t in top])
    out["shape"] = dict(frames=a.frames, H=a.height, W=a.width, latent=list(z.shape), out=list(y.shape))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
