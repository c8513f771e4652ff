"""Goldens for the frame-byte ends of the pipeline.

Executes the UNMODIFIED source text of the reference's own functions — cut out of the files under /root/reference
with `ast` (the modules themselves import packages this image does not have: imageio, diffusers, omegaconf) and
exec'd against the real numpy / torch / torchvision / einops / PIL plus an in-memory stand-in for `imageio` (a frame
list as the reader, a capture for `mimsave`):

  * `load_video_frames`   (fast_infer.py:43-91)                      frame selection + uint8 -> fp32 in [-1, 1]
  * `WanPipeline.decode_latents` (videox_fun/pipeline/pipeline_wan.py:423-428)  bf16 (x / 2 + 0.5).clamp(0, 1) -> fp32 numpy
  * `save_videos_grid`    (videox_fun/utils/utils.py:59-83)          fp32 [0, 1] -> uint8 frames
  * `save_results`        (fast_infer.py:170-180)                    the caller (single image or video)

and stores inputs and outputs in tests/golden/video_io.npz:
  * every one of the 256 byte values through load_video_frames, then the pipeline's cast to bf16 (pipeline_wan.py:397);
  * every non-NaN bf16 bit pattern as a decoder output through decode_latents + save_results -> bytes;
  * the frame indices load_video_frames picks for a set of (total_frames, source_frames, seed) cases, including the
    short-video case that pads by repeating the last frame.

    python tools/gen_golden_video_io.py
"""
import ast
import os
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VIDEOCOF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "video_io.npz")

SELECT_CASES = [  # (total_frames, source_frames, seed)
    (100, 33, 0), (100, 33, 1), (81, 81, 0), (40, 33, 3), (20, 33, 0), (1, 5, 0), (300, 33, 7), (67, 33, 2),
    (66, 33, 5), (165, 81, 4), (34, 33, 9), (33, 1, 0),
]


def cut_function(path, name, cls=None):
    """Source text of a top-level function (or of a method of `cls`) exactly as it stands in the reference file."""
    src = open(path).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    node = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    lines = src.splitlines()[node.lineno - 1:node.end_lineno]
    if cls is not None:  # dedent one level
        lines = [ln[4:] if ln.startswith("    ") else ln for ln in lines]
    return "\n".join(lines) + "\n"


class FakeReader:
    """imageio reader over an in-memory frame list; records which indices were read."""

    def __init__(self, frames):
        self.frames = frames
        self.read = []

    def count_frames(self):
        return len(self.frames)

    def get_data(self, idx):
        if idx >= len(self.frames):
            raise IndexError(idx)
        self.read.append(int(idx))
        return self.frames[idx]

    def __iter__(self):
        return iter(self.frames)

    def close(self):
        pass


def reference_functions():
    from einops import rearrange
    import torchvision
    from PIL import Image
    captured = {}
    fake_imageio = types.SimpleNamespace(
        get_reader=lambda path: fake_imageio.reader,
        mimsave=lambda path, outputs, **kw: captured.update(path=path, frames=[np.array(o) for o in outputs], kw=kw))
    ns = dict(np=np, torch=torch, Image=Image, imageio=fake_imageio, os=os, rearrange=rearrange,
              torchvision=torchvision, print=lambda *a, **k: None)
    exec(cut_function(os.path.join(REF, "fast_infer.py"), "load_video_frames"), ns)
    exec(cut_function(os.path.join(REF, "videox_fun/utils/utils.py"), "save_videos_grid"), ns)
    exec(cut_function(os.path.join(REF, "fast_infer.py"), "save_results"), ns)
    exec(cut_function(os.path.join(REF, "fast_infer.py"), "_normalize_to_01"), ns)
    exec(cut_function(os.path.join(REF, "fast_infer.py"), "save_side_by_side"), ns)
    exec(cut_function(os.path.join(REF, "videox_fun/pipeline/pipeline_wan.py"), "decode_latents", cls="WanPipeline"), ns)
    return ns, fake_imageio, captured


def all_bf16_patterns():
    """Every 16-bit pattern that is not a NaN, as int16 bits (NaN never leaves the decoder's clamp)."""
    b = np.arange(65536, dtype=np.uint32)
    nan = ((b & 0x7f80) == 0x7f80) & ((b & 0x007f) != 0)
    return b[~nan].astype(np.uint16)


def main():
    ns, fake_imageio, captured = reference_functions()
    out = {}

    # 1. frame bytes in: all 256 values, 3 channels, through load_video_frames and the cast to the VAE dtype
    rng = np.random.default_rng(0)
    T, H, W = 5, 16, 24
    frames = rng.integers(0, 256, size=(T, H, W, 3), dtype=np.uint8)
    frames.reshape(-1)[:256] = np.arange(256, dtype=np.uint8)
    fake_imageio.reader = FakeReader(list(frames))
    torch.manual_seed(0)
    video, h0, w0 = ns["load_video_frames"]("x.mp4", source_frames=T)
    assert (h0, w0) == (H, W) and tuple(video.shape) == (1, 3, T, H, W) and video.dtype == torch.float32
    out["in_frames_u8"] = frames                                                   # [T, H, W, 3]
    out["in_video_f32"] = video.numpy()                                            # [1, 3, T, H, W]
    out["in_video_bf16_bits"] = video.to(torch.bfloat16).view(torch.int16).numpy()  # pipeline_wan.py:397

    # 2. frame bytes out: every non-NaN bf16 pattern as decoder output
    pats = all_bf16_patterns()
    To, Ho, Wo = 2, 96, 128
    n = 3 * To * Ho * Wo
    fill = torch.empty(n - pats.size).uniform_(-1.25, 1.25, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    bits = np.concatenate([pats.view(np.int16), fill.view(torch.int16).numpy()])
    dec = torch.from_numpy(bits.copy()).view(torch.bfloat16).reshape(1, 3, To, Ho, Wo)
    fake_self = types.SimpleNamespace(vae=types.SimpleNamespace(
        dtype=torch.bfloat16, decode=lambda z: types.SimpleNamespace(sample=z)))
    f32 = ns["decode_latents"](fake_self, dec)                                     # numpy fp32 [1, 3, T, H, W]
    assert f32.dtype == np.float32 and f32.min() >= 0.0 and f32.max() <= 1.0
    with tempfile.TemporaryDirectory() as td:
        ns["save_results"](torch.from_numpy(f32), os.path.join(td, "o", "v.mp4"), 16)
        video_u8 = np.stack(captured["frames"])                                    # [T, H, W, 3]
        assert captured["kw"] == {"fps": 16}
        from PIL import Image
        png = os.path.join(td, "o", "i.png")
        ns["save_results"](torch.from_numpy(f32[:, :, :1]), png, 16)               # T == 1: single image branch
        image_u8 = np.array(Image.open(png))
    assert np.array_equal(image_u8, video_u8[0])
    out["out_dec_bf16_bits"] = bits.reshape(1, 3, To, Ho, Wo)
    out["out_frames_u8"] = video_u8
    out["out_video_f32_crc"] = np.array([int(np.frombuffer(f32.tobytes(), np.uint8).astype(np.uint64).sum())])

    # 3. frame selection
    sel = []
    for total, want, seed in SELECT_CASES:
        vid = [np.full((2, 2, 3), i % 256, dtype=np.uint8) for i in range(total)]
        fake_imageio.reader = FakeReader(vid)
        torch.manual_seed(seed)
        v, _, _ = ns["load_video_frames"]("x.mp4", source_frames=want)
        picked = list(fake_imageio.reader.read)
        picked += [picked[-1]] * (want - len(picked))                 # short video: the last frame is repeated
        seen = ((v[0, 0, :, 0, 0] + 1.0) * 127.5).round().to(torch.int64).tolist()   # frame i is filled with i % 256
        assert len(picked) == want and seen == [i % 256 for i in picked]
        sel.append(picked + [-1] * (96 - want))
    out["select_cases"] = np.array(SELECT_CASES, dtype=np.int64)
    out["select_picked"] = np.array(sel, dtype=np.int64)

    # 4. the comparison clip (fast_infer.py:183-206): input bytes through load_video_frames next to the result; the
    #    second case has no byte below 128, where the reference skips its [-1, 1] -> [0, 1] mapping
    with tempfile.TemporaryDirectory() as td:
        for tag, lo in (("sbs", 0), ("sbs_bright", 128)):
            src = rng.integers(lo, 256, size=(5, 16, 24, 3), dtype=np.uint8)
            src.reshape(-1)[:256 - lo] = np.arange(lo, 256, dtype=np.uint8)
            fake_imageio.reader = FakeReader(list(src))
            torch.manual_seed(0)
            vin, _, _ = ns["load_video_frames"]("x.mp4", source_frames=5)
            res = torch.from_numpy(f32[:, :, :, :24, :20])                          # result: 2 frames, 24 x 20
            ns["save_side_by_side"](vin, res, os.path.join(td, tag, "c.mp4"), 16)
            out[tag + "_in_u8"] = src
            out[tag + "_frames_u8"] = np.stack(captured["frames"])                  # [2, 16, 40, 3]
        out["sbs_result_u8"] = video_u8[None, :, :24, :20]

        # 5. a batch of clips as a grid (utils/utils.py:63: make_grid, 2-pixel black padding)
        grid_in = torch.rand(5, 3, 2, 6, 10, generator=torch.Generator().manual_seed(2))
        ns["save_videos_grid"](grid_in, os.path.join(td, "g", "g.mp4"), n_rows=3, fps=8)
        out["grid_in_f32"] = grid_in.numpy()
        out["grid_frames_u8"] = np.stack(captured["frames"])

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
