"""Frame bytes at the two ends of the pipeline, everything that needs no GPU:

  * the oracle (oracle/video_io_oracle.py) against the goldens of the executed reference functions
    (tests/golden/video_io.npz, tools/gen_golden_video_io.py) — bit-exact;
  * the per-element arithmetic of the two CUDA kernels, evaluated on the host by the library's own debug entry points
    (the functions are __host__ __device__), against the same goldens — exhaustively: all 256 input bytes, all 65 280
    non-NaN bf16 decoder outputs;
  * the host module videocof_b200/video_io.py (reference CLI function names) against the goldens, with an in-memory
    imageio stand-in;
  * the VAE / pipeline host logic for byte frames through the contract emulator.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

import vcof_emulator
from oracle import video_io_oracle as vo


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "video_io.npz"))


# ---- oracle pinned to the executed reference -------------------------------------------------------------------

def test_oracle_bytes_in_matches_reference(gold):
    f32, bits = vo.video_to_model_input(gold["in_frames_u8"])
    assert np.array_equal(f32.view(np.uint32), gold["in_video_f32"].view(np.uint32))
    assert np.array_equal(bits, gold["in_video_bf16_bits"].view(np.uint16))
    assert set(np.unique(gold["in_frames_u8"])) == set(range(256))          # every byte value is covered


def test_oracle_bytes_out_matches_reference(gold):
    bits = gold["out_dec_bf16_bits"]
    assert np.unique(bits).size == 65536 - 2 * 127                           # every non-NaN bf16 pattern is covered
    frames = vo.model_output_to_frames(bits)
    assert np.array_equal(frames[0], gold["out_frames_u8"])


def test_oracle_frame_selection_matches_reference(gold):
    for (total, want, seed), picked in zip(gold["select_cases"], gold["select_picked"]):
        torch.manual_seed(int(seed))
        start = torch.randint(0, vo.start_frame_bound(int(total), int(want)), (1,))[0].item()
        assert vo.select_frame_indices(int(total), int(want), start) == picked[:want].tolist()


# ---- the kernels' arithmetic, evaluated on the host by libvcof ------------------------------------------------

def test_kernel_arithmetic_bytes_in_exhaustive(gold):
    from videocof_b200 import _lib
    src = np.arange(256, dtype=np.uint8)
    bits = np.zeros(256, dtype=np.uint16)
    _lib.call("vcof_debug_video_bf16_host", src.ctypes.data, bits.ctypes.data, 256)
    assert np.array_equal(bits, vo.video_to_model_input(src.reshape(1, 1, 256, 1))[1].reshape(-1))
    frames = np.ascontiguousarray(gold["in_frames_u8"])
    got = np.zeros(frames.shape, dtype=np.uint16)
    _lib.call("vcof_debug_video_bf16_host", frames.ctypes.data, got.ctypes.data, frames.size)
    ref = gold["in_video_bf16_bits"].view(np.uint16)[0].transpose(1, 2, 3, 0)          # [T, H, W, 3]
    assert np.array_equal(got, ref)


def test_kernel_arithmetic_bytes_out_exhaustive(gold):
    from videocof_b200 import _lib
    bits = np.ascontiguousarray(gold["out_dec_bf16_bits"].view(np.uint16)[0].transpose(1, 2, 3, 0))   # [T, H, W, 3]
    got = np.zeros(bits.shape, dtype=np.uint8)
    _lib.call("vcof_debug_frame_u8_host", bits.ctypes.data, got.ctypes.data, bits.size)
    assert np.array_equal(got, gold["out_frames_u8"])
    nan = np.array([0x7fc0, 0xffc0, 0x7f81], dtype=np.uint16)               # NaN is pinned to black (documented)
    out = np.full(3, 7, dtype=np.uint8)
    _lib.call("vcof_debug_frame_u8_host", nan.ctypes.data, out.ctypes.data, 3)
    assert out.tolist() == [0, 0, 0]


# ---- host module with the reference CLI's function names --------------------------------------------------------

class _Reader:
    def __init__(self, frames, over_report=0):
        self.frames, self.over = frames, over_report

    def count_frames(self):
        return len(self.frames) + self.over

    def get_data(self, i):
        if i >= len(self.frames):
            raise IndexError(i)
        return self.frames[i]

    def __iter__(self):
        return iter(self.frames)

    def close(self):
        pass


@pytest.fixture
def fake_imageio(monkeypatch):
    cap = {}
    mod = types.ModuleType("imageio")
    mod.get_reader = lambda path: mod.reader
    mod.mimsave = lambda path, outputs, **kw: cap.update(path=path, frames=np.stack([np.array(o) for o in outputs]), kw=kw)
    mod.captured = cap
    monkeypatch.setitem(sys.modules, "imageio", mod)
    return mod


def test_load_video_frames_both_forms(gold, fake_imageio):
    from videocof_b200 import video_io
    fake_imageio.reader = _Reader(list(gold["in_frames_u8"]))
    torch.manual_seed(0)
    v, h, w = video_io.load_video_frames("x.mp4", source_frames=5)
    assert (h, w) == (16, 24) and v.dtype == torch.float32
    assert np.array_equal(v.numpy().view(np.uint32), gold["in_video_f32"].view(np.uint32))
    torch.manual_seed(0)
    b, h, w = video_io.load_video_frames("x.mp4", source_frames=5, as_uint8=True)
    assert b.dtype == torch.uint8 and tuple(b.shape) == (1, 5, 16, 24, 3) and (h, w) == (16, 24)
    assert np.array_equal(b[0].numpy(), gold["in_frames_u8"])


def test_load_video_frames_selection(gold, fake_imageio):
    from videocof_b200 import video_io
    for (total, want, seed), picked in zip(gold["select_cases"], gold["select_picked"]):
        clip = [np.full((2, 2, 3), i % 256, dtype=np.uint8) for i in range(int(total))]
        fake_imageio.reader = _Reader(clip)
        torch.manual_seed(int(seed))
        b, _, _ = video_io.load_video_frames("x.mp4", source_frames=int(want), as_uint8=True)
        assert b[0, :, 0, 0, 0].tolist() == [i % 256 for i in picked[:want]]
    # count_frames over-reporting: reading stops at the IndexError and the last good frame is repeated (:74-80)
    fake_imageio.reader = _Reader([np.full((2, 2, 3), i, dtype=np.uint8) for i in range(3)], over_report=2)
    torch.manual_seed(0)
    b, _, _ = video_io.load_video_frames("x.mp4", source_frames=5, as_uint8=True)
    assert b[0, :, 0, 0, 0].tolist() == [0, 1, 2, 2, 2]
    with pytest.raises(AssertionError):
        video_io.load_video_frames("x.mp4")


def test_save_results_float_and_bytes(gold, fake_imageio, tmp_path):
    from PIL import Image
    from videocof_b200 import video_io
    dec = torch.from_numpy(gold["out_dec_bf16_bits"].copy()).view(torch.bfloat16)
    f32 = (dec / 2 + 0.5).clamp(0, 1).float()                                # what decode_latents returns
    video_io.save_results(f32, str(tmp_path / "a" / "v.mp4"), 16)
    assert np.array_equal(fake_imageio.captured["frames"], gold["out_frames_u8"])
    assert fake_imageio.captured["kw"] == {"fps": 16}
    u8 = torch.from_numpy(gold["out_frames_u8"])[None]                       # what output_type="uint8" returns
    video_io.save_results(u8, str(tmp_path / "b" / "v.mp4"), 12)
    assert np.array_equal(fake_imageio.captured["frames"], gold["out_frames_u8"])
    assert fake_imageio.captured["kw"] == {"fps": 12}
    for name, clip in (("f.png", f32[:, :, :1]), ("u.png", u8[:, :1])):     # one frame -> an image
        video_io.save_results(clip, str(tmp_path / "c" / name), 16)
        assert np.array_equal(np.array(Image.open(tmp_path / "c" / name)), gold["out_frames_u8"][0])
    video_io.save_videos_grid(f32, str(tmp_path / "d" / "v.gif"), fps=10)
    assert fake_imageio.captured["kw"] == {"duration": 100.0}


def test_save_videos_grid_batch(gold, fake_imageio, tmp_path):
    from videocof_b200 import video_io
    x = torch.from_numpy(gold["grid_in_f32"])
    video_io.save_videos_grid(x, str(tmp_path / "g" / "g.mp4"), n_rows=3, fps=8)
    assert np.array_equal(fake_imageio.captured["frames"], gold["grid_frames_u8"])
    u8 = (x * 255).numpy().astype(np.uint8).transpose(0, 2, 3, 4, 1)
    video_io.save_videos_grid(torch.from_numpy(np.ascontiguousarray(u8)), str(tmp_path / "g" / "h.mp4"), n_rows=3, fps=8)
    assert np.array_equal(fake_imageio.captured["frames"], gold["grid_frames_u8"])


@pytest.mark.parametrize("tag", ["sbs", "sbs_bright"])
def test_save_side_by_side(gold, fake_imageio, tmp_path, tag):
    from videocof_b200 import video_io
    src = torch.from_numpy(gold[tag + "_in_u8"])[None]                       # uint8 [1, 5, 16, 24, 3]
    res = torch.from_numpy(gold["sbs_result_u8"])                            # uint8 [1, 2, 24, 20, 3]
    video_io.save_side_by_side(src, res, str(tmp_path / tag / "c.mp4"), 16)
    assert np.array_equal(fake_imageio.captured["frames"], gold[tag + "_frames_u8"])
    # the reference's own float form
    vin = src[0].permute(3, 0, 1, 2)[None].float() * (2.0 / 255.0) - 1.0
    dec = torch.from_numpy(gold["out_dec_bf16_bits"].copy()).view(torch.bfloat16)
    f32 = (dec / 2 + 0.5).clamp(0, 1).float()[:, :, :, :24, :20]
    video_io.save_side_by_side(vin, f32, str(tmp_path / tag / "d.mp4"), 16)
    assert np.array_equal(fake_imageio.captured["frames"], gold[tag + "_frames_u8"])
    with pytest.raises(ValueError):
        video_io.save_side_by_side(src, f32, str(tmp_path / tag / "e.mp4"), 16)


def test_missing_imageio_is_loud(monkeypatch):
    from videocof_b200 import video_io
    monkeypatch.setitem(sys.modules, "imageio", None)
    with pytest.raises(ImportError, match="imageio"):
        video_io.load_video_frames("x.mp4", source_frames=3)


# ---- VAE / pipeline host logic for byte frames (contract emulator) -------------------------------------------

@pytest.fixture(scope="module")
def vae_model():
    from oracle.vae_oracle import VAEConfig, make_vae_params
    from videocof_b200.vae import AutoencoderKLWan
    m = AutoencoderKLWan()
    m.load_state_dict(make_vae_params(VAEConfig(), seed=17), strict=True)
    return m.to(torch.bfloat16).eval()


def test_vae_byte_frames_match_float_path(vae_model, monkeypatch):
    """encode_frames(bytes) feeds the first convolution the same bits as encode(load_video_frames' fp32 clip) and
    decode_frames gives the bytes the reference's host chain makes of decode's output."""
    vcof_emulator.install(monkeypatch)
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (1, 5, 16, 16, 3), generator=g, dtype=torch.uint8)
    video = frames[0].permute(3, 0, 1, 2)[None].float() * (2.0 / 255.0) - 1.0          # fast_infer.py:86-88
    with torch.no_grad():
        mu_b = vae_model.encode_frames(frames)[0].mode()
        mu_f = vae_model.encode(video.to(torch.bfloat16))[0].mode()                     # pipeline_wan.py:397, 406
        z = torch.randn(1, 16, 2, 2, 2, generator=g).to(torch.bfloat16)
        dec = vae_model.decode(z).sample
        dec_u8 = vae_model.decode_frames(z)
    assert torch.equal(mu_b, mu_f)
    assert dec_u8.dtype == torch.uint8 and tuple(dec_u8.shape) == (1, 5, 16, 16, 3)
    ref = vo.model_output_to_frames(dec.view(torch.int16).numpy())
    assert np.array_equal(dec_u8.numpy(), ref)
    from videocof_b200._lib import VcofError
    with pytest.raises(VcofError):
        vae_model.encode_frames(video)                                       # float clip: wrong entry point
    with pytest.raises(VcofError):
        vae_model.encode_frames(frames[..., :2])                             # not RGB


def test_overlay_utils_exports():
    """`from videox_fun.utils.utils import filter_kwargs, save_videos_grid` (fast_infer.py:37) resolves through the
    overlay: the writer takes bytes, filter_kwargs behaves like the reference's (utils/utils.py:17-21)."""
    import videox_fun.utils.utils as m
    from videocof_b200 import video_io
    assert m.save_videos_grid is video_io.save_videos_grid

    class K:
        def __init__(self, a, b=2):
            pass
    assert m.filter_kwargs(K, {"a": 1, "c": 3, "self": 0}) == {"a": 1}
