"""Frame-byte conversions on the GPU through the C ABI (vcof_u8_to_cl / vcof_cl_to_u8), against
the oracle (oracle/video_io_oracle.py, pinned bit-exactly to the executed reference by tests/golden/video_io.npz).
Byte work: every comparison is bit-exact.

Small cases compare with the oracle directly; at the full 81 x 720p size the property used is that each output
element is a pure function of one input element, so the device result must equal a 256-entry (65 536-entry) table
built by the oracle, looked up on the device."""
import os

import numpy as np
import pytest
import torch

from oracle import video_io_oracle as vo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "video_io.npz"))


def bf16_from_bits(bits_u16):
    return torch.from_numpy(np.ascontiguousarray(bits_u16).view(np.int16)).view(torch.bfloat16)


def bits_of(t):
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


@pytest.mark.parametrize("Cp", [32, 8, 3, 5])
def test_u8_to_cl_golden(gold, Cp):
    from videocof_b200 import ops
    frames = torch.from_numpy(gold["in_frames_u8"]).cuda()                  # all 256 byte values
    y = ops.u8_to_cl(frames, Cp)
    assert tuple(y.shape) == (5, 16, 24, Cp) and y.dtype == torch.bfloat16
    ref = gold["in_video_bf16_bits"].view(np.uint16)[0].transpose(1, 2, 3, 0)
    assert np.array_equal(bits_of(y[..., :3]), ref)
    assert not bits_of(y[..., 3:]).any()                                    # padding channels are zero bits


@pytest.mark.parametrize("shape", [(1, 1, 1, 3), (1, 3, 7, 3), (2, 5, 9, 1), (1, 2, 33, 4), (3, 8, 8, 8)])
def test_u8_to_cl_ragged(shape):
    from videocof_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    frames = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
    C = shape[3]
    for Cp in (C, 8, 32):
        y = ops.u8_to_cl(frames.cuda(), Cp)
        _, ref = vo.video_to_model_input(frames.numpy())                    # [1, C, T, H, W]
        assert np.array_equal(bits_of(y[..., :C]), ref[0].transpose(1, 2, 3, 0)), (shape, Cp)
        assert not bits_of(y[..., C:]).any()


def test_cl_to_u8_every_bf16_pattern(gold):
    """All 65 280 non-NaN patterns, through the vector path (3 of 8 stored channels), its scalar tail, an unaligned
    view and the generic path."""
    from videocof_b200 import ops
    bits = gold["out_dec_bf16_bits"].view(np.uint16)[0].transpose(1, 2, 3, 0)          # [2, 96, 128, 3]
    want = gold["out_frames_u8"]
    x3 = bf16_from_bits(bits).cuda()
    x8 = torch.full((2, 96, 128, 8), 3.0, dtype=torch.bfloat16, device="cuda")        # junk in the unused channels
    x8[..., :3] = x3
    assert np.array_equal(ops.cl_to_u8(x8, 3).cpu().numpy(), want)                     # vector path
    assert np.array_equal(ops.cl_to_u8(x3.contiguous(), 3).cpu().numpy(), want)        # generic path (ld = 3)
    tail = x8.reshape(1, 1, -1, 8)[:, :, :1027].contiguous()                           # 1027 = 4 * 256 + 3 positions
    assert np.array_equal(ops.cl_to_u8(tail, 3).cpu().numpy().reshape(-1, 3), want.reshape(-1, 3)[:1027])
    odd = x8.reshape(-1, 8)[1:1 + 513].reshape(1, 1, 513, 8)                           # view starting 16 B in
    assert np.array_equal(ops.cl_to_u8(odd, 3).cpu().numpy().reshape(-1, 3), want.reshape(-1, 3)[1:514])
    buf = torch.zeros(4 + 513 * 8, dtype=torch.bfloat16, device="cuda")                # rows 8 B off a 16-B boundary
    shifted = buf[4:].view(1, 1, 513, 8)
    shifted.copy_(odd)
    assert shifted.data_ptr() % 16 == 8
    assert np.array_equal(ops.cl_to_u8(shifted, 3).cpu().numpy().reshape(-1, 3), want.reshape(-1, 3)[1:514])
    x32 = torch.zeros((2, 96, 128, 32), dtype=torch.bfloat16, device="cuda")
    x32[..., :3] = x3
    assert np.array_equal(ops.cl_to_u8(x32, 3).cpu().numpy(), want)                    # generic path (ld = 32)
    out = torch.empty((2, 96, 128, 3), dtype=torch.uint8, device="cuda")
    assert ops.cl_to_u8(x8, 3, out=out) is out and np.array_equal(out.cpu().numpy(), want)


def test_cl_to_u8_matches_oracle_on_random_input():
    from videocof_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(3, 17, 23, 8, generator=g) * 0.8).to(torch.bfloat16)
    got = ops.cl_to_u8(x.cuda(), 3).cpu().numpy()
    ref = vo.model_output_to_frames(bits_of(x[..., :3]).transpose(3, 0, 1, 2)[None])[0]
    assert np.array_equal(got, ref)


def test_full_size_81f_720p_tables():
    """BASELINE config size: 81 frames x 720 x 1280.  Each output element depends on one input element only, so the
    device result must equal the oracle's table looked up per element (checked on the device)."""
    from videocof_b200 import ops
    T, H, W = 81, 720, 1280
    g = torch.Generator(device="cuda").manual_seed(1)
    frames = torch.randint(0, 256, (T, H, W, 3), generator=g, dtype=torch.uint8, device="cuda")
    y = ops.u8_to_cl(frames, 32)
    lut_in = torch.from_numpy(vo.video_to_model_input(np.arange(256, dtype=np.uint8).reshape(1, 1, 256, 1))[1]
                              .reshape(-1).view(np.int16)).cuda()
    assert torch.equal(y[..., :3].contiguous().view(torch.int16), lut_in[frames.long()])
    assert int(y[..., 3:].contiguous().view(torch.int16).abs().max()) == 0
    del y
    allbits = np.arange(65536, dtype=np.uint16)
    nan = ((allbits & 0x7f80) == 0x7f80) & ((allbits & 0x7f) != 0)
    lut_out = vo.model_output_to_frames(np.where(nan, 0, allbits).astype(np.uint16).reshape(1, 1, 1, 1, -1)).reshape(-1)
    lut_out[nan] = 0                                                         # NaN -> black (documented in vcof.h)
    lut_out = torch.from_numpy(lut_out).cuda()
    for t0 in range(0, T, 27):                                               # 27-frame slabs bound the index temporaries
        x = torch.randint(0, 65536, (27, H, W, 8), generator=g, dtype=torch.int32, device="cuda").to(torch.int16) \
            .view(torch.bfloat16)                                            # arbitrary bit patterns, NaNs included
        got = ops.cl_to_u8(x, 3)
        idx = (x[..., :3].contiguous().view(torch.int16).int() & 0xffff).long()
        assert torch.equal(got, lut_out[idx])
        del x, got, idx


def test_ops_reject_bad_arguments():
    from videocof_b200 import ops
    from videocof_b200._lib import VcofError
    with pytest.raises(VcofError):
        ops.u8_to_cl(torch.zeros(1, 2, 2, 3, device="cuda"), 32)            # not bytes
    with pytest.raises(VcofError):
        ops.u8_to_cl(torch.zeros(1, 2, 2, 3, dtype=torch.uint8, device="cuda"), 2)
    with pytest.raises(VcofError):
        ops.cl_to_u8(torch.zeros(1, 2, 2, 8, device="cuda"), 3)             # fp32
    with pytest.raises(VcofError):
        ops.cl_to_u8(torch.zeros(1, 2, 4, 8, dtype=torch.bfloat16, device="cuda")[:, :, ::2], 3)


def test_ops_reject_cpu_tensor():
    from videocof_b200 import ops
    from videocof_b200._lib import VcofError
    with pytest.raises(VcofError):
        ops.u8_to_cl(torch.zeros(1, 2, 2, 3, dtype=torch.uint8), 32)
    with pytest.raises(VcofError):
        ops.cl_to_u8(torch.zeros(1, 2, 2, 8, dtype=torch.bfloat16), 3)


@pytest.fixture(scope="module")
def vae():
    from oracle.vae_oracle import VAEConfig, make_vae_params
    from videocof_b200.vae import AutoencoderKLWan
    m = AutoencoderKLWan()
    m.load_state_dict(make_vae_params(VAEConfig(), seed=17), strict=True)
    return m.to("cuda", torch.bfloat16).eval()


def test_vae_byte_frames_bit_identical_to_float_path(vae):
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (1, 9, 32, 48, 3), generator=g, dtype=torch.uint8)
    video = frames[0].permute(3, 0, 1, 2)[None].float() * (2.0 / 255.0) - 1.0          # fast_infer.py:86-88
    with torch.no_grad():
        mu_b = vae.encode_frames(frames.cuda())[0].mode()
        mu_f = vae.encode(video.cuda().to(torch.bfloat16))[0].mode()                    # pipeline_wan.py:397, 406
        z = torch.randn(1, 16, 3, 4, 6, generator=g).to(torch.bfloat16).cuda()
        dec = vae.decode(z).sample
        dec_u8 = vae.decode_frames(z)
    assert torch.equal(mu_b, mu_f)
    assert tuple(dec_u8.shape) == (1, 9, 32, 48, 3) and dec_u8.dtype == torch.uint8
    # the reference's host chain on the decoder output: decode_latents + save_videos_grid
    f32 = (dec / 2 + 0.5).clamp(0, 1).cpu().float().numpy()
    ref = (f32 * 255).astype(np.uint8).transpose(0, 2, 3, 4, 1)
    assert np.array_equal(dec_u8.cpu().numpy(), ref)
    assert np.array_equal(ref, vo.model_output_to_frames(bits_of(dec)))


def test_pipeline_bytes_in_bytes_out(vae):
    """WanPipeline(video=uint8 frames, output_type="uint8") == the float path followed by the reference's byte
    conversion, bit for bit (same seeds)."""
    from oracle.dit_oracle import DiTConfig, make_dit_params
    from videocof_b200.dit import WanTransformer3DModel
    from videocof_b200.pipeline import WanPipeline
    from videocof_b200.scheduler import FlowUniPCMultistepScheduler
    dcfg = DiTConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32)
    dit = WanTransformer3DModel(**dcfg.to_kwargs())
    dit.load_state_dict(make_dit_params(dcfg, seed=5), strict=True)
    dit = dit.to("cuda", torch.bfloat16).eval()
    g = torch.Generator().manual_seed(2)
    frames = torch.randint(0, 256, (1, 9, 32, 48, 3), generator=g, dtype=torch.uint8)
    video = frames[0].permute(3, 0, 1, 2)[None].float() * (2.0 / 255.0) - 1.0
    ctx = [torch.randn(6, 64, generator=g).bfloat16().cuda()]
    outs = {}
    for kind, vid, otype in (("float", video, "numpy"), ("bytes", frames, "uint8")):
        pipe = WanPipeline(None, None, vae, dit, FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1))
        outs[kind] = pipe(video=vid, prompt_embeds=ctx, height=32, width=48, source_frames=9, reasoning_frames=4,
                          num_inference_steps=2, guidance_scale=1.0, shift=3, repeat_rope=True, cot=True,
                          generator=torch.Generator().manual_seed(11), output_type=otype)
    f, b = outs["float"], outs["bytes"]
    assert b.videos.dtype == torch.uint8 and tuple(b.videos.shape) == (1, 10, 32, 48, 3)
    for name in ("videos", "ground_videos", "edit_videos"):
        ref = (getattr(f, name) * 255).numpy().astype(np.uint8).transpose(0, 2, 3, 4, 1)   # utils/utils.py:66
        assert np.array_equal(getattr(b, name).numpy(), ref), name
