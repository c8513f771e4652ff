"""CPU restatement (numpy, integer / byte arithmetic) of the frame-byte ends of the reference pipeline — SURVEY.md §8f
rank 4.  Test infrastructure only: the product never imports this (tests/test_boundary_cpu.py).

Follows, in /root/reference:
  * frame selection and padding   fast_infer.py:50-83    (`select_frame_indices`)
  * bytes -> model input          fast_infer.py:86-88    uint8 [T,H,W,3] -> fp32 [1,3,T,H,W] = u * fp32(2/255) - 1,
                                  then the pipeline's cast to the VAE dtype, videox_fun/pipeline/pipeline_wan.py:397
                                  (`video_to_model_input`)
  * decoder output -> bytes       videox_fun/pipeline/pipeline_wan.py:425-427  bf16 (x / 2 + 0.5).clamp(0, 1) -> fp32,
                                  videox_fun/utils/utils.py:59-68 / fast_infer.py:175-177  (x * 255).astype(uint8)
                                  (`model_output_to_frames`)

Pinned bit-exactly against the executed reference functions by tests/golden/video_io.npz
(tools/gen_golden_video_io.py): all 256 input bytes, all 65 280 non-NaN bf16 decoder outputs, 12 selection cases.
"""
import numpy as np


def f32_to_bf16_bits(x):
    """Round-to-nearest-even truncation of fp32 to bf16, as torch's .to(bfloat16); returns uint16 bit patterns."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = (u + 0x7fff + ((u >> 16) & 1)) >> 16
    return r.astype(np.uint16)


def bf16_bits_to_f32(b):
    return (np.asarray(b).astype(np.uint16).astype(np.uint32) << 16).view(np.float32)


def _bf16(x):
    """fp32 value rounded to bf16, kept as fp32 (torch's per-op rounding for bf16 tensors)."""
    return bf16_bits_to_f32(f32_to_bf16_bits(x))


def select_frame_indices(total_frames, source_frames, start_frame):
    """fast_infer.py:58-83 with the random start given: indices read from the file, the last one repeated until
    `source_frames` are present (:77-82; an empty video gives black frames there — not representable as indices)."""
    stride = max(1, total_frames // source_frames)
    picked = []
    for i in range(source_frames):
        idx = start_frame + i * stride
        if idx >= total_frames:
            break
        picked.append(idx)
    while picked and len(picked) < source_frames:
        picked.append(picked[-1])
    return picked


def start_frame_bound(total_frames, source_frames):
    """Exclusive upper bound of the torch.randint draw (:59)."""
    stride = max(1, total_frames // source_frames)
    return max(1, total_frames - stride * source_frames)


def video_to_model_input(frames_u8):
    """uint8 [T, H, W, C] -> (fp32 [1, C, T, H, W], uint16 bf16 bits of the same) (fast_infer.py:86-88, pipeline_wan.py:397)."""
    x = np.asarray(frames_u8).astype(np.float32)
    x = x * np.float32(2.0 / 255.0)          # python double scalar enters the fp32 multiply as fp32
    x = x - np.float32(1.0)
    x = np.ascontiguousarray(np.transpose(x, (3, 0, 1, 2))[None])
    return x, f32_to_bf16_bits(x)


def model_output_to_frames(dec_bf16_bits):
    """bf16 bits of the decoder output [B, C, T, H, W] -> uint8 [B, T, H, W, C]."""
    x = bf16_bits_to_f32(np.asarray(dec_bf16_bits).view(np.uint16))
    with np.errstate(over="ignore", invalid="ignore"):
        t = _bf16(x / np.float32(2.0))            # frames / 2        (bf16 op)
        t = _bf16(t + np.float32(0.5))            # + 0.5             (bf16 op)
        t = np.clip(t, np.float32(0.0), np.float32(1.0))             # .clamp(0, 1)
        y = (t * np.float32(255.0)).astype(np.uint8)                   # fp32 product, C truncation
    return np.ascontiguousarray(np.transpose(y, (0, 2, 3, 4, 1)))
