"""Frame bytes at the two ends of the pipeline with the reference CLI's function names.

The reference reads a clip with imageio, scales it to fp32 [-1, 1] on the host and ships fp32 to the GPU
(fast_infer.py:43-91); after the run it pulls fp32 [0, 1] frames back and converts them to bytes on the host
(pipeline_wan.py:425-427, utils/utils.py:59-83, fast_infer.py:170-206).  Here the bytes themselves cross PCIe
(4x less traffic each way) and both conversions are libvcof kernels fused with the VAE's layout passes
(`vcof_u8_to_cl`, `vcof_cl_to_u8`), bit-identical to the reference's chain:

    video, h, w = load_video_frames(path, source_frames=33, as_uint8=True)        # uint8 [1, T, H, W, 3], pinned
    out = pipeline(video=video, ..., output_type="uint8").videos                  # uint8 [1, T', H, W, 3]
    save_results(out, "out.mp4", fps)

Every function also accepts what the reference's function of the same name accepts (fp32 [B, C, T, H, W]) and then
does what the reference does; file decode / encode stays imageio on the host as in the reference (NVDEC / NVENC are
not reachable from this image).  Host-side code only: no kernels are launched from this module.
"""
import os

import numpy as np
import torch


def _imageio():
    try:
        import imageio
    except ImportError as e:  # the reference imports it at module scope (fast_infer.py:12)
        raise ImportError("imageio is needed to read / write video files, as in the reference CLI") from e
    return imageio


def _read_indices(total_frames, source_frames, start_frame):
    stride = max(1, total_frames // source_frames)
    return [start_frame + i * stride for i in range(source_frames) if start_frame + i * stride < total_frames]


def select_frame_indices(total_frames, source_frames, start_frame):
    """fast_infer.py:57-83 — evenly strided indices from `start_frame`; a clip that runs out repeats its last frame."""
    picked = _read_indices(total_frames, source_frames, start_frame)
    return picked + picked[-1:] * (source_frames - len(picked))


def load_video_frames(video_path, source_frames=None, as_uint8=False):
    """fast_infer.py:43-91.  -> (video, original_height, original_width).

    as_uint8=False: the reference's return value, fp32 [1, 3, T, H, W] in [-1, 1].
    as_uint8=True : the stacked byte frames uint8 [1, T, H, W, 3] (pinned when CUDA is present) for
                    `WanPipeline(video=...)` / `AutoencoderKLWan.encode_frames`; the scaling happens on the device."""
    assert source_frames is not None, "source_frames is required"
    imageio = _imageio()
    reader = imageio.get_reader(video_path)
    try:
        total_frames = reader.count_frames()
    except Exception:
        total_frames = sum(1 for _ in reader)
        reader = imageio.get_reader(video_path)
    stride = max(1, total_frames // source_frames)
    # same draw as the reference (:59), so a seeded run picks the same frames
    start_frame = torch.randint(0, max(1, total_frames - stride * source_frames), (1,))[0].item()
    frames = []
    for idx in _read_indices(total_frames, source_frames, start_frame):
        try:
            frames.append(np.asarray(reader.get_data(idx)))
        except IndexError:      # count_frames over-reported: stop and pad like the reference (:74-75)
            break
    reader.close()
    original_height = original_width = None
    if frames:
        original_height, original_width = frames[0].shape[:2]
        print(f"Original video dimensions: {original_width}x{original_height}")
    else:                       # unreadable clip: black 832x480 frames, dimensions stay None (:81-82)
        frames = [np.zeros((480, 832, 3), dtype=np.uint8)]
    frames += frames[-1:] * (source_frames - len(frames))
    print(f"Loaded {source_frames} source frames")
    stacked = torch.from_numpy(np.stack(frames))                        # [T, H, W, 3] uint8
    if as_uint8:
        video = stacked.unsqueeze(0)
        if torch.cuda.is_available():
            video = video.pin_memory()
        return video, original_height, original_width
    video = stacked.permute([3, 0, 1, 2]).unsqueeze(0).float()
    video = video * (2.0 / 255.0) - 1.0
    return video, original_height, original_width


# ---- bytes for the writers ------------------------------------------------------------------------------------

def _grid_u8(frames, n_rows, padding=2):
    """torchvision.utils.make_grid (padding 2, pad value 0) on byte images [B, H, W, C] -> [H', W', C]; one image is
    returned as it is (utils/utils.py:63)."""
    B, H, W, C = frames.shape
    if B == 1:
        return frames[0]
    xmaps = min(n_rows, B)
    ymaps = -(-B // xmaps)
    hh, ww = H + padding, W + padding
    grid = np.zeros((hh * ymaps + padding, ww * xmaps + padding, C), dtype=np.uint8)
    for k in range(B):
        y, x = divmod(k, xmaps)
        grid[y * hh + padding:y * hh + padding + H, x * ww + padding:x * ww + padding + W] = frames[k]
    return grid


def _to_u8_bthwc(videos, rescale=False):
    """Byte frames [B, T, H, W, C] from either form: uint8 [B, T, H, W, C] (already bytes) or the reference's float
    [B, C, T, H, W] in [0, 1] ([-1, 1] with rescale), converted as utils/utils.py:64-66 does."""
    if isinstance(videos, np.ndarray):
        videos = torch.from_numpy(videos)
    if videos.dtype == torch.uint8:
        if rescale:
            raise ValueError("rescale applies to float videos only")
        return videos.cpu().numpy()
    x = videos.detach().cpu().float()
    if rescale:
        x = (x + 1.0) / 2.0
    return (x * 255).numpy().astype(np.uint8).transpose(0, 2, 3, 4, 1)


def save_videos_grid(videos, path, rescale=False, n_rows=6, fps=12, imageio_backend=True,
                     color_transfer_post_process=False):
    """utils/utils.py:59-83 for uint8 [B, T, H, W, 3] (device-converted) or float [B, C, T, H, W] videos."""
    from PIL import Image
    u8 = _to_u8_bthwc(videos, rescale)
    outputs = [Image.fromarray(np.ascontiguousarray(_grid_u8(u8[:, t], n_rows))) for t in range(u8.shape[1])]
    if color_transfer_post_process:
        from videox_fun.utils.utils import color_transfer          # the reference's own (cv2) implementation
        for i in range(1, len(outputs)):
            outputs[i] = Image.fromarray(color_transfer(np.uint8(outputs[i]), np.uint8(outputs[0])))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if imageio_backend:
        imageio = _imageio()
        if path.endswith("mp4"):
            imageio.mimsave(path, outputs, fps=fps)
        else:
            imageio.mimsave(path, outputs, duration=(1000 * 1 / fps))
    else:
        if path.endswith("mp4"):
            path = path.replace(".mp4", ".gif")
        outputs[0].save(path, format="GIF", append_images=outputs, save_all=True, duration=100, loop=0)


def save_results(tensor, file_path, fps_out=16):
    """fast_infer.py:170-180: one frame -> an image file, otherwise a video."""
    from PIL import Image
    os.makedirs(os.path.dirname(file_path), exist_ok=True)
    if (tensor.shape[1] if tensor.dtype == torch.uint8 else tensor.shape[2]) == 1:
        Image.fromarray(np.ascontiguousarray(_to_u8_bthwc(tensor[:1])[0, 0])).save(file_path)
    else:
        save_videos_grid(tensor, file_path, fps=fps_out)
    print(f"Saved video → {file_path}")


def _compare_lut(rescaled):
    """What save_side_by_side makes of an input byte u that went through load_video_frames (fast_infer.py:86-88,
    183-189, utils/utils.py:66): trunc(255 * clamp((fp32(u) * fp32(2/255) - 1 [+ 1) / 2], 0, 1)), as a 256-entry table
    built with the same fp32 torch operations."""
    x = torch.arange(256, dtype=torch.uint8).float() * (2.0 / 255.0) - 1.0
    if rescaled:
        x = (x + 1.0) / 2.0
    return (x.clamp(0.0, 1.0) * 255).numpy().astype(np.uint8)


def save_side_by_side(input_tensor, sample_tensor, file_path, fps_out=16):
    """fast_infer.py:192-206: input clip and result next to each other, cropped to the common T / H / W.

    Byte inputs (uint8 [B, T, H, W, 3]) stay bytes: the input clip goes through the table of `_compare_lut`, which
    reproduces the reference's float round trip (its [-1, 1] -> [0, 1] mapping is skipped when no byte is below 128,
    fast_infer.py:184-188 — kept), the result is used as it is."""
    if input_tensor.dtype != torch.uint8 and sample_tensor.dtype != torch.uint8:
        def norm(v):
            v = v.detach().cpu()
            if float(v.min()) < 0.0 or float(v.max()) > 1.0:
                v = (v + 1.0) / 2.0
            return v.clamp(0.0, 1.0)
        a, b = norm(input_tensor), norm(sample_tensor)
        T, H, W = (min(a.shape[i], b.shape[i]) for i in (2, 3, 4))
        combined = torch.cat([a[:, :, :T, :H, :W], b[:, :, :T, :H, :W]], dim=4)
    else:
        if input_tensor.dtype != torch.uint8 or sample_tensor.dtype != torch.uint8:
            raise ValueError("save_side_by_side: pass both clips as uint8 [B, T, H, W, 3] or both as float [B, C, T, H, W]")
        a = input_tensor.cpu().numpy()
        a = _compare_lut(rescaled=bool(a.min() < 128))[a]
        b = sample_tensor.cpu().numpy()
        T, H, W = (min(a.shape[i], b.shape[i]) for i in (1, 2, 3))
        combined = torch.from_numpy(np.concatenate([a[:, :T, :H, :W], b[:, :T, :H, :W]], axis=3))
    os.makedirs(os.path.dirname(file_path), exist_ok=True)
    save_videos_grid(combined, file_path, fps=fps_out)
    print(f"Saved side-by-side video → {file_path}")
