"""Overlay of videox_fun.utils.utils: `save_videos_grid` also accepts the byte frames uint8 [B, T, H, W, 3] that
`WanPipeline(..., output_type="uint8")` returns (converted on the GPU, videocof_b200/video_io.py) and behaves like the
reference's function for float videos; everything else the reference module defines (`filter_kwargs`, the image /
video latent helpers, …) is re-exported from the reference file when a reference checkout is reachable
(VIDEOCOF_REFERENCE_ROOT, see videox_fun/__init__.py)."""
import importlib.util
import inspect
import os

from videocof_b200.video_io import save_videos_grid  # noqa: F401

from .. import REFERENCE_ROOT


def filter_kwargs(cls, kwargs):
    """reference utils/utils.py:17-21 — keep the kwargs the constructor names (needed by the CLIs before any model
    exists, so it must not depend on the reference module's own imports resolving)."""
    valid = set(inspect.signature(cls.__init__).parameters.keys()) - {"self", "cls"}
    return {k: v for k, v in kwargs.items() if k in valid}


if REFERENCE_ROOT is not None:
    _path = os.path.join(REFERENCE_ROOT, "videox_fun", "utils", "utils.py")
    if os.path.exists(_path):
        try:
            _spec = importlib.util.spec_from_file_location("videox_fun.utils._reference_utils", _path)
            _ref = importlib.util.module_from_spec(_spec)
            _spec.loader.exec_module(_ref)
            for _n in dir(_ref):
                if not _n.startswith("_") and _n != "save_videos_grid":
                    globals()[_n] = getattr(_ref, _n)
        except Exception as _exc:      # noqa: BLE001 — the reference module's own imports (diffusers, cv2 …) may be
            import warnings            # missing or of another version; what the overlay itself exports needs none of it
            warnings.warn(f"videox_fun overlay: optional re-export of {_path} skipped ({_exc!r})")
