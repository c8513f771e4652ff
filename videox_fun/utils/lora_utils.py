"""Overlay of videox_fun.utils.lora_utils: `merge_lora` / `unmerge_lora` run on the GPU through libvcof
(videocof_b200/lora.py); everything else the reference module defines is re-exported from the reference file when
a reference checkout is reachable (VIDEOCOF_REFERENCE_ROOT, see videox_fun/__init__.py)."""
import importlib.util
import os

from videocof_b200.lora import merge_lora, unmerge_lora  # noqa: F401

from .. import REFERENCE_ROOT

if REFERENCE_ROOT is not None:
    _path = os.path.join(REFERENCE_ROOT, "videox_fun", "utils", "lora_utils.py")
    if os.path.exists(_path):
        try:
            _spec = importlib.util.spec_from_file_location("videox_fun.utils._reference_lora_utils", _path)
            _ref = importlib.util.module_from_spec(_spec)
            _spec.loader.exec_module(_ref)
            for _n in dir(_ref):
                if not _n.startswith("_") and _n not in ("merge_lora", "unmerge_lora"):
                    globals()[_n] = getattr(_ref, _n)
        except Exception as _exc:      # noqa: BLE001 — the reference module's own imports (diffusers, cv2 …) may be
            import warnings            # missing or of another version; what the overlay itself exports needs none of it
            warnings.warn(f"videox_fun overlay: optional re-export of {_path} skipped ({_exc!r})")
