"""Drop-in overlay of the reference's `videox_fun` package (knightyxp/VideoCoF).

Put this repository BEFORE the reference checkout on PYTHONPATH: `videox_fun.models` /
`videox_fun.pipeline` / `videox_fun.utils.fm_solvers_unipc` / `videox_fun.utils.lora_utils` then resolve
to the B200-native implementations in `videocof_b200`, while every other sub-module the CLIs import
(`utils.fp8_optimization`, `utils.utils`, `data.dataset_image_video`, …) still resolves to the UNMODIFIED reference files through the
extended package search path.  See INTEGRATION.md.
"""
import os
import sys


def _reference_root():
    env = os.environ.get("VIDEOCOF_REFERENCE_ROOT")
    if env and os.path.isdir(os.path.join(env, "videox_fun")):
        return env
    here = os.path.dirname(os.path.abspath(__file__))
    for p in sys.path:
        cand = os.path.join(os.path.abspath(p or "."), "videox_fun")
        if os.path.isdir(cand) and os.path.abspath(cand) != here and os.path.exists(os.path.join(cand, "models")):
            return os.path.dirname(cand)
    return None


REFERENCE_ROOT = _reference_root()


def extend_with_reference(path_list, *sub):
    """Append the reference's matching package directory so un-overridden modules are found there."""
    if REFERENCE_ROOT is not None:
        cand = os.path.join(REFERENCE_ROOT, "videox_fun", *sub)
        if os.path.isdir(cand) and cand not in path_list:
            path_list.append(cand)


extend_with_reference(__path__)
