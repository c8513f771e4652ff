"""ctypes binding of libvcof.so (include/vcof.h).  Fails loudly: no fallback path."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# VCOF_LIB: another build of libvcof.so (A/B timing of two kernel revisions inside one GPU call); default: the in-tree one
LIB_PATH = os.environ.get("VCOF_LIB") or os.path.join(_HERE, "csrc", "libvcof.so")
_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float

# name -> argtypes; mirrors include/vcof.h one to one (tests check both directions)
SIGNATURES = {
    "vcof_gemm_bf16": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll,
                       c_int, c_int, c_int, c_int, c_void_p],
    "vcof_attn_fwd": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll,
                      c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p],
    "vcof_ln_modulate": [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll,
                         c_int, c_int, c_float, c_void_p],
    "vcof_rmsnorm_rope": [c_void_p, c_ll, c_void_p, c_float, c_int, c_int, c_int, c_void_p,
                          c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "vcof_rmsnorm_rope_blocked": [c_void_p, c_ll, c_void_p, c_int, c_ll, c_void_p, c_float, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "vcof_copy_blocked": [c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_int, c_int, c_void_p],
    "vcof_patchify": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vcof_unpatchify": [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vcof_conv_igemm": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                        c_void_p, c_void_p, c_void_p, c_ll, c_float, c_void_p, c_void_p, c_void_p],
    "vcof_conv_lines": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                        c_void_p, c_ll, c_float, c_void_p, c_void_p, c_void_p],
    "vcof_rms_silu_cl": [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_ll, c_int, c_int, c_void_p],
    "vcof_nchw_to_cl": [c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p, c_void_p, c_void_p],
    "vcof_cl_to_nchw": [c_void_p, c_ll, c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p],
    "vcof_softmax_rows": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_float, c_void_p],
    "vcof_vae_attn": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_float, c_void_p],
    "vcof_embed_rows": [c_void_p, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p],
    "vcof_t5_rmsnorm": [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_ll, c_int, c_float, c_void_p],
    "vcof_t5_attn": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_void_p,
                     c_int, c_int, c_int, c_int, c_void_p],
    "vcof_rmsnorm_rope_scatter": [c_void_p, c_ll, c_void_p, c_int, c_void_p, c_float, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "vcof_attn_fwd_scatter": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_ll, c_int, c_int,
                              c_int, c_int, c_int, c_float, c_void_p],
    "vcof_copy_scatter": [c_void_p, c_ll, c_void_p, c_int, c_ll, c_int, c_void_p],
    "vcof_copy_rows_scatter": [c_void_p, c_ll, c_void_p, c_int, c_ll, c_int, c_void_p],
    "vcof_cl_to_u8": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_void_p],
    "vcof_u8_to_cl": [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p],
    "vcof_debug_frame_u8_host": [c_void_p, c_void_p, c_ll],
    "vcof_debug_video_bf16_host": [c_void_p, c_void_p, c_ll],
    "vcof_linear_f32": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                        c_void_p],
}


class VcofError(RuntimeError):
    pass


def load():
    """Return the loaded CDLL; raise VcofError if libvcof.so is absent or broken."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise VcofError(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `python videocof_b200/build.py`). There is no CPU/PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.vcof_last_error.restype = ctypes.c_char_p
        lib.vcof_last_error.argtypes = []
        lib.vcof_abi_version.restype = c_int
        lib.vcof_abi_version.argtypes = []
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError -> symbol missing: loud
            fn.restype = c_int
            fn.argtypes = args
        _lib = lib
    return _lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise VcofError(f"{name} failed ({rc}): {lib.vcof_last_error().decode()}")
