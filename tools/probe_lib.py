"""ctypes binding of tests/native/libvcof_probes.so (tests/native/probes/vcof_probes.h): the hardware probes used while
designing libvcof.  They live outside the product library; error text comes from libvcof's vcof_last_error()."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "tests", "native", "libvcof_probes.so")
c_void_p, c_int = ctypes.c_void_p, ctypes.c_int
SIGNATURES = {
    "vcof_debug_tma_probe": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int,
                             c_int, c_void_p, c_int, c_int, c_int, c_void_p],
    "vcof_debug_umma_probe": [c_void_p, c_void_p, c_void_p, c_int, c_void_p],
}
_lib = None


def load():
    global _lib
    if _lib is None:
        from videocof_b200 import _lib as product
        product.load()                         # libvcof.so first: the probes link against its runtime helpers
        if not os.path.exists(PATH):
            raise product.VcofError(f"{PATH} not built: run tests/native/build.sh (or __graft_entry__.build())")
        lib = ctypes.CDLL(PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = c_int, argtypes
        _lib = lib
    return _lib


def call(name, *args):
    from videocof_b200 import _lib as product
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise product.VcofError(f"{name} failed ({rc}): {product.load().vcof_last_error().decode()}")
