"""videocof_b200 — B200-native (sm_100a) implementation of the VideoCoF denoising hot path.

Only what the path needs lives here: `csrc/` (hand-written CUDA kernels + the C ABI of
include/vcof.h), `_lib` (ctypes binding), `ops` (tensor-level wrappers), `dit` / `vae`
(host-side mirrors of the reference's videox_fun.models classes), `pipeline` (the
denoise loop of videox_fun.pipeline.WanPipeline) and `dist` (sequence-parallel plumbing).
There is no CPU fallback: every op raises if libvcof or a CUDA device is missing.
"""
__version__ = "0.1.0"
