"""Does a tcgen05 smem descriptor accept a K-major SW128 tile whose start is shifted by whole 128-byte rows inside a
resident, absolutely-swizzled buffer (the per-tap "row-shifted view" an A-resident implicit-GEMM convolution needs)?
Runs vcof_debug_umma_probe for shift = 1..7 with and without the descriptor's base-offset field and reports the max
error against A[shift:shift+128] @ B^T."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import probe_lib as _lib  # noqa: E402  (tests/native/libvcof_probes.so)


def main():
    g = torch.Generator().manual_seed(0)
    a = torch.randn(136, 64, generator=g).bfloat16().cuda()
    b = torch.randn(128, 64, generator=g).bfloat16().cuda()
    d = torch.empty(128, 128, dtype=torch.float32, device="cuda")
    for shift in range(0, 8):
        for bo in (0, 1):
            mode = (shift << 4) | (bo << 7)
            d.zero_()
            _lib.call("vcof_debug_umma_probe", a.data_ptr(), b.data_ptr(), d.data_ptr(), mode,
                      torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            ref = a[shift:shift + 128].float() @ b.float().t()
            err = float((d - ref).abs().max())
            print(json.dumps(dict(shift=shift, base_offset_field=bool(bo), max_abs_err=err, ok=err < 1e-3)), flush=True)


if __name__ == "__main__":
    main()
