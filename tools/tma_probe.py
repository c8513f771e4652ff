"""TMA feed-rate microbenchmark (vcof_debug_tma_probe): bytes/clk/SM for the box geometries libvcof uses."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videocof_b200 import _lib  # noqa: E402


def run(name, base, dims, strides, box, swz, iters, coords, step_dim, step, wrap, grid=148):
    rank = len(dims)
    cyc = torch.zeros(grid, dtype=torch.int64, device="cuda")
    L, I = ctypes.c_longlong, ctypes.c_int
    args = (base.data_ptr(), rank, (L * rank)(*dims), (L * max(rank - 1, 1))(*strides), (I * rank)(*box), swz, iters,
            (I * 5)(*(list(coords) + [0] * (5 - len(coords)))), step_dim, step, wrap, cyc.data_ptr(), grid,
            torch.cuda.current_stream().cuda_stream)
    _lib.call("vcof_debug_tma_probe", *args)
    torch.cuda.synchronize()
    _lib.call("vcof_debug_tma_probe", *args)
    torch.cuda.synchronize()
    nbytes = 2
    for b in box:
        nbytes *= b
    c = cyc.float().mean().item()
    print(json.dumps(dict(case=name, box=box, swizzle=swz, box_bytes=nbytes, rows=nbytes // (box[0] * 2),
                          row_bytes=box[0] * 2, cycles_per_box=c / iters, bytes_per_clk_per_sm=nbytes * iters / c)))


def main():
    dev = "cuda"
    iters = 4000
    # GEMM-style: [M, K] row-major, box 64 x 128 (128 B rows, contiguous K)
    a = torch.zeros(75600, 5120, dtype=torch.bfloat16, device=dev)
    run("gemm A 128x64 (128B rows, pitch 10KB)", a, (5120, 75600), (5120,), (64, 128), 128, iters, (0, 0), 1, 128, 75000)
    run("gemm B 256x64", a, (5120, 75600), (5120,), (64, 256), 128, iters, (0, 0), 1, 256, 75000)
    # conv-style channels-last [T,H,W,C]
    for C, cb, swz in ((96, 32, 64), (192, 32, 64), (192, 64, 128), (128, 64, 128), (384, 64, 128)):
        T, H, W = 5, 360, 640
        x = torch.zeros(T, H, W, C, dtype=torch.bfloat16, device=dev)
        dims = (C, W, 1, H, T)
        strides = (C, W * C, W * C, H * W * C)
        run(f"conv box C={C} chunk={cb} [c,16w,1,8h,1]", x, dims, strides, (cb, 16, 1, 8, 1), swz, iters, (0, 0, 0, 0, 0),
            3, 8, H - 8)
        run(f"conv row box C={C} chunk={cb} [c,128w,1,1h,1]", x, dims, strides, (cb, 128, 1, 1, 1), swz, iters,
            (0, 0, 0, 0, 0), 3, 1, H - 1)
    # channel-slice-major layout [T, H, C/32, W, 32]: a 16-pixel x 32-channel slice is one contiguous 1 KB run
    T, H, W, Cs = 5, 360, 640, 3
    y = torch.zeros(T, H, Cs, W, 32, dtype=torch.bfloat16, device=dev)
    run("slice-major [32, 16w, 1cs, 8h, 1t]", y, (32, W, Cs, H, T), (32, W * 32, Cs * W * 32, H * Cs * W * 32),
        (32, 16, 1, 8, 1), 64, iters, (0, 0, 0, 0, 0), 3, 8, H - 8)
    y2 = torch.zeros(T, H, 2, W, 64, dtype=torch.bfloat16, device=dev)
    run("slice-major 64ch [64, 16w, 1cs, 8h, 1t]", y2, (64, W, 2, H, T), (64, W * 64, 2 * W * 64, H * 2 * W * 64),
        (64, 16, 1, 8, 1), 128, iters, (0, 0, 0, 0, 0), 3, 8, H - 8)


if __name__ == "__main__":
    main()
