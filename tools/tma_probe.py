"""TMA feed-rate microbenchmark (vcof_debug_tma_probe): bytes/clk/SM for the box geometries libvcof uses."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from videocof_b200 import _lib as _product  # noqa: E402
import probe_lib as _lib  # noqa: E402


def run(name, base, dims, strides, box, swz, iters, coords, step_dim, step, wrap, grid=148, producers=1, flags=0):
    rank = len(dims)
    cyc = torch.zeros(grid, dtype=torch.int64, device="cuda")
    L, I = ctypes.c_longlong, ctypes.c_int
    args = (base.data_ptr(), rank, (L * rank)(*dims), (L * max(rank - 1, 1))(*strides), (I * rank)(*box), swz, iters,
            (I * 5)(*(list(coords) + [0] * (5 - len(coords)))), step_dim, step, wrap, cyc.data_ptr(), grid, producers, flags,
            torch.cuda.current_stream().cuda_stream)
    try:
        _lib.call("vcof_debug_tma_probe", *args)
        torch.cuda.synchronize()
        _lib.call("vcof_debug_tma_probe", *args)
        torch.cuda.synchronize()
    except _product.VcofError as e:           # e.g. the ring does not fit this many producers
        print(json.dumps(dict(case=name, producers=producers, skipped=str(e)[-60:])), flush=True)
        return
    nbytes = 2 * max(flags & 15, 1)
    for b in box:
        nbytes *= b
    c = cyc.float().mean().item()
    print(json.dumps(dict(case=name, box=box, swizzle=swz, box_bytes=nbytes, rows=nbytes // (box[0] * 2),
                          row_bytes=box[0] * 2, producers=producers, flags=flags, cycles_per_box=round(c / iters, 1),
                          bytes_per_clk_per_sm=round(nbytes * iters * producers / c, 1))), flush=True)


def sweep():
    """Per-box cost vs rows / row bytes / number of issuing lanes: separates a per-box latency from a per-row cost and
    shows whether boxes issued by different warps overlap."""
    dev = "cuda"
    iters = 3000
    a = torch.zeros(75600, 5120, dtype=torch.bfloat16, device=dev)
    for rows in (128, 256):
        for prod in (1, 2, 3, 4):
            run(f"2-D {rows} x 128B", a, (5120, 75600), (5120,), (64, rows), 128, iters, (0, 0), 1, rows, 60000,
                producers=prod)
    for rows in (64, 256):
        for prod in (1, 4):
            run(f"2-D {rows} x 64B", a, (5120, 75600), (5120,), (32, rows), 64, iters, (0, 0), 1, rows, 60000,
                producers=prod)
    T, H, W = 5, 360, 640
    for C, cb, swz in ((96, 32, 64), (96, 64, 128), (192, 64, 128)):
        x = torch.zeros(T, H, W, C, dtype=torch.bfloat16, device=dev)
        dims = (C, W, 1, H, T)
        strides = (C, W * C, W * C, H * W * C)
        for tg in (1, 3):
            for prod in (1, 2, 3, 4):
                run(f"conv C={C} [{cb}c,16w,1,8h,{tg}t]", x, dims, strides, (cb, 16, 1, 8, tg), swz, iters,
                    (0, 0, 0, 0, 0), 3, 8, H - 8, producers=prod)
        for prod in (1, 4):
            run(f"conv C={C} [{cb}c,16w,1,2h,1t] (32 rows)", x, dims, strides, (cb, 16, 1, 2, 1), swz, iters,
                (0, 0, 0, 0, 0), 3, 2, H - 2, producers=prod)


def sweep2():
    """Is the ~520-clock floor per box, per (lane, tensor map) or per barrier round trip?  flags: low 4 bits = boxes per
    round trip, 16 = alternate two map copies, 32 = poll with test_wait."""
    a = torch.zeros(75600, 5120, dtype=torch.bfloat16, device="cuda")
    for rows in (32, 128):
        for flags in (1, 2, 4, 2 | 16, 4 | 16, 1 | 32, 4 | 32):
            run(f"2-D {rows} x 128B flags={flags}", a, (5120, 75600), (5120,), (64, rows), 128, 3000, (0, 0), 1, rows,
                60000, producers=1, flags=flags)


def main():
    if "--sweep" in sys.argv:
        return sweep()
    if "--sweep2" in sys.argv:
        return sweep2()
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
