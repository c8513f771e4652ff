"""Per-kernel parity of libvcof (through the C ABI) against plain torch fp32 math on the GPU.
Cases and tolerances live in tools/gpu_probe.py (shared with the stand-alone probe)."""
import json

import pytest

import gpu_probe

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", gpu_probe.cases(), ids=lambda c: json.dumps(c, separators=(",", ":")))
def test_kernel_case(case):
    res = gpu_probe.run_case(case)
    assert res.get("ok"), res


def test_attention_rows_do_not_depend_on_the_query_partition():
    """A query row's output is a function of that row and the key / value sequence only: shifting which rows share a
    tile (and a warp) must not change a single bit, also when the lazy rescale fires at every key block (keys growing
    block by block).  This is what makes the sequence-parallel shards — whose tiles start at other row offsets —
    reproduce the single-GPU result bit for bit for ANY input."""
    import torch
    from videocof_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    Lq, Lk, heads = 600, 1100, 2
    q = torch.randn(Lq, heads * 128, generator=g, device="cuda").bfloat16()
    k = torch.randn(Lk, heads * 128, generator=g, device="cuda")
    k = (k * (1.0 + 0.35 * (torch.arange(Lk, device="cuda") // 128)[:, None])).bfloat16()   # maxima grow per block
    v = torch.randn(Lk, heads * 128, generator=g, device="cuda").bfloat16()
    rows = torch.randn(Lq, generator=g, device="cuda") > 0.8        # a few rows with much larger scores than their neighbours
    q[rows] *= 6
    full = ops.attention(q, k, v, heads)
    for off in (8, 40, 168, 333):
        part = ops.attention(q[off:].contiguous(), k, v, heads)
        assert torch.equal(part, full[off:]), off
    assert torch.isfinite(full.float()).all()
