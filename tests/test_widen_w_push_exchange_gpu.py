"""Scatter entry points of the push exchange (vcof_rmsnorm_rope_scatter / vcof_copy_scatter / vcof_copy_rows_scatter,
videocof_b200/dist.py) on ONE GPU: with local slabs as destinations they must reproduce, bit for bit, the column-blocked
kernels of the validated head exchange (the arithmetic is the same code; only the store addresses differ).  The
multi-GPU side — peer-mapped slabs, barriers — is `torchrun ... tools/sp_check.py push`."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rope(F, H, W, hd, dev, row_offset=0):
    from videocof_b200 import ops
    from videocof_b200.dit import rope_params
    d = hd
    freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                       rope_params(1024, 2 * (d // 6))], dim=1)
    table = torch.stack([freqs.real, freqs.imag], dim=-1).to(torch.float32).contiguous().to(dev)
    c = freqs.shape[1]
    tpos = torch.tensor([1, 2, 0, 1, 2][:F], dtype=torch.int32, device=dev)
    return ops.RopeSpec(table, tpos, F, H, W, c - 2 * (c // 3), c // 3, row_offset)


@pytest.mark.parametrize("P,heads,rope_on", [(2, 4, True), (4, 8, True), (8, 8, False), (1, 2, True)])
def test_rmsnorm_rope_scatter_equals_blocked(P, heads, rope_on):
    from videocof_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(P * 100 + heads)
    F, H, W, hd = 5, 7, 9, 128
    L, C = F * H * W + 5, heads * hd                     # 5 padding rows beyond the grid: normalised, not rotated
    x = torch.randn(L, C, generator=g).bfloat16().to(dev)
    w = (1 + 0.1 * torch.randn(C, generator=g)).bfloat16().to(dev)
    rope = _rope(F, H, W, hd, dev, row_offset=0) if rope_on else None
    keep = x.clone()
    blocked = torch.empty(P, L, C // P, dtype=torch.bfloat16, device=dev)
    ops.rmsnorm_rope_(x, w, 1e-6, hd, rope, out_blocked=blocked)
    slabs = [torch.full((L, C // P), 7.0, dtype=torch.bfloat16, device=dev) for _ in range(P)]
    ops.rmsnorm_rope_scatter(x, w, 1e-6, hd, rope, slabs)
    assert torch.equal(x, keep)                            # the input is not modified
    for b in range(P):
        assert torch.equal(slabs[b], blocked[b]), b


def test_copy_scatter_and_rows_scatter():
    from videocof_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(1)
    rows, C, P = 333, 1024, 4
    v = torch.randn(rows, C, generator=g).bfloat16().to(dev)
    slabs = [torch.zeros(rows, C // P, dtype=torch.bfloat16, device=dev) for _ in range(P)]
    ops.copy_scatter(v, slabs)
    for b in range(P):
        assert torch.equal(slabs[b], v[:, b * (C // P):(b + 1) * (C // P)])
    # strided source rows (a column slice of a wider matrix)
    wide = torch.randn(rows, 2 * C, generator=g).bfloat16().to(dev)
    ops.copy_scatter(wide[:, :C], slabs)
    for b in range(P):
        assert torch.equal(slabs[b], wide[:, b * (C // P):(b + 1) * (C // P)])
    o = torch.randn(P * rows, 256, generator=g).bfloat16().to(dev)
    back = torch.zeros(P, rows, 256, dtype=torch.bfloat16, device=dev)
    ops.copy_rows_scatter(o, [back[c] for c in range(P)])
    assert torch.equal(back.view(P * rows, 256), o)


def test_scatter_rejects_bad_destinations():
    from videocof_b200 import ops
    from videocof_b200._lib import VcofError
    dev = torch.device("cuda")
    x = torch.zeros(16, 256, dtype=torch.bfloat16, device=dev)
    w = torch.ones(256, dtype=torch.bfloat16, device=dev)
    with pytest.raises(VcofError):
        ops.rmsnorm_rope_scatter(x, w, 1e-6, 128, None, [torch.zeros(16, 64, dtype=torch.bfloat16, device=dev)] * 2)
    with pytest.raises(VcofError):
        ops.copy_rows_scatter(x, [torch.zeros(5, 256, dtype=torch.bfloat16, device=dev)] * 3)
    with pytest.raises(VcofError):
        ops.copy_scatter(x, [])


def test_scatter_rejects_cpu_slab():
    from videocof_b200 import ops
    from videocof_b200._lib import VcofError
    x = torch.zeros(16, 256, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(VcofError):
        ops.copy_scatter(x, [torch.zeros(16, 128, dtype=torch.bfloat16)] * 2)


@pytest.mark.parametrize("Lq,Lk,kv,heads,P", [(512, 640, 600, 2, 2), (1024, 1024, 1024, 3, 4), (300, 333, 333, 1, 1),
                                                (2400, 777, 777, 2, 8)])
def test_attention_scatter_equals_attention(Lq, Lk, kv, heads, P):
    """The scattered epilogue changes store addresses only: chunk c of the output must equal rows
    [c * Lq/P, (c+1) * Lq/P) of the plain kernel's output, bit for bit."""
    from videocof_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(Lq + heads)
    C = heads * 128
    q = torch.randn(Lq, C, generator=g).bfloat16().to(dev)
    k = torch.randn(Lk, C, generator=g).bfloat16().to(dev)
    v = torch.randn(Lk, C, generator=g).bfloat16().to(dev)
    ref = ops.attention(q, k, v, heads, kv_len=kv)
    rows = Lq // P
    slabs = [torch.full((rows, C), 5.0, dtype=torch.bfloat16, device=dev) for _ in range(P)]
    ops.attention_scatter(q, k, v, heads, slabs, kv_len=kv)
    for c in range(P):
        assert torch.equal(slabs[c], ref[c * rows:(c + 1) * rows]), c
