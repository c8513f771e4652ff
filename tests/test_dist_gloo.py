"""Sequence-parallel plumbing (videocof_b200.dist) on CPU with gloo, world_size 2 and 3: token sharding with
padding, K/V all-gather or head exchange (all-to-all), attention, row gather — against un-sharded attention.  The attention
function is injected (the libvcof kernel needs a GPU; the math here is the oracle's attention_ref)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _attn(q, k, v, heads, kv_len=None, out=None, **_):
    from oracle.dit_oracle import attention_ref
    L, C = q.shape
    d = C // heads
    o = attention_ref(q.view(L, heads, d), k.view(-1, heads, d), v.view(-1, heads, d), kv_len).reshape(L, C)
    if out is not None:
        out.copy_(o)
        return out
    return o


def _copy(rowmajor, blocked, to_blocked):
    """torch statement of vcof_copy_blocked (include/vcof.h)."""
    P, rows, cp = blocked.shape
    if to_blocked:
        blocked.copy_(rowmajor.view(rows, P, cp).transpose(0, 1))
        return blocked
    rowmajor.view(rows, P, cp).copy_(blocked.transpose(0, 1))
    return rowmajor


def _worker(rank, world, port, L, q, mode="gather"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VCOF_SP_MODE=mode)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from videocof_b200.dist import SequenceParallel
        torch.manual_seed(0)
        heads, d = (2, 16) if mode == "gather" else (2 * world, 8)
        C = heads * d
        seq_len = (L + world - 1) // world * world              # reference padding rule (:904-905)
        full = [torch.randn(seq_len, C) for _ in range(3)]
        for t in full:
            t[L:] = 7.0                                          # padding rows hold junk that must be masked
        sp = SequenceParallel(attn_fn=_attn, copy_fn=_copy)
        rows = seq_len // world
        sp.configure(kv_len=L, rows=rows)
        ql, kl, vl = (sp.shard(t) for t in full)
        if mode == "gather":
            out_local = sp.attention(ql.contiguous(), kl.contiguous(), vl.contiguous(), heads)
        else:
            assert sp.can_exchange_heads(heads)
            for name, t in (("q", ql), ("k", kl), ("v", vl)):
                t = t.contiguous()
                sp.start_exchange(name, sp.pack(t, sp.send_buffer(name, t)))
            out_local = sp.attention_exchanged(heads, out=torch.empty(rows, C))
        gathered = sp.all_gather_rows(out_local)
        ref = _attn(full[0], full[1], full[2], heads, kv_len=L)
        err = float((gathered[:L] - ref[:L]).abs().max())
        q.put((rank, err, tuple(gathered.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,L,mode", [(2, 37, "gather"), (3, 50, "gather"), (2, 64, "gather"),
                                           (2, 37, "heads"), (3, 50, "heads")])
def test_sequence_parallel_attention_matches_unsharded(world, L, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, L, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, shape in res:
        assert err < 1e-5, (rank, err)
        assert shape[0] == (L + world - 1) // world * world
