"""Sequence (context) parallelism for the DiT: one process per GPU, tokens sharded contiguously.

Tokens are ordered (f, h, w)-major (reference wan_transformer3d.py:879), so rank r owning rows
[r*L/P, (r+1)*L/P) owns a temporal slab.  Every DiT op except self-attention is token-local;
self-attention needs all keys/values.  Two exchange schemes, both bit-identical to one GPU:

* head exchange (default when heads % P == 0): an all-to-all turns the token-sharded [L/P, all heads] Q, K, V into
  head-sharded [L, heads/P] tensors (each launched asynchronously right after its projection, overlapping the next
  projection), attention runs over the full sequence for this rank's heads, and a second all-to-all returns the
  output to token shards.  Per layer and rank this moves 4·(L/P)·C·(P-1)/P elements over NVLink instead of the
  2·L·C·(P-1)/P of the K/V all-gather — 3.5x less at P = 8 — and the NCCL kernels hold SMs for a fraction of the time.
* K/V all-gather (fallback, any head count): each layer all-gathers the post-RMSNorm, post-RoPE K and V shards and
  runs the local queries against them.

The head output is gathered at the end so every rank runs the identical scheduler step.  This replaces the
reference's xfuser USP path (videox_fun/dist/wan_xfuser.py:68-111, wan_transformer3d.py:802-816, :949-953,
:1085-1086), which cannot run VideoCoF's chain-of-frames kwargs.

* push exchange (DEFAULT since round 2 when heads % P == 0: validated on hardware, bit-identical to one GPU at C2
  widths, and the fastest of the three at P = 2 — profiles/r2_gpurun12_multi_gpu_n2.log): the head
  exchange without a collective call.  Every rank owns receive buffers in symmetric (NVLink peer-mapped) memory; the
  kernels that PRODUCE the exchanged tensors store straight into the other ranks' buffers — the norm/RoPE kernel for Q
  and K (vcof_rmsnorm_rope_scatter), a pack kernel for V, the attention kernel's own epilogue for the output on the way
  back (vcof_attn_fwd_scatter) —
  so the transfer overlaps the producing kernel store by store and costs no NCCL launch or staging pass; two
  cross-GPU barriers per layer (stream-ordered, symmetric-memory signal pads) order producers against consumers.

`attn_fn` / `copy_fn` are injectable so the sharding / exchange logic can be exercised on CPU with gloo
(tests/test_dist_gloo.py); the product default is the libvcof tcgen05 kernel.
"""
import os

import torch
import torch.distributed as dist


def symm_alloc(shape, like, group, tag):
    """Symmetric-memory allocation for the push exchange: -> (local tensor, [the same buffer on every rank as a
    peer-mapped tensor], barrier()).  torch.distributed._symmetric_memory maps the peers' allocations over NVLink
    (CUDA VMM + fabric / fd handles); `barrier` is a stream-ordered cross-GPU barrier on the buffer's signal pad."""
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(shape, dtype=like.dtype, device=like.device)
    h = symm.rendezvous(t, group)
    peers = [h.get_buffer(r, tuple(shape), like.dtype) for r in range(h.world_size)]
    return t, peers, h.barrier


class PushBuffers:
    """Receive side of the push exchange for one (rows, C) problem.

    recv[w] [P*rows, C/P] (w in q, k, v): all tokens in global order, this rank's heads — rank s stores its rows
    [s*rows, (s+1)*rows) there; back [P, rows, C/P]: this rank's tokens, block s = the heads rank s attended over.
    dst[w][r] / dst_back[r]: this rank's slab inside rank r's buffers (a peer-mapped tensor)."""

    def __init__(self, sp, rows, C, like, alloc):
        P, cp = sp.world, C // sp.world
        mine = slice(sp.rank * rows, (sp.rank + 1) * rows)
        self.recv, self.dst = {}, {}
        for w in ("q", "k", "v"):
            local, peers, barrier = alloc((P * rows, cp), like, sp.group, w)
            self.recv[w] = local
            self.dst[w] = [peers[r][mine] for r in range(P)]
        local, peers, _ = alloc((P * rows, cp), like, sp.group, "b")
        self.back = local.view(P, rows, cp)
        self.dst_back = [peers[r][mine] for r in range(P)]
        self.barrier = barrier


class SequenceParallel:
    def __init__(self, group=None, attn_fn=None, copy_fn=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.kv_len = None
        self.rows = None
        self._kg = self._vg = None
        self._pending = {}
        self._xbuf = {}
        if attn_fn is None:
            from . import ops
            attn_fn = ops.attention
        if copy_fn is None:
            from . import ops
            copy_fn = ops.copy_blocked
        self.attn_fn = attn_fn
        self.copy_fn = copy_fn        # (rowmajor [rows,C], blocked [P,rows,C/P], to_blocked) pack / unpack kernel
        self.alloc_fn = symm_alloc    # injectable like attn_fn / copy_fn (CPU tests share plain tensors between threads)
        self._push = None
        self._push_key = None
        self._push_ok = None          # result of the one-time symmetric-memory probe (use_push)

    def configure(self, kv_len, rows):
        """kv_len: number of real (non-padding) tokens of the full sequence; rows: tokens per rank."""
        self.kv_len, self.rows = int(kv_len), int(rows)

    def shard(self, x):
        """Rows of the padded full sequence owned by this rank."""
        rows = x.shape[0] // self.world
        return x[self.rank * rows:(self.rank + 1) * rows]

    def _gather_buf(self, name, like):
        buf = getattr(self, name)
        shape = (self.world * like.shape[0],) + tuple(like.shape[1:])
        if buf is None or buf.shape != shape or buf.dtype != like.dtype or buf.device != like.device:
            buf = torch.empty(shape, dtype=like.dtype, device=like.device)
            setattr(self, name, buf)
        return buf

    def start_gather(self, which, x):
        """Launch the all-gather of this rank's K (or V) shard asynchronously: NCCL runs on its own stream
        once the producing kernels have finished, and overlaps whatever the compute stream does next (the
        V / Q projections); `attention_gathered` waits for it."""
        buf = self._gather_buf("_kg" if which == "k" else "_vg", x)
        work = dist.all_gather_into_tensor(buf, x.contiguous(), group=self.group, async_op=True)
        self._pending[which] = (buf, work)

    def attention_gathered(self, q, heads, out=None):
        (kg, wk), (vg, wv) = self._pending.pop("k"), self._pending.pop("v")
        wk.wait()
        wv.wait()
        return self.attn_fn(q, kg, vg, heads, kv_len=self.kv_len, out=out)

    def attention(self, q, k, v, heads, out=None):
        """Local queries against the all-gathered keys/values (keys beyond kv_len are padding)."""
        self.start_gather("k", k)
        self.start_gather("v", v)
        return self.attention_gathered(q, heads, out=out)

    # -- head exchange (all-to-all) ------------------------------------------------------------------
    def can_exchange_heads(self, heads):
        """Head exchange needs heads divisible by P.  Default policy (measured on B200/NVSwitch, profiles/): P >= 4 ->
        head exchange (at P = 8 the K/V all-gather moves 1.35 GB per layer and rank and leaves ~2 ms exposed);
        P = 2 -> K/V all-gather (fully hidden behind the V and Q projections).  VCOF_SP_MODE=heads|gather forces."""
        mode = os.environ.get("VCOF_SP_MODE", "auto")
        if heads % self.world != 0 or mode == "gather":
            return False
        return mode == "heads" or self.world >= 4

    def _buf(self, name, shape, like):
        buf = self._xbuf.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != like.dtype or buf.device != like.device:
            buf = torch.empty(shape, dtype=like.dtype, device=like.device)
            self._xbuf[name] = buf
        return buf

    def send_buffer(self, which, x):
        """Contiguous [P, rows, C/P] send buffer for the [rows, C] projection x: block r holds the heads that rank r
        will attend over.  The producer writes it directly (ops.rmsnorm_rope_(..., out_blocked=) for Q and K,
        `pack` for V), so the exchange needs no separate transpose pass."""
        rows, C = x.shape
        return self._buf("s" + which, (self.world, rows, C // self.world), x)

    def pack(self, x, blocked):
        return self.copy_fn(x, blocked, True)

    def start_exchange(self, which, blocked):
        """blocked [P, rows, C/P] (this rank's tokens, heads grouped by destination) -> asynchronously
        [P*rows, C/P] (all tokens in global order, this rank's heads).  The all-to-all runs on NCCL's stream and
        overlaps whatever is launched next."""
        P, rows, cp = blocked.shape
        recv = self._buf("r" + which, (P * rows, cp), blocked)
        work = dist.all_to_all_single(recv.view(-1), blocked.view(-1), group=self.group, async_op=True)
        self._pending[which] = (recv, work)

    def _head_groups(self, local_heads):
        """How the head-sharded attention is split so that the return all-to-all of one group of heads overlaps the
        attention of the next: two groups (the larger first) when there are at least two local heads;
        VCOF_SP_SPLIT=0 keeps one launch and one blocking exchange (round-1 behaviour)."""
        if local_heads < 2 or os.environ.get("VCOF_SP_SPLIT", "1") == "0":
            return [local_heads]
        return [local_heads - local_heads // 2, local_heads // 2]

    def attention_exchanged(self, heads, out):
        """Attention over the full sequence for this rank's heads/P heads, then the inverse exchange into
        out [rows, C] (this rank's tokens, all heads).  The local heads run as two launches: the all-to-all that returns
        the first group's output travels while the second group is computed, so only the smaller group's exchange is
        exposed (at P = 8 with 5 local heads: 2/5 of the bytes instead of all of them)."""
        P = self.world
        (q, wq), (k, wk), (v, wv) = self._pending.pop("q"), self._pending.pop("k"), self._pending.pop("v")
        wq.wait()
        wk.wait()
        wv.wait()
        hl = heads // P
        rows, cp = q.shape[0] // P, q.shape[1]
        hd = cp // hl
        groups = self._head_groups(hl)
        if len(groups) == 1:
            o = self._buf("o", tuple(q.shape), q)
            self.attn_fn(q, k, v, hl, kv_len=self.kv_len, out=o)
            back = self._buf("b", (P, rows, cp), q)
            dist.all_to_all_single(back.view(-1), o.view(-1), group=self.group)      # chunk r of o = rank r's tokens
            self.copy_fn(out, back, False)
            return out
        col, flights = 0, []
        for gi, hg in enumerate(groups):
            w = hg * hd
            o = self._buf(f"o{gi}", (P * rows, w), q)
            self.attn_fn(q[:, col:col + w], k[:, col:col + w], v[:, col:col + w], hg, kv_len=self.kv_len, out=o)
            back = self._buf(f"b{gi}", (P, rows, w), q)
            work = dist.all_to_all_single(back.view(-1), o.view(-1), group=self.group, async_op=True)
            flights.append((back, work, col, w))
            col += w
        ov = out.view(rows, P, cp)
        for back, work, c0, w in flights:
            work.wait()
            # block s of `back` = columns [s*cp + c0, s*cp + c0 + w) of this rank's rows (a strided device copy)
            ov[:, :, c0:c0 + w].copy_(back.permute(1, 0, 2))
        return out

    # -- push exchange (producer kernels store into the peers' receive buffers) -----------------------
    def use_push(self, heads):
        """Push exchange: default ("auto") whenever the heads divide by P and the ranks' buffers can be mapped into
        each other (torch symmetric memory over NVLink peer access; probed once — a failed probe falls back to the
        collective schemes with a note on stderr, it never raises).  Measured on 2 x B200, C2 step
        (profiles/r2_gpurun12_multi_gpu_n2.log): push 2761 ms, K/V all-gather 2792 ms, head exchange 2820 ms.
        VCOF_SP_MODE=heads|gather forces a collective scheme, =push insists (probe failures raise)."""
        mode = os.environ.get("VCOF_SP_MODE", "auto")
        if heads % self.world != 0 or mode in ("heads", "gather"):
            return False
        if mode == "push":
            return True
        return self._push_available()

    def _push_available(self):
        if self._push_ok is None:
            ok = 1
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
            try:
                if dev.type != "cuda":
                    raise RuntimeError("no CUDA device")
                probe = torch.empty(8, dtype=torch.bfloat16, device=dev)
                _t, peers, barrier = self.alloc_fn((8,), probe, self.group, "probe")
                barrier()
                ok = 1 if len(peers) == self.world else 0
            except Exception as exc:                       # noqa: BLE001 — any failure means "use the collectives"
                import sys
                print(f"videocof_b200.dist: symmetric memory unavailable ({exc!r}); using NCCL exchanges", file=sys.stderr)
                ok = 0
            # every rank must take the same path: agree on the minimum
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            self._push_ok = bool(int(flag.item()))
        return self._push_ok

    def push_buffers(self, rows, C, like):
        key = (rows, C, like.dtype, str(like.device))
        if self._push_key != key:
            self._push = PushBuffers(self, rows, C, like, self.alloc_fn)
            self._push_key = key
        return self._push

    def attention_pushed(self, heads, out, attn_scatter_fn):
        """After every rank has pushed its Q, K, V slabs: barrier, attention over the full sequence for this rank's
        heads/P heads with the epilogue storing each row chunk straight into the buffer of the rank that owns the rows
        (ops.attention_scatter: compute and return transfer in one kernel), barrier, unpack into out [rows, C]."""
        pb = self._push
        pb.barrier()
        attn_scatter_fn(pb.recv["q"], pb.recv["k"], pb.recv["v"], heads // self.world, pb.dst_back, kv_len=self.kv_len)
        pb.barrier()
        self.copy_fn(out, pb.back, False)
        return out

    def all_gather_rows(self, y):
        full = torch.empty((self.world * y.shape[0],) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
        dist.all_gather_into_tensor(full, y.contiguous(), group=self.group)
        return full
