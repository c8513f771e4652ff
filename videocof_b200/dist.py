"""Sequence (context) parallelism for the DiT: one process per GPU, tokens sharded contiguously.

Tokens are ordered (f, h, w)-major (reference wan_transformer3d.py:879), so rank r owning rows
[r*L/P, (r+1)*L/P) owns a temporal slab.  Every DiT op except self-attention is token-local;
self-attention needs all keys/values, so each layer all-gathers the post-RMSNorm, post-RoPE K and
V shards over NCCL/NVLink and runs the local queries against them; the head output is gathered at
the end so every rank runs the identical scheduler step.  This replaces the reference's xfuser
USP path (videox_fun/dist/wan_xfuser.py:68-111, wan_transformer3d.py:802-816, :949-953,
:1085-1086), which cannot run VideoCoF's chain-of-frames kwargs.

`attn_fn` is injectable so the sharding / gather logic can be exercised on CPU with gloo
(tests/test_dist_gloo.py); the product default is the libvcof tcgen05 kernel.
"""
import torch
import torch.distributed as dist


class SequenceParallel:
    def __init__(self, group=None, attn_fn=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.kv_len = None
        self.rows = None
        self._kg = self._vg = None
        self._pending = {}
        if attn_fn is None:
            from . import ops
            attn_fn = ops.attention
        self.attn_fn = attn_fn

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

    def all_gather_rows(self, y):
        full = torch.empty((self.world * y.shape[0],) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
        dist.all_gather_into_tensor(full, y.contiguous(), group=self.group)
        return full
