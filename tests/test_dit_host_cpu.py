"""Host-side logic of videocof_b200/dit.py on CPU: the libvcof entry points are replaced by the executable C-ABI
statements of tests/vcof_emulator.py, and the forward must reproduce the goldens of the executed reference
(tests/golden/dit_*.npz, tools/gen_golden.py) — patchify order, fp32 timestep path, modulation stacking, the three
RoPE position modes, CFG batch 2, padded sequences, the TeaCache gate, cfg_skip, and (gloo, world_size 2) the
token-sharded sequence-parallel forward in both exchange schemes.

Tolerances are those of tests/test_dit_gpu.py (bf16 compute vs fp32 reference): relative Frobenius < 1.5e-2 against
the oracle with bf16 rounding emulated at the reference's CUDA rounding points, < 4e-2 against the fp32 goldens."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vcof_emulator
from gen_golden import DIT_CASES, ROPE_MODES, dit_inputs
from oracle.dit_oracle import DiTConfig, dit_forward, make_dit_params


def build_model(cfg, params):
    from videocof_b200.dit import WanTransformer3DModel
    m = WanTransformer3DModel(**cfg.to_kwargs())
    m.load_state_dict(params, strict=True)
    return m.to(torch.bfloat16).eval()


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm())


def case(name, B=None):
    ckw, shape, n_ctx, B0 = DIT_CASES[name]
    B = B0 if B is None else B
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    f = shape[1]
    return cfg, params, x, ctx, t, f, f * (shape[2] // 2) * (shape[3] // 2), B, shape


@pytest.mark.parametrize("name", ["dit_tiny", "dit_tiny_b2"])
@pytest.mark.parametrize("mode", list(ROPE_MODES))
def test_forward_matches_reference_goldens(name, mode, golden_dir, monkeypatch):
    vcof_emulator.install_dit(monkeypatch)
    cfg, params, x, ctx, t, f, seq_len, B, shape = case(name)
    kw = ROPE_MODES[mode](f, B)
    model = build_model(cfg, params)
    with torch.no_grad():
        y = model(x=x.bfloat16(), t=t, context=[c.bfloat16() for c in ctx], seq_len=seq_len, **kw)
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (B,) + shape
    emu = dit_forward(params, cfg, x.bfloat16().float(), t, [c.bfloat16().float() for c in ctx], seq_len,
                      emulate_bf16=True, **kw)
    assert rel(y, emu) < 1.5e-2, ("vs bf16-emulating oracle", rel(y, emu))
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, name + ".npz"))["out_" + mode])
    assert rel(y, gold) < 4e-2, ("vs reference golden", rel(y, gold))


def test_c1_widths_two_layers(golden_dir, monkeypatch):
    """1.3B widths (C = 1536, F = 8960, 12 heads) at the C1 token count (L = 1280), chain-of-frames RoPE."""
    vcof_emulator.install_dit(monkeypatch)
    cfg, params, x, ctx, t, f, seq_len, B, _ = case("dit_c1_2layer")
    model = build_model(cfg, params)
    with torch.no_grad():
        y = model(x=x.bfloat16(), t=t, context=[c.bfloat16() for c in ctx], seq_len=seq_len, **ROPE_MODES["cot"](f, B))
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, "dit_c1_2layer.npz"))["out_cot"])
    assert rel(y, gold) < 4e-2, rel(y, gold)


def test_padded_sequence_matches_unpadded(monkeypatch):
    """seq_len > L (reference :904-910): padded rows never reach real tokens (keys beyond kv_len are masked)."""
    vcof_emulator.install_dit(monkeypatch)
    cfg, params, x, ctx, t, f, L, B, _ = case("dit_tiny")
    model = build_model(cfg, params)
    args = dict(x=x.bfloat16(), t=t, context=[c.bfloat16() for c in ctx])
    with torch.no_grad():
        a = model(seq_len=L, **args)
        b = model(seq_len=L + 37, **args)
    assert rel(a, b) < 1e-3            # not bit-equal on CPU: the host BLAS blocks a 240- and a 277-row GEMM differently


def test_output_is_fresh_and_inputs_untouched(monkeypatch):
    """The pipeline zeroes the source frames of the returned tensor in place (pipeline_wan.py:736)."""
    vcof_emulator.install_dit(monkeypatch)
    cfg, params, x, ctx, t, f, L, B, _ = case("dit_tiny")
    model = build_model(cfg, params)
    xin = x.bfloat16()
    keep = xin.clone()
    with torch.no_grad():
        y = model(x=xin, t=t, context=[c.bfloat16() for c in ctx], seq_len=L)
    y[:, :, :1] = 0
    assert torch.equal(xin, keep) and y.data_ptr() != xin.data_ptr()


def test_teacache_gate_matches_reference(golden_dir, monkeypatch):
    """Same skip decisions and outputs as the executed reference over 4 steps (wan_transformer3d.py:956-1031)."""
    from gen_golden import TEACACHE, TEACACHE_T
    vcof_emulator.install_dit(monkeypatch)
    cfg, params, x, ctx, _, f, seq_len, B, _ = case("dit_tiny")
    gold = np.load(os.path.join(golden_dir, "dit_tiny_teacache.npz"))
    model = build_model(cfg, params)
    model.enable_teacache(**TEACACHE)
    calc = []
    for i, tv in enumerate(TEACACHE_T):
        with torch.no_grad():
            y = model(x=x.bfloat16(), t=torch.tensor([tv]), context=[c.bfloat16() for c in ctx], seq_len=seq_len,
                      **ROPE_MODES["cot"](f, B))
        calc.append(bool(model.should_calc))
        assert rel(y, torch.from_numpy(gold["outs"][i])) < 4e-2, (i, rel(y, torch.from_numpy(gold["outs"][i])))
    assert calc == [bool(v) for v in gold["should_calc"]] and not all(calc)


def test_cfg_skip_drops_the_unconditional_half(monkeypatch):
    """cfg_skip (utils/cfg_optimization.py:5-38): late steps run the cond half only and duplicate it."""
    vcof_emulator.install_dit(monkeypatch)
    cfg, params, x, ctx, t, f, seq_len, _, _ = case("dit_tiny", B=2)
    model = build_model(cfg, params)
    args = dict(x=x.bfloat16(), t=t, context=[c.bfloat16() for c in ctx], seq_len=seq_len)
    with torch.no_grad():
        full = model(**args)
        model.enable_cfg_skip(0.5, 4)
        model.current_steps = 3
        skipped = model(**args)
        model.current_steps = 0
        early = model(**args)
    assert torch.equal(skipped[0], skipped[1]) and rel(skipped[1], full[1]) < 1e-3
    assert rel(early, full) < 1e-3


def test_unsupported_inputs_raise(monkeypatch):
    vcof_emulator.install_dit(monkeypatch)
    cfg, params, x, ctx, t, f, L, B, _ = case("dit_tiny")
    model = build_model(cfg, params)
    with pytest.raises(NotImplementedError):
        model(x=x.bfloat16(), t=t, context=ctx, seq_len=L, clip_fea=torch.zeros(1))
    with pytest.raises(NotImplementedError):
        model(x=x.bfloat16(), t=t[:, None].expand(-1, L), context=ctx, seq_len=L)
    with pytest.raises(AssertionError):
        model(x=x.bfloat16(), t=t, context=[c.bfloat16() for c in ctx], seq_len=L - 1)


# ---- sequence parallel (gloo, one process per rank) -------------------------------------------------------------

def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sp_worker(rank, world, port, mode, heads, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VCOF_SP_MODE=mode)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from videocof_b200 import dit, ops
        for name in vcof_emulator.DIT_OPS:
            setattr(ops, name, getattr(vcof_emulator, name))
        dit.WanTransformer3DModel._check_ready = lambda self, x: None
        ckw, shape, n_ctx, _ = DIT_CASES["dit_tiny"]
        ckw = dict(ckw, num_heads=heads, dim=heads * (ckw["dim"] // ckw["num_heads"]))
        cfg = DiTConfig(**ckw)
        params = make_dit_params(cfg, seed=11)
        shape = (shape[0], 3, 10, 6)                       # L = 3 * 5 * 3 = 45 tokens: padded to 46 / 48 rows
        x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, 1, seed=23)
        L = 45
        kw = ROPE_MODES["cot"](3, 1)
        args = dict(x=x.bfloat16(), t=t, context=[c.bfloat16() for c in ctx], seq_len=L, **kw)
        model = build_model(cfg, params)
        with torch.no_grad():
            single = model(**args)
            model.enable_multi_gpus_inference()
            if mode == "auto":
                # no symmetric memory on a CPU-only box: the one-time probe fails on every rank, the ranks agree on it
                # (all-reduce MIN) and the collective schemes serve — the default path never raises
                assert model._sp.use_push(heads) is False and model._sp._push_ok is False
            else:
                assert model._sp.world == world and model._sp.can_exchange_heads(heads) == (mode == "heads")
            sharded = model(**args)
        q.put((rank, rel(sharded, single), tuple(sharded.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode,heads", [(2, "gather", 2), (2, "heads", 2), (3, "gather", 2), (4, "heads", 4),
                                              (2, "auto", 2)])
def test_sequence_parallel_forward_matches_single_rank(world, mode, heads):
    """Every rank returns the full output, equal to the un-sharded forward (token padding, global RoPE row offset,
    K/V all-gather or head exchange, head all-gather: wan_transformer3d.py:904-905, 949-953, 1085-1086)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sp_worker, args=(r, world, port, mode, heads, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, shape in res:
        assert err < 2e-3, (rank, err)          # bf16 rounding flips from differently blocked host GEMMs only
        assert shape == (1, 16, 3, 10, 6)


# ---- push exchange (symmetric memory stands in as plain tensors shared between threads) ----------------------

class _ThreadSP:
    """videocof_b200.dist.SequenceParallel for rank `rank` of `world` THREADS of one process: the symmetric-memory
    allocator hands every rank the same list of per-rank tensors (so a store into rank r's slab is a plain tensor
    write), the cross-GPU barrier is a threading.Barrier, and the final row gather goes through a shared list."""

    def __new__(cls, rank, world, shared, barrier):
        from videocof_b200 import dist as vdist, ops

        class SP(vdist.SequenceParallel):
            def __init__(self):                      # no process group: set what __init__ would have set
                self.group, self.world, self.rank = None, world, rank
                self.kv_len = self.rows = None
                self._kg = self._vg = None
                self._pending, self._xbuf = {}, {}
                self.attn_fn, self.copy_fn = ops.attention, ops.copy_blocked
                self._push = self._push_key = None
                self.alloc_fn = self._alloc

            def _alloc(self, shape, like, group, tag):
                with shared["lock"]:
                    bufs = shared.setdefault(("buf", tag, tuple(shape)),
                                             [torch.full(shape, float("nan"), dtype=like.dtype) for _ in range(world)])
                return bufs[rank], bufs, barrier.wait

            def all_gather_rows(self, y):
                shared[("rows", rank)] = y.clone()
                barrier.wait()
                full = torch.cat([shared[("rows", r)] for r in range(world)], dim=0)
                barrier.wait()
                return full
        return SP()


@pytest.mark.parametrize("world,heads", [(2, 2), (4, 4)])
def test_push_exchange_forward_matches_single_rank(world, heads, monkeypatch):
    """VCOF_SP_MODE=push: Q, K, V land in the other ranks' receive buffers straight from the producing ops, the
    attention output is pushed back by row chunk; the sharded forward equals the un-sharded one on every rank."""
    import threading
    vcof_emulator.install_dit(monkeypatch)
    monkeypatch.setenv("VCOF_SP_MODE", "push")
    ckw, shape, n_ctx, _ = DIT_CASES["dit_tiny"]
    ckw = dict(ckw, num_heads=heads, dim=heads * (ckw["dim"] // ckw["num_heads"]))
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    shape = (shape[0], 3, 10, 6)                           # 45 tokens: padded to 46 / 48 rows
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, 1, seed=23)
    args = dict(x=x.bfloat16(), t=t, context=[c.bfloat16() for c in ctx], seq_len=45, **ROPE_MODES["cot"](3, 1))
    with torch.no_grad():
        single = build_model(cfg, params)(**args)
    shared, barrier = {"lock": threading.Lock()}, threading.Barrier(world)
    outs, errs = [None] * world, []

    def run(rank):
        try:
            model = build_model(cfg, params)
            model._sp = _ThreadSP(rank, world, shared, barrier)
            assert model._sp.use_push(heads)
            with torch.no_grad():
                outs[rank] = model(**args)
                outs[rank] = model(**args)           # second forward: the receive buffers are reused
        except Exception as e:                       # noqa: BLE001 - reported below, and the barrier is released
            errs.append((rank, repr(e)))
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errs, errs
    for r in range(world):
        assert outs[r] is not None and rel(outs[r], single) < 2e-3, (r, rel(outs[r], single))
    # every receive buffer was completely overwritten by the pushes (they start as NaN)
    for key, bufs in shared.items():
        if isinstance(key, tuple) and key[0] == "buf":
            assert all(not torch.isnan(b.float()).any() for b in bufs), key


@pytest.mark.parametrize("pad", [0, 5])
def test_batched_forward_is_bit_identical_to_the_per_sample_loop(monkeypatch, pad):
    """CFG batch 2 (pipeline_wan.py:700): the batch-aware forward stacks the samples' tokens along M so that the block
    stack streams its weights once per step (WanAttentionBlock.run_batched); row-wise the arithmetic is unchanged, so
    the output must equal the per-sample loop's (VCOF_DIT_BATCHED=0) bit for bit — also with a padded sequence."""
    vcof_emulator.install_dit(monkeypatch)
    cfg, params, x, ctx, t, f, L, B, _ = case("dit_tiny_b2")
    assert B == 2
    model = build_model(cfg, params)
    args = dict(x=x.bfloat16(), t=t, context=[c.bfloat16() for c in ctx], seq_len=L + pad, **ROPE_MODES["cot"](f, B))
    with torch.no_grad():
        monkeypatch.setenv("VCOF_DIT_BATCHED", "0")
        loop = model(**args)
        monkeypatch.setenv("VCOF_DIT_BATCHED", "1")
        batched = model(**args)
    assert torch.equal(loop, batched)


def test_context_cache_is_bit_identical_and_invalidates(monkeypatch):
    """SURVEY §8a a4 / a11 (K3 / K10): with `enable_context_cache` the text embedding and the blocks' cross-attention
    K / V are computed once per prompt embedding; a cached forward equals an uncached one bit for bit (per-sample loop
    and batched CFG path), and an in-place edit of the embedding or of a weight they were computed from misses."""
    vcof_emulator.install_dit(monkeypatch)
    cfg, params, x, ctx, t, f, seq_len, B, shape = case("dit_tiny_b2")
    model = build_model(cfg, params)
    ctx = [c.bfloat16() for c in ctx]
    kw = ROPE_MODES["cot"](f, B)

    def fwd(xx, tt):
        with torch.no_grad():
            return model(x=xx.bfloat16(), t=tt, context=ctx, seq_len=seq_len, **kw)

    calls = {"n": 0}
    orig = type(model.blocks[0]).context_kv

    def counting(self, c):
        calls["n"] += 1
        return orig(self, c)

    monkeypatch.setattr(type(model.blocks[0]), "context_kv", counting)
    layers = len(model.blocks)
    for batched in ("1", "0"):
        monkeypatch.setenv("VCOF_DIT_BATCHED", batched)
        model.disable_context_cache()
        calls["n"] = 0
        ref = [fwd(x, t), fwd(x * 0.5, t * 0.5)]
        uncached_calls = calls["n"]
        assert uncached_calls == (2 * layers if batched == "1" else 2 * B * layers)
        model.enable_context_cache()
        calls["n"] = 0
        got = [fwd(x, t), fwd(x * 0.5, t * 0.5)]
        assert calls["n"] == B * layers                      # first forward only, once per distinct embedding
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
        # an in-place edit of one embedding recomputes that embedding alone, and the result follows it
        ctx[0].mul_(0.5)
        y = fwd(x, t)
        assert calls["n"] == (B + 1) * layers
        model.disable_context_cache()
        assert torch.equal(y, fwd(x, t)) and not torch.equal(y, ref[0])
        # an in-place edit of a weight the cache depends on misses for every embedding
        model.enable_context_cache()
        fwd(x, t)
        calls["n"] = 0
        model.blocks[1].cross_attn.v.weight.detach().mul_(1.25)
        y = fwd(x, t)
        assert calls["n"] == B * layers
        model.disable_context_cache()
        assert torch.equal(y, fwd(x, t))
        ctx[0].mul_(2.0)
        model.blocks[1].cross_attn.v.weight.detach().mul_(0.8)
    # the cache is bounded
    model.enable_context_cache(max_entries=2)
    for s in (1.0, 0.9, 0.8):
        with torch.no_grad():
            model(x=x.bfloat16(), t=t, context=[c * s for c in ctx], seq_len=seq_len, **kw)
    assert len(model._ctx_cache["entries"]) == 2
