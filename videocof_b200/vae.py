"""B200-native Wan-2.1 3D causal VAE with the reference's class API (videox_fun.models.AutoencoderKLWan).

Same module tree / state-dict keys as videox_fun/models/wan_vae.py (reference) — 194 tensors such
as `model.decoder.upsamples.14.residual.6.weight` — and the same encode()/decode() contract, but
the compute is un-chunked and channels-last on libvcof kernels:

  * every CausalConv3d / Conv2d runs in the im2col-free tcgen05 implicit-GEMM kernel
    (`vcof_conv_igemm`): causal time padding, spatial padding and ragged edges come from TMA
    zero-fill; the stride-2 down-sampler reads a 5-D parity view; nearest-2x + Conv2d is four parity
    sub-convolutions with folded 2x2 taps (the 4x tensor is never materialised); the temporal
    up-sampler interleaves its two channel halves into frames on store;
  * 1x1x1 convolutions are plain tcgen05 GEMMs on the channels-last matrix;
  * RMS_norm(+SiLU) is one fused HBM pass; the d=384 single-head attention is GEMM -> softmax -> GEMM.

The reference's chunked streaming loop with a 2-frame feature cache (:520-575) is equivalent to these
un-chunked causal convolutions except for the first-frame rules of the temporal resamplers, which
are kept (oracle/vae_oracle.py proves the closed form against the reference's own loop).
"""
import math

import os

import torch
import torch.nn as nn

from . import ops
from ._lib import VcofError

__all__ = ["AutoencoderKLWan", "AutoencoderKLWan_", "CausalConv3d", "RMS_norm", "Resample", "ResidualBlock",
           "AttentionBlock", "Encoder3d", "Decoder3d", "DiagonalGaussianDistribution", "DecoderOutput",
           "AutoencoderKLOutput"]

CACHE_T = 2


# ----------------------------------------------------------------------------------------------
# parameter containers (reference names)
# ----------------------------------------------------------------------------------------------
class CausalConv3d(nn.Conv3d):
    """reference :21-40 (parameters only; padding is realised by TMA zero-fill)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._padding = (self.padding[2], self.padding[2], self.padding[1], self.padding[1], 2 * self.padding[0], 0)
        self.padding = (0, 0, 0)


class RMS_norm(nn.Module):
    """reference :43-58."""

    def __init__(self, dim, channel_first=True, images=True, bias=False):
        super().__init__()
        shape = (dim, *((1, 1) if images else (1, 1, 1))) if channel_first else (dim,)
        self.channel_first, self.scale = channel_first, dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(shape))
        self.bias = nn.Parameter(torch.zeros(shape)) if bias else 0.


class Upsample(nn.Upsample):
    pass


class Resample(nn.Module):
    """reference :70-164."""

    def __init__(self, dim, mode):
        assert mode in ("none", "upsample2d", "upsample3d", "downsample2d", "downsample3d")
        super().__init__()
        self.dim, self.mode = dim, mode
        if mode in ("upsample2d", "upsample3d"):
            self.resample = nn.Sequential(Upsample(scale_factor=(2., 2.), mode="nearest-exact"),
                                          nn.Conv2d(dim, dim // 2, 3, padding=1))
            if mode == "upsample3d":
                self.time_conv = CausalConv3d(dim, dim * 2, (3, 1, 1), padding=(1, 0, 0))
        elif mode in ("downsample2d", "downsample3d"):
            self.resample = nn.Sequential(nn.ZeroPad2d((0, 1, 0, 1)), nn.Conv2d(dim, dim, 3, stride=(2, 2)))
            if mode == "downsample3d":
                self.time_conv = CausalConv3d(dim, dim, (3, 1, 1), stride=(2, 1, 1), padding=(0, 0, 0))
        else:
            self.resample = nn.Identity()


class ResidualBlock(nn.Module):
    """reference :190-224."""

    def __init__(self, in_dim, out_dim, dropout=0.0):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.residual = nn.Sequential(RMS_norm(in_dim, images=False), nn.SiLU(),
                                      CausalConv3d(in_dim, out_dim, 3, padding=1),
                                      RMS_norm(out_dim, images=False), nn.SiLU(), nn.Dropout(dropout),
                                      CausalConv3d(out_dim, out_dim, 3, padding=1))
        self.shortcut = CausalConv3d(in_dim, out_dim, 1) if in_dim != out_dim else nn.Identity()


class AttentionBlock(nn.Module):
    """reference :227-266."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.norm = RMS_norm(dim)
        self.to_qkv = nn.Conv2d(dim, dim * 3, 1)
        self.proj = nn.Conv2d(dim, dim, 1)
        nn.init.zeros_(self.proj.weight)


class Encoder3d(nn.Module):
    """reference :269-320."""

    def __init__(self, dim=128, z_dim=4, dim_mult=(1, 2, 4, 4), num_res_blocks=2, attn_scales=(),
                 temperal_downsample=(True, True, False), dropout=0.0):
        super().__init__()
        if attn_scales:
            raise NotImplementedError("attn_scales is empty in every Wan VAE config")
        dims = [dim * u for u in [1] + list(dim_mult)]
        self.conv1 = CausalConv3d(3, dims[0], 3, padding=1)
        layers = []
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(num_res_blocks):
                layers.append(ResidualBlock(cin, cout, dropout))
                cin = cout
            if i != len(dim_mult) - 1:
                layers.append(Resample(cout, "downsample3d" if temperal_downsample[i] else "downsample2d"))
        self.downsamples = nn.Sequential(*layers)
        self.middle = nn.Sequential(ResidualBlock(cout, cout, dropout), AttentionBlock(cout),
                                    ResidualBlock(cout, cout, dropout))
        self.head = nn.Sequential(RMS_norm(cout, images=False), nn.SiLU(), CausalConv3d(cout, z_dim, 3, padding=1))


class Decoder3d(nn.Module):
    """reference :373-425."""

    def __init__(self, dim=128, z_dim=4, dim_mult=(1, 2, 4, 4), num_res_blocks=2, attn_scales=(),
                 temperal_upsample=(False, True, True), dropout=0.0):
        super().__init__()
        dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
        self.conv1 = CausalConv3d(z_dim, dims[0], 3, padding=1)
        self.middle = nn.Sequential(ResidualBlock(dims[0], dims[0], dropout), AttentionBlock(dims[0]),
                                    ResidualBlock(dims[0], dims[0], dropout))
        layers = []
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            if i in (1, 2, 3):
                cin = cin // 2
            for _ in range(num_res_blocks + 1):
                layers.append(ResidualBlock(cin, cout, dropout))
                cin = cout
            if i != len(dim_mult) - 1:
                layers.append(Resample(cout, "upsample3d" if temperal_upsample[i] else "upsample2d"))
        self.upsamples = nn.Sequential(*layers)
        self.head = nn.Sequential(RMS_norm(cout, images=False), nn.SiLU(), CausalConv3d(cout, 3, 3, padding=1))


# ----------------------------------------------------------------------------------------------
# weight packing (cached per module, rebuilt when the parameter is mutated)
# ----------------------------------------------------------------------------------------------
def _pad32(c):
    return (c + 31) // 32 * 32


def _pad16(c):
    return (c + 15) // 16 * 16


def _cached(mod, key, versions, build):
    cache = mod.__dict__.setdefault("_vcof_pack", {})
    hit = cache.get(key)
    if hit is None or hit[0] != versions:
        cache[key] = (versions, build())
    return cache[key][1]


def _ver(*tensors):
    return tuple((t.data_ptr(), t._version, str(t.device), t.dtype) for t in tensors if t is not None)


def _slice_plan(cin, cout, kt=1):
    """K-slice geometry of one convolution for vcof_conv_igemm: (kc, cin_p, tgroup).

    kc = channels per K slice = one TMA box row (32 -> 64-byte rows, 64 -> 128-byte rows).  Measured on B200
    (profiles/r1_tma_probe_v2.txt, r1_gpurun17): the kernel's operand feed is bounded by bytes in flight / slot
    latency, not by the row width, and C = 96 wastes a quarter of a 128-channel K extent, so kc = 32 is the
    default; VCOF_CONV_KC=64 selects the wide slices (kept correct by tests/test_vae_gpu.py).  Channels are
    zero-padded to a multiple of kc in the WEIGHTS only (an activation box that reads past C is zero-filled by
    TMA or multiplied by zero weights).  tgroup = k_t when two ring stages of k_t fused temporal taps fit the
    kernel's 200 KB operand ring, else 1."""
    kc = 64 if (os.environ.get("VCOF_CONV_KC") == "64" and cin >= 64) else 32
    cin_p = (cin + kc - 1) // kc * kc
    n_tile = _ntile(_pad16(cout))
    tgroup = kt if 2 * kt * (128 + n_tile) * kc * 2 <= 200 * 1024 else 1
    return kc, cin_p, tgroup


def _pack_taps(w_taps, bias, cin_p, tgroup=1, kc=32):
    """w_taps: fp32 [Cout, Cin, ntaps] (taps ordered group-major, the `tgroup` dt-taps of a group adjacent) ->
    bf16 [slices, Cout_p16, kc] with slice = ((group * cin_p/kc + chunk) * tgroup + j); bias fp32 [Cout_p16]."""
    cout, cin, ntaps = w_taps.shape
    cout_p = _pad16(cout)
    w = torch.zeros((cout_p, ntaps, cin_p), dtype=torch.float32, device=w_taps.device)
    w[:cout, :, :cin] = w_taps.permute(0, 2, 1)
    w = w.view(cout_p, ntaps // tgroup, tgroup, cin_p // kc, kc).permute(1, 3, 2, 0, 4)   # [G, cc, tg, n, kc]
    b = torch.zeros(cout_p, dtype=torch.float32, device=w_taps.device)
    if bias is not None:
        b[:cout] = bias.float()
    return w.reshape(-1, cout_p, kc).to(torch.bfloat16).contiguous(), b


def pack_conv(conv):
    """nn.Conv3d [Cout,Cin,kt,kh,kw] / nn.Conv2d [Cout,Cin,kh,kw] -> (packed weight, bias, cin_p, tgroup); taps are
    ordered (kh, kw)-major with the k_t temporal taps adjacent, so that one TMA box can feed all k_t of them."""
    def build():
        w = conv.weight.detach().float()
        if w.dim() == 4:
            w = w.unsqueeze(2)
        kt = w.shape[2]
        w = w.permute(0, 1, 3, 4, 2).reshape(w.shape[0], w.shape[1], -1)    # taps (i, j, a), a fastest
        kc, cin_p, tg = _slice_plan(w.shape[1], w.shape[0], kt)
        pw, pb = _pack_taps(w, conv.bias.detach() if conv.bias is not None else None, cin_p, tgroup=tg, kc=kc)
        return pw, pb, cin_p, tg
    return _cached(conv, "plain" + os.environ.get("VCOF_CONV_KC", ""), _ver(conv.weight, conv.bias), build)


def pack_conv_lines(conv):
    """Weight pack of the line-resident kernel (vcof_conv_lines): bf16 [cin_p/32 * kt * 9, Cout_p16, 32]
    with slice ((chunk * kt + dt) * 3 + dh) * 3 + dw; returns (pw, bias, cin_p, kt)."""
    def build():
        w = conv.weight.detach().float()
        if w.dim() == 4:
            w = w.unsqueeze(2)
        cout, cin, kt = w.shape[0], w.shape[1], w.shape[2]
        cin_p, cout_p = _pad32(cin), _pad16(cout)
        wp = torch.zeros((cout_p, cin_p, kt, 3, 3), dtype=torch.float32, device=w.device)
        wp[:cout, :cin] = w
        wp = wp.view(cout_p, cin_p // 32, 32, kt, 3, 3).permute(1, 3, 4, 5, 0, 2)       # [chunk, dt, dh, dw, n, 32]
        b = torch.zeros(cout_p, dtype=torch.float32, device=w.device)
        if conv.bias is not None:
            b[:cout] = conv.bias.detach().float()
        return wp.reshape(-1, cout_p, 32).to(torch.bfloat16).contiguous(), b, cin_p, kt
    return _cached(conv, "lines", _ver(conv.weight, conv.bias), build)


def _lines_plan(n_total, fused_act=False):
    """(n_tile, rows) of vcof_conv_lines: channels per pass and output rows per work item, favouring weight reuse over
    rows.  The kernel keeps nacc = min(5, 512 // roundup32(n_tile)) row accumulators in a TMEM ring.  Plain layers take
    as many rows as fit (4: the weights of a phase serve more rows); layers with the fused RMS_norm + SiLU epilogue
    (~2500 clk per row) leave two accumulators spare, so the next work item's first rows do not wait for the epilogue
    to drain the burst of rows that complete in an item's last phase: 96 -> 96 + act at 720p 5.1 -> 4.5 ms, while plain
    192- / 384-channel layers lose 1-8 % with fewer rows (profiles/r2_gpurun20_conv_ring_ab.log).
    VCOF_CONV_SPARE=n forces n spare accumulators everywhere (A/B runs)."""
    if n_total <= 128:
        n_tile = n_total
    elif n_total % 128 == 0:
        n_tile = 128
    elif n_total % 96 == 0:
        n_tile = 96
    else:
        n_tile = 64 if n_total % 64 == 0 else 16
    nacc = min(5, 512 // ((n_tile + 31) // 32 * 32))
    spare = os.environ.get("VCOF_CONV_SPARE")
    spare = int(spare) if spare is not None else (2 if fused_act else 0)
    rows = max(1, min(4, nacc - spare))
    return n_tile, rows


def pack_upsample_conv(conv):
    """Fold nearest-2x + Conv2d 3x3 (pad 1) into four parity 2x2 convolutions on the low-res input.

    out(2i+ph, 2j+pw) = sum_{dh,dw} W[dh,dw] * in(floor((2i+ph+dh-1)/2), floor((2j+pw+dw-1)/2)); the three
    taps along each axis collapse onto two source offsets, whose weights are summed (in fp32)."""
    def build():
        w = conv.weight.detach().float()          # [Cout, Cin, 3, 3]
        kc, cin_p, _ = _slice_plan(w.shape[1], w.shape[0])
        packs = {}
        # source offset of kernel index d for output parity p: floor((p + d - 1) / 2)
        for ph in (0, 1):
            for pw_ in (0, 1):
                offs_h = sorted({(ph + d - 1) // 2 for d in range(3)})
                offs_w = sorted({(pw_ + d - 1) // 2 for d in range(3)})
                taps, mats = [], []
                for oh in offs_h:
                    for ow in offs_w:
                        m = torch.zeros_like(w[:, :, 0, 0])
                        for dh in range(3):
                            for dw in range(3):
                                if (ph + dh - 1) // 2 == oh and (pw_ + dw - 1) // 2 == ow:
                                    m = m + w[:, :, dh, dw]
                        mats.append(m)
                        taps.append((0, ow, 0, oh, 0))
                wt = torch.stack(mats, dim=2)                      # [Cout, Cin, 4]
                packs[(ph, pw_)] = _pack_taps(wt, conv.bias.detach(), cin_p, kc=kc) + (taps,)
        return packs, cin_p
    return _cached(conv, "up" + os.environ.get("VCOF_CONV_KC", ""), _ver(conv.weight, conv.bias), build)


def _vec(mod, name, t):
    return _cached(mod, "vec_" + name, _ver(t), lambda: t.detach().float().reshape(-1).contiguous())


def _bf(mod, name, t):
    return _cached(mod, "bf_" + name, _ver(t), lambda: t.detach().to(torch.bfloat16).contiguous())


# ----------------------------------------------------------------------------------------------
# temporal sharding (multi-GPU): contiguous frame ranges per rank, 2-frame halos per causal convolution
# ----------------------------------------------------------------------------------------------
class TimeShard:
    """Temporal sharding context of one VAE call.

    Rank r owns a contiguous range of frames at every level of the network (latent frame i > 0 maps to 2 frames
    after each temporal up-sampler, so ownership stays aligned to latent frames).  Every activation tensor is
    allocated with HALO leading frames; a causal (k_t = 3) convolution first receives the last two frames of its
    input from rank r-1 into that halo (rank 0 keeps zeros = the causal padding, reference wan_vae.py:26-40) and
    then runs the ordinary TMA-tiled kernel on the haloed view.  New design: the reference has no VAE
    parallelism in-tree (the hook is the proprietary paifuser.parallel_magvit_vae)."""
    HALO = 2

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.active = self.world          # ranks that hold frames under the current plan (see plan())

    def plan(self, f):
        """Latent frames per rank for a clip of f latent frames: the first `active` = min(P, f // 2) ranks share the
        frames (a rank needs two frames of its own to fill its right neighbour's halo), the others get none and only
        join the final gather — so a 10-latent-frame clip still shards over 5 of 8 ranks instead of running whole on
        every rank.  Returns None when fewer than two ranks would be active."""
        active = min(self.world, f // 2)
        if active < 2:
            return None
        base, extra = divmod(f, active)
        self.active = active
        return [base + (1 if r < extra else 0) if r < active else 0 for r in range(self.world)]

    def exchange(self, xh, send=None):
        """xh: haloed buffer [HALO + T, H, W, C]; fills xh[:HALO] with the left neighbour's last HALO frames
        (`send` overrides what this rank ships to the right, default xh[-HALO:]).  Only the active ranks of the
        current plan take part."""
        dist = self.dist
        ops_ = []
        if self.rank >= self.active:
            return
        if self.rank > 0:
            recv = torch.empty_like(xh[:self.HALO])
            ops_.append(dist.P2POp(dist.irecv, recv, dist.get_global_rank(self.group, self.rank - 1), self.group))
        else:
            xh[:self.HALO].zero_()
        if self.rank + 1 < self.active:
            out = (xh[-self.HALO:] if send is None else send).contiguous()
            ops_.append(dist.P2POp(dist.isend, out, dist.get_global_rank(self.group, self.rank + 1), self.group))
        if ops_:
            for w in dist.batch_isend_irecv(ops_):
                w.wait()
        if self.rank > 0:
            xh[:self.HALO].copy_(recv)


    def gather_frames(self, out, counts):
        """All-gather per-rank frame ranges [C, counts[r], H, W] into the full clip on every rank."""
        pad = torch.zeros((out.shape[0], max(counts)) + tuple(out.shape[2:]), dtype=out.dtype, device=out.device)
        pad[:, :out.shape[1]] = out
        allp = torch.empty((self.world,) + tuple(pad.shape), dtype=out.dtype, device=out.device)
        self.dist.all_gather_into_tensor(allp, pad, group=self.group)
        return torch.cat([allp[r, :, :counts[r]] for r in range(self.world)], dim=1)


_SHARD = None      # active TimeShard of the current encode/decode call (None = single GPU)


def _alloc(T, H, W, C, device, zero=False):
    """Channels-last activation [T, H, W, C]; under temporal sharding it is the tail of a buffer that carries
    TimeShard.HALO extra leading frames (see _haloed)."""
    extra = TimeShard.HALO if _SHARD is not None else 0
    mk = torch.zeros if zero else torch.empty
    buf = mk((T + extra, H, W, C), dtype=torch.bfloat16, device=device)
    return buf[extra:]


def _haloed(x):
    """Zero-copy view [HALO + T, H, W, C] of a tensor made by _alloc."""
    h = TimeShard.HALO
    off = x.storage_offset() - h * x.stride(0)
    if off < 0:
        raise VcofError("activation was not allocated with a temporal halo")
    return x.as_strided((x.shape[0] + h,) + tuple(x.shape[1:]), x.stride(), off)


# ----------------------------------------------------------------------------------------------
# channels-last building blocks: x is bf16 [T, H, W, C] contiguous
# ----------------------------------------------------------------------------------------------
def _geom(T, H, W, n_total, n_tile, ot=(1, 0), oh=(1, 0), ow=(1, 0), Hs=None, Ws=None, t_stride=1, half=0,
          n_store=None):
    return [T, H, W, t_stride, n_total, n_tile, ot[0], ot[1], oh[0], oh[1], ow[0], ow[1],
            H if Hs is None else Hs, W if Ws is None else Ws, half, n_total if n_store is None else n_store]


def _view5(x):
    """(dims, strides) of the plain 5-D TMA view (C, W, 1, H, T) of a channels-last tensor."""
    T, H, W, C = x.shape
    return (C, W, 1, H, T), (x.stride(2), x.stride(1), x.stride(1), x.stride(0))


def _ntile(n_total):
    return n_total if n_total <= 384 else 384


def conv_causal(x, conv, residual=None, clamp=0.0, n_store=None, act_norm=None, want_raw=True):
    """CausalConv3d / per-frame Conv2d with 'same' padding, stride 1 (reference :21-40, :142-145).

    act_norm: an RMS_norm module whose norm+SiLU (the opening of the NEXT layer) is fused into the epilogue;
    returns (raw, act) then (raw is None when want_raw is False)."""
    T, H, W, C = x.shape
    if C != _pad32(conv.weight.shape[1]):
        raise VcofError(f"conv input has {C} channels, the layer expects {_pad32(conv.weight.shape[1])}")
    k = conv.kernel_size
    kt, kh, kw = (1, k[0], k[1]) if len(k) == 2 else k
    if kh == 3 and kw == 3 and kt in (1, 3) and _use_lines(conv):
        return _conv_causal_lines(x, conv, kt, residual, clamp, n_store, act_norm, want_raw)
    pw, pb, cin_p, tg = pack_conv(conv)
    t_shift = 0
    xin = x
    if _SHARD is not None and kt > 1:
        # temporal sharding: pull the two preceding frames from the left neighbour into the halo, then index time
        # from the start of the haloed buffer (no negative coordinates: rank 0's halo is the causal zero padding)
        xin = _haloed(x)
        _SHARD.exchange(xin)
        t_shift = TimeShard.HALO
    taps = [(0, j - kw // 2, 0, i - kh // 2, a - (kt - 1) + t_shift)
            for i in range(kh) for j in range(kw) for a in range(kt)]
    n_total = pw.shape[1]
    ns = n_total if n_store is None else n_store
    ldc = (ns + 7) // 8 * 8
    dims, strides = _view5(xin)
    geom = _geom(T, H, W, n_total, _ntile(n_total), n_store=ns)
    if act_norm is not None and n_total <= 384:
        out = _alloc(T, H, W, ldc, x.device) if want_raw else None
        act = _alloc(T, H, W, ldc, x.device)
        ops.conv_igemm(xin, dims, strides, pw, taps, cin_p, geom, pb, out, residual=residual, clamp=clamp,
                       act_out=act, act_gamma=_vec(act_norm, "gamma", act_norm.gamma), tgroup=tg)
        return out, act
    out = _alloc(T, H, W, ldc, x.device, zero=(ldc != ns))
    ops.conv_igemm(xin, dims, strides, pw, taps, cin_p, geom, pb, out, residual=residual, clamp=clamp, tgroup=tg)
    if act_norm is not None:
        return out, rms_silu(out, act_norm)
    return out


def _use_lines(conv):
    """Which kernel runs a stride-1 3x3(x3) layer.  Default ("auto"): the line-resident kernel for EVERY such layer.
    Round 1 kept the 192- / 384-channel layers on the tap-streaming kernel (several channel passes drop the fused
    RMS_norm+SiLU); since the issue loops became warp-uniform (round 2) the line kernel runs those layers at
    1.1-1.2 PFLOP/s against 0.95-1.0 and wins even with the separate norm pass (9 f x 720p decode 92.4 ms vs 97.2,
    profiles/r2_gpurun7_vae_elect.json).  VCOF_CONV_LINES=0 forces the tap-streaming kernel, "narrow" restores the
    round-1 rule (line kernel up to 128 output channels)."""
    mode = os.environ.get("VCOF_CONV_LINES", "auto")
    if mode == "0":
        return False
    if mode == "narrow":
        return _pad16(conv.weight.shape[0]) <= 128
    return True


def _conv_causal_lines(x, conv, kt, residual, clamp, n_store, act_norm, want_raw):
    """conv_causal through the line-resident kernel (vcof_conv_lines)."""
    pw, pb, cin_p, _ = pack_conv_lines(conv)
    T, H, W, C = x.shape
    t_shift = 0
    xin = x
    if _SHARD is not None and kt > 1:
        xin = _haloed(x)
        _SHARD.exchange(xin)
        t_shift = TimeShard.HALO
    n_total = pw.shape[1]
    ns = n_total if n_store is None else n_store
    ldc = (ns + 7) // 8 * 8
    dims, strides = _view5(xin)
    n_tile, rows = _lines_plan(n_total)
    fuse = act_norm is not None and n_tile == n_total
    if fuse:
        n_tile, rows = _lines_plan(n_total, fused_act=True)
    geom = [T, H, W, n_total, n_tile, rows, ns]
    t0 = -(kt - 1) + t_shift
    if fuse:
        out = _alloc(T, H, W, ldc, x.device) if want_raw else None
        act = _alloc(T, H, W, ldc, x.device)
        ops.conv_lines(xin, dims, strides, pw, cin_p, kt, t0, geom, pb, out, residual=residual, clamp=clamp,
                       act_out=act, act_gamma=_vec(act_norm, "gamma", act_norm.gamma))
        return out, act
    out = _alloc(T, H, W, ldc, x.device, zero=(ldc != ns))
    ops.conv_lines(xin, dims, strides, pw, cin_p, kt, t0, geom, pb, out, residual=residual, clamp=clamp)
    if act_norm is not None:
        return out, rms_silu(out, act_norm)
    return out


def conv1x1(x, conv, out_ld=None):
    """1x1(x1) convolution = GEMM over the channels-last matrix (reference :203-204, :509-510, :236-237)."""
    T, H, W, C = x.shape
    w = _cached(conv, "w1x1", _ver(conv.weight), lambda: conv.weight.detach().reshape(conv.weight.shape[0], -1)
                .to(torch.bfloat16).contiguous())
    b = _bf(conv, "b1x1", conv.bias)
    cout = w.shape[0]
    ld = cout if out_ld is None else out_ld
    out = _alloc(T, H, W, ld, x.device, zero=(ld != cout))
    ops.gemm(x.reshape(-1, C)[:, :w.shape[1]], w, b, "bias", out=out.reshape(-1, ld)[:, :cout])
    return out


def rms_silu(x, norm, silu=True):
    out = _alloc(*x.shape, x.device)
    return ops.rms_silu_cl(x, _vec(norm, "gamma", norm.gamma), silu=silu, out=out)


def res_block(x, blk, x_act=None, next_norm=None):
    """reference :190-224.  x_act = silu(rms_norm(x)) if the producer already fused it; with next_norm the
    block returns (y, silu(rms_norm_next(y))) from the last conv's epilogue."""
    h = x if isinstance(blk.shortcut, nn.Identity) else conv1x1(x, blk.shortcut)
    if x_act is None:
        x_act = rms_silu(x, blk.residual[0])
    _, y_act = conv_causal(x_act, blk.residual[2], act_norm=blk.residual[3], want_raw=False)
    if next_norm is None:
        return conv_causal(y_act, blk.residual[6], residual=h), None
    return conv_causal(y_act, blk.residual[6], residual=h, act_norm=next_norm)


def attn_block(x, blk):
    """reference :227-266 — per frame, single head, d = C."""
    T, H, W, C = x.shape
    n = H * W
    y = rms_silu(x, blk.norm, silu=False)
    qkv = conv1x1(y, blk.to_qkv).reshape(T, n, 3 * C)
    if C == 384 and os.environ.get("VCOF_VAE_ATTN", "fused") != "unfused":
        # one fused tcgen05 launch for all frames: scores and probabilities never reach HBM (csrc/vae_attn_sm100.cu)
        o = ops.vae_attn(qkv, C)
        a = conv1x1(o.reshape(T, H, W, C), blk.proj)
        out = _alloc(T, H, W, C, x.device)
        torch.add(x, a, out=out)
        return out
    # unfused cross-check path (round 1): GEMM with raw fp32 scores -> row softmax -> GEMM, per frame
    o = torch.empty((T, n, C), dtype=torch.bfloat16, device=x.device)
    npad = (n + 7) // 8 * 8
    for t in range(T):
        q, k, v = qkv[t, :, :C], qkv[t, :, C:2 * C], qkv[t, :, 2 * C:]
        s = torch.empty((n, npad), dtype=torch.float32, device=x.device)[:, :n]
        ops.gemm(q, k, None, "raw_f32", out=s)
        p = ops.softmax_rows(s, 1.0 / math.sqrt(C))
        vt = torch.zeros((C, npad), dtype=torch.bfloat16, device=x.device)
        vt[:, :n] = v.t()
        ops.gemm(p, vt[:, :n], None, "bias", out=o[t])
    a = conv1x1(o.reshape(T, H, W, C), blk.proj)
    out = _alloc(T, H, W, C, x.device)
    torch.add(x, a, out=out)
    return out


def downsample(x, rs):
    """reference :91-100, :147-163."""
    conv = rs.resample[1]
    pw, pb, cin_p, _ = pack_conv(conv)
    T, H, W, C = x.shape
    H2, W2 = H // 2, W // 2
    # parity view (c_inner = pw*C + c, w2, ph, h2, t): input row 2*h2 + ph, column 2*w2 + pw
    dims = (2 * C, W2, 2, H2, T)
    strides = (2 * x.stride(2), x.stride(1), 2 * x.stride(1), x.stride(0))
    taps = [((dw % 2) * C, dw // 2, dh % 2, dh // 2, 0) for dh in range(3) for dw in range(3)]
    n_total = pw.shape[1]
    y = _alloc(T, H2, W2, n_total, x.device)
    ops.conv_igemm(x, dims, strides, pw, taps, cin_p, _geom(T, H2, W2, n_total, _ntile(n_total)), pb, y)
    first = _SHARD is None or _SHARD.rank == 0
    if rs.mode == "downsample3d" and (T > 1 or not first):
        tw, tb, tcin, ttg = pack_conv(rs.time_conv)
        if _SHARD is None:
            To = (T - 1) // 2
            z = _alloc(1 + To, H2, W2, n_total, x.device)
            z[0].copy_(y[0])
            d5, s5 = _view5(y)
            ttaps = [(0, 0, 0, 0, a) for a in range(3)]
            ops.conv_igemm(y, d5, s5, tw, ttaps, tcin,
                           _geom(To, H2, W2, n_total, _ntile(n_total), ot=(1, 1), t_stride=2), tb, z, tgroup=ttg)
        else:
            # out[1+m] = conv(y[2m], y[2m+1], y[2m+2]) on GLOBAL frame indices; ownership follows the reference's
            # (1, 4, 4, ...) chunking, so a rank > 0 owns an even number of frames and needs exactly the last
            # frame of its left neighbour (second halo slot); rank 0 keeps frame 0 as is.
            yh = _haloed(y)
            _SHARD.exchange(yh)
            if first:
                To = (T - 1) // 2
                z = _alloc(1 + To, H2, W2, n_total, x.device)
                z[0].copy_(y[0])
                shift, ot = 2, (1, 1)
            else:
                To = T // 2
                z = _alloc(To, H2, W2, n_total, x.device)
                shift, ot = 1, (1, 0)
            d5, s5 = _view5(yh)
            ttaps = [(0, 0, 0, 0, a + shift) for a in range(3)]
            if To > 0:
                ops.conv_igemm(yh, d5, s5, tw, ttaps, tcin,
                               _geom(To, H2, W2, n_total, _ntile(n_total), ot=ot, t_stride=2), tb, z, tgroup=ttg)
        y = z
    return y


def upsample(x, rs):
    """reference :80-89, :107-145."""
    T, H, W, C = x.shape
    first = _SHARD is None or _SHARD.rank == 0        # does this rank hold global frame 0?
    if rs.mode == "upsample3d" and (T > 1 or not first):
        tw, tb, tcin, ttg = pack_conv(rs.time_conv)           # 2C outputs, 3 temporal taps
        n_total = tw.shape[1]
        if _SHARD is None:
            To = 1 + 2 * (T - 1)
            z = _alloc(To, H, W, C, x.device)
            z[0].copy_(x[0])
            xs = x[1:]
            d5, s5 = _view5(xs)
            ttaps = [(0, 0, 0, 0, a - 2) for a in range(3)]
            # channels [0,C) -> frame 1+2t, channels [C,2C) -> frame 2+2t
            ops.conv_igemm(xs, d5, s5, tw, ttaps, tcin,
                           _geom(T - 1, H, W, n_total, _ntile(n_total), ot=(2, 1), half=C), tb, z, tgroup=ttg)
        else:
            # The temporal conv runs over global frames 1.. with zero history (frame 0 is NOT part of it).  Rank 0
            # therefore ships [0, frame1] when it owns only two frames; every other rank ships its last two.
            xh = _haloed(x)
            send = None
            if first and T < 3:
                send = torch.zeros_like(xh[-2:])
                if T == 2:
                    send[1].copy_(x[1])
            _SHARD.exchange(xh, send=send)
            skip = 1 if first else 0                              # frame 0 bypasses the conv (:107-112)
            To = (1 if first else 0) + 2 * (T - skip)
            z = _alloc(To, H, W, C, x.device)
            if first:
                z[0].copy_(x[0])
                xh[2].zero_()     # global frame 0 is not part of the time_conv sequence: it reads as zero padding
            d5, s5 = _view5(xh)
            ttaps = [(0, 0, 0, 0, a + skip) for a in range(3)]    # own frame t (+skip) sits at xh[2 + skip + t]
            if T - skip > 0:
                ops.conv_igemm(xh, d5, s5, tw, ttaps, tcin,
                               _geom(T - skip, H, W, n_total, _ntile(n_total), ot=(2, skip), half=C), tb, z,
                               tgroup=ttg)
        x = z
        T = To
    packs, cin_p = pack_upsample_conv(rs.resample[1])
    cout = rs.resample[1].weight.shape[0]
    y = _alloc(T, 2 * H, 2 * W, cout, x.device)
    d5, s5 = _view5(x)
    for (ph, pw_), (w4, b4, taps) in packs.items():
        n_total = w4.shape[1]
        ops.conv_igemm(x, d5, s5, w4, taps, cin_p,
                       _geom(T, H, W, n_total, _ntile(n_total), oh=(2, ph), ow=(2, pw_), Hs=2 * H, Ws=2 * W,
                             n_store=cout), b4, y)
    return y


# ----------------------------------------------------------------------------------------------
# model
# ----------------------------------------------------------------------------------------------
class AutoencoderKLWan_(nn.Module):
    """reference :487-596 (encode/decode un-chunked; `scale` = [mean, 1/std])."""

    def __init__(self, dim=128, z_dim=4, dim_mult=(1, 2, 4, 4), num_res_blocks=2, attn_scales=(),
                 temperal_downsample=(True, True, False), dropout=0.0):
        super().__init__()
        self.dim, self.z_dim, self.dim_mult = dim, z_dim, list(dim_mult)
        self.num_res_blocks, self.attn_scales = num_res_blocks, list(attn_scales)
        self.temperal_downsample = list(temperal_downsample)
        self.temperal_upsample = self.temperal_downsample[::-1]
        self.encoder = Encoder3d(dim, z_dim * 2, dim_mult, num_res_blocks, attn_scales, self.temperal_downsample, dropout)
        self.conv1 = CausalConv3d(z_dim * 2, z_dim * 2, 1)
        self.conv2 = CausalConv3d(z_dim, z_dim, 1)
        self.decoder = Decoder3d(dim, z_dim, dim_mult, num_res_blocks, attn_scales, self.temperal_upsample, dropout)

    def _check(self, x):
        if not x.is_cuda:
            raise VcofError("AutoencoderKLWan needs CUDA tensors: libvcof has no CPU path")
        w = self.conv1.weight
        if w.device != x.device or w.dtype != torch.bfloat16:
            raise VcofError(f"VAE weights must be bf16 on {x.device} (got {w.dtype} on {w.device}); "
                            "call .to(device, torch.bfloat16) as the reference CLIs do (fast_infer.py:300-303)")

    @staticmethod
    def _run(x, layers, x_act=None, tail_norm=None):
        """Walk a layer list.  RMS_norm+SiLU that opens a ResidualBlock (or `tail_norm`, the head's norm) is
        produced by the previous ResidualBlock's last convolution when there is one."""
        layers = list(layers)
        for i, layer in enumerate(layers):
            nxt = layers[i + 1] if i + 1 < len(layers) else None
            next_norm = nxt.residual[0] if isinstance(nxt, ResidualBlock) else (tail_norm if nxt is None else None)
            if isinstance(layer, ResidualBlock):
                x, x_act = res_block(x, layer, x_act, next_norm)
            elif isinstance(layer, AttentionBlock):
                x, x_act = attn_block(x, layer), None
            elif isinstance(layer, Resample):
                x = downsample(x, layer) if layer.mode.startswith("down") else upsample(x, layer)
                x_act = None
            else:
                raise VcofError(f"unexpected layer {type(layer)}")
        return x, x_act

    def encode(self, x, scale, shard=None):
        """x [1, 3, T, H, W] -> [1, 2*z, f, H/8, W/8] = cat(normalised mu, logvar) (:520-548).  x may also be the byte
        frames uint8 [1, T, H, W, 3] of the clip: the [-1, 1] scaling of load_video_frames (fast_infer.py:86-88) and
        the cast to bf16 (pipeline_wan.py:397) then happen in the layout kernel (bit-identical input to the first conv).

        shard: optional TimeShard — rank r encodes the video frames of its latent-frame range (frame 0 alone, then 4
        per latent: the reference's own chunk boundaries) with 2-frame halos; the latents are all-gathered."""
        global _SHARD
        self._check(x)
        if x.shape[0] != 1:
            raise VcofError("encode expects batch 1 (the reference loops over the batch, :647-653)")
        as_bytes = x.dtype == torch.uint8
        if as_bytes and (x.dim() != 5 or x.shape[4] != 3):
            raise VcofError(f"byte frames must be uint8 [1, T, H, W, 3], got {tuple(x.shape)}")
        T, H, W = (x.shape[1], x.shape[2], x.shape[3]) if as_bytes else (x.shape[2], x.shape[3], x.shape[4])
        if (T - 1) % 4 != 0 or H % 8 or W % 8:
            raise VcofError("video must have 1+4k frames and H, W multiples of 8")
        enc = self.encoder
        f = (T - 1) // 4 + 1
        sizes = shard.plan(f) if (shard is not None and shard.world > 1) else None
        sharded = sizes is not None
        t0, t1 = 0, T
        if sharded:
            a0 = sum(sizes[:shard.rank])
            b0 = a0 + sizes[shard.rank]
            t0, t1 = (0 if a0 == 0 else 4 * a0 - 3), 4 * b0 - 3
            if sizes[shard.rank] == 0:          # idle rank of this plan: nothing to encode, join the gather
                empty = torch.empty((2 * self.z_dim, 0, H // 8, W // 8), dtype=torch.bfloat16, device=x.device)
                return shard.gather_frames(empty, sizes)[None]
        if as_bytes:
            h = ops.u8_to_cl(x[0, t0:t1].contiguous(), 32)
        else:
            h = ops.nchw_to_cl(x[0].to(torch.bfloat16)[:, t0:t1].contiguous(), 32)
        if sharded:
            _SHARD = shard
            own = _alloc(*h.shape, h.device)
            own.copy_(h)
            h = own
        try:
            h = conv_causal(h, enc.conv1)
            h, a = self._run(h, enc.downsamples)
            h, a = self._run(h, enc.middle, a, tail_norm=enc.head[0])
            if a is None:
                a = rms_silu(h, enc.head[0])
            h = conv_causal(a, enc.head[2])                        # [f_r, h, w, 32]
        finally:
            _SHARD = None
        h = conv1x1(h, self.conv1)
        z = self.z_dim
        mean = scale[0].to(x.device, torch.bfloat16).float().contiguous()
        inv_std = scale[1].to(x.device, torch.bfloat16).float().contiguous()
        mu = ops.cl_to_nchw(h, z, sub=mean, mul=inv_std)
        logvar = ops.cl_to_nchw(h[..., z:], z)
        out = torch.cat([mu, logvar], dim=0)                       # [2z, f_r, h, w]
        if sharded:
            out = shard.gather_frames(out, sizes)
        return out[None]

    def decode(self, z, scale, shard=None, as_bytes=False):
        """z [1, 16, f, h, w] (normalised) -> [1, 3, 4(f-1)+1, 8h, 8w] (:550-575); clamp is applied by the caller
        in the reference (:669) and fused into the last convolution here.  as_bytes: return the byte frames uint8
        [1, T, 8h, 8w, 3] the reference's host chain would make of that output — decode_latents' bf16
        (x / 2 + 0.5).clamp(0, 1) -> fp32 (pipeline_wan.py:425-427), save_videos_grid's (x * 255).astype(uint8)
        (utils/utils.py:66) — bit-exactly, converted on the device.

        shard: optional TimeShard — the latent frames are split into contiguous per-rank ranges, every causal
        convolution exchanges a 2-frame halo with the left neighbour, and the decoded frames are all-gathered."""
        global _SHARD
        self._check(z)
        if z.shape[0] != 1:
            raise VcofError("decode expects batch 1 (the reference loops over the batch, :667-674)")
        dec = self.decoder
        mean = scale[0].to(z.device, torch.bfloat16).float().contiguous()
        inv_std = scale[1].to(z.device, torch.bfloat16).float().contiguous()
        h = ops.nchw_to_cl(z[0].to(torch.bfloat16).contiguous(), self.z_dim, div=inv_std, add=mean)
        h = conv1x1(h, self.conv2, out_ld=32)                  # 16 -> 16, stored in a 32-channel (zero padded) tensor
        f = h.shape[0]
        sizes = shard.plan(f) if (shard is not None and shard.world > 1) else None
        sharded = sizes is not None
        if sharded:
            counts = [4 * n - (3 if r == 0 else 0) for r, n in enumerate(sizes)]
            if sizes[shard.rank] == 0:          # idle rank of this plan: nothing to decode, join the gather
                Hh, Ww = 8 * h.shape[1], 8 * h.shape[2]
                if as_bytes:
                    empty = torch.empty((1, 0, Hh, Ww, 3), dtype=torch.uint8, device=z.device)
                    return shard.gather_frames(empty, counts)[0][None]
                empty = torch.empty((3, 0, Hh, Ww), dtype=torch.bfloat16, device=z.device)
                return shard.gather_frames(empty, counts)[None]
            a = sum(sizes[:shard.rank])
            _SHARD = shard
            own = _alloc(sizes[shard.rank], h.shape[1], h.shape[2], h.shape[3], h.device)
            own.copy_(h[a:a + sizes[shard.rank]])
            h = own
        try:
            h = conv_causal(h, dec.conv1)
            h, a_ = self._run(h, dec.middle)
            h, a_ = self._run(h, dec.upsamples, a_, tail_norm=dec.head[0])
            if a_ is None:
                a_ = rms_silu(h, dec.head[0])
            h = conv_causal(a_, dec.head[2], clamp=1.0, n_store=3)  # [T, H, W, 8] (3 real channels)
        finally:
            _SHARD = None
        if as_bytes:
            out = ops.cl_to_u8(h, 3)                                 # [T_r, H, W, 3]
            if sharded:
                out = shard.gather_frames(out[None], counts)[0]
            return out[None]
        out = ops.cl_to_nchw(h, 3)                                   # [3, T_r, H, W]
        if sharded:
            out = shard.gather_frames(out, counts)
        return out[None]

    def clear_cache(self):
        """The reference's streaming caches do not exist here; kept for API compatibility (:589-596)."""


class DiagonalGaussianDistribution:
    """Enough of diffusers' class for the pipeline: `.mode()` (pipeline_wan.py:407), mean / logvar / sample."""

    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise


class AutoencoderKLOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist

    def __getitem__(self, i):
        return (self.latent_dist,)[i]


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


class _Config(dict):
    __getattr__ = dict.get


class AutoencoderKLWan(nn.Module):
    """Drop-in for the reference class (:620-705)."""

    def __init__(self, latent_channels=16, temporal_compression_ratio=4, spatial_compression_ratio=8):
        super().__init__()
        self.config = _Config(latent_channels=latent_channels, temporal_compression_ratio=temporal_compression_ratio,
                              spatial_compression_ratio=spatial_compression_ratio)
        self.latent_channels = latent_channels
        self.temporal_compression_ratio = temporal_compression_ratio
        self.spatial_compression_ratio = spatial_compression_ratio
        mean = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
                0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
        std = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
               3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]
        self.mean = torch.tensor(mean, dtype=torch.float32)
        self.std = torch.tensor(std, dtype=torch.float32)
        self.scale = [self.mean, 1.0 / self.std]
        self.model = AutoencoderKLWan_(dim=96, z_dim=latent_channels, dim_mult=[1, 2, 4, 4], num_res_blocks=2,
                                       attn_scales=[], temperal_downsample=[False, True, True], dropout=0.0)

    @property
    def dtype(self):
        return self.model.conv1.weight.dtype

    @property
    def device(self):
        return self.model.conv1.weight.device

    def _encode(self, x):
        shard = getattr(self, "_shard", None)
        return torch.cat([self.model.encode(u.unsqueeze(0), self.scale, shard=shard) for u in x], dim=0)

    def encode(self, x, return_dict=True):
        posterior = DiagonalGaussianDistribution(self._encode(x))
        return AutoencoderKLOutput(latent_dist=posterior) if return_dict else (posterior,)

    def encode_frames(self, frames, return_dict=True):
        """`encode` for byte frames uint8 [B, T, H, W, 3] (what load_video_frames stacks before its float scaling,
        fast_infer.py:86): a quarter of the fp32 bytes cross PCIe and the scaling runs inside the layout kernel."""
        if frames.dtype != torch.uint8:
            raise VcofError(f"encode_frames expects uint8 frames, got {frames.dtype}")
        return self.encode(frames, return_dict=return_dict)

    def decode_frames(self, z):
        """z [B, 16, f, h, w] -> byte frames uint8 [B, T, H, W, 3] on the device, bit-identical to the reference's
        decode -> decode_latents -> save_videos_grid chain (pipeline_wan.py:424-427, utils/utils.py:66)."""
        shard = getattr(self, "_shard", None)
        return torch.cat([self.model.decode(u.unsqueeze(0), self.scale, shard=shard, as_bytes=True) for u in z], dim=0)

    def enable_temporal_sharding(self, group=None):
        """Encode / decode with the frames sharded over the ranks of `group` (default WORLD): 2-frame halo exchange
        per causal convolution (TimeShard); single-rank behaviour is unchanged.  Clips with fewer than two latent
        frames per rank (e.g. the 1-latent grounding segment) are processed redundantly on every rank."""
        self._shard = TimeShard(group)

    def _decode(self, zs):
        shard = getattr(self, "_shard", None)
        return DecoderOutput(sample=torch.cat([self.model.decode(u.unsqueeze(0), self.scale, shard=shard)
                                               for u in zs], dim=0))

    def decode(self, z, return_dict=True):
        decoded = self._decode(z).sample
        return DecoderOutput(sample=decoded) if return_dict else (decoded,)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, additional_kwargs={}):
        """reference :684-705 — a single .pth / .safetensors state dict whose keys get the `model.` prefix."""
        import inspect
        valid = set(inspect.signature(cls.__init__).parameters) - {"self"}
        model = cls(**{k: v for k, v in additional_kwargs.items() if k in valid})
        if pretrained_model_path.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(pretrained_model_path)
        else:
            sd = torch.load(pretrained_model_path, map_location="cpu")
        m, u = model.load_state_dict({"model." + k: v for k, v in sd.items()}, strict=False)
        print(f"### missing keys: {len(m)}; \n### unexpected keys: {len(u)};")
        return model
