"""Tensor-level wrappers over the C ABI (include/vcof.h).

PyTorch is used for device memory and streams only: every function here checks its
arguments, allocates the output with torch, and launches exactly one libvcof kernel on
the current CUDA stream.  No function has a PyTorch/CPU fallback — a CPU tensor or a
missing libvcof.so raises.
"""
import ctypes as _ct
import math

import torch

from . import _lib

EPI = {"bias": 0, "bias_gelu": 1, "bias_gate_res": 2, "bias_f32": 3, "raw_f32": 4, "gate_accum": 5, "mul": 6, "add": 7}

# kernel launches since the last reset (bench.py reports it as gpu_launches)
_launches = 0


def launches():
    return _launches


def reset_launches():
    global _launches
    _launches = 0


# optional per-launch CUDA-event timing (bench.py roofline leg / tools/kbench.py)
_timing = None


def enable_timing():
    """Record a CUDA-event pair around every libvcof launch until collect_timing() is called."""
    global _timing
    _timing = []


def collect_timing():
    """-> {key: (launches, total_ms)}; key = kernel name + problem signature.  Synchronises."""
    global _timing
    rec, _timing = _timing or [], None
    torch.cuda.synchronize()
    out = {}
    for key, e0, e1 in rec:
        n, ms = out.get(key, (0, 0.0))
        out[key] = (n + 1, ms + e0.elapsed_time(e1))
    return out


def _call(name, *args, key=None):
    global _launches
    _launches += 1
    if _timing is None:
        _lib.call(name, *args)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call(name, *args)
    e1.record()
    _timing.append((key or name, e0, e1))


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name, dims=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.VcofError(f"{name}: expected a CUDA tensor (libvcof has no CPU path)")
    if t.dtype != dtype:
        raise _lib.VcofError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if dims is not None and t.dim() != dims:
        raise _lib.VcofError(f"{name}: expected {dims} dims, got shape {tuple(t.shape)}")
    if t.dim() >= 1 and t.stride(-1) != 1:
        raise _lib.VcofError(f"{name}: innermost dimension must be contiguous")
    return t


def narrow_tiles_pay(M, N, sms=148):
    """True when 128-wide GEMM tiles finish sooner than 256-wide ones: cost = waves over the SMs x tile width.  Only
    problems of at most two 256-wide waves qualify, so large GEMMs never change shape."""
    num_m = (M + 127) // 128
    t256, t128 = num_m * ((N + 255) // 256), num_m * ((N + 127) // 128)
    return N > 128 and t256 <= 2 * sms and -(-t128 // sms) * 128 < -(-t256 // sms) * 256


def gemm(a, w, bias=None, epilogue="bias", out=None, gate=None, narrow=False):
    """out = epilogue(a @ w.T + bias).  a [M,K] bf16, w [N,K] bf16 (nn.Linear layout).

    epilogue: "bias" -> bf16 [M,N]; "bias_gelu" -> bf16 gelu_tanh; "bias_f32" -> fp32 (value
    rounded through bf16); "bias_gate_res" -> `out` (fp32 [M,N], required) += gate * bf16(.);
    "gate_accum" -> `out` (bf16 [M,N], required) = bf16(float(out) + gate[n] * (a @ w.T)) (LoRA merge);
    "mul" / "add" -> `out` (bf16 [M,N], required) = bf16(float(out) * or + bf16(a @ w.T + bias)) (text encoder).
    narrow: 128-wide output tiles (VCOF_GEMM_TILE128) — same result, better SM occupancy for skinny problems.
    """
    _chk(a, torch.bfloat16, "gemm.a", 2)
    _chk(w, torch.bfloat16, "gemm.w", 2)
    M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise _lib.VcofError(f"gemm: K mismatch {K} vs {K2}")
    epi = EPI[epilogue]
    if bias is not None:
        _chk(bias, torch.bfloat16, "gemm.bias", 1)
    if gate is not None:
        _chk(gate, torch.float32, "gemm.gate", 1)
    if epi == 2:
        if out is None:
            raise _lib.VcofError("gemm: bias_gate_res needs the fp32 residual in `out`")
        _chk(out, torch.float32, "gemm.out", 2)
    elif epi in (3, 4):
        if out is None:
            out = torch.empty((M, N), dtype=torch.float32, device=a.device)
        _chk(out, torch.float32, "gemm.out", 2)
    elif epi == 5:
        if out is None or gate is None or bias is not None:
            raise _lib.VcofError("gemm: gate_accum updates `out` in place with a per-column gate and no bias")
        _chk(out, torch.bfloat16, "gemm.out", 2)
    elif epi in (6, 7):
        if out is None:
            raise _lib.VcofError(f"gemm: {epilogue} updates the bf16 tensor passed as `out` in place")
        _chk(out, torch.bfloat16, "gemm.out", 2)
    else:
        if out is None:
            out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
        _chk(out, torch.bfloat16, "gemm.out", 2)
    if tuple(out.shape) != (M, N):
        raise _lib.VcofError(f"gemm: out shape {tuple(out.shape)} != {(M, N)}")
    _call("vcof_gemm_bf16", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _p(bias),
          _p(gate), out.data_ptr(), out.stride(0), M, N, K, epi | (0x100 if narrow else 0), _stream(),
          key=f"gemm[{epilogue}] M={M} N={N} K={K}")
    return out


def attention(q, k, v, heads, kv_len=None, scale=None, out=None, v_transposed=False):
    """softmax(q k^T * scale) v per head.  q [Lq, heads*128], k [Lk, heads*128],
    v [Lk, heads*128] (or v^T [heads*128, Lk'] when v_transposed); bf16."""
    _chk(q, torch.bfloat16, "attention.q", 2)
    _chk(k, torch.bfloat16, "attention.k", 2)
    _chk(v, torch.bfloat16, "attention.v", 2)
    Lq, C = q.shape
    Lk = k.shape[0]
    hd = C // heads
    if kv_len is None:
        kv_len = Lk
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    if out is None:
        out = torch.empty((Lq, C), dtype=torch.bfloat16, device=q.device)
    _chk(out, torch.bfloat16, "attention.out", 2)
    _call("vcof_attn_fwd", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
          v.stride(0), out.data_ptr(), out.stride(0), Lq, Lk, kv_len, heads, hd, float(scale),
          1 if v_transposed else 0, _stream(), key=f"attn Lq={Lq} Lk={kv_len} heads={heads}")
    return out


def attention_scatter(q, k, v, heads, dests, kv_len=None, scale=None):
    """attention() with the output rows scattered: row chunk c (Lq / len(dests) consecutive query rows) goes to
    dests[c], dense bf16 [Lq / len(dests), heads*128] slabs (the owning ranks' receive buffers over NVLink peer memory)."""
    _chk(q, torch.bfloat16, "attention_scatter.q", 2)
    _chk(k, torch.bfloat16, "attention_scatter.k", 2)
    _chk(v, torch.bfloat16, "attention_scatter.v", 2)
    Lq, C = q.shape
    if Lq % _ndest(dests, "attention_scatter.dests"):
        raise _lib.VcofError(f"attention_scatter: {Lq} query rows do not split into {len(dests)} chunks")
    rows = Lq // len(dests)
    arr = _slabs(dests, rows, C, "attention_scatter.dests")
    Lk = k.shape[0]
    hd = C // heads
    kv_len = Lk if kv_len is None else kv_len
    scale = 1.0 / math.sqrt(hd) if scale is None else scale
    _call("vcof_attn_fwd_scatter", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
          arr, len(dests), rows, C, Lq, Lk, kv_len, heads, hd, float(scale), _stream(),
          key=f"attn Lq={Lq} Lk={kv_len} heads={heads}")


def ln_modulate(x, ln_w=None, ln_b=None, shift=None, scale=None, eps=1e-6, out=None):
    """bf16((LayerNorm(x) * ln_w + ln_b) * (1 + scale) + shift); x fp32 [L,C], vectors fp32 [C]."""
    _chk(x, torch.float32, "ln_modulate.x", 2)
    L, C = x.shape
    for n, t in (("ln_w", ln_w), ("ln_b", ln_b), ("shift", shift), ("scale", scale)):
        if t is not None:
            _chk(t, torch.float32, "ln_modulate." + n)
            if t.numel() != C or not t.is_contiguous():
                raise _lib.VcofError(f"ln_modulate.{n}: need a contiguous [{C}] vector")
    if out is None:
        out = torch.empty((L, C), dtype=torch.bfloat16, device=x.device)
    _chk(out, torch.bfloat16, "ln_modulate.out", 2)
    _call("vcof_ln_modulate", x.data_ptr(), x.stride(0), _p(ln_w), _p(ln_b), _p(shift), _p(scale),
          out.data_ptr(), out.stride(0), L, C, float(eps), _stream())
    return out


class RopeSpec:
    """Device-side description of the 3-axis rotary embedding for one sample.

    table: fp32 [1024, head_dim/2, 2] (cos, sin) built from the model's complex128 `freqs`;
    tpos : int32 [F] temporal position per latent frame (plain / paired / chain-of-frames);
    (F,H,W): patch grid; n_t/n_h: number of complex pairs on the temporal / height axes.
    """

    def __init__(self, table, tpos, F, H, W, n_t, n_h, row_offset=0):
        self.table, self.tpos = table, tpos
        self.F, self.H, self.W, self.n_t, self.n_h, self.row_offset = F, H, W, n_t, n_h, row_offset


def _chk_blocked(blocked, rows, C, what):
    _chk(blocked, torch.bfloat16, what, 3)
    nb, r, cp = blocked.shape
    if r != rows or nb * cp != C or not blocked.is_contiguous():
        raise _lib.VcofError(f"{what}: blocked buffer {tuple(blocked.shape)} does not tile [{rows}, {C}]")
    return cp, blocked.stride(0)


def rmsnorm_rope_(x, weight, eps, head_dim, rope=None, out_blocked=None):
    """WanRMSNorm over the full row, then (optionally) RoPE.  x bf16 [L,C]; in place, or — with
    out_blocked = a contiguous [P, L, C/P] buffer — out of place into that column-blocked layout (x untouched)."""
    _chk(x, torch.bfloat16, "rmsnorm_rope.x", 2)
    _chk(weight, torch.bfloat16, "rmsnorm_rope.weight", 1)
    L, C = x.shape
    if out_blocked is not None:
        cp, bs = _chk_blocked(out_blocked, L, C, "rmsnorm_rope.out_blocked")
        r = rope
        _call("vcof_rmsnorm_rope_blocked", x.data_ptr(), x.stride(0), out_blocked.data_ptr(), cp, bs,
              weight.data_ptr(), float(eps), L, C, head_dim, _p(None if r is None else r.table),
              _p(None if r is None else r.tpos), *((1, 1, 1, 0, 0, 0) if r is None else
                                                   (r.F, r.H, r.W, r.n_t, r.n_h, r.row_offset)), _stream())
        return out_blocked
    if rope is None:
        _call("vcof_rmsnorm_rope", x.data_ptr(), x.stride(0), weight.data_ptr(), float(eps), L, C,
              head_dim, None, None, 1, 1, 1, 0, 0, 0, _stream())
    else:
        _chk(rope.table, torch.float32, "rope.table")
        _chk(rope.tpos, torch.int32, "rope.tpos")
        _call("vcof_rmsnorm_rope", x.data_ptr(), x.stride(0), weight.data_ptr(), float(eps), L, C,
              head_dim, rope.table.data_ptr(), rope.tpos.data_ptr(), rope.F, rope.H, rope.W,
              rope.n_t, rope.n_h, rope.row_offset, _stream())
    return x


def copy_blocked(rowmajor, blocked, to_blocked):
    """Pack (to_blocked) a row-major bf16 [rows, C] matrix into a contiguous [P, rows, C/P] buffer, or unpack."""
    _chk(rowmajor, torch.bfloat16, "copy_blocked.rowmajor", 2)
    rows, C = rowmajor.shape
    cp, bs = _chk_blocked(blocked, rows, C, "copy_blocked.blocked")
    _call("vcof_copy_blocked", rowmajor.data_ptr(), rowmajor.stride(0), blocked.data_ptr(), bs, rows, C, cp,
          1 if to_blocked else 0, _stream())
    return blocked if to_blocked else rowmajor


def _ndest(dests, what):
    """Number of destination slabs, checked before anything is divided by it."""
    if not 1 <= len(dests) <= 16:
        raise _lib.VcofError(f"{what}: 1..16 destinations, got {len(dests)}")
    return len(dests)


def _slabs(dests, rows, cols, what):
    """Device addresses of dense bf16 [rows, cols] destination slabs (peer-mapped tensors under sequence parallelism)."""
    _ndest(dests, what)
    for d in dests:
        _chk(d, torch.bfloat16, what, 2)
        if tuple(d.shape) != (rows, cols) or not d.is_contiguous():
            raise _lib.VcofError(f"{what}: destination {tuple(d.shape)} is not a dense [{rows}, {cols}] slab")
    return (_ct.c_void_p * len(dests))(*[d.data_ptr() for d in dests])


def rmsnorm_rope_scatter(x, weight, eps, head_dim, rope, dests):
    """rmsnorm_rope_ with the result scattered: column block b of every row goes to dests[b] (dense [L, C/len(dests)]
    slabs — the receive buffers of the other ranks over NVLink peer memory); x is not modified."""
    _chk(x, torch.bfloat16, "rmsnorm_rope_scatter.x", 2)
    _chk(weight, torch.bfloat16, "rmsnorm_rope_scatter.weight", 1)
    L, C = x.shape
    arr = _slabs(dests, L, C // _ndest(dests, "rmsnorm_rope_scatter.dests"), "rmsnorm_rope_scatter.dests")
    r = rope
    _call("vcof_rmsnorm_rope_scatter", x.data_ptr(), x.stride(0), arr, len(dests), weight.data_ptr(), float(eps), L, C,
          head_dim, _p(None if r is None else r.table), _p(None if r is None else r.tpos),
          *((1, 1, 1, 0, 0, 0) if r is None else (r.F, r.H, r.W, r.n_t, r.n_h, r.row_offset)), _stream())


def copy_scatter(rowmajor, dests):
    """Column block b of the row-major bf16 [rows, C] matrix -> dests[b] (dense [rows, C/len(dests)] slabs)."""
    _chk(rowmajor, torch.bfloat16, "copy_scatter.rowmajor", 2)
    rows, C = rowmajor.shape
    arr = _slabs(dests, rows, C // _ndest(dests, "copy_scatter.dests"), "copy_scatter.dests")
    _call("vcof_copy_scatter", rowmajor.data_ptr(), rowmajor.stride(0), arr, len(dests), rows, C, _stream())


def copy_rows_scatter(src, dests):
    """Row chunk c of the bf16 [len(dests) * rows, cols] matrix -> dests[c] (dense [rows, cols] slabs)."""
    _chk(src, torch.bfloat16, "copy_rows_scatter.src", 2)
    total, cols = src.shape
    if total % _ndest(dests, "copy_rows_scatter.dests"):
        raise _lib.VcofError(f"copy_rows_scatter: {total} rows do not split into {len(dests)} chunks")
    arr = _slabs(dests, total // len(dests), cols, "copy_rows_scatter.dests")
    _call("vcof_copy_rows_scatter", src.data_ptr(), src.stride(0), arr, len(dests), total // len(dests), cols, _stream())


def patchify(x):
    """x bf16 [Cin,F,H,W] -> [F*(H/2)*(W/2), Cin*4] (column order c, ph, pw)."""
    _chk(x, torch.bfloat16, "patchify.x", 4)
    if not x.is_contiguous():
        raise _lib.VcofError("patchify.x must be contiguous")
    Cin, F, H, W = x.shape
    a = torch.empty((F * (H // 2) * (W // 2), Cin * 4), dtype=torch.bfloat16, device=x.device)
    _call("vcof_patchify", x.data_ptr(), a.data_ptr(), Cin, F, H, W, _stream())
    return a


def unpatchify(y, Cout, F, H, W, out=None):
    """y bf16 [L, 4*Cout] (column order ph, pw, c) -> [Cout, F, H, W] (latent sizes)."""
    _chk(y, torch.bfloat16, "unpatchify.y", 2)
    if out is None:
        out = torch.empty((Cout, F, H, W), dtype=torch.bfloat16, device=y.device)
    _call("vcof_unpatchify", y.data_ptr(), y.stride(0), out.data_ptr(), Cout, F, H, W, _stream())
    return out


def linear_f32(x, w, bias=None, act_in=False, act_out=False):
    """fp32 act_out(act_in(x) @ w.T + bias) with bf16 weights; act = SiLU.  x fp32 [B,K]."""
    _chk(x, torch.float32, "linear_f32.x", 2)
    _chk(w, torch.bfloat16, "linear_f32.w", 2)
    if not x.is_contiguous() or not w.is_contiguous():
        raise _lib.VcofError("linear_f32: x and w must be contiguous")
    B, K = x.shape
    N = w.shape[0]
    out = torch.empty((B, N), dtype=torch.float32, device=x.device)
    _call("vcof_linear_f32", x.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), B, N, K,
          int(act_in), int(act_out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# VAE ops (channels-last bf16 activations [T, H, W, C])
# ------------------------------------------------------------------------------------------------
def _arr(ctype, vals):
    return (ctype * len(vals))(*vals)


def conv_igemm(x, x_dims, x_strides, w, taps, cin, geom, bias, out, residual=None, clamp=0.0, act_out=None,
               act_gamma=None, tgroup=1):
    """Raw binding of vcof_conv_igemm (see include/vcof.h).  x: any bf16 CUDA tensor whose storage the
    5-D view (x_dims / x_strides, elements) addresses from x.data_ptr(); w: packed [slices, n_total, kc] with
    kc = 32 or 64 channels per K slice (passed to the kernel as geom[16])."""
    _chk(x, torch.bfloat16, "conv.x")
    _chk(w, torch.bfloat16, "conv.w", 3)
    ref_out = out if out is not None else act_out
    _chk(ref_out, torch.bfloat16, "conv.out")
    if act_out is not None:
        _chk(act_out, torch.bfloat16, "conv.act_out")
        _chk(act_gamma, torch.float32, "conv.act_gamma", 1)
        if out is not None and act_out.stride(-2) != out.stride(-2):
            raise _lib.VcofError("conv: out and act_out must share the position pitch")
    if bias is not None:
        _chk(bias, torch.float32, "conv.bias", 1)
    if residual is not None:
        _chk(residual, torch.bfloat16, "conv.residual")
    ntaps = len(taps)
    flat = [int(v) for tp in taps for v in tp]
    ldc = ref_out.stride(-2)
    _call("vcof_conv_igemm", x.data_ptr(), _arr(_ct.c_longlong, [int(v) for v in x_dims]),
          _arr(_ct.c_longlong, [int(v) for v in x_strides]), w.data_ptr(), w.shape[0] * w.shape[2],
          _arr(_ct.c_short, flat), ntaps, int(tgroup), cin, _arr(_ct.c_int, [int(v) for v in geom] + [int(w.shape[2])]), _p(bias),
          _p(residual), _p(out), ldc, float(clamp), _p(act_out), _p(act_gamma), _stream(),
          key=f"conv taps={ntaps} cin={cin} n={geom[4]} T={geom[0]} H={geom[1]} W={geom[2]}"
              + ("+act" if act_out is not None else ""))
    return ref_out


def conv_lines(x, x_dims, x_strides, w, cin, kt, t0, geom, bias, out, residual=None, clamp=0.0, act_out=None,
               act_gamma=None):
    """Raw binding of vcof_conv_lines (line-resident 3x3(x3) convolution, include/vcof.h).
    w: packed [cin/32 * kt * 9, n_total, 32]; geom: T_out, H_out, W_out, n_total, n_tile, rows, n_store."""
    _chk(x, torch.bfloat16, "conv_lines.x")
    _chk(w, torch.bfloat16, "conv_lines.w", 3)
    ref_out = out if out is not None else act_out
    _chk(ref_out, torch.bfloat16, "conv_lines.out")
    if act_out is not None:
        _chk(act_out, torch.bfloat16, "conv_lines.act_out")
        _chk(act_gamma, torch.float32, "conv_lines.act_gamma", 1)
        if out is not None and act_out.stride(-2) != out.stride(-2):
            raise _lib.VcofError("conv_lines: out and act_out must share the position pitch")
    if bias is not None:
        _chk(bias, torch.float32, "conv_lines.bias", 1)
    if residual is not None:
        _chk(residual, torch.bfloat16, "conv_lines.residual")
    if w.shape[0] != (cin // 32) * kt * 9 or w.shape[2] != 32:
        raise _lib.VcofError(f"conv_lines: weight pack {tuple(w.shape)} does not match cin={cin}, kt={kt}")
    _call("vcof_conv_lines", x.data_ptr(), _arr(_ct.c_longlong, [int(v) for v in x_dims]),
          _arr(_ct.c_longlong, [int(v) for v in x_strides]), w.data_ptr(), cin, int(kt), int(t0),
          _arr(_ct.c_int, [int(v) for v in geom]), _p(bias), _p(residual), _p(out), ref_out.stride(-2), float(clamp),
          _p(act_out), _p(act_gamma), _stream(),
          key=f"conv_lines kt={kt} cin={cin} n={geom[3]} T={geom[0]} H={geom[1]} W={geom[2]}"
              + ("+act" if act_out is not None else ""))
    return ref_out


def rms_silu_cl(x, gamma, silu=True, out=None):
    """Channels-last RMS_norm (+SiLU).  x bf16 [..., C] contiguous rows; gamma fp32 [C]."""
    _chk(x, torch.bfloat16, "rms_silu.x")
    _chk(gamma, torch.float32, "rms_silu.gamma", 1)
    C = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    npos = x.numel() // C
    _call("vcof_rms_silu_cl", x.data_ptr(), x.stride(-2), gamma.data_ptr(), out.data_ptr(), out.stride(-2),
          npos, C, 1 if silu else 0, _stream())
    return out


def nchw_to_cl(x, Cp, div=None, add=None):
    """x bf16 [C, T, H, W] -> [T, H, W, Cp] (zero-padded channels); optional x / div[c] + add[c]."""
    _chk(x, torch.bfloat16, "nchw_to_cl.x", 4)
    if not x.is_contiguous():
        raise _lib.VcofError("nchw_to_cl.x must be contiguous")
    C, T, H, W = x.shape
    y = torch.empty((T, H, W, Cp), dtype=torch.bfloat16, device=x.device)
    _call("vcof_nchw_to_cl", x.data_ptr(), y.data_ptr(), C, Cp, T * H * W, _p(div), _p(add), _stream())
    return y


def cl_to_nchw(x, C, sub=None, mul=None):
    """x bf16 [T, H, W, ld>=C] -> [C, T, H, W]; optional (x - sub[c]) * mul[c]."""
    _chk(x, torch.bfloat16, "cl_to_nchw.x", 4)
    T, H, W, ld = x.shape
    y = torch.empty((C, T, H, W), dtype=torch.bfloat16, device=x.device)
    _call("vcof_cl_to_nchw", x.data_ptr(), x.stride(2), y.data_ptr(), C, T * H * W, _p(sub), _p(mul), _stream())
    return y


def u8_to_cl(frames, Cp):
    """Byte frames uint8 [T, H, W, C] -> channels-last bf16 [T, H, W, Cp] = bf16(fp32(u) * fp32(2/255) - 1), channels
    >= C zero: load_video_frames' scaling (reference fast_infer.py:86-88) + the cast to the VAE dtype
    (pipeline_wan.py:397) + the layout pass of the encoder's first layer, bit-exact."""
    _chk(frames, torch.uint8, "u8_to_cl.frames", 4)
    if not frames.is_contiguous():
        raise _lib.VcofError("u8_to_cl.frames must be contiguous")
    T, H, W, C = frames.shape
    if Cp < C:
        raise _lib.VcofError(f"u8_to_cl: Cp={Cp} < C={C}")
    y = torch.empty((T, H, W, Cp), dtype=torch.bfloat16, device=frames.device)
    _call("vcof_u8_to_cl", frames.data_ptr(), y.data_ptr(), T * H * W, C, Cp, _stream())
    return y


def cl_to_u8(x, C, out=None):
    """Decoder output bf16 [T, H, W, ld >= C] -> byte frames uint8 [T, H, W, C] =
    trunc(255 * clamp(bf16(bf16(x / 2) + 0.5), 0, 1)): decode_latents (pipeline_wan.py:425-427) + save_videos_grid's
    uint8 conversion (utils/utils.py:66), bit-exact."""
    _chk(x, torch.bfloat16, "cl_to_u8.x", 4)
    T, H, W, ld = x.shape
    dense = (H * W * ld, W * ld, ld, 1)
    if any(x.shape[i] > 1 and x.stride(i) != dense[i] for i in range(4)):
        raise _lib.VcofError("cl_to_u8.x must be a dense channels-last tensor")
    if out is None:
        out = torch.empty((T, H, W, C), dtype=torch.uint8, device=x.device)
    else:
        _chk(out, torch.uint8, "cl_to_u8.out", 4)
        if tuple(out.shape) != (T, H, W, C) or not out.is_contiguous():
            raise _lib.VcofError("cl_to_u8.out must be a contiguous uint8 [T, H, W, C]")
    _call("vcof_cl_to_u8", x.data_ptr(), ld, out.data_ptr(), T * H * W, C, _stream())
    return out


def vae_attn(qkv, C, scale=None, out=None):
    """Fused single-head attention of the VAE AttentionBlock: qkv bf16 [T, N, 3C] (q | k | v along the last dim, the
    to_qkv output) -> bf16 [T, N, C]; one launch for all T frames."""
    _chk(qkv, torch.bfloat16, "vae_attn.qkv", 3)
    T, N, C3 = qkv.shape
    if C3 < 3 * C or qkv.stride(0) != N * qkv.stride(1):
        raise _lib.VcofError(f"vae_attn: qkv {tuple(qkv.shape)} / strides {qkv.stride()} is not a dense stack of [N, 3C] frames")
    if out is None:
        out = torch.empty((T, N, C), dtype=torch.bfloat16, device=qkv.device)
    _chk(out, torch.bfloat16, "vae_attn.out", 3)
    if tuple(out.shape) != (T, N, C) or out.stride(0) != N * out.stride(1):
        raise _lib.VcofError(f"vae_attn: out {tuple(out.shape)} / strides {out.stride()} does not match [T, N, C]")
    scale = 1.0 / math.sqrt(C) if scale is None else scale
    _call("vcof_vae_attn", qkv.data_ptr(), qkv.stride(1), out.data_ptr(), out.stride(1), T, N, C, float(scale), _stream(),
          key=f"vae_attn T={T} N={N}")
    return out


def softmax_rows(s, scale, out=None):
    """bf16 softmax(s * scale) over the last dim of fp32 s [rows, n]."""
    _chk(s, torch.float32, "softmax.s", 2)
    rows, n = s.shape
    if out is None:
        out = torch.empty((rows, (n + 7) // 8 * 8), dtype=torch.bfloat16, device=s.device)[:, :n]
    _call("vcof_softmax_rows", s.data_ptr(), s.stride(0), out.data_ptr(), out.stride(0), rows, n, float(scale),
          _stream())
    return out


# ------------------------------------------------------------------------------------------------
# umT5 text encoder ops (bf16 activations [tokens, C])
# ------------------------------------------------------------------------------------------------
def embed_rows(ids, table, out=None):
    """table[ids] for int64 ids [n] on the device; table bf16 [vocab, C]."""
    _chk(ids, torch.int64, "embed_rows.ids", 1)
    _chk(table, torch.bfloat16, "embed_rows.table", 2)
    if not ids.is_contiguous():
        raise _lib.VcofError("embed_rows.ids must be contiguous")
    n, (vocab, C) = ids.numel(), table.shape
    if out is None:
        out = torch.empty((n, C), dtype=torch.bfloat16, device=table.device)
    _chk(out, torch.bfloat16, "embed_rows.out", 2)
    _call("vcof_embed_rows", ids.data_ptr(), table.data_ptr(), table.stride(0), vocab, out.data_ptr(), out.stride(0),
          n, C, _stream())
    return out


def t5_rmsnorm(x, weight, eps=1e-6, out=None):
    """T5LayerNorm: bf16(w * bf16(x * rsqrt(mean(x^2) + eps))); x bf16 [rows, C], weight bf16 [C]."""
    _chk(x, torch.bfloat16, "t5_rmsnorm.x", 2)
    _chk(weight, torch.bfloat16, "t5_rmsnorm.weight", 1)
    rows, C = x.shape
    if weight.numel() != C:
        raise _lib.VcofError(f"t5_rmsnorm: weight has {weight.numel()} entries for C={C}")
    if out is None:
        out = torch.empty((rows, C), dtype=torch.bfloat16, device=x.device)
    _chk(out, torch.bfloat16, "t5_rmsnorm.out", 2)
    _call("vcof_t5_rmsnorm", x.data_ptr(), x.stride(0), weight.data_ptr(), out.data_ptr(), out.stride(0), rows, C,
          float(eps), _stream())
    return out


def t5_attention(q, k, v, bias_rel, B, L, heads, key_mask=None, out=None):
    """softmax(q k^T + bias) v per head, unscaled.  q/k/v bf16 [B*L, heads*d]; bias_rel fp32 [heads, 2L-1] indexed by
    (key - query) + L - 1; key_mask int32 [B, L] (0 = masked) or None."""
    for n, t in (("q", q), ("k", k), ("v", v)):
        _chk(t, torch.bfloat16, "t5_attention." + n, 2)
        if t.shape[0] != B * L:
            raise _lib.VcofError(f"t5_attention.{n}: {t.shape[0]} rows for B={B}, L={L}")
    _chk(bias_rel, torch.float32, "t5_attention.bias_rel", 2)
    if bias_rel.shape[0] != heads or bias_rel.shape[1] < 2 * L - 1:
        raise _lib.VcofError(f"t5_attention.bias_rel: shape {tuple(bias_rel.shape)} for heads={heads}, L={L}")
    if key_mask is not None:
        _chk(key_mask, torch.int32, "t5_attention.key_mask", 2)
        if tuple(key_mask.shape) != (B, L) or not key_mask.is_contiguous():
            raise _lib.VcofError(f"t5_attention.key_mask: need a contiguous [{B}, {L}] tensor")
    C = q.shape[1]
    if out is None:
        out = torch.empty((B * L, C), dtype=torch.bfloat16, device=q.device)
    _chk(out, torch.bfloat16, "t5_attention.out", 2)
    _call("vcof_t5_attn", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
          out.data_ptr(), out.stride(0), bias_rel.data_ptr(), bias_rel.stride(0), _p(key_mask), B, L, heads,
          C // heads, _stream(), key=f"t5_attn B={B} L={L} heads={heads}")
    return out
