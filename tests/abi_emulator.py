"""Dry run of `-m gpu` test files on a CPU-only box (test infrastructure only; never imported by the product).

The GPU box is reached a few times per round, so a GPU test that fails on a typo — a wrong argument order in a
ctypes call, a shape slip in the test itself — wastes a whole call.  `pytest --gpu-dryrun` (tests/conftest.py -> install() below) lets the *unchanged* test functions
run here:

  * tensors asked onto "cuda" stay on the CPU (a TorchFunctionMode rewrites the device argument; `.cuda()` is the
    identity) and `videocof_b200.ops` accepts them;
  * `videocof_b200._lib.call` first calls the REAL libvcof.so, whose argument validation (include/vcof.h) runs before
    anything touches a device: a validation failure (-1) is raised exactly as on the GPU, a CUDA failure (-2, no
    device here) hands the call to the executable statement of the entry point's contract below;
  * the statements decode the raw C arguments (pointers, pitches, sizes) in the order include/vcof.h declares them —
    written from the header, not from ops.py, so a marshalling slip on either side shows up as a wrong result —
    and do the arithmetic with tests/vcof_emulator.py.

Every entry point of include/vcof.h has a statement here (the debug probes excepted).

What a dry run proves: the test's own logic, ops.py's marshalling, the library's argument checks.  What it cannot:
the kernels.  tests/test_gpu_dryrun_cpu.py lists the test files that are dry-run.
"""
import ctypes
import types

import torch
from torch.overrides import TorchFunctionMode

import vcof_emulator as emu

_ITEM = {torch.bfloat16: 2, torch.float32: 4, torch.int32: 4, torch.uint8: 1, torch.int64: 8}


def _flat(ptr, n, dtype):
    """n elements of `dtype` at the host address `ptr`, sharing memory with whoever owns it."""
    if n == 0:
        return torch.empty(0, dtype=dtype)
    buf = (ctypes.c_char * (n * _ITEM[dtype])).from_address(ptr)
    return torch.frombuffer(buf, dtype=dtype, count=n)


def _mat(ptr, rows, cols, ld, dtype=torch.bfloat16):
    """[rows, cols] matrix with row pitch ld (elements) at ptr."""
    return _flat(ptr, (rows - 1) * ld + cols, dtype).as_strided((rows, cols), (ld, 1))


def _ptr_array(arr, n):
    """A `void* const*` argument as ctypes handed it over (an array instance or an address)."""
    if isinstance(arr, ctypes.Array):
        return [int(arr[i]) for i in range(n)]
    return [int(v) for v in (ctypes.c_void_p * n).from_address(int(arr))]


def _rope(table, tpos, F, H, W, n_t, n_h, row_offset, head_dim):
    if not table:
        return None
    return types.SimpleNamespace(table=_flat(table, 1024 * (head_dim // 2) * 2, torch.float32).view(1024, head_dim // 2, 2),
                                 tpos=_flat(tpos, F, torch.int32), F=F, H=H, W=W, n_t=n_t, n_h=n_h,
                                 row_offset=row_offset)


# ---- one statement per entry point; parameter names and order are those of include/vcof.h ---------------------------
def vcof_attn_fwd(q, ldq, k, ldk, v, ldv, out, ldo, Lq, Lk, kv_len, heads, head_dim, softmax_scale, v_transposed, stream):
    C = heads * head_dim
    vt = _mat(v, C, Lk, ldv) if v_transposed else _mat(v, Lk, C, ldv)
    emu.attention(_mat(q, Lq, C, ldq), _mat(k, Lk, C, ldk), vt, heads, kv_len=kv_len, scale=softmax_scale,
                  out=_mat(out, Lq, C, ldo), v_transposed=bool(v_transposed))


def vcof_attn_fwd_scatter(q, ldq, k, ldk, v, ldv, out_chunks, n_chunks, rows_per_chunk, ldo, Lq, Lk, kv_len, heads,
                          head_dim, softmax_scale, stream):
    C = heads * head_dim
    o = emu.attention(_mat(q, Lq, C, ldq), _mat(k, Lk, C, ldk), _mat(v, Lk, C, ldv), heads, kv_len=kv_len,
                      scale=softmax_scale)
    for c, p in enumerate(_ptr_array(out_chunks, n_chunks)):
        r0, r1 = c * rows_per_chunk, min(Lq, (c + 1) * rows_per_chunk)
        if r1 > r0:
            _mat(p, r1 - r0, C, ldo).copy_(o[r0:r1])


def vcof_rmsnorm_rope(x, ldx, weight, eps, L, C, head_dim, rope_table, tpos, F, H, W, n_t, n_h, row_offset, stream):
    xm = _mat(x, L, C, ldx)
    xm.copy_(emu._rmsnorm_rope(xm, _flat(weight, C, torch.bfloat16), eps, head_dim,
                               _rope(rope_table, tpos, F, H, W, n_t, n_h, row_offset, head_dim)))


def vcof_rmsnorm_rope_blocked(x, ldx, y, cols_per_block, block_stride, weight, eps, L, C, head_dim, rope_table, tpos, F, H,
                              W, n_t, n_h, row_offset, stream):
    r = emu._rmsnorm_rope(_mat(x, L, C, ldx), _flat(weight, C, torch.bfloat16), eps, head_dim,
                          _rope(rope_table, tpos, F, H, W, n_t, n_h, row_offset, head_dim))
    for b in range(C // cols_per_block):
        _mat(y + 2 * b * block_stride, L, cols_per_block, cols_per_block).copy_(
            r[:, b * cols_per_block:(b + 1) * cols_per_block])


def vcof_rmsnorm_rope_scatter(x, ldx, block_ptrs, n_blocks, weight, eps, L, C, head_dim, rope_table, tpos, F, H, W, n_t,
                              n_h, row_offset, stream):
    r = emu._rmsnorm_rope(_mat(x, L, C, ldx), _flat(weight, C, torch.bfloat16), eps, head_dim,
                          _rope(rope_table, tpos, F, H, W, n_t, n_h, row_offset, head_dim))
    cp = C // n_blocks
    for b, p in enumerate(_ptr_array(block_ptrs, n_blocks)):
        _mat(p, L, cp, cp).copy_(r[:, b * cp:(b + 1) * cp])


def vcof_copy_blocked(rowmajor, ld, blocked, block_stride, rows, C, cols_per_block, to_blocked, stream):
    rm = _mat(rowmajor, rows, C, ld)
    for b in range(C // cols_per_block):
        blk = _mat(blocked + 2 * b * block_stride, rows, cols_per_block, cols_per_block)
        cols = rm[:, b * cols_per_block:(b + 1) * cols_per_block]
        blk.copy_(cols) if to_blocked else cols.copy_(blk)


def vcof_copy_scatter(rowmajor, ld, block_ptrs, n_blocks, rows, C, stream):
    rm = _mat(rowmajor, rows, C, ld)
    cp = C // n_blocks
    for b, p in enumerate(_ptr_array(block_ptrs, n_blocks)):
        _mat(p, rows, cp, cp).copy_(rm[:, b * cp:(b + 1) * cp])


def vcof_copy_rows_scatter(src, ld, chunk_ptrs, n_chunks, rows, cols, stream):
    s = _mat(src, n_chunks * rows, cols, ld)
    for c, p in enumerate(_ptr_array(chunk_ptrs, n_chunks)):
        _mat(p, rows, cols, cols).copy_(s[c * rows:(c + 1) * rows])


def vcof_cl_to_u8(x, ldx, out, npos, C, stream):
    res = emu.cl_to_u8(_mat(x, npos, C, ldx).view(1, 1, npos, C), C)          # the library's host evaluation
    _flat(out, npos * C, torch.uint8).copy_(res.reshape(-1))


def vcof_u8_to_cl(frames, y, npos, C, Cp, stream):
    res = emu.u8_to_cl(_flat(frames, npos * C, torch.uint8).view(1, 1, npos, C), Cp)
    _flat(y, npos * Cp, torch.bfloat16).copy_(res.reshape(-1))


def vcof_gemm_bf16(a, lda, w, ldw, bias, gate, out, ldo, M, N, K, epilogue, stream):
    epi = {v: k for k, v in _EPI.items()}[epilogue & 0xff]            # 0x100 = VCOF_GEMM_TILE128: same result
    f32_out = epi in ("bias_gate_res", "bias_f32", "raw_f32")
    emu.gemm(_mat(a, M, K, lda), _mat(w, N, K, ldw), _flat(bias, N, torch.bfloat16) if bias else None, epi,
             out=_mat(out, M, N, ldo, torch.float32 if f32_out else torch.bfloat16),
             gate=_flat(gate, N, torch.float32) if gate else None)


def vcof_ln_modulate(x, ldx, ln_w, ln_b, shift, scale, out, ldo, L, C, eps, stream):
    vec = lambda p: _flat(p, C, torch.float32) if p else None
    emu.ln_modulate(_mat(x, L, C, ldx, torch.float32), vec(ln_w), vec(ln_b), vec(shift), vec(scale), eps,
                    out=_mat(out, L, C, ldo))


def vcof_patchify(x, a, Cin, F, H, W, stream):
    src = _flat(x, Cin * F * H * W, torch.bfloat16).view(Cin, F, H, W)
    _flat(a, Cin * F * H * W, torch.bfloat16).view(-1, Cin * 4).copy_(emu.patchify(src))


def vcof_unpatchify(y, ldy, out, Cout, F, H, W, stream):
    L = F * (H // 2) * (W // 2)
    emu.unpatchify(_mat(y, L, 4 * Cout, ldy), Cout, F, H, W,
                   out=_flat(out, Cout * F * H * W, torch.bfloat16).view(Cout, F, H, W))


def vcof_linear_f32(x, w, bias, out, B, N, K, act_in, act_out, stream):
    res = emu.linear_f32(_mat(x, B, K, K, torch.float32), _mat(w, N, K, K), _flat(bias, N, torch.bfloat16) if bias else None,
                         act_in=bool(act_in), act_out=bool(act_out))
    _mat(out, B, N, N, torch.float32).copy_(res)


def vcof_softmax_rows(s, lds, p, ldp, rows, n, scale, stream):
    _mat(p, rows, n, ldp).copy_(emu.softmax_rows(_mat(s, rows, n, lds, torch.float32), scale))


def vcof_vae_attn(qkv, ld, out, ldo, T, N, C, softmax_scale, stream):
    x = _flat(qkv, (T * N - 1) * ld + 3 * C, torch.bfloat16).as_strided((T, N, 3 * C), (N * ld, ld, 1))
    y = _flat(out, (T * N - 1) * ldo + C, torch.bfloat16).as_strided((T, N, C), (N * ldo, ldo, 1))
    emu.vae_attn(x, C, softmax_scale, out=y)


def vcof_rms_silu_cl(x, ldx, gamma, y, ldy, npos, C, silu, stream):
    emu.rms_silu_cl(_mat(x, npos, C, ldx), _flat(gamma, C, torch.float32), bool(silu), out=_mat(y, npos, C, ldy))


def vcof_nchw_to_cl(x, y, C, Cp, thw, div, add, stream):
    src = _flat(x, C * thw, torch.bfloat16).view(C, 1, 1, thw)
    vec = lambda p: _flat(p, C, torch.float32) if p else None
    _flat(y, thw * Cp, torch.bfloat16).view(1, 1, thw, Cp).copy_(emu.nchw_to_cl(src, Cp, vec(div), vec(add)))


def vcof_cl_to_nchw(x, ldx, y, C, thw, sub, mul, stream):
    vec = lambda p: _flat(p, C, torch.float32) if p else None
    res = emu.cl_to_nchw(_mat(x, thw, ldx, ldx).view(1, 1, thw, ldx), C, vec(sub), vec(mul))
    _flat(y, C * thw, torch.bfloat16).view(C, 1, 1, thw).copy_(res)


def vcof_embed_rows(ids, table, ldt, vocab, out, ldo, n, C, stream):
    emu.embed_rows(_flat(ids, n, torch.int64), _mat(table, vocab, C, ldt), out=_mat(out, n, C, ldo))


def vcof_t5_rmsnorm(x, ldx, weight, y, ldy, rows, C, eps, stream):
    emu.t5_rmsnorm(_mat(x, rows, C, ldx), _flat(weight, C, torch.bfloat16), eps, out=_mat(y, rows, C, ldy))


def vcof_t5_attn(q, ldq, k, ldk, v, ldv, out, ldo, bias_rel, bias_ld, key_mask, B, L, heads, head_dim, stream):
    C = heads * head_dim
    bias = _mat(bias_rel, heads, 2 * L - 1, bias_ld, torch.float32)
    emu.t5_attention(_mat(q, B * L, C, ldq), _mat(k, B * L, C, ldk), _mat(v, B * L, C, ldv), bias, B, L, heads,
                     key_mask=_flat(key_mask, B * L, torch.int32) if key_mask else None, out=_mat(out, B * L, C, ldo))


def _ints(ptr, n, ctype):
    if isinstance(ptr, ctypes.Array):
        return [int(ptr[i]) for i in range(n)]
    return [int(v) for v in (ctype * n).from_address(int(ptr))]


def _conv_common(x, x_dims, x_strides, geom_out, ldc, bias, residual, out, act_out, act_gamma, n_total, n_store):
    """Tensors of a convolution call: the input storage the 5-D view addresses, and the output-addressed buffers."""
    dims, strides = _ints(x_dims, 5, ctypes.c_longlong), _ints(x_strides, 4, ctypes.c_longlong)
    extent = dims[0] + sum((d - 1) * st for d, st in zip(dims[1:], strides))
    xs = _flat(x, extent, torch.bfloat16)
    Tt, Hs, Ws = geom_out
    buf = lambda p: _flat(p, Tt * Hs * Ws * ldc, torch.bfloat16).view(Tt, Hs, Ws, ldc) if p else None
    return (xs, dims, strides, _flat(bias, n_total, torch.float32) if bias else None, buf(residual), buf(out), buf(act_out),
            _flat(act_gamma, n_store, torch.float32) if act_gamma else None)


def vcof_conv_igemm(x, x_dims, x_strides, w, k_total, taps, ntaps, tgroup, cin, geom, bias, residual, out, ldc, clamp,
                    act_out, act_gamma, stream):
    g = _ints(geom, 17, ctypes.c_int)
    T, H, W, t_stride, n_total, n_tile, ot_mul, ot_add, oh_mul, oh_add, ow_mul, ow_add, Hs, Ws, half, n_store, kc = g
    Tt = (T - 1) * ot_mul + ot_add + 1 + (1 if half > 0 else 0)
    xs, dims, strides, b, res, o, act, gamma = _conv_common(x, x_dims, x_strides, (Tt, Hs, Ws), ldc, bias, residual, out,
                                                            act_out, act_gamma, n_total, n_store)
    flat = _ints(taps, 5 * ntaps, ctypes.c_short)
    tap_list = [tuple(flat[5 * i:5 * i + 5]) for i in range(ntaps)]
    wt = _flat(w, (k_total // kc) * n_total * kc, torch.bfloat16).view(k_total // kc, n_total, kc)
    emu.conv_igemm(xs, dims, strides, wt, tap_list, cin, g[:16], b, o, residual=res, clamp=clamp, act_out=act,
                   act_gamma=gamma, tgroup=tgroup)


def vcof_conv_lines(x, x_dims, x_strides, w, cin, kt, t0, geom, bias, residual, out, ldc, clamp, act_out, act_gamma,
                    stream):
    g = _ints(geom, 7, ctypes.c_int)
    T, H, W, n_total, n_tile, rows, n_store = g
    xs, dims, strides, b, res, o, act, gamma = _conv_common(x, x_dims, x_strides, (T, H, W), ldc, bias, residual, out,
                                                            act_out, act_gamma, n_total, n_store)
    slices = (cin // 32) * kt * 9
    wt = _flat(w, slices * n_total * 32, torch.bfloat16).view(slices, n_total, 32)
    emu.conv_lines(xs, dims, strides, wt, cin, kt, t0, g, b, o, residual=res, clamp=clamp, act_out=act, act_gamma=gamma)


_EPI = {"bias": 0, "bias_gelu": 1, "bias_gate_res": 2, "bias_f32": 3, "raw_f32": 4, "gate_accum": 5, "mul": 6, "add": 7}


STATEMENTS = {f.__name__: f for f in (vcof_attn_fwd, vcof_attn_fwd_scatter, vcof_rmsnorm_rope, vcof_rmsnorm_rope_blocked,
                                      vcof_rmsnorm_rope_scatter, vcof_copy_blocked, vcof_copy_scatter,
                                      vcof_copy_rows_scatter, vcof_cl_to_u8, vcof_u8_to_cl, vcof_gemm_bf16,
                                      vcof_ln_modulate, vcof_patchify, vcof_unpatchify, vcof_linear_f32, vcof_softmax_rows, vcof_vae_attn,
                                      vcof_rms_silu_cl, vcof_nchw_to_cl, vcof_cl_to_nchw, vcof_embed_rows, vcof_t5_rmsnorm,
                                      vcof_t5_attn, vcof_conv_igemm, vcof_conv_lines)}


def _is_cuda(d):
    if isinstance(d, torch.device):
        return d.type == "cuda"
    return isinstance(d, str) and d.split(":")[0] == "cuda"


class _CudaIsCpu(TorchFunctionMode):
    def __torch_function__(self, func, types_, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if getattr(func, "__name__", "") == "cuda" and args and isinstance(args[0], torch.Tensor):
            return args[0]
        if _is_cuda(kwargs.get("device")):
            kwargs["device"] = "cpu"
        args = tuple("cpu" if _is_cuda(a) else a for a in args)
        return func(*args, **kwargs)


# ops.py functions replaced wholesale by tests/vcof_emulator.py instead of going through their wrapper: none any more
OPS_LEVEL = ()


def install(monkeypatch):
    """From here on "cuda" means the CPU and libvcof calls go real-validation-then-statement (module docstring).  Meant
    for a whole pytest process (`pytest --gpu-dryrun`, tests/conftest.py): module-scoped fixtures move models to
    "cuda" too.  Returns the object whose __exit__ undoes the device redirection."""
    from videocof_b200 import _lib, ops

    real_call = _lib.call

    def call(name, *args):
        if name.endswith("_host"):          # host-side debug entries (vcof_debug_*_host) run for real
            return real_call(name, *args)
        lib = _lib.load()
        rc = getattr(lib, name)(*args)
        msg = lib.vcof_last_error().decode()
        if rc == -1 and "entry point unavailable" not in msg:
            raise _lib.VcofError(f"{name} failed ({rc}): {msg}")
        assert rc != 0, f"{name} succeeded without a GPU?"
        if name not in STATEMENTS:
            raise NotImplementedError(f"no pointer-level statement for {name}")
        STATEMENTS[name](*[0 if a is None else a for a in args])

    monkeypatch.setattr(_lib, "call", call)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    for name in OPS_LEVEL:
        monkeypatch.setattr(ops, name, getattr(emu, name))
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    real_gen = torch.Generator
    monkeypatch.setattr(torch, "Generator", lambda device="cpu": real_gen("cpu"))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda device=None: types.SimpleNamespace(
        multi_processor_count=148, name="dry run (CPU)", total_memory=180 << 30, major=10, minor=0))
    mode = _CudaIsCpu()
    mode.__enter__()
    return mode
