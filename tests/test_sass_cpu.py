"""Static properties of the built kernels (tools/sass_report.py: `cuobjdump -res-usage` / `-sass` of libvcof.so, no GPU):
the hot kernels take the hardware path DESIGN.md §3 says they take — tcgen05.mma (UTCHMMA) fed by TMA (UTMALDG) with
accumulators read back from TMEM (LDTM) — and none of the default kernels spills registers."""
import pytest

import sass_report


@pytest.fixture(scope="module")
def ks():
    import __graft_entry__ as g
    g.build()
    return sass_report.kernels()


def test_every_default_kernel_is_spill_free(ks):
    # opt-in experiments (VCOF_ATTN_SPEC / SPLIT_S) and the debug probes may use a few stack slots
    optin = lambda n: n.startswith(("tma_probe", "umma_probe")) or (
        n.startswith("attn_fwd_kernel") and n not in DEFAULT_ATTN and not n.endswith(", 0, 1>"))
    for name, k in ks.items():
        if optin(name):
            continue
        assert k["STACK"] <= 8 and k["LDL"] + k["STL"] <= 2, (name, k)     # attention: one 8-byte slot, one LDL / STL


DEFAULT_ATTN = ("attn_fwd_kernel<0, 0, 0, 2, 0, 0>", "attn_fwd_kernel<1, 0, 0, 2, 0, 0>")


@pytest.mark.parametrize("name", DEFAULT_ATTN + ("attn_fwd_kernel<0, 0, 0, 2, 0, 1>",))
def test_attention_runs_on_tcgen05_tma_tmem(ks, name):
    k = ks[name]
    assert k["UTCHMMA"] > 0 and k["UTMALDG"] > 0 and k["LDTM"] > 0 and k["STTM"] > 0 and k["MUFU.EX2"] > 0, k
    assert k["HMMA"] == 0 and k["REG"] <= 168, k        # 3 x 128 threads x 168 registers = one CTA per SM's file


def test_gemm_and_conv_run_on_tcgen05_tma(ks):
    names = [n for n in ks if n.startswith(("gemm_bf16_kernel<", "gemm2cta_bf16_kernel<", "conv_igemm_kernel",
                                            "conv_lines_kernel"))]
    assert len(names) >= 3 * 8 + 3 + 2
    for n in names:
        k = ks[n]
        assert k["UTCHMMA"] > 0 and k["UTMALDG"] > 0 and k["LDTM"] > 0 and k["HMMA"] == 0 and k["STACK"] == 0, (n, k)


def test_text_encoder_attention_is_warp_level_mma(ks):
    for d in (16, 32, 64, 128):
        assert ks[f"t5_attn_mma_kernel<{d}>"]["HMMA"] > 0
        assert ks[f"t5_attn_kernel<{d}>"]["HMMA"] == 0       # the CUDA-core cross-check kernel


def test_scatter_variants_share_the_default_arithmetic(ks):
    """The push-exchange instantiations differ from the validated kernels in their store addresses only: same register
    budget, same tensor / TMA / exponential instruction counts."""
    a, b = ks["attn_fwd_kernel<0, 0, 0, 2, 0, 0>"], ks["attn_fwd_kernel<0, 0, 0, 2, 0, 1>"]
    for key in ("REG", "STACK", "UTCHMMA", "UTMALDG", "LDTM", "STTM", "MUFU.EX2", "SYNCS"):
        assert a[key] == b[key], key
    # the scatter variant of the norm / RoPE kernel carries a table of destination pointers: a few registers more, no stack
    # (both variants are validated bit-exactly against each other on hardware: tests/test_widen_w_push_exchange_gpu.py)
    r0, r1 = ks["rmsnorm_rope_kernel<0>"], ks["rmsnorm_rope_kernel<1>"]
    assert abs(r1["REG"] - r0["REG"]) <= 16 and r0["STACK"] == 0 and r1["STACK"] == 0, (r0, r1)
