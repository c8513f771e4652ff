"""`-m gpu` test files, dry-run on the CPU (`pytest --gpu-dryrun`, tests/abi_emulator.py): the unchanged test functions
run with "cuda" redirected to the CPU, libvcof's real argument validation in front of every call and the contract
statements of the C ABI behind it.  Catches what does not need a GPU to be wrong — a test's own shape or index slip,
a marshalling slip in videocof_b200/ops.py, an argument the library rejects — before a GPU call is spent on it.
It says nothing about the kernels: those are the `-m gpu` runs proper.

Deselected: tests that need a real "this tensor lives on the CPU" answer (`cpu` in their name: under the dry run every
tensor claims to be a CUDA tensor), the tcgen05 descriptor probe (no contract to state), the full-size halves (75 600
tokens / 720p: GPU only).  tests/test_lora_gpu.py is not dry-run (merge_lora asks the device object for its type)."""
import os
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

FILES = {
    # file: (-k expression, minimum number of tests that must have run)
    # the two whole-VAE / whole-pipeline byte tests (85 s of emulated convolutions) are left to the GPU run: the file
    # has been green on hardware since round 2; `pytest --gpu-dryrun tests/test_widen_video_io_gpu.py` dry-runs them all
    "test_widen_video_io_gpu.py": ("not cpu and not full_size and not pipeline_bytes and not vae_byte_frames", 12),
    "test_widen_w_push_exchange_gpu.py": ("not cpu", 10),                 # idem
    "test_widen_x_full_size_gpu.py": ("c1 or small", 12),                 # idem; the c2 / 720p half needs the GPU
    "test_dit_gpu.py": ("not cpu", 10),
    "test_kernels_gpu.py": ("not umma and not cpu", 60),
    "test_vae_gpu.py": ("not cpu", 27),
    "test_pipeline_gpu.py": ("not cpu", 1),
    "test_widen_text_encoder_gpu.py": ("not cpu", 23),
}


@pytest.fixture(scope="module")
def runs():
    """All files at once, one pytest process each (the dry run patches torch process-wide)."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    procs = {name: subprocess.Popen([sys.executable, "-m", "pytest", os.path.join(HERE, name), "--gpu-dryrun", "-q", "-x",
                                     "-k", expr, "-p", "no:cacheprovider"], stdout=subprocess.PIPE,
                                    stderr=subprocess.STDOUT, text=True, cwd=os.path.dirname(HERE), env=env)
             for name, (expr, _) in FILES.items()}
    yield procs
    for p in procs.values():
        if p.poll() is None:
            p.kill()


@pytest.mark.skipif(torch.cuda.is_available(), reason="a GPU is present: run the tests for real")
@pytest.mark.parametrize("name", list(FILES))
def test_gpu_file_dry_runs(runs, name):
    out, _ = runs[name].communicate(timeout=1500)
    tail = out[-3000:]
    assert runs[name].returncode == 0, tail
    passed = int(out.rsplit(" passed", 1)[0].rsplit(None, 1)[-1])
    assert passed >= FILES[name][1], tail


def test_statements_follow_the_header():
    """Each contract statement takes the parameters of its entry point, by name and in the header's order (so the
    statements are written against include/vcof.h, not against ops.py), and matches the ctypes binding's arity."""
    import inspect
    import re

    import abi_emulator
    from videocof_b200 import _lib
    text = open(os.path.join(os.path.dirname(HERE), "include", "vcof.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {m.group(1): m.group(2) for m in re.finditer(r"\bint\s+(vcof_\w+)\s*\(([^)]*)\)\s*;", text)}
    without = {"vcof_abi_version"}
    for name, params in decls.items():
        if name in without or name.startswith("vcof_debug"):
            continue
        assert name in abi_emulator.STATEMENTS, f"no contract statement for {name}"
        want = [re.sub(r"\[.*\]", "", p.strip()).split()[-1].lstrip("*") for p in params.split(",")]
        got = list(inspect.signature(abi_emulator.STATEMENTS[name]).parameters)
        assert got == want, (name, got, want)
        assert len(_lib.SIGNATURES[name]) == len(want), name
    assert set(abi_emulator.STATEMENTS) <= set(decls)
