import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


_DRYRUN = []


def pytest_addoption(parser):
    parser.addoption("--gpu-dryrun", action="store_true",
                     help="run `-m gpu` tests on a CPU-only box against the contract statements of tests/abi_emulator.py "
                          "(checks the tests' own logic, ops.py's marshalling and libvcof's argument validation; "
                          "not the kernels)")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu`")
    if config.getoption("--gpu-dryrun"):
        import torch
        if torch.cuda.is_available():
            raise pytest.UsageError("--gpu-dryrun is for CPU-only boxes; with a GPU run the tests for real")
        import abi_emulator
        mp = pytest.MonkeyPatch()
        _DRYRUN.extend([abi_emulator.install(mp), mp])


def pytest_unconfigure(config):
    if _DRYRUN:
        mode, mp = _DRYRUN
        mode.__exit__(None, None, None)
        mp.undo()
        _DRYRUN.clear()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
