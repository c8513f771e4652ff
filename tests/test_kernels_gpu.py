"""Per-kernel parity of libvcof (through the C ABI) against plain torch fp32 math on the GPU.
Cases and tolerances live in tools/gpu_probe.py (shared with the stand-alone probe)."""
import json

import pytest

import gpu_probe

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", gpu_probe.cases(), ids=lambda c: json.dumps(c, separators=(",", ":")))
def test_kernel_case(case):
    res = gpu_probe.run_case(case)
    assert res.get("ok"), res
