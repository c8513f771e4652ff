"""Multi-GPU parity (needs >= 2 CUDA devices; skipped on a 1-GPU box): the sequence-parallel DiT and the
frame-sharded VAE must reproduce the single-GPU results exactly (tools/sp_check.py, tools/vae_shard_check.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(script, n):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(_port()), os.path.join(ROOT, "tools", script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines, p.stdout[-800:] + p.stderr[-1500:]
    return json.loads(lines[-1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sequence_parallel_dit_matches_single_gpu():
    r = _torchrun("sp_check.py", 2)
    assert r["sp_check"] == "ok", r


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_frame_sharded_vae_matches_single_gpu():
    r = _torchrun("vae_shard_check.py", 2)
    assert r["vae_shard_check"] == "ok", r
