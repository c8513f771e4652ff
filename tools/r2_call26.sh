#!/bin/bash
# One-GPU call: context-cache tests + ncu --set full of the fused VAE attention and of the line conv kernel (with the
# accumulator ring).      gpurun --timeout 900 -- 'bash tools/r2_call26.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dit_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2c26_pytest.log
N="ncu --set full --clock-control none --import-source on"
cat > /tmp/vae_attn_once.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from videocof_b200 import ops
qkv = torch.randn(3, 14400, 1152, device="cuda").bfloat16()
for _ in range(3):
    ops.vae_attn(qkv, 384)
torch.cuda.synchronize()
PY
timeout 200 $N -k regex:vae_attn -s 2 -c 1 -o gpurun_out/r2c_vae_attn python /tmp/vae_attn_once.py > gpurun_out/r2c26_ncu_vae_attn.log 2>&1
timeout 300 $N -k regex:conv_lines -s 1 -c 7 -o gpurun_out/r2c_conv_lines python tools/vae_bench.py --frames 9 > gpurun_out/r2c26_ncu_conv.log 2>&1
ls -la gpurun_out/r2c_* gpurun_out/r2c26_*
tail -n 2 gpurun_out/r2c26_ncu_vae_attn.log gpurun_out/r2c26_ncu_conv.log
