#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_widen_v_vae_attn_gpu.py -q -x 2>&1 | tail -12
timeout 300 python -m pytest tests/test_vae_gpu.py tests/test_kernels_gpu.py tests/test_pipeline_gpu.py -q -x 2>&1 | tail -3
timeout 200 python tools/vae_bench.py --frames 9 2>&1 | tail -1 | cut -c1-1500
VCOF_VAE_ATTN=unfused timeout 200 python tools/vae_bench.py --frames 9 2>&1 | tail -1 | cut -c1-700
