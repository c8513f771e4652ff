#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_vae_gpu.py tests/test_widen_x_full_size_gpu.py tests/test_pipeline_gpu.py tests/test_widen_video_io_gpu.py -q -x 2>&1 | tail -3
for sp in 1 0 2; do echo "spare=$sp"; VCOF_CONV_SPARE=$sp timeout 200 python tools/vae_bench.py --frames 9 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('enc', round(d['enc']['ms'],2), 'dec', round(d['dec']['ms'],2)); print(d['dec']['top'][:4])"; done
VCOF_CONV_SPARE=0 timeout 200 python -m pytest tests/test_vae_gpu.py -q -x 2>&1 | tail -1
VCOF_CONV_SPARE=2 timeout 200 python -m pytest tests/test_vae_gpu.py -q -x 2>&1 | tail -1
