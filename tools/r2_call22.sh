#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_vae_gpu.py tests/test_widen_x_full_size_gpu.py tests/test_pipeline_gpu.py -q -x 2>&1 | tail -2
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('enc', round(d['enc']['ms'],2), 'dec', round(d['dec']['ms'],2)); print([ (k[0][-40:],k[1],k[2]) for k in d['dec']['top'][:4]])"; }
for i in 1 2; do
echo "old kernel (4 accumulators, rows=4)"; VCOF_LIB=$PWD/tests/native/old/libvcof.so VCOF_CONV_SPARE=0 timeout 200 python tools/vae_bench.py --frames 9 2>&1 | show
echo "new kernel, default plan"; timeout 200 python tools/vae_bench.py --frames 9 2>&1 | show
done
