#!/bin/bash
# 8-GPU run of the finished tree: default bench line (push exchange; pipeline and C5 legs) + frame-sharded VAE parity.
#     gpurun --gpus 8 --timeout 700 -- 'bash tools/r2_call28.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c28_bench_n$N.out 2>&1
echo "bench rc=$?"
grep '^{' gpurun_out/r2c28_bench_n$N.out > gpurun_out/r2c28_bench_n$N.json
cut -c1-2600 gpurun_out/r2c28_bench_n$N.json
timeout 200 $TR --master-port 29516 tools/vae_shard_check.py 2>&1 | grep -v "Warning\|warn\|^\*\*\*\|OMP_NUM_THREADS\|^$" | tail -n 3 | tee gpurun_out/r2c28_vae_shard_n$N.log
