#!/bin/bash
# Multi-GPU call of round 2 (N = 2 by default; N=8 bash tools/r2_multi_gpu_call.sh for the confirmation run):
#     gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_multi_gpu_call.sh'
# 1. parity of all three exchange schemes against the single-GPU forward (the push exchange has never run on hardware),
# 2. frame-sharded VAE parity, 3. A/B of the exchange schemes on the c2 step, 4. the 2-GPU test file.
# A hang in the push exchange's cross-GPU barrier must not take the box down: every step runs under `timeout`.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
OUT=gpurun_out/r2_multi_gpu_n$N.log
: > "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
step() { echo "### $*" | tee -a "$OUT"; timeout "${T:-300}" "$@" 2>&1 | tail -25 | tee -a "$OUT"; echo "rc=${PIPESTATUS[0]}" | tee -a "$OUT"; }

step $TR --master-port 29511 tools/sp_check.py                      # heads + gather: validated in round 1, must stay ok
T=120 step $TR --master-port 29512 tools/sp_check.py push           # + push exchange (symmetric memory)
step $TR --master-port 29513 tools/vae_shard_check.py
for mode in auto push; do
  echo "### VCOF_SP_MODE=$mode bench" | tee -a "$OUT"
  VCOF_SP_MODE=$mode T=600 step $TR --master-port 29514 bench.py --gpus $N --steps 2 --warmup 3 --no-pipeline --no-cpu-baseline
done
[ "$N" = 2 ] && T=600 step python -m pytest tests/test_multi_gpu.py -q
echo done | tee -a "$OUT"
