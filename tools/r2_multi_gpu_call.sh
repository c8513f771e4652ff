#!/bin/bash
# Multi-GPU call of round 2 (N = 2 by default; `N=8 bash tools/r2_multi_gpu_call.sh` for the confirmation run):
#     gpurun --gpus 2 --timeout 1200 -- 'bash tools/r2_multi_gpu_call.sh'
# 1. parity of all three exchange schemes against the single-GPU forward, toy widths and C2 widths (75 600 tokens),
# 2. frame-sharded VAE parity, 3. A/B of the exchange schemes on the c2 step (latents_sha256 must equal the N = 1 one),
# 4. the 2-GPU test file.  Every step runs under `timeout`: a hang in a cross-GPU barrier must not take the box down.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
OUT=gpurun_out/r2_multi_gpu_n$N.log
: > "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
step() { echo "### $*" | tee -a "$OUT"; timeout "${T:-300}" "$@" 2>&1 | grep -v "Warning\|warn" | tail -${TAIL:-6} | tee -a "$OUT"; echo "rc=${PIPESTATUS[0]}" | tee -a "$OUT"; }

T=200 step $TR --master-port 29511 tools/sp_check.py push                 # toy widths: heads, gather, push
T=300 step $TR --master-port 29512 tools/sp_check.py push c2              # C2 widths and token count, 1 layer
T=300 step $TR --master-port 29513 tools/vae_shard_check.py
for mode in ${MODES:-auto push heads}; do
  echo "### VCOF_SP_MODE=$mode bench" | tee -a "$OUT"
  VCOF_SP_MODE=$mode T=600 TAIL=2 step $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_FLAGS:-}
done
[ "$N" = 2 ] && T=600 step python -m pytest tests/test_multi_gpu.py -q
echo done | tee -a "$OUT"
