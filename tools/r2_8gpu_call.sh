#!/bin/bash
# 8-GPU confirmation call (charged 8x): parity of the three exchange schemes at C2 widths, then the C2 step bench with
# the default exchange (push) incl. the pipeline and the C5 (321-frame) legs, then the head exchange for comparison.
#     gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_8gpu_call.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
OUT=gpurun_out/r2_multi_gpu_n$N.log
: > "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
step() { echo "### $*" | tee -a "$OUT"; timeout "${T:-300}" "$@" 2>&1 | grep -v "Warning\|warn\|^\*\*\*\|OMP_NUM_THREADS\|^$" | tail -${TAIL:-4} | tee -a "$OUT"; echo "rc=${PIPESTATUS[0]}" | tee -a "$OUT"; }
T=240 step $TR --master-port 29512 tools/sp_check.py push c2
echo "### default exchange (auto -> push)" | tee -a "$OUT"
T=420 TAIL=2 step $TR --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline
echo "### VCOF_SP_MODE=heads" | tee -a "$OUT"
VCOF_SP_MODE=heads T=300 TAIL=2 step $TR --master-port 29515 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-c5 --no-pipeline
echo done | tee -a "$OUT"
