#!/bin/bash
# First GPU call of a kernel-tuning session: parity of the core kernels for every queued variant, then the A/B timings
# that decide them.  Steps 1-4 are Python-free (~1 minute of box time); 6-7 import torch:
#     gpurun --timeout 600 -- 'bash tools/r2_first_call.sh'
# Results land in gpurun_out/r2_first_call.jsonl (one JSON object per line, knobs recorded in each line).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_first_call.jsonl
: > "$OUT"
export KBENCH_OUT="$OUT"
run() { echo "### $*" | tee -a "$OUT"; timeout 120 "$@" | tee -a "$OUT"; }

# 1. parity gates (default build first: it must stay green)
run tests/native/selftest_core
VCOF_ATTN_SPEC=1 run tests/native/selftest_core attn
VCOF_ATTN_SPEC=2 run tests/native/selftest_core attn
VCOF_GEMM_2CTA=1 run tests/native/selftest_core gemm

# 1b. byte-frame kernels written at the end of round 1 without GPU time left (bit-exact vs the host evaluation); the
#     Python side of the same row: python -m pytest tests/test_widen_video_io_gpu.py -q
run tests/native/selftest_core frames

# 2. attention, C2 shape on 8 heads (1/5 of a launch: same per-SM work, 5x less box time)
run tests/native/kbench attn 75600 75600 8 3
VCOF_ATTN_SPEC=1 run tests/native/kbench attn 75600 75600 8 3
VCOF_ATTN_SPEC=2 run tests/native/kbench attn 75600 75600 8 3

# 3. GEMM, the three DiT shapes; pair kernel with the occupancy-sized grid, then a sweep of the pair count
for shape in "75600 5120 5120 0" "75600 13824 5120 1" "75600 5120 13824 2"; do
  run tests/native/kbench gemm $shape 10
  VCOF_GEMM_2CTA=1 run tests/native/kbench gemm $shape 10
done
for pairs in 60 64 68 70 72 74; do
  VCOF_GEMM_2CTA=1 VCOF_GEMM_2CTA_PAIRS=$pairs run tests/native/kbench gemm 75600 5120 5120 0 10
done

# 4. text encoder end to end
run tests/native/selftest_t5 --bench "$OUT"
# 5. (needs `gpurun --gpus 2`, not part of this 1-GPU call) the push exchange of the sequence-parallel path:
#    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sp_check.py push
#    VCOF_SP_MODE=push python -m torch.distributed.run ... bench.py --gpus 2 --steps 2 --warmup 3 --no-pipeline --no-cpu-baseline
# 6. library baseline on the same GPU (plain PyTorch ops + flash-attn 2, one c2 block; the only step that imports torch):
#    what the reference's own CUDA path would make of this B200, next to bench.py's numbers for libvcof
echo "### python tools/torch_block_bench.py" | tee -a "$OUT"
timeout 600 python tools/torch_block_bench.py | tee -a "$OUT"
# 7. the GPU tests written after the round-1 budget was spent (each dry-run on the CPU beforehand)
timeout 900 python -m pytest tests/test_widen_video_io_gpu.py tests/test_widen_w_push_exchange_gpu.py \
    tests/test_widen_x_full_size_gpu.py -q 2>&1 | tail -15 | tee -a "$OUT"
echo done
