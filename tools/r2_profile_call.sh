#!/bin/bash
# One-GPU profiling call (ncu --set full on the hot kernels at workload shapes + launch list of a bench step).
#     gpurun --timeout 1200 -- 'bash tools/r2_profile_call.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 200 $N -k regex:attn_fwd -c 1 -o gpurun_out/r2b_attn tests/native/kbench attn 75600 75600 40 1 > gpurun_out/r2b_ncu_attn.log 2>&1
timeout 200 $N -k regex:gemm2cta -c 1 -o gpurun_out/r2b_gemm2cta tests/native/kbench gemm 75600 5120 5120 0 1 > gpurun_out/r2b_ncu_gemm2.log 2>&1
timeout 200 $N -k regex:gemm_bf16 -c 1 -o gpurun_out/r2b_gemm_ffn2 tests/native/kbench gemm 75600 5120 13824 2 1 > gpurun_out/r2b_ncu_gemm1.log 2>&1
timeout 400 $N -k regex:conv_lines -s 1 -c 7 -o gpurun_out/r2b_conv_lines python tools/vae_bench.py --frames 9 > gpurun_out/r2b_ncu_conv.log 2>&1
timeout 300 $N -k regex:rmsnorm_rope -c 1 -o gpurun_out/r2b_rmsrope python tools/kbench.py rms --L 75600 --C 5120 > gpurun_out/r2b_ncu_rms.log 2>&1
# launch list of one bench step (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-pipeline --no-cpu-baseline --no-gpu-reference > gpurun_out/r2b_launch_bench.log 2>&1
ls -la gpurun_out/r2b_*
