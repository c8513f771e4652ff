#!/bin/bash
cd "$(dirname "$0")/.."
N="ncu --set full --clock-control none --import-source on"
VCOF_GEMM_2CTA=1 timeout 200 $N -k regex:gemm2cta -c 1 -o gpurun_out/r2c_gemm2cta_ffn2 tests/native/kbench gemm 75600 5120 13824 2 1 > gpurun_out/r2c_ncu_g2f2.log 2>&1
VCOF_GEMM_2CTA=1 timeout 200 $N -k regex:gemm2cta -c 1 -o gpurun_out/r2c_gemm2cta_ffn1 tests/native/kbench gemm 75600 13824 5120 1 1 > gpurun_out/r2c_ncu_g2f1.log 2>&1
timeout 200 $N -k regex:gemm_bf16 -c 1 -o gpurun_out/r2c_gemm1cta_ffn1 tests/native/kbench gemm 75600 13824 5120 1 1 > gpurun_out/r2c_ncu_g1f1.log 2>&1
ls -la gpurun_out/r2c_*
