#!/bin/bash
cd "$(dirname "$0")/.."
for shape in "75600 5120 5120 0" "75600 13824 5120 1" "75600 5120 13824 2"; do
  for g in 4 8 16 32 64; do
    for pair in 0 1; do
      echo "group_m=$g pair=$pair $(VCOF_GEMM_GROUP_M=$g VCOF_GEMM_2CTA=$pair tests/native/kbench gemm $shape 10 | cut -c1-140)"
    done
  done
done
