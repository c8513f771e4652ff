#!/bin/bash
# Builds the stand-alone native parity check against the in-tree libvcof.so (run __graft_entry__.build() first).
set -e
cd "$(dirname "$0")"
nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a selftest_t5.cu -o selftest_t5 \
     -L../../videocof_b200/csrc -lvcof -Xlinker -rpath -Xlinker '$ORIGIN/../../videocof_b200/csrc'
echo built tests/native/selftest_t5
