#!/bin/bash
# Builds the stand-alone native parity check (selftest_t5) and kernel timer (kbench) against the in-tree libvcof.so (run __graft_entry__.build() first).
set -e
cd "$(dirname "$0")"
NVCC="${NVCC:-$(command -v nvcc || echo /usr/local/cuda/bin/nvcc)}"
# hardware probes (descriptor conventions, TMA feed rates): their own library, not part of the product .so
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
     probes/tma_probe.cu probes/umma_probe.cu -o libvcof_probes.so \
     -L../../videocof_b200/csrc -lvcof -Xlinker -rpath -Xlinker '$ORIGIN/../../videocof_b200/csrc'
echo built tests/native/libvcof_probes.so
for t in selftest_t5 selftest_core kbench; do
  "$NVCC" -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a $t.cu -o $t \
       -L../../videocof_b200/csrc -lvcof -Xlinker -rpath -Xlinker '$ORIGIN/../../videocof_b200/csrc'
  echo built tests/native/$t
done
