#!/bin/bash
# 4-GPU call: frame-sharded VAE with idle ranks (bit-identical to one GPU), then the pipeline leg
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29513 tools/vae_shard_check.py 2>&1 | grep -v "Warning\|warn\|^\*\*\*\|OMP_NUM\|^$" | tail -3
