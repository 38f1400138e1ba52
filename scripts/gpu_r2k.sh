#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/exchange_tune.py nvls p2p 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tee gpurun_out/k_exchange_tune_n$N.log
