#!/usr/bin/env bash
# round-2 multi-GPU call: N-GPU parity tests + bench.py under torchrun (N = number of visible GPUs)
set -u
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
echo "== $N GPUs"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_gpu_z_multi.py -x -q 2>&1 | tail -30 | tee gpurun_out/c2_tests_multi_n$N.log
fi
echo "== bench N=$N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c2_bench_n$N.json 2> gpurun_out/c2_bench_n$N.err
tail -c 9000 gpurun_out/c2_bench_n$N.json; tail -15 gpurun_out/c2_bench_n$N.err
