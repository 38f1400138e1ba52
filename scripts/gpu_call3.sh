#!/usr/bin/env bash
# round-2 validation call on a 2-GPU box: full single-GPU suite, 2-GPU parity tests, bench N=1 and N=2, ranking benchmarks
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== $N GPUs"
echo "== tests (all, $N GPUs visible)"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/c3_tests.log
echo "== bench N=1"; CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c3_bench_n1.json 2> gpurun_out/c3_bench_n1.err; tail -c 6000 gpurun_out/c3_bench_n1.json; tail -5 gpurun_out/c3_bench_n1.err
if [ "$N" -ge 2 ]; then
echo "== bench N=$N"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c3_bench_n$N.json 2> gpurun_out/c3_bench_n$N.err
tail -c 9000 gpurun_out/c3_bench_n$N.json; tail -15 gpurun_out/c3_bench_n$N.err
echo "== bench N=$N nvls"; KGE_B200_DP_MODE=nvls timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus $N --steps 20 --warmup 5 --no-extra > gpurun_out/c3_bench_n${N}_nvls.json 2> gpurun_out/c3_bench_n${N}_nvls.err
tail -c 2500 gpurun_out/c3_bench_n${N}_nvls.json; tail -5 gpurun_out/c3_bench_n${N}_nvls.err
fi
echo "== rbench auto"; CUDA_VISIBLE_DEVICES=0 timeout 300 python scripts/rbench.py 2>&1 | tee gpurun_out/c3_rbench_auto.log
echo "== evaluate launch list"; CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/c3_launches_evaluate.csv python scripts/rbench.py one ComplEx 200 14505 1024 > /dev/null 2>&1
python scripts/ncu_summary.py launches gpurun_out/c3_launches_evaluate.csv gpurun_out/c3_launch_list_evaluate_summary.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 200 python scripts/rbench.py one ComplEx 200 14505 1024"; cat gpurun_out/c3_launch_list_evaluate_summary.csv
