#!/usr/bin/env bash
# round-2 single-GPU validation call: full parity suite (tensor-core ranking on), the side-sorted train-kernel variant
# under the training tests, kernel A/B, bench, ranking micro-benchmarks + launch list, exhaustive sqrt check
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt 2>&1
echo "== sqrt check"; ([ -x scripts/check_sqrt ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/check_sqrt scripts/check_sqrt.cu) && timeout 120 scripts/check_sqrt 2>&1 | tee gpurun_out/c1_check_sqrt.log
echo "== tests (main lib)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/c1_tests_main.log
echo "== tests (sorted variant)"; KGE_B200_LIB=$PWD/_variants/libkge_sorted.so timeout 900 python -m pytest tests/test_gpu_parity.py -q \
  -k "forward_backward or wide_rows or negative_groups or edge_shapes or philox or external or full_size or lazy or train_steps or linearity" 2>&1 | tail -8 | tee gpurun_out/c1_tests_sorted.log
echo "== kbench main"; timeout 600 python scripts/kbench.py cfg2 cfg2u cfg3 cfg4 cfg4c cfg5w big 2>&1 | tee gpurun_out/c1_kbench_main.log
echo "== kbench sorted"; KGE_B200_LIB=$PWD/_variants/libkge_sorted.so timeout 600 python scripts/kbench.py cfg2 cfg2u cfg3 cfg4 cfg4c cfg5w big 2>&1 | tee gpurun_out/c1_kbench_sorted.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -c 7000 gpurun_out/c1_bench.json; tail -5 gpurun_out/c1_bench.err
echo "== rbench auto"; timeout 300 python scripts/rbench.py 2>&1 | tee gpurun_out/c1_rbench_auto.log
echo "== evaluate launch list"; timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/c1_launches_evaluate.csv python scripts/rbench.py one ComplEx 200 14505 1024 > /dev/null 2>&1
python scripts/ncu_summary.py launches gpurun_out/c1_launches_evaluate.csv gpurun_out/c1_launch_list_evaluate_summary.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 200 python scripts/rbench.py one ComplEx 200 14505 1024"; cat gpurun_out/c1_launch_list_evaluate_summary.csv
