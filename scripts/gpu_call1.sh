#!/usr/bin/env bash
# round-2 GPU call 1 (1 GPU): full parity suite (exact ranking mode), the side-sorted train-kernel variant under the same
# suite, kernel A/B, bench, and -- LAST, each under its own timeout -- the new tensor-core ranking path
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt 2>&1
echo "== tests (main lib, exact ranking)"; KGE_B200_RANK_MODE=exact timeout 1200 python -m pytest tests -m gpu -x -q -k "not tc_filter and not tensor_core" 2>&1 | tail -25 | tee gpurun_out/c1_tests_main.log
echo "== tests (sorted variant)"; KGE_B200_RANK_MODE=exact KGE_B200_LIB=$PWD/_variants/libkge_sorted.so timeout 900 python -m pytest tests/test_gpu_parity.py -q \
  -k "forward_backward or wide_rows or negative_groups or edge_shapes or philox or external or full_size or lazy or train_steps or linearity" 2>&1 | tail -8 | tee gpurun_out/c1_tests_sorted.log
echo "== kbench main"; timeout 600 python scripts/kbench.py cfg2 cfg2u cfg3 cfg4 cfg4c cfg5w big 2>&1 | tee gpurun_out/c1_kbench_main.log
echo "== kbench sorted"; KGE_B200_LIB=$PWD/_variants/libkge_sorted.so timeout 600 python scripts/kbench.py cfg2 cfg2u cfg3 cfg4 cfg4c cfg5w big 2>&1 | tee gpurun_out/c1_kbench_sorted.log
echo "== bench (exact ranking)"; KGE_B200_RANK_MODE=exact timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -c 7000 gpurun_out/c1_bench.json; tail -5 gpurun_out/c1_bench.err
echo "== rbench exact"; KGE_B200_RANK_MODE=exact timeout 300 python scripts/rbench.py 2>&1 | tee gpurun_out/c1_rbench_exact.log
echo "== TC probe test"; timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -k "tc_filter_error_bound" 2>&1 | tail -30 | tee gpurun_out/c1_tc_probe.log
echo "== TC rank tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core or ranks_bit_exact or ranks_subset or ranks_full_size" 2>&1 | tail -30 | tee gpurun_out/c1_tc_ranks.log
echo "== rbench auto"; timeout 200 python scripts/rbench.py 2>&1 | tee gpurun_out/c1_rbench_auto.log
