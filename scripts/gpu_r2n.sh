#!/usr/bin/env bash
# Round-2 call N: dynamic assignment of positives to warps (A/B against the static stride), ncu of the new kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_z_multi.py 2>&1 | tail -8 | tee gpurun_out/n_tests.log
for sched in static dynamic; do
  echo "== $sched"; KGE_B200_TRAIN_SCHED=$sched timeout 300 python scripts/kbench.py cfg2 cfg2u cfg3 cfg4 big cfg5w 2>&1 | tee gpurun_out/n_kbench_$sched.log
done
bash scripts/ncu_all.sh r2n "train_cfg3 train_cfg4 rank_transe rank_rotate" > gpurun_out/n_ncu.log 2>&1
ls gpurun_out | grep r2n
