#!/usr/bin/env bash
# Round-2 call S: the RotatE fast path (kge_train_rot_kernel): parity, A/B against the general kernel, cfg4-shaped sweeps
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/s_tests.log
for kern in general fast general fast; do
  K=""; [ $kern = general ] && K=general
  echo "== $kern"; KGE_B200_TRAIN_KERNEL=$K timeout 300 python scripts/kbench.py cfg4 cfg1 2>&1 | tee -a gpurun_out/s_kbench_$kern.log
done
echo "== fast path, group size sweep"
for g in 4 5 6 7 8; do echo -n "G=$g  "; KGE_B200_ROT_G=$g timeout 120 python scripts/kbench.py cfg4 2>&1 | tee -a gpurun_out/s_kbench_rot_sweep.log; done
