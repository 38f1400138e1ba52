#!/usr/bin/env bash
# Round-2 call O: dynamic assignment of positives with the float counter (draw latency hidden), A/B against the static stride
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_z_multi.py 2>&1 | tail -8 | tee gpurun_out/o_tests.log
for sched in static dynamic static dynamic; do
  echo "== $sched"; KGE_B200_TRAIN_SCHED=$sched timeout 300 python scripts/kbench.py cfg2 cfg2u cfg3 cfg4 big cfg5w cfg1 2>&1 | tee -a gpurun_out/o_kbench_$sched.log
done
