#!/usr/bin/env bash
# Round-2 call L: RotatE scorer diet / two-value butterfly / DistMult NIT=4 fast path: parity + A/B against the previous commit
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_z_multi.py 2>&1 | tail -15 | tee gpurun_out/l_tests.log
for lib in base main; do
  L=""; [ $lib != main ] && L=$PWD/_variants/libkge_$lib.so
  echo "== $lib"; KGE_B200_LIB=$L timeout 300 python scripts/kbench.py cfg4 cfg1 cfg3 cfg4c 2>&1 | tee gpurun_out/l_kbench_$lib.log
done
echo "== main, cfg3 on the fast path with 6 warps"
KGE_B200_RES_MIN_WARPS=6 timeout 300 python scripts/kbench.py cfg3 2>&1 | tee gpurun_out/l_kbench_cfg3_res6.log
echo "== main, cfg4 group sweep"
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/l_kbench_cfg4_sweep.log
import sys; sys.path.insert(0, 'scripts')
import kbench
for g in (0, 14, 12, 10, 9, 8, 7, 6, 5, 4):
    kbench.run("cfg4", neg_group=g)
PY
