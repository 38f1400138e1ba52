#!/usr/bin/env bash
# Round-2 call V: the next positive's triple requested a positive ahead (fast paths): suite + A/B against the previous commit
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/v_tests.log
for lib in prev main prev main; do
  L=""; [ $lib != main ] && L=$PWD/_variants/libkge_$lib.so
  echo "== $lib"; KGE_B200_LIB=$L timeout 300 python scripts/kbench.py cfg2 cfg2u cfg3 cfg4 big 2>&1 | tee -a gpurun_out/v_kbench_$lib.log
done
