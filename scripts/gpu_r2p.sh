#!/usr/bin/env bash
# Round-2 call P: chunked draws A/B, the full bench line, smoke, ncu of the headline kernel + launch lists
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_z_multi.py 2>&1 | tail -5 | tee gpurun_out/p_tests.log
for lib in main chunk2 chunk3 main; do
  L=""; [ $lib != main ] && L=$PWD/_variants/libkge_$lib.so
  echo "== $lib"; KGE_B200_LIB=$L timeout 300 python scripts/kbench.py cfg2 cfg2u cfg3 2>&1 | tee -a gpurun_out/p_kbench_$lib.log
done
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/p_smoke.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/p_bench_line.json 2> gpurun_out/p_bench.err; tail -c 600 gpurun_out/p_bench_line.json; tail -3 gpurun_out/p_bench.err
bash scripts/ncu_all.sh r2p "launches_bench launches_evaluate train_cfg2" > gpurun_out/p_ncu.log 2>&1
ls gpurun_out | grep r2p
