#!/usr/bin/env bash
# 1-GPU call: split-scatter A/B of the fast path, the no-scatter profiling build, bench-like kbench modes, bench under a launch list
set -u
mkdir -p gpurun_out
echo "== tests (train)"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "resident_fast_path or forward_backward or train_steps or external or lazy or full_size" > gpurun_out/d_tests.log 2>&1; tail -5 gpurun_out/d_tests.log
for sp in 0 1 2; do echo "== kbench split=$sp"; KGE_B200_SCATTER_SPLIT=$sp timeout 300 python scripts/kbench.py cfg2 cfg2u big 2>&1 | tee gpurun_out/d_kbench_split$sp.log; done
echo "== kbench noscatter build"; KGE_B200_LIB=$PWD/_variants/libkge_noscatter.so timeout 300 python scripts/kbench.py cfg2 big 2>&1 | tee gpurun_out/d_kbench_noscatter.log
echo "== kbench noscatter general"; KGE_B200_TRAIN_KERNEL=general KGE_B200_LIB=$PWD/_variants/libkge_noscatter.so timeout 300 python scripts/kbench.py cfg2 2>&1 | tee gpurun_out/d_kbench_noscatter_general.log
echo "== kbench with optimizer in the loop"; KBENCH_OPT=1 timeout 300 python scripts/kbench.py cfg2 2>&1 | tee gpurun_out/d_kbench_opt.log
echo "== kbench with optimizer, no sync"; KBENCH_OPT=1 KBENCH_NOSYNC=1 timeout 300 python scripts/kbench.py cfg2 2>&1 | tee -a gpurun_out/d_kbench_opt.log
echo "== bench launch list"; timeout 400 ncu --clock-control none --metrics gpu__time_duration.sum -c 120 --csv --log-file gpurun_out/d_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
python scripts/ncu_summary.py launches gpurun_out/d_launches_bench.csv gpurun_out/d_launch_list_bench_summary.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 120 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra"; cat gpurun_out/d_launch_list_bench_summary.csv
for sp in 0 1; do echo "== bench split=$sp"; KGE_B200_SCATTER_SPLIT=$sp timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])"; done
