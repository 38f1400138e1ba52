#!/usr/bin/env bash
# 1-GPU call: fast path with the relation row in registers (12 warps), general-kernel KEEP variant A/B, bench per-step debug
set -u
mkdir -p gpurun_out
echo "== tests (single GPU)"; timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_z_multi.py > gpurun_out/f_tests.log 2>&1; tail -4 gpurun_out/f_tests.log
echo "== kbench fast path"; timeout 600 python scripts/kbench.py cfg2 cfg2u big 2>&1 | tee gpurun_out/f_kbench_fast.log
echo "== kbench general kernel: main vs KEEP variant"
timeout 600 python scripts/kbench.py cfg3 cfg4 cfg4c cfg5w 2>&1 | tee gpurun_out/f_kbench_general_main.log
KGE_B200_LIB=$PWD/_variants/libkge_keep.so timeout 600 python scripts/kbench.py cfg3 cfg4 cfg4c cfg5w 2>&1 | tee gpurun_out/f_kbench_general_keep.log
echo "== bench debug"; KGE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; tail -3 gpurun_out/f_bench.err | cut -c1-700
python -c "import sys,json; d=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['arm'])"
echo "== bench, no flush between steps (diagnostic)"; KGE_BENCH_NOFLUSH=1 KGE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2>&1 >/dev/null | tail -3 | cut -c1-700
