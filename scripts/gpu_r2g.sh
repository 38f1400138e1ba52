#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
KGE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; grep -E "probe|per-step kernel" gpurun_out/g_bench.err | cut -c1-300
echo "== fast loss math variant"
KGE_B200_LIB=$PWD/_variants/libkge_fastloss.so KGE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/g_bench_fastloss.json 2> gpurun_out/g_bench_fastloss.err; grep -E "probe|per-step kernel" gpurun_out/g_bench_fastloss.err | cut -c1-300
