#!/usr/bin/env bash
# 1-GPU call: hot-entity privatisation of the fast path
set -u
mkdir -p gpurun_out
echo "== tests (train)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_facade.py -x -q > gpurun_out/h_tests.log 2>&1; tail -4 gpurun_out/h_tests.log
echo "== kbench"; timeout 300 python scripts/kbench.py cfg2 cfg2u big 2>&1 | tee gpurun_out/h_kbench.log
for hot in 1 0; do
echo "== bench hot=$hot"; KGE_BENCH_HOT=$hot KGE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/h_bench_hot$hot.json 2> gpurun_out/h_bench_hot$hot.err; grep -E "per-step kernel|probe same engine, bench batches, no opt|uniform" gpurun_out/h_bench_hot$hot.err | cut -c1-300
python -c "import sys,json; d=json.loads(open('gpurun_out/h_bench_hot$hot.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])"
done
