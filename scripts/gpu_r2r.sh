#!/usr/bin/env bash
# Round-2 call R: [ent | rel] as one block (one optimizer launch per step), warm cfg5 ranking measurement: tests, smoke, bench (driver-like and default)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r_tests.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r_smoke.log
pick() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g  ms/step %.4f  kernel %.4f  e2e %.4g launches/step %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['arm']['launches_per_step'])); [print(k, json.dumps(v)[:700]) for k, v in d.get('extra', {}).items()]"; }
echo "== bench --steps 20 --warmup 5"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r_bench_line_20.json 2> gpurun_out/r_bench_20.err; pick < gpurun_out/r_bench_line_20.json; tail -2 gpurun_out/r_bench_20.err
echo "== bench (defaults)"; timeout 900 python bench.py --no-extra > gpurun_out/r_bench_line.json 2> gpurun_out/r_bench.err; pick < gpurun_out/r_bench_line.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-600 | tee gpurun_out/r_bench_reference_line.json
