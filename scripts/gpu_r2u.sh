#!/usr/bin/env bash
# Round-2 call U: hot hint ignored under the dynamic assignment, early s/p/o loads in the RotatE fast path: suite, smoke, A/B, bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/u_tests.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/u_smoke.log
for lib in prev main prev main; do
  L=""; [ $lib != main ] && L=$PWD/_variants/libkge_$lib.so
  echo "== $lib"; KGE_B200_LIB=$L timeout 300 python scripts/kbench.py cfg4 cfg2 2>&1 | tee -a gpurun_out/u_kbench_$lib.log
done
pick() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g  ms/step %.4f  kernel %.4f  e2e %.4g launches/step %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['arm']['launches_per_step'])); [print(k, json.dumps(v)[:600]) for k, v in d.get('extra', {}).items()]"; }
echo "== bench --steps 20 --warmup 5"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/u_bench_line_20.json 2> gpurun_out/u_bench_20.err; pick < gpurun_out/u_bench_line_20.json; tail -2 gpurun_out/u_bench_20.err
echo "== bench (defaults)"; timeout 900 python bench.py --no-extra --no-cpu > gpurun_out/u_bench_line.json 2> gpurun_out/u_bench.err; pick < gpurun_out/u_bench_line.json
