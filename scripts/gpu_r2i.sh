#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for lib in main hot1; do
  L=""; [ $lib != main ] && L=$PWD/_variants/libkge_$lib.so
  echo "== $lib kbench"; KGE_B200_LIB=$L timeout 300 python scripts/kbench.py cfg2 cfg2u 2>&1 | tee gpurun_out/i_kbench_$lib.log
  echo "== $lib bench"; KGE_B200_LIB=$L timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])"
done
