#!/usr/bin/env bash
# round-2 single-GPU evidence call: full GPU suite, kernel micro-benchmarks, bench line, ranking benchmarks, then the ncu
# captures of every kernel family (scripts/ncu_all.sh).  Everything lands in gpurun_out/ (prefix a_).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/a_tests.log
echo "== kbench"; timeout 600 python scripts/kbench.py cfg2 cfg2u cfg3 cfg4 cfg4c cfg5w big 2>&1 | tee gpurun_out/a_kbench.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -c 7000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
echo "== bench reference arm"; timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; tail -c 2000 gpurun_out/a_bench_ref.json
echo "== rbench auto"; timeout 300 python scripts/rbench.py 2>&1 | tee gpurun_out/a_rbench_auto.log
echo "== rbench exact"; KGE_B200_RANK_MODE=exact timeout 300 python scripts/rbench.py 2>&1 | tee gpurun_out/a_rbench_exact.log
echo "== ncu"; timeout 1500 bash scripts/ncu_all.sh r2a 2>&1 | tail -40
