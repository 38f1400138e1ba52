#!/usr/bin/env bash
# Round-2 call T: ncu of the RotatE fast path + launch list of a bench step with one optimizer launch, bench probes (why the
# kernel takes 147 us in bench.py's process and 138 us in kbench)
set -u
mkdir -p gpurun_out
bash scripts/ncu_all.sh r2t "launches_bench train_cfg4 train_cfg2" > gpurun_out/t_ncu.log 2>&1
ls gpurun_out | grep r2t
echo "== probes"
KGE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2> gpurun_out/t_bench_probes.log | tail -c 300; grep -i "probe\|debug" gpurun_out/t_bench_probes.log | head -40
echo "== probes, no hot hint"
KGE_BENCH_HOT=0 KGE_BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2> gpurun_out/t_bench_probes_nohot.log | tail -c 300; grep -i "probe\|debug" gpurun_out/t_bench_probes_nohot.log | head -40
echo "== kbench with the optimizer between the steps"
KBENCH_OPT=1 timeout 200 python scripts/kbench.py cfg2 cfg2u 2>&1 | tee gpurun_out/t_kbench_opt.log
