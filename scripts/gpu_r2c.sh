#!/usr/bin/env bash
# 1-GPU call: parity tests, then A/B of the resident trilinear fast path (kge_train_res.cu) against the general kernel
set -u
mkdir -p gpurun_out
echo "== tests (single GPU)"; timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_z_multi.py > gpurun_out/c_tests.log 2>&1; tail -12 gpurun_out/c_tests.log
echo "== kbench fast path"; timeout 600 python scripts/kbench.py cfg2 cfg2u big cfg3 2>&1 | tee gpurun_out/c_kbench_fast.log
echo "== kbench general"; KGE_B200_TRAIN_KERNEL=general timeout 600 python scripts/kbench.py cfg2 cfg2u big 2>&1 | tee gpurun_out/c_kbench_general.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; tail -c 2500 gpurun_out/c_bench.json; tail -5 gpurun_out/c_bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu source-level capture, cfg2 fast path"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kge_train_res_kernel -s 4 -c 1 -f -o gpurun_out/c_train_res_cfg2 \
  python scripts/kbench.py one cfg2 0 > gpurun_out/c_train_res_cfg2.stdout 2>&1
ncu -i gpurun_out/c_train_res_cfg2.ncu-rep --page source --csv --print-source sass > gpurun_out/c_train_res_cfg2_source_sass.csv 2>/dev/null
ncu -i gpurun_out/c_train_res_cfg2.ncu-rep --page raw --csv > gpurun_out/c_train_res_cfg2_raw.csv 2>/dev/null
python scripts/ncu_summary.py full gpurun_out/c_train_res_cfg2_raw.csv gpurun_out/c_train_res_cfg2_ncu_full_summary.json "ncu --set full --clock-control none --import-source on -k regex:kge_train_res_kernel -s 4 -c 1 python scripts/kbench.py one cfg2 0"
rm -f gpurun_out/c_train_res_cfg2_raw.csv
fi
