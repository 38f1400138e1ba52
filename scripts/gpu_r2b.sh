#!/usr/bin/env bash
# round-2 multi-GPU evidence call (run with gpurun --gpus N): full GPU suite with N GPUs visible (multi-GPU parity tests
# included), bench.py under torchrun in the p2p / nvls / nccl exchange modes, and one source-level ncu capture of the
# cfg2 train kernel on GPU 0 (per-instruction stall samples).  Outputs: gpurun_out/b_*.
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== $N GPUs"
echo "== tests (all, $N GPUs visible)"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/b_tests_n$N.log
run_bench() {  # run_bench <tag> [env...]
  local tag=$1; shift
  echo "== bench N=$N $tag"
  env "$@" timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 ${EXTRA_FLAGS:-} > gpurun_out/b_bench_n${N}_$tag.json 2> gpurun_out/b_bench_n${N}_$tag.err
  tail -c 9000 gpurun_out/b_bench_n${N}_$tag.json; tail -5 gpurun_out/b_bench_n${N}_$tag.err
}
run_bench p2p KGE_B200_DP_MODE=auto
EXTRA_FLAGS=--no-extra run_bench nvls KGE_B200_DP_MODE=nvls
EXTRA_FLAGS=--no-extra run_bench nccl KGE_B200_DP_MODE=nccl
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu source-level capture, cfg2 train kernel"
CUDA_VISIBLE_DEVICES=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:kge_train_kernel -s 4 -c 1 -f -o gpurun_out/b_train_cfg2_src \
  python scripts/kbench.py one cfg2 0 > gpurun_out/b_train_cfg2_src.stdout 2>&1
ncu -i gpurun_out/b_train_cfg2_src.ncu-rep --page source --csv --print-source sass > gpurun_out/b_train_cfg2_source_sass.csv 2>/dev/null
ls -la gpurun_out/b_train_cfg2_src.ncu-rep gpurun_out/b_train_cfg2_source_sass.csv
fi
