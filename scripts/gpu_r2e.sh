#!/usr/bin/env bash
# multi-GPU call (gpurun --gpus N): multi-GPU parity tests (full log kept), bench N=1 with/without the NVML sampler (host
# enqueue time), bench at N GPUs in p2p / nvls modes with the exchange-kernel phase stamps
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== multi-GPU tests ($N GPUs)"; timeout 1200 python -m pytest tests/test_gpu_z_multi.py -q > gpurun_out/e_tests_multi_n$N.log 2>&1; grep -E "max err|passed|failed|Error|assert" gpurun_out/e_tests_multi_n$N.log | tail -30
pick() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g  ms/step %.4f  kernel %.4f  e2e %.4g  arm %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], json.dumps(d['arm'])[-160:])); print(json.dumps(d.get('extra', {}).get('exchange_phases_us')))"; }
echo "== bench N=1 sampler on"; CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2>/dev/null | pick
echo "== bench N=1 sampler off"; KGE_BENCH_SAMPLER=0 CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2>/dev/null | pick
for mode in auto nvls; do
echo "== bench N=$N $mode"
KGE_B200_DP_MODE=$mode timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 5 ${EXTRA_FLAGS:-} > gpurun_out/e_bench_n${N}_$mode.json 2> gpurun_out/e_bench_n${N}_$mode.err
pick < gpurun_out/e_bench_n${N}_$mode.json; tail -3 gpurun_out/e_bench_n${N}_$mode.err
EXTRA_FLAGS=--no-extra
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e_bench_n*_auto.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    for k,v in d.get('extra',{}).items():
        print(k, json.dumps(v)[:900])
PY
