#!/usr/bin/env bash
# Round-2 call W (gpurun --gpus 2): final state of the round: the whole GPU suite (incl. the multi-GPU parity tests), smoke, bench at N=2
# with the full extra block
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== GPU suite ($N GPUs)"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/w_tests_n$N.log 2>&1; grep -E "max err|passed|failed|Error|assert" gpurun_out/w_tests_n$N.log | tail -20
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/w_smoke.log
pick() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g  ms/step %.4f  kernel %.4f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])); print(json.dumps(d['arm'])[:1200]); [print(k, json.dumps(v)[:1500]) for k, v in d.get('extra', {}).items()]"; }
echo "== bench N=$N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/w_bench_n${N}.json 2> gpurun_out/w_bench_n${N}.err
pick < gpurun_out/w_bench_n${N}.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/w_bench_n${N}.err | tail -4
