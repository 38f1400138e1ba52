#!/usr/bin/env bash
# N-GPU call (gpurun --gpus 8): bench.py under torchrun with the full extra block (cfg5 AS STATED at N=8: 10 M entities,
# k=1000, eta=50, lazy Adam, row-sharded, + 1,024-triple full-entity ranking), then the p2p exchange for comparison
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== $N GPUs"; nvidia-smi --query-gpu=name,memory.total --format=csv | head -3
echo "== row-sharded parity test (2 of the GPUs)"; timeout 600 python -m pytest tests/test_gpu_z_multi.py -q -k row_sharded > gpurun_out/j_tests_sharded.log 2>&1; grep -E "max err|passed|failed|Error" gpurun_out/j_tests_sharded.log | tail -12
pick() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g  ms/step %.4f  kernel %.4f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])); print(json.dumps(d['arm'])[:1500]); [print(k, json.dumps(v)[:1600]) for k, v in d.get('extra', {}).items()]"; }
run() { # run <tag> <extra flags> [env...]
  local tag=$1 flags=$2; shift 2
  echo "== bench N=$N $tag"
  env "$@" timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 $flags > gpurun_out/j_bench_n${N}_$tag.json 2> gpurun_out/j_bench_n${N}_$tag.err
  pick < gpurun_out/j_bench_n${N}_$tag.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/j_bench_n${N}_$tag.err | tail -4
}
run auto "" KGE_B200_DP_MODE=auto
run p2p "--no-extra --no-cpu" KGE_B200_DP_MODE=p2p
