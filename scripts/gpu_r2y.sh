#!/usr/bin/env bash
# Round-2 call Y (gpurun --gpus 2): exchange kernel on a co-resident grid (occupancy-sized) against the previous commit
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
pick() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g  ms/step %.4f  kernel %.4f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])); print(json.dumps(d['arm']['exchange_phases_us'])[:700])"; }
for lib in prev main; do
  L=""; [ $lib != main ] && L=$PWD/_variants/libkge_$lib.so
  echo "== bench N=$N $lib"
  KGE_B200_LIB=$L timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/y_bench_n${N}_$lib.json 2> gpurun_out/y_bench_n${N}_$lib.err
  pick < gpurun_out/y_bench_n${N}_$lib.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/y_bench_n${N}_$lib.err | tail -3
done
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_gpu_z_multi.py -q > gpurun_out/y_tests_multi_n$N.log 2>&1; grep -E "passed|failed|Error|assert" gpurun_out/y_tests_multi_n$N.log | tail -8
