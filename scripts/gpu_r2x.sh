#!/usr/bin/env bash
# Round-2 call X: compute-sanitizer over the kernels written this round (RotatE fast path, DistMult NIT=4 fast path, packed pair
# ranking kernel, dynamic assignment, single-buffered groups): memcheck on their tests, racecheck on a subset
set -u
mkdir -p gpurun_out
SEL='fast_path_equals_general or packed_pair or dynamic_assignment or single_buffered or forward_backward_vs_oracle'
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" > gpurun_out/x_memcheck.log 2>&1; echo "exit $?" | tee -a gpurun_out/x_memcheck.log; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/x_memcheck.log | tail -8
echo "== racecheck"; timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "fast_path_equals_general and RotatE or packed_pair and RotatE-7" > gpurun_out/x_racecheck.log 2>&1; echo "exit $?" | tee -a gpurun_out/x_racecheck.log; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/x_racecheck.log | tail -8
