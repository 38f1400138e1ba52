#!/usr/bin/env bash
# Round-2 call Z: last check of the final library on one GPU (the world-1 exchange test runs the re-gridded kernel)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/z_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/z_smoke.log
