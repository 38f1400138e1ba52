#!/usr/bin/env bash
# Round-2 call M: packed pair ranking kernel (TransE / RotatE), single-buffer groups + 12 warps for RotatE training
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_z_multi.py 2>&1 | tail -8 | tee gpurun_out/m_tests.log
echo "== parity file with single-buffered groups"
KGE_B200_TRAIN_NBUF=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/m_tests_nbuf1.log
echo "== rbench pair kernel"
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/m_rbench_pair.log
import sys; sys.path.insert(0, 'scripts')
import rbench
rbench.run("TransE", 400, 14505, 237, 1024)
rbench.run("RotatE", 200, 14505, 237, 1024)
rbench.run("RotatE", 200, 123182, 37, 1024)
PY
echo "== rbench tile kernel"
KGE_B200_RANK_KERNEL=tile timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/m_rbench_tile.log
import sys; sys.path.insert(0, 'scripts')
import rbench
rbench.run("TransE", 400, 14505, 237, 1024)
rbench.run("RotatE", 200, 14505, 237, 1024)
PY
echo "== cfg4 sweep (nbuf, G, warps cap)"
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/m_kbench_cfg4_sweep.log
import os, sys; sys.path.insert(0, 'scripts')
import kbench
for nbuf, gs in ((2, (3, 4, 5, 6, 8)), (1, (5, 6, 8, 10, 12, 15))):
    for w in (8, 10, 12):
        for g in gs:
            os.environ["KGE_B200_TRAIN_NBUF"] = str(nbuf); os.environ["KGE_B200_TRAIN_WARPS"] = str(w)
            print("nbuf=%d warps<=%d" % (nbuf, w), end="  ")
            kbench.run("cfg4", neg_group=g, iters=10)
PY
echo "== defaults"
timeout 300 python scripts/kbench.py cfg4 cfg4c cfg5w cfg3 2>&1 | tee gpurun_out/m_kbench_main.log
