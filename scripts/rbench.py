#!/usr/bin/env python
"""Ranking micro-benchmark (development aid): times kge_rank for one side, b queries against all
entities of a cfg2/cfg4-shaped table.  CUDA events, L2 flushed between iterations."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ampligraph_b200.engine import KGEEngine  # noqa: E402


def run(model, k, E, R, b, filt=0, iters=10):
    rng = np.random.default_rng(0)
    eng = KGEEngine(model, k, 1, E, R)
    eng.init_glorot_uniform(1)
    t = torch.as_tensor(np.stack([rng.integers(0, E, b), rng.integers(0, R, b), rng.integers(0, E, b)], 1).astype(np.int32)).cuda()
    off = idx = None
    if filt:
        off = torch.arange(0, (b + 1) * filt, filt, dtype=torch.int64).cuda()
        idx = torch.as_tensor(rng.integers(0, E, b * filt).astype(np.int32)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = torch.zeros(b, dtype=torch.int32, device="cuda")
    for side in ("s", "o"):
        for _ in range(2):
            eng.rank(t, side, "worst", off, idx, out=out)
        torch.cuda.synchronize()
        ms = []
        for i in range(iters):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.rank(t, side, "worst", off, idx, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        med = float(np.median(ms))
        flops = 2.0 * b * E * eng.ld
        print("%-8s k=%-4d E=%-7d b=%-5d side=%s filt=%-3d  med %8.1f us  %7.2f TFLOP/s(fma-equiv)  %9.1f Mscores/s  table stream %6.1f GB/s"
              % (model, k, E, b, side, filt, med * 1e3, flops / med / 1e9, b * E / med / 1e3, E * eng.ld * 4 / med / 1e6), flush=True)
    eng.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), 237, int(sys.argv[5]), iters=3)
        sys.exit(0)
    for model, k in (("ComplEx", 200), ("DistMult", 400), ("TransE", 400), ("RotatE", 200), ("HolE", 200)):
        run(model, k, 14505, 237, 1024)
    run("ComplEx", 200, 14505, 237, 4096)
    run("ComplEx", 200, 14505, 237, 1024, filt=50)
    run("ComplEx", 200, 14505, 237, 64)
    run("ComplEx", 200, 1000000, 237, 1024)
