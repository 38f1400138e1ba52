#!/usr/bin/env python
"""Kernel micro-benchmark (development aid): times kge_train_step alone for the BASELINE
config shapes under different residency groups.  CUDA events, L2 flushed
between iterations.  Usage: python scripts/kbench.py [cfg2 cfg3 cfg4 ...]"""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ampligraph_b200.engine import KGEEngine  # noqa: E402

CFGS = {
    "cfg1": dict(model="TransE", k=50, eta=2, E=1000, R=10, B=1000, loss="pairwise"),
    "cfg2": dict(model="ComplEx", k=200, eta=10, E=14505, R=237, B=27212, loss="self_adversarial"),
    "cfg2u": dict(model="ComplEx", k=200, eta=10, E=14505, R=237, B=27212, loss="self_adversarial", uniform=True),
    "cfg3": dict(model="DistMult", k=400, eta=20, E=40943, R=11, B=8684, loss="pairwise"),
    "cfg4": dict(model="RotatE", k=200, eta=30, E=123182, R=37, B=10791, loss="self_adversarial"),
    "cfg4c": dict(model="ComplEx", k=200, eta=30, E=123182, R=37, B=10791, loss="self_adversarial"),
    "cfg5w": dict(model="ComplEx", k=1000, eta=50, E=200000, R=1000, B=8192, loss="self_adversarial", uniform=True),
    "big": dict(model="ComplEx", k=200, eta=10, E=2000000, R=1000, B=65536, loss="self_adversarial", uniform=True),
}


def triples(c, rng):
    E, R, B = c["E"], c["R"], c["B"]
    if c.get("uniform"):
        s, o = rng.integers(0, E, B), rng.integers(0, E, B)
    else:
        w = 1.0 / np.arange(1, E + 1)
        cdf = np.cumsum(w / w.sum())
        perm = rng.permutation(E)
        s, o = perm[np.searchsorted(cdf, rng.random(B))], perm[np.searchsorted(cdf, rng.random(B))]
    return np.stack([s, rng.integers(0, R, B), o], 1).astype(np.int32)


def run(name, neg_group=0, iters=20):
    c = CFGS[name]
    rng = np.random.default_rng(0)
    eng = KGEEngine(c["model"], c["k"], c["eta"], c["E"], c["R"], loss=c["loss"], neg_group=neg_group)
    eng.init_glorot_uniform(1)
    t = torch.as_tensor(triples(c, rng)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for i in range(3):
        eng.forward_backward(t, None, seed=1, step=i)
    torch.cuda.synchronize()
    ms = []
    for i in range(iters):
        flush.fill_(i & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.forward_backward(t, None, seed=1, step=10 + i)
        e1.record()
        if os.environ.get("KBENCH_OPT") == "1":  # the full step of bench.py: dense optimizer after the kernel, outside the events
            eng.apply_gradients()
        if os.environ.get("KBENCH_NOSYNC") != "1":
            torch.cuda.synchronize()
        ms.append((e0, e1))
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ms])
    row_bytes = eng.ld * 4
    alg = 2 * (3 + c["eta"]) * row_bytes * c["B"]
    g = eng.g_ent.double().abs().sum().item()
    print("%-6s %-9s G=%-2d  med %8.1f us  min %8.1f us  %7.1f GB/s alg  %6.2f Mpos/s  %8.1f Mtriples/s  |g|=%.6e"
          % (name, os.path.basename(os.environ.get("KGE_B200_LIB", "main")).replace("libkge_", "").replace(".so", ""), neg_group, np.median(ms) * 1e3, ms.min() * 1e3, alg / np.median(ms) / 1e6,
             c["B"] / np.median(ms) / 1e3, c["B"] * (1 + c["eta"]) / np.median(ms) / 1e3, g), flush=True)
    eng.close()


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "one":  # one <cfg> <G>   (for ncu)
        run(args[1], int(args[2]), iters=5)
        sys.exit(0)
    if args and args[0] == "sweep":
        for n, gs in (("cfg4", (0, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3)), ("cfg4c", (0, 15, 13, 11, 10, 9, 8, 7, 6, 5)),
                      ("cfg3", (0, 13, 11, 10, 9, 7, 5)), ("cfg5w", (0,))):
            for g in gs:
                run(n, neg_group=g)
        sys.exit(0)
    names = args or ["cfg2", "cfg2u", "cfg3", "cfg4", "cfg1", "big"]
    for n in names:
        run(n)
