"""TEST INFRASTRUCTURE (see oracle/__init__.py): independent restatement of the corruption stream of libkge_b200.

The reference draws its corruptions with two stateful `tf.random.uniform` ops
(ampligraph/latent_features/layers/corruption_generation/CorruptionGenerationLayerTrain.py:55-74); that stream
lives in TensorFlow and cannot be reproduced (RNG parity unpinned).  The product replaces it by the
counter-based Philox4x32-10 generator of Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3"
(SC'11).  This file states that generator from the paper, in plain Python integers, and is pinned by the Random123
known-answer vectors (tests/test_oracle.py); the library's device and host draws are then checked against it.
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = [int(x) & MASK for x in ctr]
    k0, k1 = [int(x) & MASK for x in key]
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = (p1 >> 32) ^ c1 ^ k0, p1 & MASK, (p0 >> 32) ^ c3 ^ k1, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def corruption_stream(triples, eta, n_ent, seed, step):
    """[eta*B, 3] corruptions: row r = j*B+i is the j-th corruption of positive i (tile order,
    CorruptionGenerationLayerTrain.py:52); counter = (r, step), key = seed; keep_subj = x & 1,
    replacement = (y * n_ent) >> 32; exactly one side replaced, relation kept (:77-88)."""
    t = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    B = t.shape[0]
    out = np.empty((B * eta, 3), dtype=np.int32)
    for r in range(B * eta):
        x, y, _, _ = philox4x32_10((r & MASK, r >> 32, step & MASK, step >> 32), (seed & MASK, seed >> 32))
        keep, repl = x & 1, (y * int(n_ent)) >> 32
        s, p, o = t[r % B]
        out[r] = (s, p, repl) if keep else (repl, p, o)
    return out
