"""Pin the oracle (oracle/kge_oracle.c and oracle/ref_step.py) against the
reference's own golden vectors (tests/golden/reference_kats.json, lifted from
the reference's test files by tests/golden/extract_reference_kats.py)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import c_oracle, ref_step

MODELS = ["TransE", "DistMult", "ComplEx", "HolE", "RotatE"]


def _case_arrays(c):
    return (np.array(c["e_s"], np.float32), np.array(c["e_p"], np.float32), np.array(c["e_o"], np.float32))


def test_scoring_kats_c_oracle(kats):
    assert len(kats["scoring"]) == 15
    for c in kats["scoring"]:
        e_s, e_p, e_o = _case_arrays(c)
        mrs = c["max_rel_size"] or 1
        if c["form"] == "triple":
            got = c_oracle.score_rows(c["model"], e_s, e_p, e_o, mrs)
        else:
            side = "s" if c["form"] == "sub_diag" else "o"
            got = np.diag(c_oracle.corruption_scores(c["model"], side, e_s, e_p, e_o,
                                                     np.array(c["ent_matrix"], np.float32), mrs))
        got = np.around(got, c["round_decimals"])
        assert (got == np.array(c["expected"], np.float32)).all(), (c["source"], got)


def test_scoring_kats_torch_oracle(kats):
    for c in kats["scoring"]:
        e_s, e_p, e_o = [torch.tensor(a) for a in _case_arrays(c)]
        if c["form"] == "triple":
            got = ref_step.compute_scores(c["model"], e_s, e_p, e_o, c["max_rel_size"]).numpy()
        else:
            side = "s" if c["form"] == "sub_diag" else "o"
            got = np.diag(ref_step.corruption_scores(c["model"], side, e_s, e_p, e_o,
                                                     torch.tensor(c["ent_matrix"], dtype=torch.float32),
                                                     c["max_rel_size"]).numpy())
        got = np.around(got, c["round_decimals"])
        assert (got == np.array(c["expected"], np.float32)).all(), (c["source"], got)


def test_rank_kats(kats):
    r = kats["ranks"]
    e_s, e_p, e_o = _case_arrays(r)
    cand = np.array(r["ent_matrix"], np.float32)
    assert len(r["cases"]) == 6
    for case in r["cases"]:
        sides = [s for s in ("s", "o") if s in case["corrupt_side"]]
        filt = case["filters"]
        got = []
        for n, side in enumerate(sides):
            # AbstractScoringLayer.py:262 / :375-378: filter_index 0 for the first side present, 1 for 'o' of 's,o'
            f = filt[n] if filt else None
            got.append(c_oracle.ranks_side(r["model"], side, case["comparison_type"], e_s, e_p, e_o, cand,
                                           case["start_ent_id"], case["end_ent_id"], f))
        assert (np.array(got) == np.array(case["expected"], np.int32)).all(), (case, got)


def test_loss_kats(kats):
    assert len(kats["losses"]) == 10
    for c in kats["losses"]:
        p = dict(c["params"])
        pos = torch.tensor(c["pos_score"], dtype=torch.float32)
        neg = torch.tensor(c["corr_score"], dtype=torch.float32)
        per = ref_step.per_positive_loss(c["loss"], pos, neg.reshape(c["eta"], -1),
                                         margin=p.get("margin"), alpha=p.get("alpha"),
                                         reduction=p.get("reduction", "sum"))
        tot = ref_step.total_loss(c["loss"], pos, neg, c["eta"], margin=p.get("margin"),
                                  alpha=p.get("alpha"), reduction=p.get("reduction", "sum"))
        assert np.allclose(per.numpy(), c["expected_per_positive"], atol=c["tol"]), (c["source"], per)
        assert abs(float(tot) - sum(c["expected_per_positive"])) < c["tol"], c["source"]


def test_lookup_kat(kats):
    c = kats["lookup"]
    ent = np.array(c["ent_emb"], np.float32)
    rel = np.array(c["rel_emb"], np.float32)
    t = np.array(c["sample"], np.int32)
    L = c_oracle.lib()
    for col, table, exp in ((0, ent, c["expected_s_p_o"][0]), (1, rel, c["expected_s_p_o"][1]),
                            (2, ent, c["expected_s_p_o"][2])):
        ids = np.ascontiguousarray(t[:, col])
        out = np.empty((len(ids), 3), np.float32)
        L.kgeo_lookup(table.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(3), 3,
                      ids.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(len(ids)),
                      out.ctypes.data_as(C.POINTER(C.c_float)))
        assert (out == np.array(exp, np.float32)).all()


def test_corruption_structure_kat(kats):
    """The ids come from TF's RNG; what is pinned is the layout: derive the two
    random draws from the reference's expected output and check that the
    restated arithmetic (CorruptionGenerationLayerTrain.py:52-94) reproduces it."""
    c = kats["corruption"]
    pos = np.array(c["pos"], np.int32)
    exp = np.array(c["expected_with_tf_seed_0"], np.int32)
    B, eta = len(pos), c["eta"]
    keep, repl = np.zeros(B * eta, np.uint8), np.zeros(B * eta, np.int32)
    for r in range(B * eta):
        i = r % B  # tile order: row j*B+i
        assert exp[r, 1] == pos[i, 1]
        if exp[r, 0] == pos[i, 0]:
            keep[r], repl[r] = 1, exp[r, 2]
        else:
            assert exp[r, 2] == pos[i, 2]
            keep[r], repl[r] = 0, exp[r, 0]
    assert (repl < c["ent_size"]).all()
    assert (c_oracle.corrupt(pos, eta, keep, repl) == exp).all()
    got = ref_step.corrupt(torch.tensor(pos, dtype=torch.long), eta, torch.tensor(keep),
                           torch.tensor(repl, dtype=torch.long))
    assert (got.numpy() == exp).all()


def test_canonical_sincos_close_to_libm():
    x = np.concatenate([np.linspace(-20, 20, 20001),
                        np.random.default_rng(0).normal(0, 50, 5000)]).astype(np.float32)
    s, c = c_oracle.sincos(x)
    rs, rc = np.sin(x.astype(np.float64)), np.cos(x.astype(np.float64))
    ulp_s = np.spacing(np.maximum(np.abs(rs), 1e-30).astype(np.float32))
    ulp_c = np.spacing(np.maximum(np.abs(rc), 1e-30).astype(np.float32))
    assert np.max(np.abs(s - rs) / ulp_s) <= 1.0
    assert np.max(np.abs(c - rc) / ulp_c) <= 1.0


@pytest.mark.parametrize("model", MODELS)
def test_c_oracle_matches_torch_oracle_random(model):
    """The two restatements (canonical-order C, broadcast torch) agree to fp32 noise."""
    rng = np.random.default_rng(3)
    k = 16
    K = k if model in ("TransE", "DistMult") else 2 * k
    E, R, b = 50, 5, 12
    ent = rng.uniform(-0.5, 0.5, (E, K)).astype(np.float32)
    rel = rng.uniform(-0.5, 0.5, (R, K)).astype(np.float32)
    t = np.stack([rng.integers(0, E, b), rng.integers(0, R, b), rng.integers(0, E, b)], 1).astype(np.int32)
    sc = c_oracle.score_triples(model, ent, rel, t)
    te, tr = torch.tensor(ent), torch.tensor(rel)
    tt = torch.tensor(t, dtype=torch.long)
    e_s, e_p, e_o = te[tt[:, 0]], tr[tt[:, 1]], te[tt[:, 2]]
    ref = ref_step.compute_scores(model, e_s, e_p, e_o, R).numpy()
    assert np.allclose(sc, ref, rtol=1e-5, atol=1e-5)
    for side in ("s", "o"):
        m = c_oracle.corruption_scores(model, side, e_s.numpy(), e_p.numpy(), e_o.numpy(), ent, R)
        mr = ref_step.corruption_scores(model, side, e_s, e_p, e_o, te, R).numpy()
        assert np.allclose(m, mr, rtol=1e-5, atol=1e-5)


def test_rank_invariants_random():
    """worst >= middle >= best; filtered <= unfiltered; entity shards are additive
    (ScoringBasedEmbeddingModel.py:1449-1452)."""
    rng = np.random.default_rng(5)
    E, R, K, b = 40, 4, 8, 10
    ent = np.round(rng.uniform(-1, 1, (E, K)), 1).astype(np.float32)  # coarse values -> ties
    rel = np.round(rng.uniform(-1, 1, (R, K)), 1).astype(np.float32)
    t = np.stack([rng.integers(0, E, b), rng.integers(0, R, b), rng.integers(0, E, b)], 1).astype(np.int32)
    filt = [sorted(set(rng.integers(0, E, rng.integers(0, 6)).tolist())) for _ in range(b)]
    for model in MODELS[:2]:
        for side in ("s", "o"):
            w = c_oracle.rank_triples(model, side, "worst", ent, rel, t)
            m = c_oracle.rank_triples(model, side, "middle", ent, rel, t)
            be = c_oracle.rank_triples(model, side, "best", ent, rel, t)
            assert (w >= m).all() and (m >= be).all()
            wf = c_oracle.rank_triples(model, side, "worst", ent, rel, t, filters=filt)
            assert (wf <= w).all()
            h = E // 2
            w0 = c_oracle.rank_triples(model, side, "worst", ent, rel, t, filters=filt, start_id=0, n_cand=h)
            w1 = c_oracle.rank_triples(model, side, "worst", ent, rel, t, filters=filt, start_id=h, n_cand=E - h)
            assert (w0 + w1 == wf).all()


def test_legacy_adam_matches_closed_form():
    opt = ref_step.LegacyOptimizer("adam", 0.001)
    var = torch.zeros(3)
    g = torch.tensor([1.0, -2.0, 0.0])
    opt.step({"v": (var, g)})
    assert np.allclose(var.numpy(), [-0.001, 0.001, 0.0], atol=1e-6)
    lr_t = 0.001 * math.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(var.numpy()[0], -lr_t * 0.1 / (math.sqrt(0.001) + 1e-7), rtol=1e-5)


def _closed_form_dscores(name, P, N, margin, alpha, reduction):
    """SURVEY.md 8(a) "Gradients the fused kernel must reproduce", stated in numpy (float64), independently of autograd.
    P [B], N [eta, B] -> (dP [B], dN [eta, B])."""
    eta = N.shape[0]
    w = 1.0 if reduction == "sum" else 1.0 / eta
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    if name == "pairwise":
        dN = w * ((margin - P + N) >= 0)
        return -dN.sum(0), dN
    if name == "nll":
        w2 = 1.0 if reduction == "sum" else 1.0 / (2 * eta)
        inN, inP = (N >= -75) & (N <= 75), (P >= -75) & (P <= 75)
        return -w2 * eta * sig(-P) * inP, w2 * sig(N) * inN
    if name == "absolute_margin":
        return np.full_like(P, -w * eta), w * ((margin + N) >= 0)
    if name == "self_adversarial":
        e = np.exp(alpha * N - (alpha * N).max(0))
        p = e / e.sum(0)
        l = -np.log1p(np.exp(N + margin))  # log sigmoid(-N - margin)
        S = (p * l).sum(0)
        return -sig(-(margin + P)), w * (p * sig(N + margin) - alpha * p * (l - S))
    if name == "multiclass_nll":
        inN, inP = (N >= -75) & (N <= 75), (P >= -75) & (P <= 75)
        Nc, Pc = np.clip(N, -75, 75), np.clip(P, -75, 75)
        D = w * np.exp(Nc).sum(0) + np.exp(Pc)
        return (np.exp(Pc) / D - 1.0) * inP, w * np.exp(Nc) / D * inN
    raise AssertionError(name)


@pytest.mark.parametrize("reduction", ["sum", "mean"])
@pytest.mark.parametrize("name", ["pairwise", "nll", "absolute_margin", "self_adversarial", "multiclass_nll"])
def test_closed_form_loss_gradients_match_autograd(name, reduction):
    """The dL/dscore formulas the CUDA kernel implements == autograd through the op-for-op loss
    restatement that the reference's loss KATs pin (test_loss_kats)."""
    rng = np.random.default_rng(5)
    eta, B = 7, 33
    P = rng.normal(0, 3, B)
    N = rng.normal(0, 3, (eta, B))
    P[0], N[0, 1], N[1, 2] = 80.0, -90.0, 76.0  # outside the +-75 clip of nll / multiclass_nll
    margin, alpha = {"pairwise": 1.5, "absolute_margin": 0.7, "self_adversarial": 3.0}.get(name, 1.0), 0.5
    tp = torch.tensor(P, dtype=torch.float64, requires_grad=True)
    tn = torch.tensor(N, dtype=torch.float64, requires_grad=True)
    kw = {"reduction": reduction}
    if name in ("pairwise", "absolute_margin", "self_adversarial"):
        kw["margin"] = margin
    if name == "self_adversarial":
        kw["alpha"] = alpha
    ref_step.per_positive_loss(name, tp, tn, **kw).sum().backward()
    dP, dN = _closed_form_dscores(name, P, N, margin, alpha, reduction)
    assert np.allclose(tp.grad.numpy(), dP, rtol=1e-9, atol=1e-12)
    assert np.allclose(tn.grad.numpy(), dN, rtol=1e-9, atol=1e-12)


def test_legacy_sgd_and_adagrad_closed_form():
    """tf.keras.optimizers.legacy SGD (with momentum) and Adagrad (initial accumulator 0.1, eps 1e-7) update rules."""
    g = torch.tensor([1.0, -2.0, 0.5])
    var = torch.zeros(3)
    opt = ref_step.LegacyOptimizer("sgd", 0.1, momentum=0.9)
    opt.step({"v": (var, g)})
    opt.step({"v": (var, g)})
    # accum_1 = -lr g ; accum_2 = 0.9 accum_1 - lr g ; var = accum_1 + accum_2
    assert np.allclose(var.numpy(), (-0.1 * g - (0.9 * 0.1 + 0.1) * g).numpy(), rtol=1e-6)
    var = torch.zeros(3)
    opt = ref_step.LegacyOptimizer("adagrad", 0.05)
    opt.step({"v": (var, g)})
    want = -0.05 * g.numpy() / (np.sqrt(0.1 + g.numpy() ** 2) + 1e-7)
    assert np.allclose(var.numpy(), want, rtol=1e-6)


RANDOM123_PHILOX4X32_10_KATS = [  # Random123 kat_vectors: (counter, key) -> output
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers_oracle_and_library():
    """The corruption RNG is the published Philox4x32-10: both the oracle's restatement and the library's
    host entry point reproduce the Random123 known-answer vectors."""
    from oracle import philox
    from ampligraph_b200 import _lib
    lib = _lib.load()
    for ctr, key, want in RANDOM123_PHILOX4X32_10_KATS:
        assert philox.philox4x32_10(ctr, key) == want
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        lib.kge_philox4x32_10(c, k, o)
        assert tuple(o) == want


def test_host_corruption_replay_matches_oracle_stream():
    """kge_host_corruptions (the CPU replay of what the fused kernel draws) == the oracle's statement of the stream,
    and has the structure of CorruptionGenerationLayerTrain.py:52-88."""
    from oracle import philox
    from ampligraph_b200 import _lib
    rng = np.random.default_rng(3)
    E, B, eta = 1000, 37, 5
    t = np.stack([rng.integers(0, E, B), rng.integers(0, 7, B), rng.integers(0, E, B)], 1).astype(np.int32)
    for seed, step in ((0, 0), (5, 3), (2 ** 40 + 17, 2 ** 33 + 1)):
        got = _lib.host_corruptions(t, eta, E, seed=seed, step=step)
        assert (got == philox.corruption_stream(t, eta, E, seed, step)).all()
        tiled = np.tile(t, (eta, 1))
        assert (got[:, 1] == tiled[:, 1]).all()
        assert ((got[:, 0] == tiled[:, 0]) | (got[:, 2] == tiled[:, 2])).all()
        assert got.min() >= 0 and got[:, [0, 2]].max() < E
    assert (_lib.host_corruptions(t, eta, E, seed=5, step=3) != _lib.host_corruptions(t, eta, E, seed=5, step=4)).any()
    with pytest.raises(ValueError):
        _lib.host_corruptions(t, 0, E)
