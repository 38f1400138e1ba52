"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle.

Run on the B200 box with `pytest -m gpu`.  Tolerances: fp32 scores / gradients /
losses within 1e-4 relative (BASELINE.json north_star); ranks bit-exact.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MODELS = ["TransE", "DistMult", "ComplEx", "HolE", "RotatE"]
LOSSES = ["pairwise", "nll", "absolute_margin", "self_adversarial", "multiclass_nll"]
RTOL = 1e-4


def _engine(*a, **kw):
    from ampligraph_b200.engine import KGEEngine
    return KGEEngine(*a, **kw)


def _tables(model, E, R, k, rng, scale=None):
    K = k if model in ("TransE", "DistMult") else 2 * k
    lim_e = scale or np.sqrt(6.0 / (E + K))
    lim_r = scale or np.sqrt(6.0 / (R + K))
    ent = rng.uniform(-lim_e, lim_e, (E, K)).astype(np.float32)
    rel = rng.uniform(-lim_r, lim_r, (R, K)).astype(np.float32)
    return ent, rel


def _triples(E, R, n, rng):
    return np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)


def _negatives(E, B, eta, rng):
    return rng.integers(0, E, B * eta).astype(np.int32), rng.integers(0, 2, B * eta).astype(np.uint8)


def _dev(a):
    return torch.as_tensor(a).cuda().contiguous()


def _dense(eng, table):
    """padded [rows, ld] -> dense [rows, internal_k] (host-side slicing, independent of kge_unpack_rows)."""
    t = table.detach().cpu().numpy()
    k, kp = eng.k, eng.kp
    if eng.internal_k == k:
        return t[:, :k].copy()
    return np.concatenate([t[:, :k], t[:, kp:kp + k]], axis=1)


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() <= rtol * scale + 1e-7


# ---------------------------------------------------------------------------
def test_library_loads_and_layout():
    from ampligraph_b200 import _lib
    lib = _lib.load()
    assert lib.kge_abi_version() == 1
    for model, k, ld in (("TransE", 50, 52), ("DistMult", 400, 400), ("ComplEx", 3, 8), ("RotatE", 200, 400)):
        eng = _engine(model, k, 2, 10, 3)
        assert eng.ld == ld and eng.internal_k == (k if model in ("TransE", "DistMult") else 2 * k)
        eng.close()


@pytest.mark.parametrize("model", MODELS)
def test_pack_unpack_roundtrip(model):
    rng = np.random.default_rng(0)
    eng = _engine(model, 7, 2, 33, 5)
    ent, rel = _tables(model, 33, 5, 7, rng)
    eng.set_embeddings(ent, rel)
    assert (_dense(eng, eng.ent) == ent).all() and (_dense(eng, eng.rel) == rel).all()
    e2, r2 = eng.get_embeddings()
    assert (e2.cpu().numpy() == ent).all() and (r2.cpu().numpy() == rel).all()
    pad = eng.ent.cpu().numpy()[:, 7:8]
    assert (pad == 0).all()


def test_reference_scoring_kats_on_gpu(kats):
    """The reference's own golden vectors, through kge_score_triples."""
    from oracle import c_oracle  # checker only
    for c in kats["scoring"]:
        if c["form"] != "triple":
            continue
        e_s, e_p, e_o = (np.array(c[x], np.float32) for x in ("e_s", "e_p", "e_o"))
        n = len(e_s)
        R = c["max_rel_size"] or n
        rel = np.zeros((max(R, n), e_p.shape[1]), np.float32)
        rel[:n] = e_p
        eng = _engine(c["model"], c["k"], 1, 2 * n, max(R, n))
        # RotatE's phase normalisation uses n_rel == max_rel_size of the KAT
        assert eng.n_rel == (c["max_rel_size"] or n)
        eng.set_embeddings(np.concatenate([e_s, e_o]), rel[:eng.n_rel])
        t = np.stack([np.arange(n), np.arange(n), n + np.arange(n)], 1).astype(np.int32)
        got = np.around(eng.score(_dev(t)).cpu().numpy(), c["round_decimals"])
        assert (got == np.array(c["expected"], np.float32)).all(), (c["source"], got)
        eng.close()


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("k", [5, 50, 200])
def test_score_triples_vs_oracle(model, k):
    from oracle import c_oracle
    rng = np.random.default_rng(1)
    E, R, n = 300, 7, 1000
    ent, rel = _tables(model, E, R, k, rng, scale=0.5)
    t = _triples(E, R, n, rng)
    eng = _engine(model, k, 2, E, R)
    eng.set_embeddings(ent, rel)
    got = eng.score(_dev(t)).cpu().numpy()
    ref = c_oracle.score_triples(model, ent, rel, t)
    assert np.allclose(got, ref, rtol=RTOL, atol=1e-5), np.abs(got - ref).max()
    eng.close()


# ---------------------------------------------------------------------------
def _ref(model, k, ent, rel, eta, loss, lp, **kw):
    from oracle import ref_step
    p = dict(lp)
    return ref_step.RefStep(model, ent.shape[1], ent, rel, eta, loss=loss,
                            loss_params={k_: v for k_, v in p.items()}, **kw)


def _corruption_tensor(t, neg_ent, neg_keep, eta):
    from oracle import c_oracle
    return c_oracle.corrupt(t, eta, neg_keep, neg_ent)


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("loss", LOSSES)
@pytest.mark.parametrize("reduction", ["sum", "mean"])
def test_forward_backward_vs_oracle(model, loss, reduction):
    rng = np.random.default_rng(hash((model, loss, reduction)) % 2**32)
    E, R, k, eta, B = 64, 5, 12, 5, 97
    ent, rel = _tables(model, E, R, k, rng, scale=0.6)
    t = _triples(E, R, B, rng)
    neg_ent, neg_keep = _negatives(E, B, eta, rng)
    lp = {"reduction": reduction}
    if loss in ("pairwise", "absolute_margin"):
        lp["margin"] = 0.5
    if loss == "self_adversarial":
        lp.update(margin=2.0, alpha=0.7)
    eng = _engine(model, k, eta, E, R, loss=loss, loss_params=lp)
    eng.set_embeddings(ent, rel)
    sp = torch.empty(B, device="cuda")
    sn = torch.empty(B * eta, device="cuda")
    eng.forward_backward(_dev(t), (_dev(neg_ent), _dev(neg_keep)), scores_pos=sp, scores_neg=sn)
    rs = _ref(model, k, ent, rel, eta, loss, lp)
    rl, rsp, rsn, g_ent, g_rel = rs.loss_and_grads(t, _corruption_tensor(t, neg_ent, neg_keep, eta))
    assert np.allclose(sp.cpu().numpy(), rsp.numpy(), rtol=RTOL, atol=1e-5)
    assert np.allclose(sn.cpu().numpy(), rsn.numpy(), rtol=RTOL, atol=1e-5)
    assert abs(eng.read_loss() - float(rl)) <= RTOL * abs(float(rl)) + 1e-5
    assert _close(_dense(eng, eng.g_ent), g_ent.numpy()), np.abs(_dense(eng, eng.g_ent) - g_ent.numpy()).max()
    assert _close(_dense(eng, eng.g_rel), g_rel.numpy()), np.abs(_dense(eng, eng.g_rel) - g_rel.numpy()).max()
    eng.close()


@pytest.mark.parametrize("model,k,eta,B", [("ComplEx", 200, 10, 300), ("DistMult", 400, 20, 200),
                                           ("RotatE", 200, 30, 150), ("TransE", 50, 2, 1000),
                                           ("HolE", 130, 7, 123)])
def test_forward_backward_baseline_shapes(model, k, eta, B):
    """BASELINE.json config shapes (k, eta) on a reduced KG."""
    rng = np.random.default_rng(7)
    E, R = 500, 11
    loss = {"ComplEx": "self_adversarial", "RotatE": "self_adversarial"}.get(model, "pairwise")
    ent, rel = _tables(model, E, R, k, rng)
    t = _triples(E, R, B, rng)
    neg_ent, neg_keep = _negatives(E, B, eta, rng)
    eng = _engine(model, k, eta, E, R, loss=loss)
    eng.set_embeddings(ent, rel)
    eng.forward_backward(_dev(t), (_dev(neg_ent), _dev(neg_keep)))
    rs = _ref(model, k, ent, rel, eta, loss, {})
    rl, _, _, g_ent, g_rel = rs.loss_and_grads(t, _corruption_tensor(t, neg_ent, neg_keep, eta))
    assert abs(eng.read_loss() - float(rl)) <= RTOL * abs(float(rl))
    assert _close(_dense(eng, eng.g_ent), g_ent.numpy())
    assert _close(_dense(eng, eng.g_rel), g_rel.numpy())
    eng.close()


@pytest.mark.parametrize("model,loss,eta,B", [("ComplEx", "self_adversarial", 40, 37), ("TransE", "multiclass_nll", 70, 9),
                                              ("RotatE", "self_adversarial", 33, 21), ("DistMult", "nll", 1, 1),
                                              ("HolE", "pairwise", 1, 300)])
def test_edge_shapes_eta_over_32_and_tiny_batches(model, loss, eta, B):
    """eta > 32 (several corruptions per lane in the loss, more than one ballot word), eta = 1, B = 1."""
    rng = np.random.default_rng(47)
    E, R, k = 90, 3, 10
    ent, rel = _tables(model, E, R, k, rng, scale=0.4)
    t = _triples(E, R, B, rng)
    neg_ent, neg_keep = _negatives(E, B, eta, rng)
    eng = _engine(model, k, eta, E, R, loss=loss)
    eng.set_embeddings(ent, rel)
    sn = torch.empty(B * eta, device="cuda")
    sp = torch.empty(B, device="cuda")
    eng.forward_backward(_dev(t), (_dev(neg_ent), _dev(neg_keep)), scores_pos=sp, scores_neg=sn)
    rs = _ref(model, k, ent, rel, eta, loss, {})
    rl, rsp, rsn, g_ent, g_rel = rs.loss_and_grads(t, _corruption_tensor(t, neg_ent, neg_keep, eta))
    assert np.allclose(sn.cpu().numpy(), rsn.numpy(), rtol=RTOL, atol=1e-5)
    assert abs(eng.read_loss() - float(rl)) <= RTOL * abs(float(rl)) + 1e-5
    assert _close(_dense(eng, eng.g_ent), g_ent.numpy()) and _close(_dense(eng, eng.g_rel), g_rel.numpy())
    # an empty batch is a no-op
    eng.g_ent.zero_(); eng.g_rel.zero_()
    eng.forward_backward(torch.empty((0, 3), dtype=torch.int32, device="cuda"), None)
    assert eng.read_loss() == 0 and (eng.g_ent == 0).all()
    assert eng.score(torch.empty((0, 3), dtype=torch.int32, device="cuda")).numel() == 0
    assert eng.rank(torch.empty((0, 3), dtype=torch.int32, device="cuda"), "s").numel() == 0
    eng.close()


@pytest.mark.parametrize("model,k", [("TransE", 600), ("DistMult", 1100), ("ComplEx", 520), ("HolE", 1000), ("RotatE", 700)])
@pytest.mark.parametrize("group", [0, 3])
def test_wide_rows_column_windows(model, k, group):
    """half-rows wider than 512 floats are processed in column windows (cfg5 has k=1000), with and
    without negative groups; compared with the oracle."""
    rng = np.random.default_rng(41)
    E, R, eta, B = 60, 4, 7, 41
    loss = "self_adversarial" if model in ("ComplEx", "RotatE") else "nll"
    ent, rel = _tables(model, E, R, k, rng, scale=0.05)
    t = _triples(E, R, B, rng)
    neg_ent, neg_keep = _negatives(E, B, eta, rng)
    eng = _engine(model, k, eta, E, R, loss=loss, neg_group=group)
    eng.set_embeddings(ent, rel)
    sp = torch.empty(B, device="cuda")
    sn = torch.empty(B * eta, device="cuda")
    eng.forward_backward(_dev(t), (_dev(neg_ent), _dev(neg_keep)), scores_pos=sp, scores_neg=sn)
    rs = _ref(model, k, ent, rel, eta, loss, {})
    rl, rsp, rsn, g_ent, g_rel = rs.loss_and_grads(t, _corruption_tensor(t, neg_ent, neg_keep, eta))
    assert np.allclose(sp.cpu().numpy(), rsp.numpy(), rtol=RTOL, atol=1e-5)
    assert np.allclose(sn.cpu().numpy(), rsn.numpy(), rtol=RTOL, atol=1e-5)
    assert abs(eng.read_loss() - float(rl)) <= RTOL * abs(float(rl)) + 1e-5
    assert _close(_dense(eng, eng.g_ent), g_ent.numpy()) and _close(_dense(eng, eng.g_rel), g_rel.numpy())
    eng.close()


@pytest.mark.parametrize("model", ["ComplEx", "TransE", "RotatE"])
@pytest.mark.parametrize("group", [1, 3])
def test_negative_groups_match_resident(model, group):
    """eta negatives processed in groups of G (two-pass path) == all-resident path."""
    rng = np.random.default_rng(11)
    E, R, k, eta, B = 80, 4, 20, 7, 64
    ent, rel = _tables(model, E, R, k, rng, scale=0.5)
    t = _triples(E, R, B, rng)
    neg = _negatives(E, B, eta, rng)
    outs = []
    for g in (0, group):
        eng = _engine(model, k, eta, E, R, loss="self_adversarial", neg_group=g)
        eng.set_embeddings(ent, rel)
        eng.forward_backward(_dev(t), (_dev(neg[0]), _dev(neg[1])))
        outs.append((_dense(eng, eng.g_ent), _dense(eng, eng.g_rel), eng.read_loss()))
        eng.close()
    assert _close(outs[1][0], outs[0][0]) and _close(outs[1][1], outs[0][1])
    assert abs(outs[1][2] - outs[0][2]) <= RTOL * abs(outs[0][2])


@pytest.mark.parametrize("model,k,group", [("ComplEx", 20, 3), ("RotatE", 24, 2), ("TransE", 600, 3), ("ComplEx", 520, 0),
                                           ("DistMult", 1100, 1)])
def test_row_stash_matches_second_gather(model, k, group):
    """kge_set_row_stash: the gradient pass re-reads the replaced rows from the stash the score pass filled; same
    gradients as gathering them twice, and the stash holds exactly the gathered table rows."""
    rng = np.random.default_rng(23)
    E, R, eta, B = 90, 4, 7, 70
    ent, rel = _tables(model, E, R, k, rng, scale=0.3 if k < 100 else 0.05)
    t = _triples(E, R, B, rng)
    neg = _negatives(E, B, eta, rng)
    outs = []
    for stash in (False, True):
        eng = _engine(model, k, eta, E, R, loss="self_adversarial", neg_group=group)
        eng.set_embeddings(ent, rel)
        if stash:
            st = eng.ensure_row_stash(B)
            assert st is not None, "expected a non-resident geometry"
            st.fill_(float("nan"))
        eng.forward_backward(_dev(t), (_dev(neg[0]), _dev(neg[1])))
        outs.append((_dense(eng, eng.g_ent), _dense(eng, eng.g_rel), eng.read_loss()))
        if stash:
            torch.cuda.synchronize()
            want = eng.ent[_dev(neg[0]).long().view(eta, B).t().reshape(-1)]  # stash row i*eta+j <- tile row j*B+i
            assert torch.equal(st[:B * eta], want)
        eng.close()
    assert _close(outs[1][0], outs[0][0]) and _close(outs[1][1], outs[0][1])
    assert abs(outs[1][2] - outs[0][2]) <= RTOL * abs(outs[0][2])


def test_philox_corruptions_structure_and_parity():
    """In-kernel Philox negatives: structure of A3 + the fused kernel really uses that stream."""
    rng = np.random.default_rng(13)
    E, R, k, eta, B = 1000, 10, 16, 4, 500
    ent, rel = _tables("DistMult", E, R, k, rng, scale=0.5)
    t = _triples(E, R, B, rng)
    eng = _engine("DistMult", k, eta, E, R, loss="nll")
    eng.set_embeddings(ent, rel)
    corr = eng.generate_corruptions(_dev(t), seed=5, step=3).cpu().numpy()
    assert corr.shape == (B * eta, 3)
    tiled = np.tile(t, (eta, 1))
    assert (corr[:, 1] == tiled[:, 1]).all()  # relation kept
    s_same, o_same = corr[:, 0] == tiled[:, 0], corr[:, 2] == tiled[:, 2]
    assert (s_same | o_same).all()  # at most one side replaced
    assert corr.min() >= 0 and corr.max() < E
    frac_obj = (s_same & ~o_same).mean()
    assert 0.4 < frac_obj < 0.6  # keep_subj ~ Bernoulli(1/2)
    repl = np.where(~s_same, corr[:, 0], corr[:, 2])
    assert abs(repl.mean() - (E - 1) / 2) < 0.05 * E  # uniform replacement ids
    corr2 = eng.generate_corruptions(_dev(t), seed=5, step=4).cpu().numpy()
    assert (corr2 != corr).any()  # the stream advances with the step
    # the device draw is the published Philox4x32-10 stream: bit-identical to the library's CPU replay and to the
    # oracle's independent statement of it (oracle/philox.py, pinned by the Random123 known answers)
    from ampligraph_b200 import _lib
    from oracle import philox
    assert (corr == _lib.host_corruptions(t, eta, E, seed=5, step=3)).all()
    assert (corr[:64] == philox.corruption_stream(t, eta, E, 5, 3)[:64]).all()
    # gradients of the Philox path == oracle fed with the materialised corruptions
    eng.forward_backward(_dev(t), None, seed=5, step=3)
    rs = _ref("DistMult", k, ent, rel, eta, "nll", {})
    rl, _, _, g_ent, g_rel = rs.loss_and_grads(t, corr)
    assert abs(eng.read_loss() - float(rl)) <= RTOL * abs(float(rl))
    assert _close(_dense(eng, eng.g_ent), g_ent.numpy()) and _close(_dense(eng, eng.g_rel), g_rel.numpy())
    eng.close()


@pytest.mark.parametrize("opt,params", [("adam", {}), ("adam", {"learning_rate": 0.01, "beta_1": 0.8}),
                                        ("sgd", {"learning_rate": 0.05}), ("sgd", {"learning_rate": 0.05, "momentum": 0.9}),
                                        ("adagrad", {"learning_rate": 0.1})])
@pytest.mark.parametrize("reg", [None, {"p": 2, "lambda": 1e-3}, {"p": 3, "lambda": 1e-2}])
def test_train_steps_vs_oracle(opt, params, reg):
    """several full steps (fwd+bwd+optimizer+regulariser) track the op-for-op restatement."""
    rng = np.random.default_rng(17)
    model, E, R, k, eta, B = "ComplEx", 50, 4, 10, 3, 40
    ent, rel = _tables(model, E, R, k, rng, scale=0.5)
    eng = _engine(model, k, eta, E, R, loss="multiclass_nll", optimizer=opt, optimizer_params=params, regularizer=reg)
    eng.set_embeddings(ent, rel)
    rs = _ref(model, k, ent, rel, eta, "multiclass_nll", {}, optimizer=opt, optimizer_params=dict(params),
              regularizer=({"p": reg["p"], "lam": reg["lambda"]} if reg else None))
    for step in range(5):
        t = _triples(E, R, B, rng)
        neg_ent, neg_keep = _negatives(E, B, eta, rng)
        eng.train_step(_dev(t), (_dev(neg_ent), _dev(neg_keep)))
        ref_loss = rs.train_step(t, _corruption_tensor(t, neg_ent, neg_keep, eta))
        got_loss = eng.read_loss()
        assert abs(got_loss - ref_loss) <= 2e-4 * abs(ref_loss), (step, got_loss, ref_loss)
        assert (eng.g_ent == 0).all() and (eng.g_rel == 0).all()  # accumulators cleared for the next step
    assert _close(_dense(eng, eng.ent), rs.ent.detach().numpy(), rtol=5e-4)
    assert _close(_dense(eng, eng.rel), rs.rel.detach().numpy(), rtol=5e-4)
    eng.close()


@pytest.mark.parametrize("opt", ["lazy_adam", "lazy_adagrad", "lazy_sgd"])
def test_lazy_optimizer_matches_restatement(opt):
    """opt-in lazy optimizer (rows touched by the step only) vs the oracle's lazy restatement; untouched
    rows must not move at all."""
    rng = np.random.default_rng(43)
    model, E, R, k, eta, B = "DistMult", 400, 6, 12, 3, 25   # few positives: most rows stay untouched
    ent, rel = _tables(model, E, R, k, rng, scale=0.5)
    eng = _engine(model, k, eta, E, R, loss="nll", optimizer=opt, optimizer_params={"learning_rate": 0.01})
    eng.set_embeddings(ent, rel)
    rs = _ref(model, k, ent, rel, eta, "nll", {}, optimizer=opt, optimizer_params={"learning_rate": 0.01})
    touched = np.zeros(E, bool)
    for step in range(4):
        t = _triples(E, R, B, rng)
        neg_ent, neg_keep = _negatives(E, B, eta, rng)
        eng.train_step(_dev(t), (_dev(neg_ent), _dev(neg_keep)), step=step)
        rs.train_step(t, _corruption_tensor(t, neg_ent, neg_keep, eta))
        touched[t[:, 0]] = touched[t[:, 2]] = touched[neg_ent] = True
        assert (eng.g_ent == 0).all()
    got = _dense(eng, eng.ent)
    assert _close(got, rs.ent.detach().numpy(), rtol=5e-4) and _close(_dense(eng, eng.rel), rs.rel.detach().numpy(), rtol=5e-4)
    assert (~touched).sum() > 50 and (got[~touched] == ent[~touched]).all()
    eng.close()


def test_external_loss_two_phase():
    """LossFunctionWrapper path (loss_functions.py:657): scores out, dL/dscore back in."""
    rng = np.random.default_rng(19)
    model, E, R, k, eta, B = "HolE", 40, 3, 9, 4, 33
    ent, rel = _tables(model, E, R, k, rng, scale=0.5)
    t = _triples(E, R, B, rng)
    neg_ent, neg_keep = _negatives(E, B, eta, rng)
    eng = _engine(model, k, eta, E, R)
    eng.set_embeddings(ent, rel)
    from ampligraph_b200 import _lib
    sp = torch.empty(B, device="cuda")
    sn = torch.empty(B * eta, device="cuda")
    negs = (_dev(neg_ent), _dev(neg_keep))
    eng.forward_backward(_dev(t), negs, mode=_lib.STEP_FORWARD_ONLY, scores_pos=sp, scores_neg=sn)
    assert (eng.g_ent == 0).all()

    def user_loss(scores_pos, scores_neg):  # the docstring example of the reference (:660-668)
        neg_exp, pos_exp = torch.exp(scores_neg), torch.exp(scores_pos)
        return -torch.log(pos_exp / (neg_exp.sum(0) + pos_exp))

    spg, sng = sp.clone().requires_grad_(True), sn.clone().requires_grad_(True)
    user_loss(spg, sng.reshape(eta, -1)).sum().backward()
    eng.forward_backward(_dev(t), negs, mode=_lib.STEP_BACKWARD_EXT, dpos=spg.grad.contiguous(), dneg=sng.grad.contiguous())
    from oracle import ref_step
    rs = _ref(model, k, ent, rel, eta, "pairwise", {})
    rsp, rsn = rs.forward(t, _corruption_tensor(t, neg_ent, neg_keep, eta))
    user_loss(rsp, rsn.reshape(eta, -1)).sum().backward()
    assert _close(_dense(eng, eng.g_ent), rs.ent.grad.numpy()) and _close(_dense(eng, eng.g_rel), rs.rel.grad.numpy())
    eng.close()


def test_gradient_linearity_full_size():
    """cfg2 shape at full size (14.5k entities, B=27,212): grads(A u B) == grads(A) + grads(B),
    loss additive -- a size-independent property the oracle cannot check in seconds."""
    rng = np.random.default_rng(23)
    model, E, R, k, eta, B = "ComplEx", 14505, 237, 200, 10, 27212
    eng = _engine(model, k, eta, E, R, loss="self_adversarial")
    eng.init_glorot_uniform(seed=1)
    t = _dev(_triples(E, R, B, rng))
    neg_ent, neg_keep = _negatives(E, B, eta, rng)
    ne, nk = _dev(neg_ent), _dev(neg_keep)
    eng.forward_backward(t, (ne, nk))
    g_all, gr_all, loss_all = eng.g_ent.clone(), eng.g_rel.clone(), eng.read_loss()
    eng.g_ent.zero_(); eng.g_rel.zero_()
    h = B // 2
    idx = torch.arange(B, device="cuda")
    for sel in (idx[:h], idx[h:]):
        rows = (torch.arange(eta, device="cuda")[:, None] * B + sel[None, :]).reshape(-1)
        eng.forward_backward(t[sel].contiguous(), (ne[rows].contiguous(), nk[rows].contiguous()))
    loss_parts = eng.read_loss()
    assert abs(loss_parts - loss_all) <= 1e-5 * abs(loss_all)
    scale = g_all.abs().max().item()
    assert (eng.g_ent - g_all).abs().max().item() <= 1e-4 * scale
    assert (eng.g_rel - gr_all).abs().max().item() <= 1e-4 * gr_all.abs().max().item()
    assert (eng.g_ent[:, eng.k:eng.kp] == 0).all()
    eng.close()


# ---------------------------------------------------------------------------
def test_reference_rank_kats_on_gpu(kats):
    r = kats["ranks"]
    e_s, e_p, e_o = (np.array(r[x], np.float32) for x in ("e_s", "e_p", "e_o"))
    cand = np.array(r["ent_matrix"], np.float32)
    # entity table = candidates (ids 0..3) followed by the subject/object rows of the two test triples
    ent = np.concatenate([cand, e_s, e_o])
    n = len(e_s)
    t = np.stack([4 + np.arange(n), np.arange(n), 4 + n + np.arange(n)], 1).astype(np.int32)
    eng = _engine(r["model"], r["k"], 1, len(ent), n)
    eng.set_embeddings(ent, e_p)
    for case in r["cases"]:
        sides = [s for s in ("s", "o") if s in case["corrupt_side"]]
        got = []
        for j, side in enumerate(sides):
            off = idx = None
            if case["filters"]:
                f = case["filters"][j]
                off = _dev(np.concatenate([[0], np.cumsum([len(x) for x in f])]).astype(np.int64))
                idx = _dev(np.concatenate(f).astype(np.int32))
            got.append(eng.rank(_dev(t), side, case["comparison_type"], off, idx, cand_begin=0,
                                n_cand=case["end_ent_id"]).cpu().numpy())
        assert (np.array(got) == np.array(case["expected"], np.int32)).all(), (case, got)
    eng.close()


def _filters(E, b, rng, maxlen=8):
    lists = [sorted(set(rng.integers(0, E, rng.integers(0, maxlen)).tolist())) for _ in range(b)]
    off = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    idx = np.concatenate([np.asarray(x, np.int32) for x in lists]) if off[-1] else np.zeros(0, np.int32)
    return lists, off, idx.astype(np.int32)


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("k", [6, 50])
def test_ranks_bit_exact_vs_oracle(model, k):
    from oracle import c_oracle
    rng = np.random.default_rng(29)
    E, R, b = 700, 6, 70
    ent, rel = _tables(model, E, R, k, rng, scale=0.5)
    ent = np.round(ent, 2)  # coarse values -> plenty of quantisation ties
    t = _triples(E, R, b, rng)
    lists, off, idx = _filters(E, b, rng)
    eng = _engine(model, k, 1, E, R)
    eng.set_embeddings(ent, rel)
    for side in ("s", "o"):
        for strategy in ("worst", "best", "middle"):
            got = eng.rank(_dev(t), side, strategy).cpu().numpy()
            ref = c_oracle.rank_triples(model, side, strategy, ent, rel, t)
            assert (got == ref).all(), (model, side, strategy, np.flatnonzero(got != ref)[:5])
            gotf = eng.rank(_dev(t), side, strategy, _dev(off), _dev(idx)).cpu().numpy()
            reff = c_oracle.rank_triples(model, side, strategy, ent, rel, t, filters=lists)
            assert (gotf == reff).all(), (model, side, strategy, "filtered")
    eng.close()


@pytest.mark.parametrize("model", ["ComplEx", "RotatE", "TransE"])
def test_ranks_subset_and_shards(model):
    from oracle import c_oracle
    rng = np.random.default_rng(31)
    E, R, k, b = 1000, 5, 32, 40
    ent, rel = _tables(model, E, R, k, rng, scale=0.5)
    t = _triples(E, R, b, rng)
    eng = _engine(model, k, 1, E, R)
    eng.set_embeddings(ent, rel)
    # entities_subset (ScoringBasedEmbeddingModel.py:1634-1643): filter ids are subset positions
    subset = np.sort(rng.choice(E, 300, replace=False)).astype(np.int32)
    lists, off, idx = _filters(len(subset), b, rng)
    for side in ("s", "o"):
        got = eng.rank(_dev(t), side, "worst", _dev(off), _dev(idx), cand_ids=_dev(subset)).cpu().numpy()
        ref = c_oracle.rank_triples(model, side, "worst", ent, rel, t, filters=lists, cand_ids=subset)
        assert (got == ref).all()
    # row shards accumulate into the same output (ScoringBasedEmbeddingModel.py:1449-1452)
    lists, off, idx = _filters(E, b, rng)
    full = c_oracle.rank_triples(model, "o", "worst", ent, rel, t, filters=lists)
    out = torch.zeros(b, dtype=torch.int32, device="cuda")
    bounds = [0, 257, 600, E]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        eng.rank(_dev(t), "o", "worst", _dev(off), _dev(idx), cand_begin=lo, n_cand=hi - lo, out=out)
    assert (out.cpu().numpy() == full).all()
    eng.close()


def test_ranks_full_size_invariants():
    """cfg2 table size, 1,024 queries: worst >= middle >= best, filtered <= unfiltered,
    shard additivity, and a 64-query slice bit-exact against the oracle."""
    from oracle import c_oracle
    rng = np.random.default_rng(37)
    model, E, R, k, b = "ComplEx", 14505, 237, 200, 1024
    eng = _engine(model, k, 1, E, R)
    eng.init_glorot_uniform(seed=3)
    t = _triples(E, R, b, rng)
    lists, off, idx = _filters(E, b, rng, maxlen=40)
    td = _dev(t)
    w = eng.rank(td, "s", "worst").cpu().numpy()
    m = eng.rank(td, "s", "middle").cpu().numpy()
    be = eng.rank(td, "s", "best").cpu().numpy()
    wf = eng.rank(td, "s", "worst", _dev(off), _dev(idx)).cpu().numpy()
    assert (w >= m).all() and (m >= be).all() and (wf <= w).all() and (w >= 1).all()
    out = torch.zeros(b, dtype=torch.int32, device="cuda")
    for lo, hi in ((0, 5000), (5000, 9999), (9999, E)):
        eng.rank(td, "s", "worst", _dev(off), _dev(idx), cand_begin=lo, n_cand=hi - lo, out=out)
    assert (out.cpu().numpy() == wf).all()
    ent, rel = (x.cpu().numpy() for x in eng.get_embeddings())
    ref = c_oracle.rank_triples(model, "s", "worst", ent, rel, t[:64], filters=lists[:64])
    assert (wf[:64] == ref).all()
    eng.close()


def test_error_convention():
    from ampligraph_b200 import _lib
    with pytest.raises(ValueError):
        _engine("NoSuchModel", 4, 1, 10, 2)
    with pytest.raises(ValueError):
        _engine("TransE", 4, 1, 10, 2, loss="nope")
    eng = _engine("TransE", 4, 1, 10, 2)
    with pytest.raises(ValueError):  # invalid corrupt side id
        _lib.check(eng.lib.kge_rank(eng.h, 7, 0, None, None, None, 0, None, 0, 0, None, None, 0, None, None))
    with pytest.raises(ValueError):  # gradient buffers are mandatory outside FORWARD_ONLY
        _lib.check(eng.lib.kge_train_step(eng.h, 0, eng.ent.data_ptr(), eng.rel.data_ptr(), None, None,
                                          eng.ent.data_ptr(), 1, None, None, 0, 0, None, None, None, None, None, None))
    eng.close()
