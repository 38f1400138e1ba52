"""KGEEngine: device-memory owner + thin call layer over the C-ABI.

PyTorch tensors hold every byte of device memory (tables, gradient accumulators,
optimizer slots, batches, outputs); all arithmetic happens in libkge_b200.so.
This is the object the reference-facing facade
(ampligraph_b200.latent_features.ScoringBasedEmbeddingModel) drives from its
fit / predict / evaluate loops, mirroring train_function / predict_function /
test_function of the reference (models/ScoringBasedEmbeddingModel.py:443, :1719, :1387).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_NVTX = os.environ.get("KGE_B200_NVTX", "0") == "1"


def _traced(name):
    """KGE_B200_NVTX=1: wrap the call in an NVTX range, so that `ncu --nvtx --nvtx-include "kge.train_step/"` (or an
    nsys timeline) attributes kernels to the step phase that launched them.  Off (default): the method is untouched."""
    def deco(fn):
        if not _NVTX:
            return fn

        def wrapped(*a, **k):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


class KGEEngine:
    def __init__(self, scoring_type, k, eta, n_ent, n_rel, loss="pairwise", loss_params=None,
                 optimizer="adam", optimizer_params=None, regularizer=None, device=0, neg_group=0,
                 table_alloc=None, ent_rows=None, max_rel_size=None, rank_mode=None, rank_pair_cap=0):
        """regularizer: None | {"p":, "lambda": [, "p2":, "lambda2":]} | a pair [entities, relations] of those
        (EmbeddingLookupLayer.py:131-155).  max_rel_size: RotatE phase normalisation when it differs from the
        relation-table rows (RotatE.py:96).  rank_mode: 'auto' (tensor-core filter + exact refine for the bilinear
        models) | 'exact' (FP32 canonical chain for every pair); env KGE_B200_RANK_MODE overrides the default."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("ampligraph_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        if scoring_type not in _lib.SCORING:
            raise ValueError("Unknown scoring_type: %r" % (scoring_type,))
        if loss not in _lib.LOSSES:
            raise ValueError("Could not interpret loss identifier: %r" % (loss,))
        lp = dict(loss_params or {})
        reduction = lp.get("reduction", "sum")
        if reduction not in _lib.REDUCTIONS:
            raise AssertionError("Invalid value for reduction!")  # loss_functions.py:95-98
        default_margin = 3.0 if loss == "self_adversarial" else 1.0  # loss_functions.py:23,:29
        self.scoring_type, self.k, self.eta = scoring_type, int(k), int(eta)
        self.n_ent, self.n_rel = int(n_ent), int(n_rel)
        # rows of the LOCAL entity table: n_ent, or one row shard of it (parallel.ShardedTrainer)
        self.ent_rows = int(ent_rows) if ent_rows is not None else self.n_ent
        self.device = torch.device("cuda", device)
        cfg = _lib.KgeConfig(C.sizeof(_lib.KgeConfig), _lib.SCORING[scoring_type], int(k), int(eta), int(n_ent),
                             int(n_rel), _lib.LOSSES[loss], _lib.REDUCTIONS[reduction],
                             float(lp.get("margin", default_margin)), float(lp.get("alpha", 0.5)), int(device),
                             int(neg_group), int(max_rel_size or 0),
                             _lib.RANK_MODES[rank_mode or os.environ.get("KGE_B200_RANK_MODE", "auto")], int(rank_pair_cap))
        h = C.c_void_p()
        _lib.check(self.lib.kge_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.internal_k = self.lib.kge_internal_k(h)
        self.kp = self.lib.kge_half_stride(h)
        self.ld = self.lib.kge_row_stride(h)
        with torch.cuda.device(self.device):
            z = lambda rows: torch.zeros((rows, self.ld), dtype=torch.float32, device=self.device)
            if table_alloc is not None:  # e.g. symmetric (peer-mappable) memory for the multi-GPU path
                self.ent, self.rel, self.g_ent, self.g_rel = table_alloc(self.ent_rows, self.n_rel, self.ld, self.device)
                for t_ in (self.ent, self.rel, self.g_ent, self.g_rel):
                    assert t_.is_contiguous() and t_.dtype == torch.float32
            else:
                # one [ent | rel] block per kind (rows have the same width): the dense optimizer then updates both tables
                # in ONE launch (apply_gradients) -- one launch gap less per step
                self._pblock, self._gblock = z(self.ent_rows + self.n_rel), z(self.ent_rows + self.n_rel)
                self.ent, self.rel = self._pblock[:self.ent_rows], self._pblock[self.ent_rows:]
                self.g_ent, self.g_rel = self._gblock[:self.ent_rows], self._gblock[self.ent_rows:]
            self.loss_acc = torch.zeros(2, dtype=torch.float64, device=self.device)  # [batch loss, reg loss]
        self.set_optimizer(optimizer, optimizer_params, regularizer)
        self.launches = 0  # kernels launched by this engine (bench.py reports it)
        self._rank_ws = None  # caller-owned ranking workspace (kge_rank_workspace_bytes), grown on demand

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.kge_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- optimizer state (optimizers.py:255-291) ---------------------------
    def set_optimizer(self, name="adam", params=None, regularizer=None):
        name = name.lower()
        # 'lazy_<name>' (extension): update only the rows touched by the step (kge_optimizer_step_lazy)
        self.lazy = name.startswith("lazy_")
        if self.lazy:
            name = name[len("lazy_"):]
        if name not in _lib.OPTIMIZERS:
            raise ValueError("Could not interpret optimizer identifier: %r" % (name,))
        p = dict(params or {})
        regs = list(regularizer) if isinstance(regularizer, (list, tuple)) else [regularizer, regularizer]
        assert len(regs) == 2, "Incorrect length for regularizer. Expected 2, got {}".format(len(regs))

        def cfg_for(reg):
            reg = dict(reg or {})
            return _lib.KgeOptimizerConfig(
                C.sizeof(_lib.KgeOptimizerConfig), _lib.OPTIMIZERS[name], float(p.get("learning_rate", 0.001)),
                float(p.get("beta_1", 0.9)), float(p.get("beta_2", 0.999)), float(p.get("epsilon", 1e-7)),
                float(p.get("momentum", 0.0)), float(p.get("initial_accumulator_value", 0.1)),
                int(reg.get("p", 2)) if reg else 0, float(reg.get("lambda", 1e-5)),
                int(reg.get("p2", 0)) if reg else 0, float(reg.get("lambda2", 0.0)))
        self.opt_cfgs = {"ent": cfg_for(regs[0]), "rel": cfg_for(regs[1])}
        self.opt_cfg = self.opt_cfgs["ent"]  # hyper-parameters are common to both tables; only the regulariser differs
        self.opt_name = name
        self.t = 0
        # slots as [ent | rel] blocks, like the tables (see __init__): slots[key] are row slices of them
        mk = lambda v=0.0: torch.full((self.ent_rows + self.n_rel, self.ld), v, dtype=torch.float32, device=self.device)
        cut = lambda blk: {"ent": None if blk is None else blk[:self.ent_rows], "rel": None if blk is None else blk[self.ent_rows:]}
        b0 = b1 = None
        if name == "adam":
            b0, b1 = mk(), mk()
        elif name == "adagrad":
            b0 = mk(self.opt_cfg.initial_accumulator_value)
        elif self.opt_cfg.momentum != 0.0:
            b0 = mk()
        self._sblocks = (b0, b1)
        self.slots = {key: [cut(b0)[key], cut(b1)[key]] for key in ("ent", "rel")}
        self.stamps = {"ent": None, "rel": None}
        if self.lazy:
            self.stamps = {"ent": torch.zeros(self.ent_rows, dtype=torch.int32, device=self.device),
                           "rel": torch.zeros(self.n_rel, dtype=torch.int32, device=self.device)}
        _lib.check(self.lib.kge_set_row_stamps(self.h, _ptr(self.stamps["ent"]), _ptr(self.stamps["rel"])))
        self._last_step = 0

    # -- tables ------------------------------------------------------------
    def ensure_row_stash(self, B):
        """Local [B*eta, ld] stash the gradient pass re-reads instead of gathering the replaced rows a second time
        (kge_set_row_stash); a no-op when the kernel keeps every row window resident."""
        if self.lib.kge_rows_resident(self.h) != 0:
            return None
        need = int(B) * self.eta
        if getattr(self, "row_stash", None) is None or self.row_stash.shape[0] < need:
            self.row_stash = torch.empty((need, self.ld), dtype=torch.float32, device=self.device)
            _lib.check(self.lib.kge_set_row_stash(self.h, _ptr(self.row_stash), need))
        return self.row_stash

    def set_hot_entities(self, triples=None, ids=None):
        """Hot-entity hint (kge_set_hot_entities): the two most frequent subject/object ids of `triples` (numpy / torch
        [N,3] of indexed triples) or explicit `ids`; None/empty clears it."""
        if ids is None and triples is not None:
            t = triples.detach().cpu().numpy() if torch.is_tensor(triples) else np.asarray(triples)
            cnt = np.bincount(np.concatenate([t[:, 0], t[:, 2]]).astype(np.int64), minlength=self.n_ent)
            top = np.argsort(-cnt, kind="stable")[:2]
            ids = [int(e) for e in top if cnt[e] > 0]
        arr = np.asarray(ids if ids is not None else [], dtype=np.int32)
        _lib.check(self.lib.kge_set_hot_entities(self.h, arr.ctypes.data_as(C.c_void_p), int(arr.size)))
        self.hot_entities = arr.tolist()

    def set_embeddings(self, ent_dense=None, rel_dense=None):
        """dense [rows, internal_k] (numpy / torch) -> padded device layout."""
        for dense, table, rows in ((ent_dense, self.ent, self.ent_rows), (rel_dense, self.rel, self.n_rel)):
            if dense is None:
                continue
            d = torch.as_tensor(np.ascontiguousarray(dense, dtype=np.float32) if not torch.is_tensor(dense) else dense)
            d = d.to(self.device, torch.float32).contiguous()
            if tuple(d.shape) != (rows, self.internal_k):
                raise ValueError("expected shape %s, got %s" % ((rows, self.internal_k), tuple(d.shape)))
            _lib.check(self.lib.kge_pack_rows(self.h, _ptr(d), _ptr(table), rows, self._stream()))
            torch.cuda.current_stream(self.device).synchronize()  # d may be freed after return

    def get_embeddings(self):
        out = []
        for table, rows in ((self.ent, self.ent_rows), (self.rel, self.n_rel)):
            d = torch.empty((rows, self.internal_k), dtype=torch.float32, device=self.device)
            _lib.check(self.lib.kge_unpack_rows(self.h, _ptr(table), _ptr(d), rows, self._stream()))
            out.append(d)
        return out[0], out[1]

    def init_glorot_uniform(self, seed=0, only=None):
        if only in (None, "ent"):
            _lib.check(self.lib.kge_init_glorot_uniform(self.h, _ptr(self.ent), self.ent_rows, int(seed) * 2 + 0, self._stream()))
        if only in (None, "rel"):
            _lib.check(self.lib.kge_init_glorot_uniform(self.h, _ptr(self.rel), self.n_rel, int(seed) * 2 + 1, self._stream()))

    def init_table(self, which, kind, a=0.0, b=0.0, seed=0):
        """kge_init_table on the 'ent' or 'rel' table: kind in _lib.INIT_KINDS (uniform [a,b) | normal (mean a, stddev b) |
        truncated_normal | constant a)."""
        table, rows = (self.ent, self.ent_rows) if which == "ent" else (self.rel, self.n_rel)
        _lib.check(self.lib.kge_init_table(self.h, _ptr(table), rows, _lib.INIT_KINDS[kind], float(a), float(b),
                                           int(seed) * 2 + (0 if which == "ent" else 1), self._stream()))

    # -- training ------------------------------------------------------------
    @_traced("kge.train_step")
    def forward_backward(self, triples, negatives=None, seed=0, step=0, mode=_lib.STEP_FUSED,
                         scores_pos=None, scores_neg=None, dpos=None, dneg=None):
        """train_step up to tape.gradient: accumulates into g_ent/g_rel and loss_acc[0]."""
        assert triples.dtype == torch.int32 and triples.is_cuda and triples.is_contiguous()
        B = triples.shape[0]
        neg_ent = neg_keep = None
        if negatives is not None:
            neg_ent, neg_keep = negatives
            assert neg_ent.dtype == torch.int32 and neg_keep.dtype == torch.uint8
            assert neg_ent.numel() == B * self.eta and neg_keep.numel() == B * self.eta
        _lib.check(self.lib.kge_train_step(
            self.h, mode, _ptr(self.ent), _ptr(self.rel), _ptr(self.g_ent), _ptr(self.g_rel), _ptr(triples), B,
            _ptr(neg_ent), _ptr(neg_keep), int(seed), int(step), _ptr(self.loss_acc), _ptr(scores_pos),
            _ptr(scores_neg), _ptr(dpos), _ptr(dneg), self._stream()))
        self._last_step = int(step)
        self.launches += 2 if self.scoring_type == "RotatE" else 1

    def _one_block(self):
        """True when tables, gradients and slots still ARE the [ent | rel] blocks allocated here (callers may rebind them: the
        multi-GPU trainers do) and both tables share one regulariser: the dense optimizer can take them in one launch."""
        pb, gb = getattr(self, "_pblock", None), getattr(self, "_gblock", None)
        if self.lazy or pb is None:
            return False
        off = self.ent_rows * self.ld * 4

        def is_split(blk, a, b):
            if blk is None:
                return a is None and b is None
            return a is not None and b is not None and a.data_ptr() == blk.data_ptr() and b.data_ptr() == blk.data_ptr() + off

        ce, cr = self.opt_cfgs["ent"], self.opt_cfgs["rel"]
        same_reg = all(getattr(ce, f) == getattr(cr, f) for f, _ in ce._fields_)
        return (same_reg and is_split(pb, self.ent, self.rel) and is_split(gb, self.g_ent, self.g_rel)
                and is_split(self._sblocks[0], self.slots["ent"][0], self.slots["rel"][0])
                and is_split(self._sblocks[1], self.slots["ent"][1], self.slots["rel"][1]))

    @_traced("kge.optimizer_step")
    def apply_gradients(self):
        """optimizer.apply_gradients on both tables (dense semantics) + LP regulariser."""
        self.t += 1
        if self._one_block():
            _lib.check(self.lib.kge_optimizer_step(
                self.h, C.byref(self.opt_cfgs["ent"]), self.t, _ptr(self._pblock), _ptr(self._gblock), _ptr(self._sblocks[0]),
                _ptr(self._sblocks[1]), self.ent_rows + self.n_rel, C.c_void_p(self.loss_acc.data_ptr() + 8), self._stream()))
            self.launches += 1
            return
        for key, table, grad, rows in (("ent", self.ent, self.g_ent, self.ent_rows),
                                       ("rel", self.rel, self.g_rel, self.n_rel)):
            s0, s1 = self.slots[key]
            if self.lazy:
                _lib.check(self.lib.kge_optimizer_step_lazy(
                    self.h, C.byref(self.opt_cfgs[key]), self.t, _ptr(table), _ptr(grad), _ptr(s0), _ptr(s1), rows,
                    _ptr(self.stamps[key]), self.lib.kge_step_stamp(self._last_step),
                    C.c_void_p(self.loss_acc.data_ptr() + 8), self._stream()))
                continue
            _lib.check(self.lib.kge_optimizer_step(
                self.h, C.byref(self.opt_cfgs[key]), self.t, _ptr(table), _ptr(grad), _ptr(s0), _ptr(s1), rows,
                C.c_void_p(self.loss_acc.data_ptr() + 8), self._stream()))
        self.launches += 2

    def train_step(self, triples, negatives=None, seed=0, step=None):
        """One reference train_step (models/ScoringBasedEmbeddingModel.py:370-429)."""
        self.forward_backward(triples, negatives, seed, self.t if step is None else step)
        self.apply_gradients()

    def read_loss(self, reset=True):
        """(batch loss + regulariser loss) accumulated since the last reset; synchronises."""
        v = float(self.loss_acc.cpu().sum())  # one 16-byte D2H copy (synchronising), summed on the host: no reduce kernel
        if reset:
            self.loss_acc.zero_()
        return v

    def generate_corruptions(self, triples, seed=0, step=0):
        out = torch.empty((triples.shape[0] * self.eta, 3), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.kge_generate_corruptions(self.h, _ptr(triples), triples.shape[0], int(seed), int(step),
                                                     _ptr(out), self._stream()))
        return out

    # -- inference -------------------------------------------------------------
    @_traced("kge.score_triples")
    def score(self, triples):
        assert triples.dtype == torch.int32 and triples.is_cuda and triples.is_contiguous()
        out = torch.empty(triples.shape[0], dtype=torch.float32, device=self.device)
        _lib.check(self.lib.kge_score_triples(self.h, _ptr(self.ent), _ptr(self.rel), _ptr(triples), triples.shape[0],
                                              _ptr(out), self._stream()))
        self.launches += 2 if self.scoring_type == "RotatE" else 1
        return out

    def rank_workspace(self, b, n_cand):
        """The caller-owned scratch kge_rank needs (kge_rank_workspace_bytes): one cached torch buffer, grown on demand.
        Growing frees the old buffer through torch's stream-ordered caching allocator, so no synchronisation is needed."""
        need = int(self.lib.kge_rank_workspace_bytes(self.h, int(b), int(n_cand)))
        if self._rank_ws is None or self._rank_ws.numel() < need:
            raw = torch.empty(max(need, 1024) + 1024, dtype=torch.uint8, device=self.device)
            off = (-raw.data_ptr()) % 1024  # the library wants a 1024-byte aligned base (operand tiles of the tensor-core pass)
            self._rank_ws_raw, self._rank_ws = raw, raw[off:off + max(need, 1024)]
        return self._rank_ws, need

    @_traced("kge.rank")
    def rank(self, triples, side, strategy="worst", filt_off=None, filt_idx=None, cand_ids=None, cand_begin=0,
             n_cand=None, out=None, counts=None):
        """get_ranks for one side.  Default: returns/accumulates int32 rank counts in `out` (caller adds 1).
        counts=[b,3] int32: accumulate the RAW counters {greater, equal, filtered} there instead (candidate
        partitions must do this and call finalize_ranks once: 'middle' is not additive over partitions)."""
        assert triples.dtype == torch.int32 and triples.is_cuda and triples.is_contiguous()
        b = triples.shape[0]
        if out is None and counts is None:
            out = torch.zeros(b, dtype=torch.int32, device=self.device)
        if n_cand is None:
            n_cand = cand_ids.numel() if cand_ids is not None else self.n_ent - cand_begin
        n_filt = int(filt_idx.numel()) if filt_idx is not None else 0
        ws, ws_bytes = self.rank_workspace(b, n_cand)
        _lib.check(self.lib.kge_rank(self.h, _lib.SIDES[side], _lib.STRATEGIES[strategy], _ptr(self.ent), _ptr(self.rel),
                                     _ptr(triples), b, _ptr(cand_ids), int(cand_begin), int(n_cand), _ptr(filt_off),
                                     _ptr(filt_idx), n_filt, _ptr(out) if counts is None else None, _ptr(counts),
                                     _ptr(ws), ws_bytes, self._stream()))
        self.launches += 5
        return out if counts is None else counts

    def finalize_ranks(self, counts, strategy="worst", out=None):
        """ranks += f_strategy(greater, equal) - filtered, once, from raw counters summed over candidate partitions."""
        b = counts.shape[0]
        if out is None:
            out = torch.zeros(b, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.kge_rank_finalize(self.h, _ptr(counts), b, _lib.STRATEGIES[strategy], _ptr(out), self._stream()))
        self.launches += 1
        return out

    def rank_filter_probe(self, triples, side, cand_ids=None, cand_begin=0, n_cand=None):
        """Diagnostic (kge_rank_filter_probe): (approximate tensor-core scores, assumed error bound), both [b, n_cand]."""
        b = triples.shape[0]
        if n_cand is None:
            n_cand = cand_ids.numel() if cand_ids is not None else self.n_ent - cand_begin
        approx = torch.empty((b, int(n_cand)), dtype=torch.float32, device=self.device)
        delta = torch.empty_like(approx)
        ws, ws_bytes = self.rank_workspace(b, n_cand)
        _lib.check(self.lib.kge_rank_filter_probe(self.h, _lib.SIDES[side], _ptr(self.ent), _ptr(self.rel), _ptr(triples), b,
                                                  _ptr(cand_ids), int(cand_begin), int(n_cand), _ptr(approx), _ptr(delta),
                                                  _ptr(ws), ws_bytes, self._stream()))
        return approx, delta

    def corruption_scores(self, triples, side, cand_ids=None, cand_begin=0, n_cand=None):
        """_get_subject_corruption_scores / _get_object_corruption_scores: fp32 [b, n_cand], canonical summation order."""
        assert triples.dtype == torch.int32 and triples.is_cuda and triples.is_contiguous()
        b = triples.shape[0]
        if n_cand is None:
            n_cand = cand_ids.numel() if cand_ids is not None else self.n_ent - cand_begin
        out = torch.empty((b, int(n_cand)), dtype=torch.float32, device=self.device)
        ws, ws_bytes = self.rank_workspace(b, n_cand)
        _lib.check(self.lib.kge_corruption_scores(self.h, _lib.SIDES[side], _ptr(self.ent), _ptr(self.rel), _ptr(triples), b,
                                                  _ptr(cand_ids), int(cand_begin), int(n_cand), _ptr(out), _ptr(ws), ws_bytes,
                                                  self._stream()))
        self.launches += 3
        return out
