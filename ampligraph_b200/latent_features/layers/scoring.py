"""Scoring-layer registry of the facade (mirror of
ampligraph/latent_features/layers/scoring/AbstractScoringLayer.py:15-54).

`SCORING_LAYER_REGISTRY[name](k)` gives an object with the reference's plugin surface
(AbstractScoringLayer.py:103-156): `internal_k`, `_compute_scores`, `_get_subject_corruption_scores`,
`_get_object_corruption_scores`, `get_ranks`.  The arithmetic is in csrc/kge_train.cu / kge_rank.cu /
kge_misc.cu, selected by `kernel_id`; these methods take EMBEDDINGS like the reference's (`triples` is the
list [e_s, e_p, e_o] of [n, internal_k] arrays produced by the lookup layer, `ent_matrix` a [m, internal_k]
slice of the entity table), stage them in a scratch engine and call kge_score_triples /
kge_corruption_scores / kge_rank through the C-ABI.  The model itself (models.py) never goes through them:
it keeps ids on the device and calls the same entry points on the resident tables.
"""
import numpy as np

SCORING_LAYER_REGISTRY = {}
COMPARISION_PRECISION = 1e3  # AbstractScoringLayer.py:11


def register_layer(name, external_params=None, class_params=None):
    def insert_in_registry(cls):
        assert name not in SCORING_LAYER_REGISTRY, "Scoring Layer with name {} already exists!".format(name)
        SCORING_LAYER_REGISTRY[name] = cls
        cls.name = name
        cls.external_params = external_params or []
        cls.class_params = class_params or {}
        return cls
    return insert_in_registry


def _np(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


class AbstractScoringLayer:
    kernel_id = None
    max_rel_size = None  # RotatE only

    def __init__(self, k):
        self.k = k
        self.internal_k = k

    # -- staging: embeddings -> a scratch engine whose tables are [ent_matrix | e_s | e_o] and e_p ------------------
    def _stage(self, triples, ent_matrix=None):
        import torch
        from ...engine import KGEEngine
        if self.kernel_id is None:
            raise NotImplementedError("scoring layer %r has no CUDA kernel" % self.name)
        e_s, e_p, e_o = (_np(x).reshape(-1, self.internal_k) for x in triples)
        n = e_s.shape[0]
        cand = _np(ent_matrix).reshape(-1, self.internal_k) if ent_matrix is not None else np.zeros((0, self.internal_k), np.float32)
        m = cand.shape[0]
        eng = KGEEngine(self.name, self.k, 1, m + 2 * n, max(n, 1), max_rel_size=self.max_rel_size)
        eng.set_embeddings(np.concatenate([cand, e_s, e_o]), e_p if n else np.zeros((1, self.internal_k), np.float32))
        ids = np.stack([m + np.arange(n), np.arange(n), m + n + np.arange(n)], 1).astype(np.int32)
        return eng, torch.as_tensor(ids).to(eng.device), n, m

    def call(self, triples):
        return self._compute_scores(triples)

    __call__ = call

    def _compute_scores(self, triples):
        """fp32 [n]: score of each (e_s, e_p, e_o) row (TransE.py:37, DistMult.py:34, ComplEx.py:39, HolE.py:31, RotatE.py:62)."""
        eng, ids, n, _ = self._stage(triples)
        try:
            return eng.score(ids).cpu().numpy() if n else np.zeros(0, np.float32)
        finally:
            eng.close()

    def _corruption_scores(self, triples, ent_matrix, side):
        eng, ids, n, m = self._stage(triples, ent_matrix)
        try:
            if n == 0 or m == 0:
                return np.zeros((n, m), np.float32)
            return eng.corruption_scores(ids, side, cand_begin=0, n_cand=m).cpu().numpy()
        finally:
            eng.close()

    def _get_subject_corruption_scores(self, triples, ent_matrix):
        """fp32 [n, m]: every row of ent_matrix substituted as the subject (TransE.py:56-84 etc.)."""
        return self._corruption_scores(triples, ent_matrix, "s")

    def _get_object_corruption_scores(self, triples, ent_matrix):
        """fp32 [n, m]: every row of ent_matrix substituted as the object (TransE.py:86-114 etc.)."""
        return self._corruption_scores(triples, ent_matrix, "o")

    def get_ranks(self, triples, ent_matrix, start_ent_id, end_ent_id, filters, mapping_dict=None, corrupt_side="s,o",
                  comparison_type="worst"):
        """AbstractScoringLayer.get_ranks (:156-422): int32 [sides, n] rank counts (the caller adds 1,
        ScoringBasedEmbeddingModel.py:1684).  `filters`: one list per corrupted side of n id arrays (known true
        entities); ids are remapped through `mapping_dict` (entities_subset, :266-275) when it is non-empty, clipped to
        [start_ent_id, end_ent_id] (:280-288) and shifted to positions in `ent_matrix`."""
        import torch
        assert comparison_type in ("worst", "best", "middle"), "Invalid value for ranking_strategy"
        sides = [s for s in ("s", "o") if s in corrupt_side]
        assert sides, "Invalid value for corrupt_side"
        eng, ids, n, m = self._stage(triples, ent_matrix)
        try:
            out = np.zeros((len(sides), n), np.int32)
            if n == 0:
                return out
            for j, side in enumerate(sides):
                off = idx = None
                if filters is not None and len(filters) > j and len(filters[j]) > 0:
                    lists = []
                    for f in filters[j]:
                        f = np.asarray(f, dtype=np.int64).reshape(-1)
                        if mapping_dict:
                            f = np.asarray([mapping_dict.get(int(x), -1) for x in f], dtype=np.int64)
                            f = f[f >= 0]
                        f = f[(f >= start_ent_id) & (f <= end_ent_id)] - start_ent_id
                        lists.append(f[(f >= 0) & (f < m)])
                    off = torch.as_tensor(np.concatenate([[0], np.cumsum([len(f) for f in lists])]).astype(np.int64)).to(eng.device)
                    flat = np.concatenate(lists).astype(np.int32) if int(off[-1]) else np.zeros(0, np.int32)
                    idx = torch.as_tensor(flat).to(eng.device)
                out[j] = eng.rank(ids, side, comparison_type, off, idx, cand_begin=0, n_cand=m).cpu().numpy()
            return out
        finally:
            eng.close()


@register_layer("TransE")
class TransE(AbstractScoringLayer):
    kernel_id = 0


@register_layer("DistMult")
class DistMult(AbstractScoringLayer):
    kernel_id = 1


@register_layer("ComplEx")
class ComplEx(AbstractScoringLayer):
    kernel_id = 2

    def __init__(self, k):
        super().__init__(k)
        self.internal_k = 2 * k  # ComplEx.py:37


@register_layer("HolE")
class HolE(ComplEx):
    kernel_id = 3


@register_layer("RotatE")
class RotatE(AbstractScoringLayer):
    kernel_id = 4

    def __init__(self, k, max_rel_size=None):
        super().__init__(k)
        self.internal_k = 2 * k  # RotatE.py:57
        self.max_rel_size = max_rel_size  # set by the model at build time (ScoringBasedEmbeddingModel.py:338)


@register_layer("Random")
class Random(AbstractScoringLayer):
    """Registry parity with layers/scoring/Random.py:23-82 (uniform-random scores, the reference's test baseline).
    It has no arithmetic to accelerate, so there is no kernel behind it: building a model with it raises."""
    kernel_id = None
