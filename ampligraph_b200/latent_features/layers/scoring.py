"""Scoring-layer registry of the facade (mirror of
ampligraph/latent_features/layers/scoring/AbstractScoringLayer.py:15-54).

Each entry is a light descriptor: the arithmetic (_compute_scores, corruption scores,
get_ranks) is in csrc/kge_train.cu and csrc/kge_rank.cu, selected by `kernel_id`.
"""
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


class AbstractScoringLayer:
    kernel_id = None

    def __init__(self, k):
        self.k = k
        self.internal_k = k


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
        self.max_rel_size = max_rel_size


@register_layer("Random")
class Random(AbstractScoringLayer):
    """Registry parity with layers/scoring/Random.py:23-82 (uniform-random scores, the reference's test baseline).
    It has no arithmetic to accelerate, so there is no kernel behind it: building a model with it raises."""
    kernel_id = None

