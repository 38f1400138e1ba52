"""Mirror of ampligraph.latent_features for the hot path (facade over KGEEngine)."""
from . import loss_functions, optimizers, regularizers  # noqa: F401
from .layers.scoring import SCORING_LAYER_REGISTRY  # noqa: F401
from .loss_functions import LOSS_REGISTRY  # noqa: F401


def __getattr__(name):
    # the model pulls in torch; keep `import ampligraph_b200.latent_features` light
    if name in ("ScoringBasedEmbeddingModel", "EarlyStopping"):
        from . import models
        return getattr(models, name)
    raise AttributeError(name)
