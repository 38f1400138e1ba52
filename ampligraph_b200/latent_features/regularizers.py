"""Regulariser selection (mirror of ampligraph/latent_features/regularizers.py).

LP_regularizer = lambda * sum(|x|^p) over the WHOLE table (regularizers.py:14-37); its
loss and dense gradient are fused into the optimizer kernel (csrc/kge_optim.cu).
"""


class LPRegularizer:
    __name__ = "LP"

    def __init__(self, regularizer_parameters=None):
        rp = dict(regularizer_parameters or {})
        self.p = int(rp.get("p", 2))
        self.lam = float(rp.get("lambda", 0.00001))

    def kernel_params(self):
        return {"p": self.p, "lambda": self.lam}


def LP_regularizer(regularizer_parameters=None):
    return LPRegularizer(regularizer_parameters)


def get(identifier, hyperparams=None):
    """regularizers.get (:40-73): 'LP' | 'l3' | LPRegularizer | None."""
    hyperparams = dict(hyperparams or {})
    if identifier is None or isinstance(identifier, LPRegularizer):
        return identifier
    if isinstance(identifier, str) and identifier == "l3":
        hyperparams["p"] = 3
        return LPRegularizer(hyperparams)
    if isinstance(identifier, str) and identifier == "LP":
        return LPRegularizer(hyperparams)
    if isinstance(identifier, str) and identifier in ("l2", "L2"):  # Keras 'l2': 0.01 * sum(x^2)
        return LPRegularizer({"p": 2, "lambda": hyperparams.get("lambda", 0.01)})
    if isinstance(identifier, str) and identifier in ("l1", "L1"):
        return LPRegularizer({"p": 1, "lambda": hyperparams.get("lambda", 0.01)})
    raise ValueError("Could not interpret regularizer identifier: %r" % (identifier,))
