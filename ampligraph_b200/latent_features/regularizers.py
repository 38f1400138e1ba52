"""Regulariser selection (mirror of ampligraph/latent_features/regularizers.py + what
`tf.keras.regularizers.get` resolves in EmbeddingLookupLayer.set_regularizer, layers/encoding/EmbeddingLookupLayer.py:131-155).

LP_regularizer = lambda * sum(|x|^p) over the WHOLE table (regularizers.py:14-37); its
loss and dense gradient are fused into the optimizer kernel (csrc/kge_optim.cu).  Keras' 'l1', 'l2' and
'l1_l2' are the same thing with p = 1 / 2 (/ both) and Keras' default factor 0.01.  `compile` accepts one
regulariser for both tables or a list [entities, relations], exactly like the reference.
"""


class LPRegularizer:
    __name__ = "LP"

    def __init__(self, regularizer_parameters=None):
        rp = dict(regularizer_parameters or {})
        self.p = int(rp.get("p", 2))
        self.lam = float(rp.get("lambda", 0.00001))
        self.p2 = int(rp.get("p2", 0))        # optional second term (Keras L1L2)
        self.lam2 = float(rp.get("lambda2", 0.0))
        if self.p < 1 or self.p2 < 0:
            raise ValueError("LP regularizer: p must be >= 1")

    def kernel_params(self):
        d = {"p": self.p, "lambda": self.lam}
        if self.p2:
            d.update(p2=self.p2, lambda2=self.lam2)
        return d


def LP_regularizer(regularizer_parameters=None):
    return LPRegularizer(regularizer_parameters)


def get(identifier, hyperparams=None):
    """regularizers.get (:40-73) + tf.keras.regularizers.get: 'LP' | 'l3' | 'l1' | 'l2' | 'l1_l2' |
    {'class_name': 'L1'|'L2'|'L1L2', 'config': {...}} | LPRegularizer | None."""
    hyperparams = dict(hyperparams or {})
    if identifier is None or isinstance(identifier, LPRegularizer):
        return identifier
    if isinstance(identifier, dict):
        cfg = dict(identifier.get("config", {}))
        name = str(identifier.get("class_name", "")).lower()
        if name == "l1":
            return LPRegularizer({"p": 1, "lambda": cfg.get("l1", 0.01)})
        if name == "l2":
            return LPRegularizer({"p": 2, "lambda": cfg.get("l2", 0.01)})
        if name == "l1l2":
            return LPRegularizer({"p": 1, "lambda": cfg.get("l1", 0.0), "p2": 2, "lambda2": cfg.get("l2", 0.0)})
        raise ValueError("Could not interpret regularizer identifier: %r" % (identifier,))
    if isinstance(identifier, str):
        key = identifier.lower()
        if identifier == "l3":
            hyperparams["p"] = 3
            return LPRegularizer(hyperparams)
        if identifier == "LP":
            return LPRegularizer(hyperparams)
        if key == "l2":  # Keras 'l2': 0.01 * sum(x^2)
            return LPRegularizer({"p": 2, "lambda": hyperparams.get("lambda", 0.01)})
        if key == "l1":
            return LPRegularizer({"p": 1, "lambda": hyperparams.get("lambda", 0.01)})
        if key == "l1_l2":
            return LPRegularizer({"p": 1, "lambda": hyperparams.get("l1", 0.01), "p2": 2, "lambda2": hyperparams.get("l2", 0.01)})
    raise ValueError("Could not interpret regularizer identifier: %r" % (identifier,))


def get_pair(identifier, hyperparams=None):
    """One regulariser for both tables, or a list of two [entities, relations] (EmbeddingLookupLayer.py:145-155)."""
    if isinstance(identifier, (list, tuple)):
        assert len(identifier) == 2, "Incorrect length for regularizer. Expected 2, got {}".format(len(identifier))
        return [get(identifier[0], hyperparams), get(identifier[1], hyperparams)]
    r = get(identifier, hyperparams)
    return [r, r]
