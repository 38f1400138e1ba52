"""Optimizer selection of the facade (mirror of ampligraph/latent_features/optimizers.py).

The reference wraps tf.keras.optimizers.legacy.* (optimizers.py:255-291); here the
update itself is csrc/kge_optim.cu and this wrapper only carries name +
hyper-parameters.  Supported: 'sgd' (optional momentum), 'adam', 'adagrad'.
"""
SUPPORTED = ("sgd", "adam", "adagrad",
             # extension (not in the reference): touched-rows-only updates for tables too large for the dense rule
             "lazy_sgd", "lazy_adam", "lazy_adagrad")


class OptimizerWrapper:
    def __init__(self, name="adam", hyperparams=None):
        name = name.lower()
        if name not in SUPPORTED:
            raise ValueError("Could not interpret optimizer identifier: ", name)
        self.name = name
        self.hyperparams = dict(hyperparams or {})
        self.hyperparams.setdefault("learning_rate", 0.001)  # optimizers.py:284
        # adam has beta_1/beta_2 slots (optimizers.py:119-120)
        self.number_hyperparams = 2 if name.endswith("adam") else 1

    def get_hyperparam_count(self):
        return self.number_hyperparams

    def get_config(self):
        return dict(self.hyperparams, name=self.name)


def get(identifier, hyperparams=None):
    """optimizers.get (:255-291): wrapper instance | name | object with get_config()."""
    if isinstance(identifier, OptimizerWrapper):
        return identifier
    if isinstance(identifier, str):
        return OptimizerWrapper(identifier, dict(hyperparams or {}))
    cfg = getattr(identifier, "get_config", None)
    if callable(cfg):  # e.g. a Keras optimizer instance: take its config
        c = dict(cfg())
        name = c.pop("name", type(identifier).__name__)
        keep = {k: c[k] for k in ("learning_rate", "beta_1", "beta_2", "epsilon", "momentum",
                                  "initial_accumulator_value") if k in c}
        return OptimizerWrapper(str(name), keep)
    raise ValueError("Could not interpret optimizer identifier: ", identifier)
