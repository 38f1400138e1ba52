"""Loss registry of the facade (mirror of ampligraph/latent_features/loss_functions.py).

The arithmetic of the five registered losses lives in the fused CUDA kernel
(csrc/kge_train.cu: loss_and_dscores); these classes only carry the name and the
hyper-parameters across the C-ABI, with the reference's defaults and error
behaviour.  A user callable is wrapped in LossFunctionWrapper and evaluated with
torch autograd on the kernel's score outputs (two-phase step).
"""
LOSS_REGISTRY = {}

DEFAULT_MARGIN = 1  # loss_functions.py:23
DEFAULT_ALPHA_ADVERSARIAL = 0.5  # :26
DEFAULT_MARGIN_ADVERSARIAL = 3  # :29
DEFAULT_REDUCTION = "sum"  # :44


def register_loss(name, external_params=None):
    def insert_in_registry(cls):
        LOSS_REGISTRY[name] = cls
        cls.name = name
        cls.external_params = external_params or []
        return cls
    return insert_in_registry


class Loss:
    """Base class: validates `reduction` and stores hyper-parameters (loss_functions.py:72-122)."""
    name = None

    def __init__(self, hyperparam_dict=None, verbose=False):
        hyperparam_dict = dict(hyperparam_dict or {})
        self._loss_parameters = {"reduction": hyperparam_dict.get("reduction", DEFAULT_REDUCTION)}
        assert self._loss_parameters["reduction"] in ["sum", "mean"], "Invalid value for reduction!"
        self._init_hyperparams(hyperparam_dict)

    def _init_hyperparams(self, hyperparam_dict):
        pass

    def kernel_params(self):
        """(registry name, params dict) handed to kge_create."""
        return self.name, dict(self._loss_parameters)


@register_loss("pairwise", ["margin"])
class PairwiseLoss(Loss):
    def _init_hyperparams(self, h):
        self._loss_parameters["margin"] = h.get("margin", DEFAULT_MARGIN)


@register_loss("nll")
class NLLLoss(Loss):
    pass


@register_loss("absolute_margin", ["margin"])
class AbsoluteMarginLoss(Loss):
    def _init_hyperparams(self, h):
        self._loss_parameters["margin"] = h.get("margin", DEFAULT_MARGIN)


@register_loss("self_adversarial", ["margin", "alpha"])
class SelfAdversarialLoss(Loss):
    def _init_hyperparams(self, h):
        self._loss_parameters["margin"] = h.get("margin", DEFAULT_MARGIN_ADVERSARIAL)
        self._loss_parameters["alpha"] = h.get("alpha", DEFAULT_ALPHA_ADVERSARIAL)


@register_loss("multiclass_nll")
class NLLMulticlass(Loss):
    pass


class LossFunctionWrapper(Loss):
    """User callable `(scores_pos [B], scores_neg [eta, B]) -> [B]` on torch tensors
    (the reference passes TF tensors, loss_functions.py:657-717)."""

    def __init__(self, user_defined_loss, name=None):
        super().__init__()
        self._user_losses = user_defined_loss
        self.name = name

    def kernel_params(self):
        return None, dict(self._loss_parameters)


def get(identifier, hyperparams=None):
    """loss_functions.get (:720-766): instance | registry name | callable."""
    if isinstance(identifier, Loss):
        return identifier
    if isinstance(identifier, str):
        if identifier not in LOSS_REGISTRY:
            raise ValueError("Could not interpret loss identifier:", identifier)
        return LOSS_REGISTRY[identifier](hyperparams or {})
    if callable(identifier):
        return LossFunctionWrapper(identifier, getattr(identifier, "__name__", "user_loss"))
    raise ValueError("Could not interpret loss identifier:", identifier)
