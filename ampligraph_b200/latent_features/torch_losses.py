"""The five registered losses on torch tensors (device side), used ONLY when the fused in-kernel
loss cannot be (FocusE re-weights the scores before the loss, ScoringBasedEmbeddingModel.py:396-406):
the kernel returns scores, autograd differentiates these few elementwise ops, and the kernel's
BACKWARD_EXT mode turns dL/dscore into embedding gradients.  Formulas: loss_functions.py
:286-308 (pairwise), :360-382 (nll), :442-464 (absolute_margin), :540-574 (self_adversarial),
:630-654 (multiclass_nll); reduction over the eta corruptions :124-129."""
import torch

_LO, _HI = -75.0, 75.0


def _reduce(x, reduction):
    return x.sum(0) if reduction == "sum" else x.mean(0)


def per_positive_loss(name, scores_pos, scores_neg, params):
    """scores_pos [B], scores_neg [eta, B] -> [B]."""
    red = params.get("reduction", "sum")
    if name == "pairwise":
        return _reduce(torch.clamp(params.get("margin", 1) - scores_pos + scores_neg, min=0), red)
    if name == "nll":
        sn, sp = torch.clamp(scores_neg, _LO, _HI), torch.clamp(scores_pos, _LO, _HI)
        sc = torch.cat([-sp.expand_as(sn), sn], 0)
        return _reduce(torch.log(1 + torch.exp(sc)), red)
    if name == "absolute_margin":
        return _reduce(torch.clamp(params.get("margin", 1) + scores_neg, min=0) - scores_pos, red)
    if name == "self_adversarial":
        m, a = params.get("margin", 3), params.get("alpha", 0.5)
        p_neg = torch.softmax(a * scores_neg, dim=0)
        return -torch.nn.functional.logsigmoid(m + scores_pos) - _reduce(
            p_neg * torch.nn.functional.logsigmoid(-scores_neg - m), red)
    if name == "multiclass_nll":
        sp, sn = torch.clamp(scores_pos, _LO, _HI), torch.clamp(scores_neg, _LO, _HI)
        pos_exp = torch.exp(sp)
        return -torch.log(pos_exp / (_reduce(torch.exp(sn), red) + pos_exp))
    raise ValueError("Could not interpret loss identifier:", name)


class _Softplus9999(torch.autograd.Function):
    """focusE 'softplus' (ScoringBasedEmbeddingModel.py:500-510): log(1 + 9999 e^x), grad 1 - 1/(1 + 9999 e^x)."""

    @staticmethod
    def forward(ctx, x):
        e = 9999 * torch.exp(x)
        ctx.save_for_backward(e)
        return torch.log(1 + e)

    @staticmethod
    def backward(ctx, dy):
        (e,) = ctx.saved_tensors
        return dy * (1 - 1 / (1 + e))


def focuse_non_linearity(name):
    if name == "linear":
        return lambda x: x
    if name == "tanh":
        return torch.tanh
    if name == "sigmoid":
        return torch.sigmoid
    if name == "softplus":
        return _Softplus9999.apply
    raise ValueError("Invalid focusE non-linearity")
