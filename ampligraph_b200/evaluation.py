"""Rank metrics (mirror of ampligraph/evaluation/metrics.py:18-259): O(n) numpy on the
int32 ranks the ranking kernel returns."""
import numpy as np


def _flat(ranks):
    r = np.asarray(ranks).reshape(-1)
    return r


def mrr_score(ranks):
    """Mean reciprocal rank (metrics.py:155)."""
    r = _flat(ranks)
    return float(np.sum(1.0 / r) / len(r))


def mr_score(ranks):
    """Mean rank (metrics.py:196)."""
    r = _flat(ranks)
    return float(np.sum(r) / len(r))


def hits_at_n_score(ranks, n):
    """Hits@n (metrics.py:18)."""
    r = _flat(ranks)
    return float(np.sum(r <= n) / len(r))


def rank_score(y_true, y_pred, pos_lab=1):
    """Rank of the positive among y_pred (metrics.py:87)."""
    y_true, y_pred = np.asarray(y_true), np.asarray(y_pred)
    idx = np.argsort(y_pred)[::-1]
    return int(np.where(y_true[idx] == pos_lab)[0][0] + 1)
