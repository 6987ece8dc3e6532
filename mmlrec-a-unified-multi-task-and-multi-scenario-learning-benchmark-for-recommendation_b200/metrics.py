"""Device-side evaluation metrics with the semantics of the sklearn functions the reference calls
(``/root/reference/model/basemodel.py:616-635``: ``roc_auc_score``, ``log_loss``, ``mean_squared_error`` and subset
``accuracy_score`` at 0.5).  Predictions stay in HBM: no device -> host copy of ``[N, T]`` arrays and no per-batch host
synchronisation.  Built from torch's CUDA primitives (sort / scan); metrics are outside the
timed training step.  Accumulation is fp64, so the values agree with sklearn to ~1e-12 (tests/test_metrics_gpu.py)."""
from __future__ import annotations

import torch


def _as_2d(t: torch.Tensor) -> torch.Tensor:
    return t.reshape(t.shape[0], -1)


def binary_auc(y_true: torch.Tensor, y_score: torch.Tensor) -> torch.Tensor:
    """Area under the ROC curve of one label column: the Mann-Whitney statistic with ties counted half, which is what
    sklearn's trapezoid over the distinct thresholds integrates to.  Raises ValueError for a single-class column
    (as sklearn does)."""
    y = y_true.reshape(-1).to(torch.float64)
    s = y_score.reshape(-1)
    order = torch.argsort(s, stable=True)
    s, y = s[order], y[order]
    boundary = torch.ones_like(s, dtype=torch.bool)
    boundary[1:] = s[1:] != s[:-1]
    gid = boundary.cumsum(0) - 1                       # tie-group id of every sample
    n_groups = int(gid[-1].item()) + 1
    pos_g = torch.zeros(n_groups, dtype=torch.float64, device=s.device).index_add_(0, gid, y)
    neg_g = torch.zeros(n_groups, dtype=torch.float64, device=s.device).index_add_(0, gid, 1.0 - y)
    n_pos, n_neg = pos_g.sum(), neg_g.sum()
    if float(n_pos) == 0.0 or float(n_neg) == 0.0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    neg_below = neg_g.cumsum(0) - neg_g
    return (pos_g * (neg_below + 0.5 * neg_g)).sum() / (n_pos * n_neg)


def roc_auc_score(y_true: torch.Tensor, y_score: torch.Tensor) -> float:
    """sklearn.metrics.roc_auc_score for a binary vector or an [N, T] multilabel indicator matrix (macro average)."""
    yt, ys = _as_2d(y_true), _as_2d(y_score)
    return float(torch.stack([binary_auc(yt[:, t], ys[:, t]) for t in range(yt.shape[1])]).mean().item())


def accuracy_score(y_true: torch.Tensor, y_pred: torch.Tensor) -> float:
    """The reference's ``_accuracy_score``: labels = pred > 0.5, subset accuracy for multilabel (all columns equal)."""
    yt, yp = _as_2d(y_true), _as_2d(y_pred)
    hit = ((yp > 0.5).to(yt.dtype) == yt).all(dim=1)
    return float(hit.to(torch.float64).mean().item())


def log_loss(y_true: torch.Tensor, y_pred: torch.Tensor) -> float:
    """sklearn.metrics.log_loss for binary labels (probabilities clipped to [eps, 1 - eps] with float64 eps)."""
    y, p = y_true.reshape(-1).to(torch.float64), y_pred.reshape(-1).to(torch.float64)
    eps = torch.finfo(torch.float64).eps
    p = p.clamp(eps, 1 - eps)
    return float((-(y * p.log() + (1 - y) * (1 - p).log())).mean().item())


def mean_squared_error(y_true: torch.Tensor, y_pred: torch.Tensor) -> float:
    d = _as_2d(y_true).to(torch.float64) - _as_2d(y_pred).to(torch.float64)
    return float((d * d).mean(dim=0).mean().item())
