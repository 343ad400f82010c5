"""On-device evaluation: AUROC / average precision by sort + prefix sums, no host round trip
of the score vector.  Matches sklearn.metrics.roc_auc_score / average_precision_score (ties handled by
grouping equal scores), which is what run.py:236-240 and src/utils.py:232-233 call on the host."""
from __future__ import annotations

import torch


def _grouped_counts(scores: torch.Tensor, labels: torch.Tensor):
    """Descending distinct thresholds with cumulative true/false positive counts (float64, exact integers)."""
    s, order = torch.sort(scores.reshape(-1).double(), descending=True, stable=True)
    y = labels.reshape(-1)[order].double()
    last = torch.ones_like(s, dtype=torch.bool)
    last[:-1] = s[1:] != s[:-1]                      # last element of every run of equal scores
    tp = torch.cumsum(y, 0)[last]
    fp = torch.cumsum(1.0 - y, 0)[last]
    return tp, fp


def roc_auc(scores: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Area under the ROC curve (trapezoid over distinct thresholds); labels in {0,1}."""
    tp, fp = _grouped_counts(scores, labels)
    p, n = tp[-1], fp[-1]
    tp0 = torch.cat([tp.new_zeros(1), tp])
    fp0 = torch.cat([fp.new_zeros(1), fp])
    area = torch.sum((fp0[1:] - fp0[:-1]) * (tp0[1:] + tp0[:-1]) * 0.5)
    return area / (p * n)


def average_precision(scores: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """sum_k (R_k - R_{k-1}) P_k over distinct thresholds (sklearn's definition)."""
    tp, fp = _grouped_counts(scores, labels)
    precision = tp / (tp + fp)
    recall = tp / tp[-1]
    r0 = torch.cat([recall.new_zeros(1), recall])
    return torch.sum((r0[1:] - r0[:-1]) * precision)
