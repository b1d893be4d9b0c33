"""Evaluation metrics of the reference on the GPU ("next").

``dscc`` is the reference's headline accuracy number: the Spearman correlation between the
strict-upper-triangle wish distances and the reconstructed pairwise distances
(``scipy.stats.spearmanr``, HiC-GNN_main.py:135-139, HiC_GAT_generalize_directly.py:242) with
AVERAGE ranks for ties -- every zero-contact pair has wish distance exactly 1.0, so ties are the
rule, not the exception.  The distances come from the library's pair-distance kernel; the rank
transform is sort-based torch glue (evaluation runs once per model, not per step).
"""
from __future__ import annotations

import torch

from . import ops


def average_ranks(v: torch.Tensor) -> torch.Tensor:
    """``scipy.stats.rankdata(v, method="average")`` as f64, on the device of ``v``."""
    sv, order = torch.sort(v)
    _, inverse, counts = torch.unique_consecutive(sv, return_inverse=True, return_counts=True)
    last = torch.cumsum(counts, 0).to(torch.float64)          # 1-based rank of the last member of each tie group
    avg = last - (counts.to(torch.float64) - 1.0) / 2.0
    ranks = torch.empty(v.numel(), dtype=torch.float64, device=v.device)
    ranks[order] = avg[inverse]
    return ranks


def _pearson(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a = a - a.mean()
    b = b - b.mean()
    return (a * b).sum() / torch.sqrt((a * a).sum() * (b * b).sum())


def upper_pairs(coords: torch.Tensor, truth: torch.Tensor):
    """(dist_truth, dist_out) over i<j like the reference's ``triu_indices`` gathers
    (HiC_GAT_generalize_directly.py:210-214); ``truth`` is the dense N x N wish matrix (any float
    dtype) or a full-matrix :class:`ops.WishTarget`."""
    if isinstance(truth, ops.WishTarget):
        if truth.r0 != 0 or truth.r1 != truth.n:
            raise RuntimeError("dscc needs the full target (gather the row blocks first)")
        truth = truth.dense()
    n = truth.shape[0]
    idx = torch.triu_indices(n, n, 1, device=truth.device)
    d = ops.pairdist(coords.detach().float())
    return truth[idx[0], idx[1]], d[idx[0], idx[1]]


def dscc(coords: torch.Tensor, truth) -> float:
    """Distance Spearman correlation coefficient of a reconstructed structure (CUDA tensors)."""
    t, d = upper_pairs(coords, truth)
    return float(_pearson(average_ranks(t.double()), average_ranks(d.double())))


def pearson(coords: torch.Tensor, truth) -> float:
    """``scipy.stats.pearsonr`` of the same two vectors (HiC_GAT_generalize_directly.py:220)."""
    t, d = upper_pairs(coords, truth)
    return float(_pearson(t.double(), d.double()))
