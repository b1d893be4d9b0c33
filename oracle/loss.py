"""Oracle (test infrastructure): the loss / metric glue of the reference loops.

Each function is the body of one reference loop, op for op (``torch.cdist``,
``MSELoss``, ``triu_indices`` gathers, scipy correlations on detached numpy).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.stats import pearsonr, spearmanr
from torch.nn import MSELoss


def triu_pairs(truth: torch.Tensor, coords: torch.Tensor):
    """``HiC_GAT_generalize_directly.py:210-214``: flatten the strict upper triangle."""
    n = truth.shape[0]
    idx = torch.triu_indices(n, n, offset=1)
    dist_truth = truth[idx[0, :], idx[1, :]]
    dist_out = torch.cdist(coords, coords)[idx[0, :], idx[1, :]]
    return dist_truth, dist_out


def mse_loss(coords: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    """``HiC-GNN_main.py:126-127`` with ``out = cdist(coords, coords)`` (``models.py:39``):
    mean over the FULL N x N matrix, diagonal and both triangles."""
    out = torch.cdist(coords, coords, p=2)
    return MSELoss()(out.float(), truth.float())


def mse_pearson_loss(coords: torch.Tensor, truth: torch.Tensor):
    """``HiC_GAT_generalize_directly.py:206-225``.  Returns (total, mse, PearsonR, alpha).
    ``PearsonR`` is a Python float: it shifts the value but carries no gradient."""
    mse = mse_loss(coords, truth)
    dist_truth, dist_out = triu_pairs(truth, coords)
    r, _ = pearsonr(dist_truth.detach().numpy(), dist_out.detach().numpy())
    alpha = min(1.0, 0.1 + (1.0 / (mse.item() + 1e-6)))
    total = mse + alpha * (1 - r)
    return total, mse, float(r), alpha


def mse_spearman_loss(coords: torch.Tensor, truth: torch.Tensor, alpha: float = 1.0):
    """``combined_loss_training.py:119-142``.  Returns (total, mse, SpRho, dRMSD)."""
    mse = mse_loss(coords, truth)
    dist_truth, dist_out = triu_pairs(truth, coords)
    rho, _ = spearmanr(dist_truth.detach().numpy(), dist_out.detach().numpy())
    if np.isnan(rho):
        rho = 0
    drmsd = torch.sqrt(torch.mean(torch.pow(dist_truth - dist_out, 2))).item()
    total = mse + alpha * (1 - rho)
    return total, mse, float(rho), drmsd


def contrastive_loss(coords: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    """``train_and_test_same_res_GAT_node2vec.py:108-134``: ``0.1 * mean |t - d|`` over
    ``i < j`` (result is f64 because ``dist_truth`` is)."""
    dist_truth, dist_out = triu_pairs(truth, coords)
    return 0.0 + 0.1 * torch.mean(torch.abs(dist_truth - dist_out))


def dscc(coords: torch.Tensor, truth: torch.Tensor) -> float:
    """``HiC-GNN_main.py:135-139``: Spearman correlation of upper-triangle distances."""
    dist_truth, dist_out = triu_pairs(truth, coords)
    return float(spearmanr(dist_truth.detach().numpy(), dist_out.detach().numpy())[0])


def pearson(coords: torch.Tensor, truth: torch.Tensor) -> float:
    dist_truth, dist_out = triu_pairs(truth, coords)
    return float(pearsonr(dist_truth.detach().numpy(), dist_out.detach().numpy())[0])


def mse_loss_rows(coords: torch.Tensor, truth_rows: torch.Tensor, r0: int, r1: int) -> torch.Tensor:
    """Rows ``[r0, r1)`` of :func:`mse_loss` (same ops: ``torch.cdist`` -> ``MSELoss``), for map
    sizes whose N x N matrices the reference formulation cannot hold (it needs >125 GB of host
    memory at 50k loci).  Summing ``value * (r1-r0)/N`` over a row partition gives ``mse_loss``;
    ``bench.py`` times it on a bounded row sample as the CPU baseline."""
    out = torch.cdist(coords[r0:r1], coords, p=2)
    return MSELoss()(out.float(), truth_rows.float())
