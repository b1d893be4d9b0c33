"""Oracle (test infrastructure): the reference's full-batch training loops.

``mode`` selects which script's loop body is replayed:

* ``"mse"``          -- ``HiC-GNN_main.py:117-132`` (also ``train_and_test_on_same_res.py``)
* ``"mse_pearson"``  -- ``HiC_GAT_generalize_directly.py:186-243``
* ``"mse_spearman"`` -- ``combined_loss_training.py:96-152``
* ``"contrastive"``  -- ``train_and_test_same_res_GAT_node2vec.py:93-143``

``as_written=True`` replays every op of the script body (second forward, per-iteration
``triu_indices`` / scipy correlations); ``False`` keeps only the ops that determine the
parameter trajectory (one forward, the differentiable loss, backward, Adam) -- the
trajectories are identical because the dropped terms are constants w.r.t. autograd.
"""
from __future__ import annotations

import torch
from scipy.stats import spearmanr
from torch.optim import Adam

from . import loss as L


def train(
    model,
    x: torch.Tensor,
    edge_index,
    truth: torch.Tensor,
    mode: str = "mse",
    lr: float = 1e-3,
    thresh: float = 1e-8,
    max_steps: int | None = None,
    alpha: float = 1.0,
    as_written: bool = True,
):
    """Returns ``(loss_history, extras)``; stops when ``|old - new| <= thresh`` (the
    reference's ``while lossdiff > thresh``) or after ``max_steps``."""
    optimizer = Adam(model.parameters(), lr=lr)
    old, diff = 1.0, 1.0
    hist, extras = [], []
    step = 0
    while diff > thresh and (max_steps is None or step < max_steps):
        model.train()
        optimizer.zero_grad()
        if mode == "mse":
            out = model(x, edge_index)  # HiC-GNN_main.py:126
            total = torch.nn.MSELoss()(out.float(), truth.float())
            extra = {}
        elif mode == "mse_pearson":
            if as_written:
                out = model(x, edge_index)  # forward #1 (:206)
                coords = model.get_model(x, edge_index)  # forward #2 (:213)
                mse = torch.nn.MSELoss()(out.float(), truth.float())
                dist_truth, dist_out = L.triu_pairs(truth, coords)
                from scipy.stats import pearsonr

                r, _ = pearsonr(dist_truth.detach().numpy(), dist_out.detach().numpy())
                a = min(1.0, 0.1 + (1.0 / (mse.item() + 1e-6)))
                total = mse + a * (1 - r)
                rho = spearmanr(dist_truth, dist_out.detach().numpy())[0]  # :242
                extra = {"mse": mse.item(), "pearson": float(r), "alpha": a, "dscc": float(rho)}
            else:
                coords = model.get_model(x, edge_index)
                total, mse, r, a = L.mse_pearson_loss(coords, truth)
                extra = {"mse": mse.item(), "pearson": r, "alpha": a}
        elif mode == "mse_spearman":
            coords = model.get_model(x, edge_index)
            if as_written:
                model(x, edge_index)
            total, mse, rho, drmsd = L.mse_spearman_loss(coords, truth, alpha)
            extra = {"mse": mse.item(), "dscc": rho, "drmsd": drmsd}
        elif mode == "contrastive":
            if as_written:
                model(x, edge_index)  # unused forward #1 (:106)
            coords = model.get_model(x, edge_index)
            total = L.contrastive_loss(coords, truth)
            extra = {}
        else:
            raise ValueError(mode)
        value = float(total.detach())
        diff = abs(old - value)
        total.backward()
        optimizer.step()
        old = value
        hist.append(old)
        extras.append(extra)
        step += 1
    return hist, extras
