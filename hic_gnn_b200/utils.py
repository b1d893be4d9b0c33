"""``utils.py`` of the reference, GPU edition: ``load_input`` (utils.py:29-73) and
``cont2dist`` (utils.py:75-80) with the same names, arguments and return fields.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import ops
from .graph import CSRGraph


@dataclass
class Data:
    """The fields of ``torch_geometric.data.Data`` the loops touch (utils.py:65-72)."""

    x: torch.Tensor
    edge_index: CSRGraph
    y: torch.Tensor
    edge_attr: None = None


def convert_to_matrix(adj, device="cuda") -> torch.Tensor:
    """``utils.convert_to_matrix`` (utils.py:10-26): 3-column ``bin_i bin_j count`` list -> dense
    symmetric f64 matrix on the GPU ("next").  Same semantics as the
    reference loop: bins are compacted in sorted order, a later record overwrites an earlier one
    for the same (i, j), symmetrisation is ``triu(mat) + tril(mat.T, 1)`` (utils.py:21: the diagonal
    doubles and the first super-diagonal receives both orientations), all-zero columns and the
    matching rows are dropped (utils.py:22-24).  Torch glue around device memory: no O(E*N) loop."""
    a = torch.as_tensor(np.asarray(adj) if not torch.is_tensor(adj) else adj, dtype=torch.float64).to(device)
    ids = torch.unique(torch.cat((a[:, 0], a[:, 1])))
    size = ids.numel()
    ii = torch.searchsorted(ids, a[:, 0].contiguous())
    jj = torch.searchsorted(ids, a[:, 1].contiguous())
    lin = ii * size + jj
    # "last record wins": stable sort by cell, keep the last member of every run
    order = torch.argsort(lin, stable=True)
    lin_s = lin[order]
    last = torch.ones_like(lin_s, dtype=torch.bool)
    last[:-1] = lin_s[1:] != lin_s[:-1]
    mat = torch.zeros(size * size, dtype=torch.float64, device=a.device)
    mat[lin_s[last]] = a[:, 2][order][last]
    mat = mat.view(size, size)
    mat = torch.triu(mat) + torch.tril(mat.t(), 1)
    keep = ~(mat == 0).all(dim=0)
    return mat[keep][:, keep].contiguous()


def load_input(input, features, device="cuda") -> Data:
    """``utils.load_input`` (utils.py:29-73).  ``input``: dense N x N contact matrix (numpy or
    tensor); the 3-column list form must be densified by the caller (``convert_to_matrix`` is
    host-side pre-processing outside the hot path).  The graph is built on the GPU by the CSR
    kernels, bit-exact with the networkx + SparseTensor path of the reference."""
    adj = torch.as_tensor(np.asarray(input) if not torch.is_tensor(input) else input, dtype=torch.float64)
    if adj.dim() == 2 and adj.shape[1] == 3 and adj.shape[0] != 3:  # utils.py:30-31: list input
        adj = convert_to_matrix(adj, device)
    if adj.dim() != 2 or adj.shape[0] != adj.shape[1]:
        raise ValueError("load_input expects a 3-column contact list or the dense N x N matrix")
    adj = adj.to(device).clone()
    adj.fill_diagonal_(0)  # utils.py:33
    rowptr, col, val = ops.csr_from_dense(adj)
    x = torch.as_tensor(features).to(device)
    return Data(x=x, edge_index=CSRGraph(rowptr, col, val, adj.shape[0]), y=adj)


def cont2dist(adj: torch.Tensor, factor: float) -> torch.Tensor:
    """``utils.cont2dist`` (utils.py:75-80): f64 N x N wish distances on the GPU."""
    return ops.cont2dist(adj.contiguous(), factor, want_f64=True, want_f32=False)[0]


def wish_target(adj: torch.Tensor, factor: float) -> ops.WishTarget:
    """Same values as ``cont2dist(adj, factor).float()`` (the per-iteration cast at
    HiC-GNN_main.py:127) written once, directly in the layout the loss kernel streams."""
    return ops.cont2dist(adj.contiguous(), factor, want_f64=False, want_f32=True)[1]


def sparse_wish_target(data: Data, factor: float) -> ops.SparseWishTarget:
    """Implicit form of :func:`wish_target` for sparse maps (row f-4): wish distances only at the
    stored pairs of ``data.edge_index``; every other pair is 1.0, the diagonal 0.  Same loss, no N x N
    array."""
    return ops.SparseWishTarget.from_graph(data.edge_index, data.y, factor)
