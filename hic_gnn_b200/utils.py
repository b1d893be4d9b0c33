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


def load_input(input, features, device="cuda") -> Data:
    """``utils.load_input`` (utils.py:29-73).  ``input``: dense N x N contact matrix (numpy or
    tensor); the 3-column list form must be densified by the caller (``convert_to_matrix`` is
    host-side pre-processing outside the hot path).  The graph is built on the GPU by the CSR
    kernels, bit-exact with the networkx + SparseTensor path of the reference."""
    adj = torch.as_tensor(np.asarray(input) if not torch.is_tensor(input) else input, dtype=torch.float64)
    if adj.dim() != 2 or adj.shape[0] != adj.shape[1]:
        raise ValueError("load_input expects the dense N x N matrix (run convert_to_matrix on list input first)")
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
