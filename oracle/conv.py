"""Oracle (test infrastructure): the two graph-conv layers on the hot path.

* ``SAGEConv`` follows the reference's own ``layers.py:12-83`` (incl. the ``x.long()``
  truncation of the root term at ``layers.py:64`` and the per-call re-normalisation at
  ``layers.py:41-54``); the sparse ops it calls (``SparseTensor.sum``, ``matmul``) are
  torch-sparse 0.6.11 (un-vendored) and are restated from their published behaviour.
* ``GATConv`` restates torch-geometric 1.7.2's layer (un-vendored; constructor call sites
  ``models.py:619`` / ``models.py:1013``): SURVEY.md Appendix A.3.  **Parity unpinned**
  (no reference test or fixture exercises it).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .graph import CSR, set_diag


def _glorot_(t: torch.Tensor) -> None:
    """torch_geometric.nn.inits.glorot: U(-s, s), s = sqrt(6 / (size(-2) + size(-1)))."""
    s = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-s, s)


class SAGEConv(nn.Module):
    """``layers.py:12-83``.  ``reset_parameters`` is referenced but never called
    (``layers.py:34``), so both Linears keep ``nn.Linear``'s default init."""

    def __init__(self, in_channels: int, out_channels: int, trunc_root: bool = True):
        super().__init__()
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)  # layers.py:30
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)  # layers.py:32
        self.trunc_root = trunc_root

    @staticmethod
    def adjust_weights(csr: CSR) -> torch.Tensor:
        """``layers.py:41-54``: ``D^-1 A`` with ``D = diag(column sums)``; returns the
        normalised edge values in CSR order (f32)."""
        n = csr.n
        # adj_t.sum(dim=0): per-column sums, accumulated in storage (row-major) order, f32
        sum_vec = torch.zeros(n, dtype=csr.value.dtype).index_add_(0, csr.col, csr.value)
        inverse = torch.divide(torch.ones(n), sum_vec).float()  # layers.py:44-51
        # matmul(diag(inverse), adj_t): entry (i,j) -> inverse[i] * A[i,j]
        return inverse[csr.row] * csr.value

    def aggregate(self, x: torch.Tensor, csr: CSR) -> torch.Tensor:
        """``layers.py:75-79``: ``matmul(norm_mat, x, reduce='add')``."""
        norm_val = self.adjust_weights(csr)
        out = torch.zeros(csr.n, x.shape[1], dtype=x.dtype)
        return out.index_add_(0, csr.row, norm_val.to(x.dtype).unsqueeze(1) * x[csr.col])

    def forward(self, x: torch.Tensor, csr: CSR) -> torch.Tensor:
        out = self.aggregate(x, csr)  # layers.py:62
        out = self.lin_l(out.float())  # layers.py:63
        x_r = x.long() if self.trunc_root else x  # layers.py:64
        out = out + self.lin_r(x_r.float())  # layers.py:66-67
        return out


class GATConv(nn.Module):
    """torch-geometric 1.7.2 ``GATConv(in, C, heads=H, concat=True)`` fed a SparseTensor.

    ``lin_r`` IS ``lin_l`` (one shared bias-free Linear; both keys appear in the
    state_dict), ``add_self_loops=True`` via ``set_diag``, edge values are ignored,
    ``negative_slope=0.2``, softmax ``exp(e - max) / (sum + 1e-16)``, attention dropout 0.
    """

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.lin_l = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.lin_r = self.lin_l
        self.att_l = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_r = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels))
        self.negative_slope = 0.2
        self.reset_parameters()

    def reset_parameters(self) -> None:
        _glorot_(self.lin_l.weight)
        _glorot_(self.lin_r.weight)  # same tensor, re-drawn (RNG order of 1.7.2)
        _glorot_(self.att_l)
        _glorot_(self.att_r)
        with torch.no_grad():
            self.bias.zero_()

    def _project(self, x):
        H, C = self.heads, self.out_channels
        x_l = self.lin_l(x).view(-1, H, C)
        alpha_l = (x_l * self.att_l).sum(dim=-1)  # source term, indexed by col j
        alpha_r = (x_l * self.att_r).sum(dim=-1)  # target term, indexed by row i
        return x_l, alpha_l, alpha_r

    def forward(self, x: torch.Tensor, csr: CSR, dense: bool = False) -> torch.Tensor:
        """``dense=False``: literal PyG message/aggregate order (materialises E x H x C).
        ``dense=True``: mathematically identical masked-dense formulation for large,
        near-dense graphs (N x N x H intermediates only)."""
        H, C = self.heads, self.out_channels
        n = x.shape[0]
        x_l, alpha_l, alpha_r = self._project(x)
        g = set_diag(csr)
        if dense:
            mask = torch.zeros(n, n, dtype=torch.bool)
            mask[g.row, g.col] = True
            e = alpha_l.t().unsqueeze(1) + alpha_r.t().unsqueeze(2)  # [H, i, j]
            e = F.leaky_relu(e, self.negative_slope)
            e = e.masked_fill(~mask.unsqueeze(0), float("-inf"))
            m = e.max(dim=2, keepdim=True).values
            p = (e - m).exp()
            a = p / (p.sum(dim=2, keepdim=True) + 1e-16)
            out = torch.einsum("hij,jhc->ihc", a, x_l)
        else:
            row, col = g.row, g.col
            e = F.leaky_relu(alpha_l[col] + alpha_r[row], self.negative_slope)  # [E,H]
            idx = row.unsqueeze(1).expand(-1, H)
            m = torch.full((n, H), float("-inf"), dtype=e.dtype).scatter_reduce(
                0, idx, e, reduce="amax", include_self=True
            )
            p = (e - m[row]).exp()
            s = torch.zeros(n, H, dtype=e.dtype).index_add_(0, row, p)
            a = p / (s[row] + 1e-16)
            msg = x_l[col] * a.unsqueeze(-1)  # [E,H,C]  (what PyG materialises)
            out = torch.zeros(n, H, C, dtype=x_l.dtype).index_add_(0, row, msg)
        return out.reshape(n, H * C) + self.bias

    def forward_rows(self, x: torch.Tensor, csr: CSR, r1: int) -> torch.Tensor:
        """Literal PyG message/aggregate order restricted to the TARGET rows ``[0, r1)`` (all source nodes):
        output rows ``[0, r1)`` of :meth:`forward`.  Used by ``bench.py``'s CPU baseline on maps whose
        E x H x C message tensor (what PyG materialises) does not fit host memory: cost is proportional to the
        edges of the slice, so a slice timing extrapolates by edge count."""
        H, C = self.heads, self.out_channels
        x_l, alpha_l, alpha_r = self._project(x)
        g = set_diag(csr)  # recomputed on every forward, like GATConv does
        k1 = int(g.rowptr[r1])
        row, col = g.row[:k1], g.col[:k1]
        e = F.leaky_relu(alpha_l[col] + alpha_r[row], self.negative_slope)
        idx = row.unsqueeze(1).expand(-1, H)
        m = torch.full((r1, H), float("-inf"), dtype=e.dtype).scatter_reduce(0, idx, e, reduce="amax", include_self=True)
        p = (e - m[row]).exp()
        s = torch.zeros(r1, H, dtype=e.dtype).index_add_(0, row, p)
        a = p / (s[row] + 1e-16)
        msg = x_l[col] * a.unsqueeze(-1)
        out = torch.zeros(r1, H, C, dtype=x_l.dtype).index_add_(0, row, msg)
        return out.reshape(r1, H * C) + self.bias

    def attention(self, x: torch.Tensor, csr: CSR) -> torch.Tensor:
        """Per-edge attention coefficients [nnz+N, H] in ``set_diag`` CSR order."""
        x_l, alpha_l, alpha_r = self._project(x)
        g = set_diag(csr)
        row, col = g.row, g.col
        H = self.heads
        e = F.leaky_relu(alpha_l[col] + alpha_r[row], self.negative_slope)
        idx = row.unsqueeze(1).expand(-1, H)
        m = torch.full((g.n, H), float("-inf"), dtype=e.dtype).scatter_reduce(
            0, idx, e, reduce="amax", include_self=True
        )
        p = (e - m[row]).exp()
        s = torch.zeros(g.n, H, dtype=e.dtype).index_add_(0, row, p)
        return p / (s[row] + 1e-16)
