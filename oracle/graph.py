"""Oracle (test infrastructure): contact list -> dense matrix -> symmetric CSR graph.

Follows reference ``utils.py:10-26`` (``convert_to_matrix``) and ``utils.py:29-73``
(``load_input``); the ``SparseTensor(...).to_symmetric()`` / ``set_diag`` semantics are
torch-sparse 0.6.11's published behaviour.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


def convert_to_matrix(adj: np.ndarray) -> np.ndarray:
    """``utils.py:10-26``: 3-column ``bin_i bin_j count`` list -> dense symmetric matrix.

    Restated without the per-record ``np.argwhere`` loop (``utils.py:17-20``) but with the
    same "last record wins" assignment order, the same ``triu + tril(mat.T, 1)``
    symmetrisation (``utils.py:21`` -- note ``k=1``: the diagonal is doubled and the first
    super-diagonal receives ``mat[i,i+1] + mat[i+1,i]``), and the same all-zero-column
    removal (``utils.py:22-24``).
    """
    adj = np.asarray(adj, dtype=np.float64)
    ids = np.unique(np.concatenate((adj[:, 0], adj[:, 1])))
    size = len(ids)
    mat = np.zeros((size, size))
    ii = np.searchsorted(ids, adj[:, 0])
    jj = np.searchsorted(ids, adj[:, 1])
    # numpy fancy assignment applies duplicates in order => last record wins, like the loop.
    mat[ii, jj] = adj[:, 2]
    mat = np.triu(mat) + np.tril(mat.T, 1)
    zero_cols = np.argwhere(np.all(mat[..., :] == 0, axis=0))
    mat = np.delete(mat, zero_cols, axis=1)
    mat = np.delete(mat, zero_cols, axis=0)
    return mat


@dataclass
class CSR:
    """What ``data.edge_index`` (a symmetric ``SparseTensor``) holds: ``utils.py:70-71``."""

    rowptr: torch.Tensor  # int64 [N+1]
    col: torch.Tensor  # int64 [nnz]
    value: torch.Tensor  # float32 [nnz]
    n: int

    @property
    def row(self) -> torch.Tensor:
        counts = self.rowptr[1:] - self.rowptr[:-1]
        return torch.repeat_interleave(torch.arange(self.n, dtype=torch.int64), counts)

    def to_dense(self) -> torch.Tensor:
        out = torch.zeros(self.n, self.n, dtype=self.value.dtype)
        out[self.row, self.col] = self.value
        return out


def symmetric_csr_from_dense(adj_mat: np.ndarray) -> CSR:
    """Graph part of ``load_input`` (``utils.py:33-71``) as a closed form.

    networkx (``utils.py:37-39``) visits the non-zero entries row-major and, the graph being
    undirected, a later ``(j,i)`` overwrites the weight stored for ``{i,j}``: an edge exists
    iff ``A[i,j] != 0 or A[j,i] != 0`` and carries ``A[max,min]`` when that is non-zero, else
    ``A[min,max]``.  Self loops are masked out (``utils.py:59-63``; the diagonal was zeroed
    at ``:33`` anyway).  ``SparseTensor(...).to_symmetric()`` (``utils.py:70-71``) then emits
    both directions, row-major sorted, weight cast f64 -> f32 (``utils.py:52``).
    """
    a = np.array(adj_mat, dtype=np.float64, copy=True)
    np.fill_diagonal(a, 0)
    n = a.shape[0]
    lower = np.tril(a, -1)
    upper_t = np.triu(a, 1).T  # upper_t[j,i] = A[i,j] for i<j  (indexed [max,min])
    w_low = np.where(lower != 0, lower, upper_t)  # [max,min] -> weight of edge {min,max}
    sym = w_low + w_low.T
    mask = sym != 0
    # NaNs compare != 0 and would become edges in networkx too; keep that behaviour.
    counts = mask.sum(axis=1)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    rows, cols = np.nonzero(mask)  # row-major order == sorted CSR
    vals = sym[rows, cols]
    return CSR(
        rowptr=torch.from_numpy(rowptr),
        col=torch.from_numpy(cols.astype(np.int64)),
        value=torch.tensor(vals, dtype=torch.float),
        n=n,
    )


@dataclass
class Data:
    """Mirror of the ``torch_geometric.data.Data`` fields the loops use (``utils.py:65``)."""

    x: torch.Tensor
    edge_index: CSR
    y: torch.Tensor
    edge_attr: None = None


def load_input(input, features) -> Data:
    """``utils.py:29-73``."""
    adj_mat = np.asarray(input)
    if adj_mat.shape[1] == 3:
        adj_mat = convert_to_matrix(adj_mat)
    adj_mat = np.array(adj_mat, dtype=np.float64, copy=True)
    np.fill_diagonal(adj_mat, 0)
    truth = torch.tensor(adj_mat, dtype=torch.double)
    csr = symmetric_csr_from_dense(adj_mat)
    return Data(x=torch.tensor(features), edge_index=csr, y=truth)


def set_diag(csr: CSR) -> CSR:
    """torch-sparse ``set_diag`` as GATConv calls it (Appendix A.2): drop existing diagonal
    entries, insert ``(i,i)`` with value 1 for every ``i``, keep row-major sorted order."""
    n = csr.n
    row = csr.row
    keep = row != csr.col
    row = torch.cat([row[keep], torch.arange(n, dtype=torch.int64)])
    col = torch.cat([csr.col[keep], torch.arange(n, dtype=torch.int64)])
    val = torch.cat([csr.value[keep], torch.ones(n, dtype=csr.value.dtype)])
    order = torch.argsort(row * n + col, stable=True)
    row, col, val = row[order], col[order], val[order]
    counts = torch.bincount(row, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return CSR(rowptr=rowptr, col=col, value=val, n=n)
