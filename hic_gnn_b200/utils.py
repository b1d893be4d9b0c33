"""``utils.py`` of the reference, GPU edition: ``load_input`` (utils.py:29-73) and
``cont2dist`` (utils.py:75-80) with the same names, arguments and return fields, plus the helpers
either side of the hot path: ``convert_to_matrix`` (utils.py:10-26), ``domain_alignment`` /
``domain_alignment_filtered`` (utils.py:83-146) and ``WritePDB`` (utils.py:149-192).
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


@dataclass
class SparseData:
    """:class:`Data` of a map that was never densified (:func:`load_input_sparse`): ``y`` does not exist; ``value64`` holds the
    f64 contact value of every stored entry of ``edge_index`` (what ``y[i, j]`` would be), ``bins`` the genomic bin ids kept."""

    x: torch.Tensor
    edge_index: CSRGraph
    value64: torch.Tensor
    bins: torch.Tensor
    y: None = None
    edge_attr: None = None


def load_input_sparse(contacts, features=None, normalize: bool = False, device="cuda") -> SparseData:
    """``utils.load_input`` for a 3-column ``bin_i bin_j count`` list WITHOUT the dense N x N matrix (SURVEY.md section 8 row
    f-2): the composition ``convert_to_matrix`` (utils.py:10-26) -> [``KRnorm``, r_utils.R:1-93, if ``normalize``] -> graph half of
    ``load_input`` (utils.py:33-71), evaluated on the records themselves.  O(E log E) device work (sort / unique / searchsorted +
    the library's CSR kernels); the reference needs an O(E N) Python loop and 8 N^2 bytes (20 GB at 50k loci).

    Same semantics, record for record: bins are compacted in sorted order; a later record overwrites an earlier one for the same
    (i, j); ``triu(mat) + tril(mat.T, 1)`` keeps the records with i <= j (mirrored), doubles the diagonal, and DROPS records below
    the first sub-diagonal; bins whose column is all-zero afterwards are removed; the diagonal is zeroed; an edge {i, j} exists
    iff its value is non-zero and carries ``float32(value)`` in both directions.  Records ON the first sub-diagonal would make the
    reference's matrix asymmetric (they are added onto the super-diagonal only): such a list is refused here -- densify it with
    ``convert_to_matrix`` / ``load_input``, whose targets know how to handle ``T != T^T``.
    Bit-exact with the reference's ``load_input(list, ...)`` (tests/golden/reference_golden_list.npz)."""
    from . import kr as _kr

    a = torch.as_tensor(np.asarray(contacts) if not torch.is_tensor(contacts) else contacts, dtype=torch.float64).to(device)
    n, rowptr, row, col, val, bins = _csr_arrays_from_list(a)
    graph = CSRGraph(rowptr, col.to(torch.int64).contiguous(), val.to(torch.float32), n)
    if normalize:                                          # the Rscript step between the two halves (HiC_GAT_generalize_directly.py:133)
        if n and int((rowptr[1:] == rowptr[:-1]).sum()):
            raise RuntimeError("load_input_sparse(normalize=True): a bin is left without contacts after the diagonal is removed; KRnorm drops "
                               "such rows (r_utils.R:3-10) -- remove them from the list")
        r32, c32 = graph.i32()
        val = _kr.kr_norm_csr(r32, c32, val)
        live = val != 0                                    # entries that round to zero vanish from the graph, like in the dense path
        if not bool(live.all()):
            row, col, val = row[live], col[live], val[live].contiguous()
            rowptr = torch.searchsorted(row.contiguous(), torch.arange(n + 1, device=a.device)).to(torch.int64)
        graph = CSRGraph(rowptr, col.to(torch.int64).contiguous(), val.to(torch.float32), n)
    x = None if features is None else torch.as_tensor(features).to(device)
    return SparseData(x=x, edge_index=graph, value64=val, bins=bins)


def _csr_arrays_from_list(a: torch.Tensor):
    """Device-agnostic core of :func:`load_input_sparse` (pure torch, so the CPU suite can pin it against the reference's own
    output): ``(n, rowptr int64[n+1], row, col, value f64[nnz], kept bin ids)`` of the symmetric, diagonal-free CSR graph."""
    if a.dim() != 2 or a.shape[1] != 3:
        raise ValueError("load_input_sparse expects a 3-column contact list")
    ids = torch.unique(torch.cat((a[:, 0], a[:, 1])))
    size = ids.numel()
    ii = torch.searchsorted(ids, a[:, 0].contiguous())
    jj = torch.searchsorted(ids, a[:, 1].contiguous())
    # "last record wins" per cell (utils.py:17-20): stable sort by cell, keep the last member of every run
    lin = ii * size + jj
    order = torch.argsort(lin, stable=True)
    lin_s = lin[order]
    last = torch.ones_like(lin_s, dtype=torch.bool)
    last[:-1] = lin_s[1:] != lin_s[:-1]
    ci, cj, cv = ii[order][last], jj[order][last], a[:, 2][order][last]
    if bool(((ci == cj + 1) & (cv != 0)).any()):
        raise NotImplementedError("the list holds records on the first sub-diagonal (bin_i one bin after bin_j): the reference's "
                                  "triu(mat) + tril(mat.T, 1) (utils.py:21) makes such a map asymmetric; use convert_to_matrix / load_input")
    up = ci < cj                                           # i > j + 1 is dropped by tril(mat.T, 1); i == j is the (doubled) diagonal
    ui, uj, uv = ci[up], cj[up], cv[up]
    # all-zero columns of the symmetrised matrix (utils.py:22-24): column c is non-zero iff some upper record touching c, or its diagonal, is
    nz = torch.zeros(size, dtype=torch.bool, device=a.device)
    live = uv != 0
    nz[ui[live]] = True
    nz[uj[live]] = True
    dg = (ci == cj) & (cv != 0)
    nz[ci[dg]] = True
    newid = torch.cumsum(nz, 0) - 1
    n = int(nz.sum())
    keep = live & nz[ui] & nz[uj]                          # edges: non-zero value (load_input: sym != 0), diagonal zeroed (utils.py:33)
    ui, uj, uv = newid[ui[keep]], newid[uj[keep]], uv[keep]
    # both directions, row-major sorted (SparseTensor(...).to_symmetric(), utils.py:70-71)
    row = torch.cat((ui, uj))
    col = torch.cat((uj, ui))
    val = torch.cat((uv, uv))
    order = torch.argsort(row * n + col)
    row, col, val = row[order], col[order], val[order].contiguous()
    rowptr = torch.searchsorted(row.contiguous(), torch.arange(n + 1, device=a.device)).to(torch.int64)
    return n, rowptr, row, col, val, ids[nz]


def cont2dist(adj: torch.Tensor, factor: float) -> torch.Tensor:
    """``utils.cont2dist`` (utils.py:75-80): f64 N x N wish distances on the GPU."""
    return ops.cont2dist(adj.contiguous(), factor, want_f64=True, want_f32=False)[0]


def wish_target(adj: torch.Tensor, factor: float) -> ops.WishTarget:
    """Same values as ``cont2dist(adj, factor).float()`` (the per-iteration cast at
    HiC-GNN_main.py:127) written once, directly in the layout the loss kernel streams."""
    return ops.cont2dist(adj.contiguous(), factor, want_f64=False, want_f32=True)[1]


def sparse_wish_target(data, factor: float) -> ops.SparseWishTarget:
    """Implicit form of :func:`wish_target` for sparse maps (row f-4): wish distances only at the
    stored pairs of ``data.edge_index``; every other pair is 1.0, the diagonal 0.  Same loss, no N x N
    array.  ``data``: a :class:`Data` (values gathered from the dense ``y``) or a :class:`SparseData` (its f64 values)."""
    if isinstance(data, SparseData):
        return ops.SparseWishTarget.from_values(data.edge_index, data.value64, factor)
    return ops.SparseWishTarget.from_graph(data.edge_index, data.y, factor)


# ------------------------------------------------------------------------------ generalisation helpers (rows f-3 / f-4)
def _bin_ids(lst, device):
    a = torch.as_tensor(np.asarray(lst) if not torch.is_tensor(lst) else lst, dtype=torch.float64).to(device)
    idx = torch.unique(a[:, 0]).long()                    # np.unique(list[:, 0]).astype(int), utils.py:84,87
    return idx, int((idx[1:] - idx[:-1]).min())           # utils.py:85,88


def domain_alignment(list1, list2, embeddings1, embeddings2, device="cuda", _filtered: bool = False) -> torch.Tensor:
    """``utils.domain_alignment`` (utils.py:83-108): orthogonal Procrustes fit of ``embeddings2`` (the map to
    generalise to, contact list ``list2``) onto ``embeddings1`` (the map the network was trained on) over the bins
    the two resolutions share, applied to all of ``embeddings2``.  Bin matching (``unique`` / ``isin``), the
    cross-covariance GEMM, the SVD and the final product run on the GPU in f64 (library calls around device
    memory: this is data preparation, not the hot path); returns a CUDA f64 tensor."""
    e1 = torch.as_tensor(np.asarray(embeddings1) if not torch.is_tensor(embeddings1) else embeddings1).to(device=device, dtype=torch.float64)
    e2 = torch.as_tensor(np.asarray(embeddings2) if not torch.is_tensor(embeddings2) else embeddings2).to(device=device, dtype=torch.float64)
    idx1, diff1 = _bin_ids(list1, device)
    idx2, diff2 = _bin_ids(list2, device)
    bins = int(diff1 / (2 * diff2))                       # utils.py:90
    a_rows, b_rows = [], []
    for i in range(bins + 1):                             # utils.py:95-100
        shifted = idx2 + i * diff2
        aidx = torch.nonzero(torch.isin(shifted, idx1)).flatten()
        bidx = torch.nonzero(torch.isin(idx1, shifted)).flatten()
        if _filtered:                                     # utils.py:126-131
            aidx, bidx = aidx[aidx < e2.shape[0]], bidx[bidx < e1.shape[0]]
            if aidx.numel() == 0 or bidx.numel() == 0:
                continue
        elif (aidx.numel() and int(aidx.max()) >= e2.shape[0]) or (bidx.numel() and int(bidx.max()) >= e1.shape[0]):
            # numpy raises here in the reference; checked on the host so that no device-side assert can fire
            raise IndexError("domain_alignment: the contact lists name more bins than the embeddings have rows")
        a_rows.append(aidx)
        b_rows.append(bidx)
    if not a_rows:
        raise ValueError("No valid alignment indices found. Check your input data.")  # utils.py:135-136
    a, b = e2[torch.cat(a_rows)], e1[torch.cat(b_rows)]
    if a.shape != b.shape:
        raise ValueError(f"shapes {tuple(a.shape)} and {tuple(b.shape)} of the matched embeddings differ")  # scipy raises alike
    u, _, vh = torch.linalg.svd(a.t() @ b)                # scipy.linalg.orthogonal_procrustes: svd((B^T A)^T), R = U V^T
    return e2 @ (u @ vh)                                  # utils.py:106


def domain_alignment_filtered(list1, list2, embeddings1, embeddings2, device="cuda") -> torch.Tensor:
    """``utils.domain_alignment_filtered`` (utils.py:111-146): same, with out-of-range rows dropped."""
    return domain_alignment(list1, list2, embeddings1, embeddings2, device, _filtered=True)


def WritePDB(positions, pdb_file, ctype="0"):
    """``utils.WritePDB`` (utils.py:149-192): byte-identical output (one leading blank line, ``ATOM`` records with
    ``%.3f`` coordinates in 8-wide fields, ``CONECT i i+1`` records -- including the dangling last one unless
    ``ctype == "1"`` --, ``END`` without a newline).  Accepts a CUDA / CPU tensor or an array (one device->host copy)."""
    pos = positions.detach().cpu().numpy() if torch.is_tensor(positions) else np.asarray(positions)
    n = len(pos)
    out = ["\n"]
    out += ["ATOM  %5s   CA MET %-6s   %8s%8s%8s  0.20 10.00\n" % (i, "B%d" % i, "%.3f" % p[0], "%.3f" % p[1], "%.3f" % p[2]) for i, p in enumerate(pos, 1)]
    last = n - 1 if ctype == "1" else n
    out += ["CONECT%5s%5s\n" % (i, i + 1) for i in range(1, last + 1)]
    out.append("END")
    with open(pdb_file, "w") as f:
        f.write("".join(out))
