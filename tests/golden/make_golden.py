#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (``python tests/golden/make_golden.py``); it needs
``/root/reference``.  The committed ``*.npz`` / ``*.json`` files are what the tests read --
nothing under tests/ touches /root/reference at run time.

How the reference is run although torch_geometric / torch_sparse are absent: the
reference's OWN modules (``utils.py``, ``layers.py``, ``models.py``) are imported unmodified
with tiny stand-ins registered in ``sys.modules`` for the un-vendored third-party names they
import.  So everything recorded here that is computed by reference code proper
(``convert_to_matrix``, the networkx edge walk and dtype casts of ``load_input``,
``cont2dist``, ``SAGEConv.forward`` / ``adjust_weights``, the MLP heads of the three live
networks) is a true reference output; what the stand-ins compute (``SparseTensor`` sort /
``to_symmetric`` / ``sum`` / ``matmul``, ``GATConv``) follows the published semantics of
torch-sparse 0.6.11 / torch-geometric 1.7.2 and is flagged "unpinned" in oracle/__init__.py.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


# --------------------------------------------------------------------------- stand-ins
class _Storage:
    def __init__(self, st):
        self._st = st

    def row(self):
        return self._st._row

    def col(self):
        return self._st._col

    def value(self):
        return self._st._value

    def rowptr(self):
        n = self._st._sizes[0]
        ptr = torch.zeros(n + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(torch.bincount(self._st._row, minlength=n), 0)
        return ptr


class SparseTensor:
    """torch-sparse 0.6.11 semantics for the members the reference touches."""

    def __init__(self, row=None, col=None, value=None, sparse_sizes=None, **_):
        if sparse_sizes is None:
            n = int(max(row.max(), col.max())) + 1
            sparse_sizes = (n, n)
        m = sparse_sizes[1]
        key = row * m + col
        if not bool((key[1:] >= key[:-1]).all()):  # ctor sorts row-major when unsorted
            perm = torch.argsort(key, stable=True)
            row, col = row[perm], col[perm]
            value = None if value is None else value[perm]
        self._row, self._col, self._value, self._sizes = row, col, value, tuple(sparse_sizes)
        self.storage = _Storage(self)

    def sizes(self):
        return list(self._sizes)

    def to_symmetric(self, reduce="sum"):
        n = max(self._sizes)
        row = torch.cat([self._row, self._col])
        col = torch.cat([self._col, self._row])
        val = torch.cat([self._value, self._value])
        key = row * n + col
        perm = torch.argsort(key, stable=True)
        key, val = key[perm], val[perm]
        uniq, inv = torch.unique_consecutive(key, return_inverse=True)
        out = torch.zeros(len(uniq), dtype=val.dtype).index_add_(0, inv, val)
        return SparseTensor(row=uniq // n, col=uniq % n, value=out, sparse_sizes=(n, n))

    def sum(self, dim=None):
        assert dim == 0
        return torch.zeros(self._sizes[1], dtype=self._value.dtype).index_add_(0, self._col, self._value)


def sparse_matmul(a, b, reduce="sum"):
    if isinstance(b, SparseTensor):  # diag(a) @ b, as layers.py:53 uses it
        assert bool((a._row == a._col).all()) and len(a._row) == a._sizes[0]
        return SparseTensor(row=b._row, col=b._col, value=a._value[b._row] * b._value, sparse_sizes=b._sizes)
    assert reduce in ("sum", "add")
    out = torch.zeros(a._sizes[0], b.shape[1], dtype=b.dtype)
    return out.index_add_(0, a._row, a._value.to(b.dtype).unsqueeze(1) * b[a._col])


class Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", **_):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kw):
        return self.message_and_aggregate(edge_index, kw["x"])


def _install_stubs():
    from oracle.conv import GATConv as OracleGAT
    from oracle.graph import CSR

    class GATConv(OracleGAT):  # PyG signature -> oracle layer
        def __init__(self, in_channels, out_channels, heads=1, concat=True, **_):
            assert concat
            super().__init__(in_channels, out_channels, heads=heads)

        def forward(self, x, edge_index, **_):
            st = edge_index
            csr = CSR(rowptr=st.storage.rowptr(), col=st._col, value=st._value, n=st._sizes[0])
            return super().forward(x, csr)

    tg = types.ModuleType("torch_geometric")
    tg_data = types.ModuleType("torch_geometric.data")
    tg_data.Data = Data
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_nn.MessagePassing, tg_nn.GATConv, tg_nn.GCNConv = MessagePassing, GATConv, GATConv
    tg_typing = types.ModuleType("torch_geometric.typing")
    for name in ("OptPairTensor", "Adj", "Size", "OptTensor"):
        setattr(tg_typing, name, object)
    ts = types.ModuleType("torch_sparse")
    ts.SparseTensor, ts.matmul = SparseTensor, sparse_matmul
    sys.modules.update(
        {
            "torch_geometric": tg,
            "torch_geometric.data": tg_data,
            "torch_geometric.nn": tg_nn,
            "torch_geometric.typing": tg_typing,
            "torch_sparse": ts,
        }
    )


def read_pdb(path):
    xyz = []
    for line in open(path):
        if line.startswith("ATOM"):
            xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return np.asarray(xyz)


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    import utils as ref_utils  # the reference's utils.py, unmodified
    import layers as ref_layers
    import models as ref_models

    from oracle import kr as okr

    out = {}
    meta = {"generated_from": REF, "torch": torch.__version__, "numpy": np.__version__}
    for tag, fname in (("1mb", "GM12878_1mb_chr19_list.txt"), ("500kb", "GM12878_500kb_chr19_list.txt")):
        lst = np.loadtxt(f"{REF}/Data/{fname}")
        out[f"{tag}_list"] = lst
        mat = ref_utils.convert_to_matrix(lst)  # reference output
        out[f"{tag}_matrix"] = mat
        m0 = mat.copy()
        np.fill_diagonal(m0, 0)  # HiC-GNN_main.py:80
        normed = okr.kr_norm(m0)  # oracle KR (R unavailable) -- an INPUT to what follows
        out[f"{tag}_kr_oracle"] = normed
        g = torch.Generator().manual_seed(7)
        feats = (0.25 * torch.randn(mat.shape[0], 512, generator=g)).numpy()
        data = ref_utils.load_input(normed.copy(), feats)  # reference output
        st = data.edge_index
        out[f"{tag}_csr_rowptr"] = st.storage.rowptr().numpy()
        out[f"{tag}_csr_col"] = st.storage.col().numpy()
        out[f"{tag}_csr_val"] = st.storage.value().numpy()
        out[f"{tag}_y"] = data.y.numpy()
        assert st.storage.value().dtype == torch.float32 and st.storage.col().dtype == torch.int64
        for f in (0.4, 0.5, 1.0):
            out[f"{tag}_wish_{f}"] = ref_utils.cont2dist(data.y.clone(), f).numpy()  # reference output

    # asymmetric / sparse matrix: pins the networkx "lower triangle wins" + zero handling
    rng = np.random.default_rng(5)
    a = rng.random((23, 23)) * (rng.random((23, 23)) < 0.3)
    a[4, :] = 0
    a[:, 4] = 0  # an isolated locus
    out["asym_matrix"] = a
    data = ref_utils.load_input(a.copy(), np.zeros((23, 4), dtype=np.float32))
    st = data.edge_index
    out["asym_csr_rowptr"] = st.storage.rowptr().numpy()
    out["asym_csr_col"] = st.storage.col().numpy()
    out["asym_csr_val"] = st.storage.value().numpy()
    out["asym_y"] = data.y.numpy()
    out["asym_wish_1.0"] = ref_utils.cont2dist(data.y.clone(), 1.0).numpy()

    # duplicate / unordered / lower-triangle records for convert_to_matrix
    lst = np.array(
        [[0, 0, 5.0], [0, 100, 2.0], [100, 300, 3.0], [300, 100, 7.0], [0, 100, 4.0], [500, 500, 0.0], [300, 300, 1.0], [100, 0, 9.0]]
    )
    out["dup_list"] = lst
    out["dup_matrix"] = ref_utils.convert_to_matrix(lst)

    # shipped known answers
    out["pdb_1mb"] = read_pdb(f"{REF}/Outputs/GM12878_1mb_chr19_list_structure.pdb")
    out["pdb_500kb"] = read_pdb(f"{REF}/Outputs/GM12878_500kb_chr19_list_generalized_structure.pdb")
    log1 = open(f"{REF}/Outputs/GM12878_1mb_chr19_list_log.txt").read().split("\n")
    log2 = open(f"{REF}/Outputs/GM12878_500kb_chr19_list_generalized_log.txt").read().split("\n")
    meta["log_1mb"] = {"conversion": float(log1[0].split(":")[1]), "dscc": float(log1[1].split(":")[1]), "mse": float(log1[2].split(":")[1])}
    meta["log_500kb"] = {"conversion": float(log2[0].split(":")[1]), "dscc": float(log2[1].split(":")[1])}
    sd = torch.load(f"{REF}/Outputs/GM12878_1mb_chr19_list_weights.pt", map_location="cpu")
    meta["net_state_dict"] = {k: {"shape": list(v.shape), "dtype": str(v.dtype), "sum": float(v.double().sum()), "absmax": float(v.abs().max())} for k, v in sd.items()}

    # reference models / SAGEConv forward (reference code proper; sparse + GAT ops via stand-ins)
    from oracle import models as omodels

    n = 37
    g = torch.Generator().manual_seed(11)
    dense = torch.rand(n, n, generator=g, dtype=torch.double) * (torch.rand(n, n, generator=g) < 0.4)
    dense = ((dense + dense.t()) / 2).numpy()
    x = 0.25 * torch.randn(n, 512, generator=g)
    x[:, :7] *= 9.0  # some |x| > 1 so that the x.long() root term (layers.py:64) is exercised
    data = ref_utils.load_input(dense.copy(), x.numpy())
    out["model_adj"] = dense
    out["model_x"] = x.numpy()
    for cls in ("Net", "GATNetSelectiveResidualsUpdated", "GATNetHeadsChanged3LayersLeakyReLUv2"):
        torch.manual_seed(42)  # combined_loss_training.py:17
        ref_model = getattr(ref_models, cls)()
        torch.manual_seed(42)
        ora_model = getattr(omodels, cls)()
        missing = ora_model.load_state_dict(ref_model.state_dict(), strict=True)
        with torch.no_grad():
            coords = ref_model.get_model(data.x, data.edge_index)
            distm = ref_model(data.x, data.edge_index)
        out[f"model_{cls}_coords"] = coords.numpy()
        out[f"model_{cls}_dist"] = distm.numpy()
        meta[f"model_{cls}_keys"] = list(ref_model.state_dict().keys())
        meta[f"model_{cls}_nparams"] = sum(p.numel() for p in ref_model.parameters())
    # reference SAGEConv alone (layers.py) with weights that make both terms visible
    torch.manual_seed(3)
    conv = ref_layers.SAGEConv(512, 512)
    with torch.no_grad():
        out["sage_out"] = conv(data.x, data.edge_index).numpy()
        out["sage_norm_val"] = conv.adjust_weights(data.edge_index).storage.value().numpy()

    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "reference_golden.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "reference_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
