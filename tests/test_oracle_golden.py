"""Pin the oracle against the reference's own outputs (tests/golden/make_golden.py) and the
reference's shipped known answers.  CPU only."""
import numpy as np
import pytest
import torch
from scipy.stats import spearmanr

from oracle import conv as oconv
from oracle import graph as ograph
from oracle import kr as okr
from oracle import models as omodels
from oracle import wish as owish


@pytest.mark.parametrize("tag,n", [("1mb", 58), ("500kb", 114)])
def test_convert_to_matrix_matches_reference(golden, tag, n):
    g, _ = golden
    mat = ograph.convert_to_matrix(g[f"{tag}_list"])
    assert mat.shape == (n, n)
    np.testing.assert_array_equal(mat, g[f"{tag}_matrix"])  # bit-exact


def test_convert_to_matrix_duplicates_and_lower_records(golden):
    g, _ = golden
    np.testing.assert_array_equal(ograph.convert_to_matrix(g["dup_list"]), g["dup_matrix"])


@pytest.mark.parametrize("tag", ["1mb", "500kb", "asym"])
def test_load_input_csr_bit_exact(golden, tag):
    g, _ = golden
    src = g["asym_matrix"] if tag == "asym" else g[f"{tag}_kr_oracle"]
    data = ograph.load_input(src.copy(), np.zeros((src.shape[0], 2), dtype=np.float32))
    csr = data.edge_index
    assert csr.rowptr.dtype == torch.int64 and csr.col.dtype == torch.int64 and csr.value.dtype == torch.float32
    np.testing.assert_array_equal(csr.rowptr.numpy(), g[f"{tag}_csr_rowptr"])
    np.testing.assert_array_equal(csr.col.numpy(), g[f"{tag}_csr_col"])
    np.testing.assert_array_equal(csr.value.numpy().view(np.uint32), g[f"{tag}_csr_val"].view(np.uint32))
    np.testing.assert_array_equal(data.y.numpy(), g[f"{tag}_y"])


@pytest.mark.parametrize("tag,factors", [("1mb", (0.4, 0.5, 1.0)), ("500kb", (0.4, 0.5, 1.0)), ("asym", (1.0,))])
def test_cont2dist_bit_exact(golden, tag, factors):
    g, _ = golden
    y = torch.tensor(g[f"{tag}_y"])
    for f in factors:
        got = owish.cont2dist(y.clone(), f).numpy()
        np.testing.assert_array_equal(got.view(np.uint64), g[f"{tag}_wish_{f}"].view(np.uint64))
    if tag == "asym":  # zero contacts -> exactly 1.0, diagonal -> 0
        w = g["asym_wish_1.0"]
        assert w.max() == 1.0 and (np.diag(w) == 0).all() and (w[g["asym_y"] == 0].ravel()[1:] >= 0).all()


def test_kr_balances_and_known_answer_1mb(golden):
    """KR -> cont2dist(0.4) vs cdist(shipped PDB / 100) reproduces the shipped log."""
    g, meta = golden
    m = g["1mb_matrix"].copy()
    np.fill_diagonal(m, 0)
    normed = okr.kr_norm(m)
    np.testing.assert_array_equal(normed, g["1mb_kr_oracle"])
    assert np.abs(normed.sum(1) - 1).max() < 5e-5
    assert np.abs(okr.kr_norm(m, literal_typo=False) - normed).max() <= 1.0000001e-6
    truth = owish.cont2dist(torch.tensor(normed), meta["log_1mb"]["conversion"])
    coords = torch.tensor(g["pdb_1mb"] / 100.0)
    d = torch.cdist(coords, coords)
    iu = np.triu_indices(58, 1)
    rho = spearmanr(truth.numpy()[iu], d.numpy()[iu])[0]
    assert abs(rho - meta["log_1mb"]["dscc"]) < 1e-5  # PDB is rounded to 3 decimals
    mse = float(((d.float() - truth.float()) ** 2).mean())
    assert abs(mse - meta["log_1mb"]["mse"]) / meta["log_1mb"]["mse"] < 2e-3


def test_known_answer_500kb(golden):
    g, meta = golden
    truth = owish.cont2dist(torch.tensor(g["500kb_kr_oracle"]), meta["log_500kb"]["conversion"])
    coords = torch.tensor(g["pdb_500kb"] / 100.0)
    d = torch.cdist(coords, coords)
    iu = np.triu_indices(114, 1)
    rho = spearmanr(truth.numpy()[iu], d.numpy()[iu])[0]
    assert abs(rho - meta["log_500kb"]["dscc"]) < 1e-5


def test_state_dict_layout_matches_shipped_weights(golden):
    _, meta = golden
    sd = omodels.Net().state_dict()
    ref = meta["net_state_dict"]
    assert sorted(sd.keys()) == sorted(ref.keys())
    for k, v in sd.items():
        assert list(v.shape) == ref[k]["shape"], k
    assert sum(p.numel() for p in omodels.Net().parameters()) == 697475


@pytest.mark.parametrize(
    "cls,nparams",
    [("Net", 697475), ("GATNetSelectiveResidualsUpdated", 601475), ("GATNetHeadsChanged3LayersLeakyReLUv2", 411651)],
)
def test_models_match_reference_forward(golden, cls, nparams):
    """Reference models.py forward/get_model (run for real in make_golden.py) vs oracle."""
    g, meta = golden
    data = ograph.load_input(g["model_adj"].copy(), g["model_x"])
    torch.manual_seed(42)
    model = getattr(omodels, cls)()
    assert list(model.state_dict().keys()) == meta[f"model_{cls}_keys"]
    assert sum(p.numel() for p in model.parameters()) == nparams == meta[f"model_{cls}_nparams"]
    with torch.no_grad():
        coords = model.get_model(data.x, data.edge_index)
        dist = model(data.x, data.edge_index)
    np.testing.assert_allclose(coords.numpy(), g[f"model_{cls}_coords"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dist.numpy(), g[f"model_{cls}_dist"], rtol=1e-5, atol=1e-6)


def test_sageconv_matches_reference_layer(golden):
    g, _ = golden
    data = ograph.load_input(g["model_adj"].copy(), g["model_x"])
    torch.manual_seed(3)
    conv = oconv.SAGEConv(512, 512)
    with torch.no_grad():
        out = conv(data.x, data.edge_index)
    np.testing.assert_allclose(out.numpy(), g["sage_out"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(oconv.SAGEConv.adjust_weights(data.edge_index).numpy(), g["sage_norm_val"])
    # the root term really is integer-truncated (layers.py:64)
    with torch.no_grad():
        no_trunc = oconv.SAGEConv(512, 512, trunc_root=False)
        no_trunc.load_state_dict(conv.state_dict())
        assert (no_trunc(data.x, data.edge_index) - out).abs().max() > 1e-3


def test_gat_dense_and_edge_formulations_agree(golden):
    g, _ = golden
    data = ograph.load_input(g["model_adj"].copy(), g["model_x"])
    torch.manual_seed(0)
    conv = oconv.GATConv(512, 256, heads=2)
    x = data.x.clone().requires_grad_(True)
    a = conv(x, data.edge_index, dense=False)
    ga = torch.autograd.grad(a.square().sum(), [x, conv.att_l, conv.lin_l.weight])
    b = conv(x, data.edge_index, dense=True)
    gb = torch.autograd.grad(b.square().sum(), [x, conv.att_l, conv.lin_l.weight])
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    for u, v in zip(ga, gb):
        torch.testing.assert_close(u, v, rtol=1e-4, atol=1e-5)
    att = conv.attention(x, data.edge_index)
    gd = ograph.set_diag(data.edge_index)
    sums = torch.zeros(gd.n, 2).index_add_(0, gd.row, att)
    torch.testing.assert_close(sums, torch.ones_like(sums), rtol=1e-5, atol=1e-6)
