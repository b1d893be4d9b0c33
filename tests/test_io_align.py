"""Rows either side of the hot path: Procrustes ``domain_alignment`` and the PDB
writer, against outputs of the reference's own ``utils.py`` (tests/golden/make_golden_io.py)."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def io_golden():
    return np.load(os.path.join(HERE, "golden", "reference_golden_io.npz"))


ALIGN_CASES = [("align_fit", "1mb", "500kb", "align_emb1", "align_emb2", False), ("align_fit_filtered", "1mb", "500kb", "align_emb1", "align_emb2", True),
               ("align_fit_swapped", "500kb", "1mb", "align_emb2", "align_emb1", True)]


@pytest.mark.parametrize("key,t1,t2,e1,e2,filtered", ALIGN_CASES)
def test_oracle_domain_alignment_matches_reference(golden, io_golden, key, t1, t2, e1, e2, filtered):
    from oracle import align

    arrays, _ = golden
    got = align.domain_alignment(arrays[f"{t1}_list"], arrays[f"{t2}_list"], io_golden[e1], io_golden[e2], filtered=filtered)
    assert got.shape == io_golden[key].shape
    assert np.abs(got - io_golden[key]).max() < 1e-12


def test_oracle_alignment_recovers_a_planted_rotation(golden):
    """Property: embeddings2 = embeddings1 (interleaved) @ Q  =>  the fit maps them back onto embeddings1."""
    from oracle import align

    arrays, _ = golden
    l1, l2 = arrays["1mb_list"], arrays["500kb_list"]
    rng = np.random.default_rng(0)
    idx1, idx2 = np.unique(l1[:, 0]).astype(int), np.unique(l2[:, 0]).astype(int)
    e1 = rng.standard_normal((len(idx1), 16))
    q, _ = np.linalg.qr(rng.standard_normal((16, 16)))
    e2 = rng.standard_normal((len(idx2), 16))
    a_rows, b_rows = align.matched_rows(l1, l2)
    e2[a_rows] = e1[b_rows] @ q.T       # matched bins carry the rotated trained embedding
    fit = align.domain_alignment(l1, l2, e1, e2)
    assert np.abs(fit[a_rows] - e1[b_rows]).max() < 1e-10


@pytest.mark.parametrize("pos_key,text_key,ctype", [("pdb_1mb", "pdb_text_1mb", "0"), ("pdb_500kb", "pdb_text_500kb", "0"),
                                                    ("pdb_rand_pos", "pdb_rand_text_c0", "0"), ("pdb_rand_pos", "pdb_rand_text_c1", "1")])
def test_write_pdb_is_byte_identical(golden, io_golden, tmp_path, pos_key, text_key, ctype):
    """Oracle restatement and the shipped host writer, against the reference's bytes; the two chr19 cases are the
    reference's SHIPPED Outputs/*_structure.pdb files (make_golden_io.py asserts WritePDB(read(shipped)) == shipped)."""
    from hic_gnn_b200 import utils
    from oracle import align

    arrays, _ = golden
    pos = arrays[pos_key] if pos_key in arrays.files else io_golden[pos_key]
    want = io_golden[text_key].tobytes()
    assert align.write_pdb(pos, ctype).encode() == want
    path = tmp_path / "s.pdb"
    utils.WritePDB(pos, str(path), ctype)
    assert path.read_bytes() == want
    utils.WritePDB(torch.as_tensor(pos), str(path), ctype)   # tensors are accepted too
    assert path.read_bytes() == want


def test_write_pdb_empty_and_single(tmp_path):
    from hic_gnn_b200 import utils
    from oracle import align

    for pos in (np.zeros((0, 3)), np.array([[1.0, -2.5, 3.14159]])):
        p = tmp_path / "e.pdb"
        utils.WritePDB(pos, str(p))
        assert p.read_text() == align.write_pdb(pos)


@pytest.mark.gpu
@pytest.mark.parametrize("key,t1,t2,e1,e2,filtered", ALIGN_CASES)
def test_gpu_domain_alignment_matches_reference(golden, io_golden, key, t1, t2, e1, e2, filtered):
    from hic_gnn_b200 import utils

    arrays, _ = golden
    fn = utils.domain_alignment_filtered if filtered else utils.domain_alignment
    got = fn(arrays[f"{t1}_list"], arrays[f"{t2}_list"], io_golden[e1], io_golden[e2])
    assert got.is_cuda and got.dtype == torch.float64
    assert np.abs(got.cpu().numpy() - io_golden[key]).max() < 1e-9
    # tensors on the device are accepted as well
    got2 = fn(torch.as_tensor(arrays[f"{t1}_list"]).cuda(), torch.as_tensor(arrays[f"{t2}_list"]).cuda(), torch.as_tensor(io_golden[e1]).cuda(),
              torch.as_tensor(io_golden[e2]).cuda())
    assert torch.equal(got, got2)


@pytest.mark.gpu
def test_gpu_domain_alignment_rejects_mismatched_bins(golden):
    from hic_gnn_b200 import utils

    arrays, _ = golden
    with pytest.raises(IndexError):
        utils.domain_alignment(arrays["1mb_list"], arrays["500kb_list"], np.zeros((3, 8)), np.zeros((114, 8)))
    with pytest.raises(ValueError):  # the filtered variant drops the out-of-range rows, then the two sides differ in length
        utils.domain_alignment_filtered(arrays["1mb_list"], arrays["500kb_list"], np.zeros((3, 8)), np.zeros((114, 8)))


@pytest.mark.gpu
def test_gpu_alignment_of_512d_embeddings_reaches_the_same_optimum(golden):
    """The reference's real shape: 512-d embeddings, 113 matched bin pairs drawn from 58 trained rows => the
    cross-covariance has rank <= 58 and the minimiser of ||A R - B|| is not unique (LAPACK and cuSOLVER pick different
    null-space bases).  What IS determined: R is orthogonal and the residual is the minimum."""
    from hic_gnn_b200 import utils
    from oracle import align

    arrays, _ = golden
    l1, l2 = arrays["1mb_list"], arrays["500kb_list"]
    rng = np.random.default_rng(5)
    e1, e2 = rng.standard_normal((58, 512)), rng.standard_normal((114, 512))
    a_rows, b_rows = align.matched_rows(l1, l2)
    want = align.domain_alignment(l1, l2, e1, e2)
    got = utils.domain_alignment(l1, l2, e1, e2).cpu().numpy()
    res_want = np.linalg.norm(want[a_rows] - e1[b_rows])
    res_got = np.linalg.norm(got[a_rows] - e1[b_rows])
    assert abs(res_got - res_want) < 1e-9 * res_want
    # an orthogonal map: all pairwise inner products of the rows are preserved whatever the null-space basis
    assert np.abs(got @ got.T - e2 @ e2.T).max() < 1e-8


@pytest.mark.gpu
def test_generalisation_flow_matches_oracle(golden, tmp_path):
    """BASELINE.json configs[1] (HiC_GAT_generalize_directly.py:312-365) from a shared state_dict and a shared aligned
    embedding: 500 kb graph + GAT net forward + dSCC + PDB on the GPU against the oracle's CPU flow."""
    from hic_gnn_b200 import metrics, models as gmodels, utils as gutils
    from oracle import align, graph as ograph, loss as oloss, models as omodels, wish as owish

    arrays, _ = golden
    l1, l2 = arrays["1mb_list"], arrays["500kb_list"]
    normed2 = arrays["500kb_kr_oracle"]
    rng = np.random.default_rng(9)
    e1 = 0.25 * rng.standard_normal((58, 512))
    e2 = 0.25 * rng.standard_normal((normed2.shape[0], 512))
    fit = align.domain_alignment(l1, l2, e1, e2)                     # shared: see the rank remark above
    torch.manual_seed(42)
    om = omodels.GATNetSelectiveResidualsUpdated()
    om.eval()
    odata = ograph.load_input(normed2.copy(), fit)
    truth = owish.cont2dist(odata.y.clone(), 1.0)
    with torch.no_grad():
        ocoords = om.get_model(odata.x.float(), odata.edge_index)
    want = oloss.dscc(ocoords, truth)

    gm = gmodels.GATNetSelectiveResidualsUpdated().cuda()
    gm.load_state_dict(om.state_dict())
    gm.eval()
    gdata = gutils.load_input(normed2.copy(), fit)
    target = gutils.wish_target(gdata.y, 1.0)
    with torch.no_grad():
        gcoords = gm.get_model(gdata.x.float(), gdata.edge_index)
    scale = float(ocoords.abs().max())
    assert float((gcoords.cpu() - ocoords).abs().max()) < 1e-4 * scale
    assert abs(metrics.dscc(gcoords, target) - want) < 1e-3           # north_star: final dSCC within 1e-3
    p1, p2 = tmp_path / "g.pdb", tmp_path / "o.pdb"
    gutils.WritePDB(gcoords * 100, str(p1))
    p2.write_text(align.write_pdb((ocoords * 100).numpy()))
    a, b = p1.read_text().splitlines(), p2.read_text().splitlines()
    assert len(a) == len(b) and [x for x in a if not x.startswith("ATOM")] == [y for y in b if not y.startswith("ATOM")]
    xa = np.array([[float(x[30:38]), float(x[38:46]), float(x[46:54])] for x in a if x.startswith("ATOM")])
    xb = np.array([[float(y[30:38]), float(y[38:46]), float(y[46:54])] for y in b if y.startswith("ATOM")])
    assert xa.shape == (normed2.shape[0], 3) and np.abs(xa - xb).max() <= 1e-4 * 100 * scale + 1.5e-3   # %.3f of values within 1e-4 relative
    assert [x[:30] for x in a] == [y[:30] for y in b]                                                    # record names / serials / residue ids


@pytest.mark.gpu
def test_generalisation_example_runs_end_to_end(monkeypatch, tmp_path):
    import importlib.util
    import sys

    root = os.path.dirname(HERE)
    spec = importlib.util.spec_from_file_location("chr19_generalize", os.path.join(root, "examples", "chr19_generalize.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = tmp_path / "gen.pdb"
    monkeypatch.setattr(sys, "argv", ["chr19_generalize.py", "--steps", "40", "--out", str(out)])
    hist, coords, d1, d2 = mod.main()
    assert len(hist) <= 40 and hist[-1] < hist[0] and coords.shape == (114, 3) and bool(torch.isfinite(coords).all())
    assert -1.0 <= d1 <= 1.0 and -1.0 <= d2 <= 1.0
    assert out.read_text().count("ATOM") == 114
