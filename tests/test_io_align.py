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
