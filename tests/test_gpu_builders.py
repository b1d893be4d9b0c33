"""GPU parity: CSR graph build + wish-distance builder vs the oracle / reference golden vectors.
Integer and index work is checked BIT-EXACT."""
import numpy as np
import pytest
import torch

from helpers import small_map

pytestmark = pytest.mark.gpu


def _csr_oracle(adj_np):
    from oracle import graph as ograph

    return ograph.symmetric_csr_from_dense(adj_np)


def _check_csr(adj_np):
    from hic_gnn_b200 import ops

    want = _csr_oracle(adj_np)
    a = torch.tensor(adj_np, dtype=torch.float64, device="cuda")
    a.fill_diagonal_(0)
    rowptr, col, val = ops.csr_from_dense(a)
    assert rowptr.dtype == torch.int64 and col.dtype == torch.int64 and val.dtype == torch.float32
    assert torch.equal(rowptr.cpu(), want.rowptr)
    assert torch.equal(col.cpu(), want.col)
    assert np.array_equal(val.cpu().numpy().view(np.uint32), want.value.numpy().view(np.uint32))
    return rowptr, col, val, want


@pytest.mark.parametrize("tag", ["1mb", "500kb", "asym"])
def test_csr_matches_reference_golden(golden, tag):
    g, _ = golden
    src = g["asym_matrix"] if tag == "asym" else g[f"{tag}_kr_oracle"]
    rowptr, col, val, _ = _check_csr(src)
    # and directly against what the reference's load_input produced
    assert np.array_equal(rowptr.cpu().numpy(), g[f"{tag}_csr_rowptr"])
    assert np.array_equal(col.cpu().numpy(), g[f"{tag}_csr_col"])
    assert np.array_equal(val.cpu().numpy().view(np.uint32), g[f"{tag}_csr_val"].view(np.uint32))


@pytest.mark.parametrize("n,density", [(1, 1.0), (2, 1.0), (33, 0.5), (257, 0.1), (1000, 0.02), (2048, 0.9)])
def test_csr_random_maps_bit_exact(n, density):
    rng = np.random.default_rng(n)
    a = rng.random((n, n)) * (rng.random((n, n)) < density)  # asymmetric on purpose
    if n > 4:
        a[3, :] = 0
        a[:, 3] = 0  # isolated locus -> empty row
    _check_csr(a)


def test_self_loops_perm_and_sage_weights():
    from hic_gnn_b200.graph import CSRGraph
    from hic_gnn_b200 import ops
    from oracle import conv as oconv
    from oracle import graph as ograph

    adj = small_map(300, 0.3, seed=3)
    a = adj.cuda()
    rowptr, col, val = ops.csr_from_dense(a)
    g = CSRGraph(rowptr, col, val, 300)
    want = ograph.symmetric_csr_from_dense(adj.numpy())
    sl = ograph.set_diag(want)
    r32, c32, perm = g.with_self_loops()
    assert torch.equal(r32.cpu().long(), sl.rowptr) and torch.equal(c32.cpu().long(), sl.col)
    # perm maps entry (i,j) to entry (j,i)
    row = sl.row
    p = perm.cpu().long()
    assert torch.equal(row[p], sl.col) and torch.equal(sl.col[p], row)
    # the fused builder with_self_loops=1 gives the same pattern
    rp2, c2, v2 = ops.csr_from_dense(a, with_self_loops=True)
    assert torch.equal(rp2.cpu(), sl.rowptr) and torch.equal(c2.cpu(), sl.col)
    assert torch.equal(v2.cpu(), sl.value)
    norm, norm_t = g.sage_weights()
    want_norm = oconv.SAGEConv.adjust_weights(want)
    assert np.array_equal(norm.cpu().numpy().view(np.uint32), want_norm.numpy().view(np.uint32))  # bit-exact
    dense_t = torch.zeros(300, 300)
    dense_t[want.row, want.col] = want_norm
    assert torch.equal(norm_t.cpu(), dense_t.t()[want.row, want.col])


@pytest.mark.parametrize("tag,factors", [("1mb", (0.4, 0.5, 1.0)), ("500kb", (0.4, 0.5, 1.0)), ("asym", (1.0,))])
def test_cont2dist_matches_reference_golden(golden, tag, factors):
    from hic_gnn_b200 import utils

    g, _ = golden
    y = torch.tensor(g[f"{tag}_y"], device="cuda")
    for f in factors:
        got = utils.cont2dist(y, f).cpu().numpy()
        want = g[f"{tag}_wish_{f}"]
        if f == 1.0:  # reciprocal, divide: IEEE on both sides -> bit-exact
            exact = want
        elif f == 0.5:
            # ATen maps pow(., 0.5) to sqrt; its AVX-512 f64 sqrt (Sleef u05) is NOT correctly
            # rounded (28 of 3364 elements of the 1 Mb golden are 1-2 ulp off IEEE sqrt), so the
            # golden is matched to 2 ulp and the correctly rounded IEEE chain is matched bit for bit.
            np.testing.assert_allclose(got, want, rtol=4.5e-16, atol=0)
            with np.errstate(divide="ignore"):
                s = np.sqrt(1.0 / g[f"{tag}_y"])
            np.fill_diagonal(s, 0.0)
            mx = s[np.isfinite(s)].max()
            exact = np.where(np.isinf(s), mx, s) / mx
        else:  # generic pow: CUDA's f64 pow vs the host libm, <= a few ulp
            np.testing.assert_allclose(got, want, rtol=1e-14, atol=0)
            exact = None
        if exact is not None:
            assert np.array_equal(got.view(np.uint64), exact.view(np.uint64))
        tgt = utils.wish_target(y, f)
        assert tgt.pitch % 4 == 0 and tgt.data.shape == (y.shape[0], tgt.pitch)
        if exact is not None:
            assert np.array_equal(tgt.dense().cpu().numpy(), exact.astype(np.float32))
        else:
            np.testing.assert_allclose(tgt.dense().cpu().numpy(), want.astype(np.float32), rtol=2e-7)
        assert float(tgt.data[:, y.shape[0]:].abs().sum()) == 0.0  # padding stays zero


def test_cont2dist_zero_contacts_and_row_blocks():
    from hic_gnn_b200 import ops
    from oracle.wish import cont2dist as ocont

    adj = small_map(257, 0.2, seed=9)  # many zero contacts -> wish distance exactly 1
    want = ocont(adj.clone(), 1.0)
    a = adj.cuda()
    full, tgt = ops.cont2dist(a, 1.0, want_f64=True, want_f32=True)
    assert torch.equal(full.cpu(), want)
    assert float(full.max()) == 1.0 and float(full.diagonal().abs().max()) == 0.0
    # row-sharded build: local max, max-reduce, apply == full build
    blocks, maxes = [], []
    for r0, r1 in [(0, 100), (100, 257)]:
        mx = torch.empty(1, dtype=torch.float64, device="cuda")
        maxes.append(mx)
    gmax = []

    def reduce_factory():
        def red(m):
            gmax.append(m.clone())
        return red

    for r0, r1 in [(0, 100), (100, 257)]:
        ops.cont2dist(a[r0:r1].contiguous(), 1.0, want_f64=False, want_f32=True, r0=r0, r1=r1, max_reduce=reduce_factory())
    true_max = torch.stack(gmax).max()
    for r0, r1 in [(0, 100), (100, 257)]:
        _, t = ops.cont2dist(a[r0:r1].contiguous(), 1.0, want_f64=False, want_f32=True, r0=r0, r1=r1, max_reduce=lambda m: m.fill_(float(true_max)))
        blocks.append(t.dense())
    assert torch.equal(torch.cat(blocks).cpu(), want.float())


@pytest.mark.parametrize("tag", ["1mb", "500kb", "dup"])
def test_convert_to_matrix_matches_reference_golden(golden, tag):
    """Row f-2: contact list -> dense matrix on the GPU, against the reference's own
    utils.convert_to_matrix outputs (golden) incl. duplicate records ("last record wins")."""
    from hic_gnn_b200 import utils

    g, _ = golden
    if f"{tag}_list" not in g:
        pytest.skip("no list fixture for this tag")
    got = utils.convert_to_matrix(g[f"{tag}_list"]).cpu().numpy()
    assert np.array_equal(got, g[f"{tag}_matrix"])
