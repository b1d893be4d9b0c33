"""Row f-4: the implicit (sparse) target against the streamed dense-target kernel and the oracle on the same
maps: identical loss / moments / gradients without any N x N array."""
import pytest
import torch

from helpers import random_coords, rel_err, small_map, wish_from_map

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _setup(n, density, seed):
    from hic_gnn_b200 import utils

    adj = small_map(n, density, seed=seed)
    x = torch.zeros(n, 4)
    data = utils.load_input(adj.numpy().copy(), x.numpy())
    dense = utils.wish_target(data.y, 1.0)
    sparse = utils.sparse_wish_target(data, 1.0)
    return adj, data, dense, sparse


@pytest.mark.parametrize("n,density", [(58, 1.0), (200, 0.05), (777, 0.2), (1500, 0.02), (3001, 0.01)])
@pytest.mark.parametrize("mode", ["mse", "mse_moments", "mse_moments_full", "contrastive"])
def test_sparse_target_matches_dense_target(n, density, mode):
    import hic_gnn_b200 as hg

    adj, data, dense, sparse = _setup(n, density, seed=n)
    # the implicit target is exactly the dense one: edges carry the same f32 values, the rest is 1.0 / 0.0
    rebuilt = torch.ones(n, n, device="cuda")
    rebuilt.fill_diagonal_(0)
    g = data.edge_index
    rebuilt[g.storage.row(), g.col] = sparse.tval
    assert torch.equal(rebuilt, dense.dense())
    coords = random_coords(n, seed=n + 1).cuda()
    c1 = coords.clone().requires_grad_(True)
    l1, m1 = hg.pairwise_loss(c1, dense, mode)
    (g1,) = torch.autograd.grad(l1, c1)
    c2 = coords.clone().requires_grad_(True)
    l2, m2 = hg.pairwise_loss(c2, sparse, mode)
    (g2,) = torch.autograd.grad(l2, c2)
    assert abs(float(l1) - float(l2)) <= 1e-6 * abs(float(l1))
    assert rel_err(m2, m1) < 1e-6
    assert rel_err(g2, g1) < TOL


def test_sparse_target_matches_oracle_and_row_blocks():
    import hic_gnn_b200 as hg
    from hic_gnn_b200 import ops
    from oracle import loss as oloss

    n = 600
    adj, data, dense, sparse = _setup(n, 0.05, seed=3)
    truth = wish_from_map(adj, 1.0)
    coords = random_coords(n, seed=9)
    c = coords.clone().requires_grad_(True)
    want = oloss.mse_loss(c, truth)
    (gw,) = torch.autograd.grad(want, c)
    cg = coords.cuda().requires_grad_(True)
    loss, m = hg.pairwise_loss(cg, sparse, "mse_moments")
    (gg,) = torch.autograd.grad(loss, cg)
    assert abs(float(loss) - float(want)) / float(want) < TOL
    assert rel_err(gg, gw) < TOL
    from scipy.stats import pearsonr

    tt, dd = oloss.triu_pairs(truth, coords)
    assert abs(float(ops.pearson_from_moments(m, n * (n - 1) / 2)) - pearsonr(tt.numpy(), dd.numpy())[0]) < 1e-6
    # row blocks (what the ranks of a sharded run compute) add up to the full launch
    mode = ops._MODES["mse_moments_full"]
    m_full, g_full = ops.pairloss_raw(coords.cuda(), sparse, mode, 4.0 / n**2, 0.0)
    m_sum, g_sum = torch.zeros_like(m_full), torch.zeros_like(g_full)
    for r0, r1 in [(0, 100), (100, 100), (100, 333), (333, 600)]:
        mb, gb = ops.pairloss_raw(coords.cuda(), sparse.rows(r0, r1), mode, 4.0 / n**2, 0.0)
        m_sum += mb
        g_sum += gb
    assert rel_err(m_sum, m_full) < 1e-7 and rel_err(g_sum, g_full) < TOL
