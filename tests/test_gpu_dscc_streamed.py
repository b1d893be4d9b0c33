"""SURVEY.md section 8 row f-3 at scale: the streamed dSCC (histogram mid-ranks, no N x N matrix, no index arrays, no sort) against
scipy.stats.spearmanr -- what the reference calls (HiC-GNN_main.py:135-139) -- and against the exact sort-based GPU evaluation."""
import numpy as np
import pytest
import torch

from helpers import random_coords, small_map, wish_from_map

pytestmark = pytest.mark.gpu


def _structured_coords(truth, seed):
    """Coordinates whose distances correlate with the wish distances (a noisy MDS-like embedding), so that rho is not ~0."""
    n = truth.shape[0]
    g = torch.Generator().manual_seed(seed)
    t = torch.linspace(0, 6.0, n)
    c = torch.stack([torch.cos(t) * (1 + 0.1 * t), torch.sin(t) * (1 + 0.1 * t), 0.2 * t], dim=1)
    return (c + 0.05 * torch.randn(n, 3, generator=g)).float()


@pytest.mark.parametrize("n,density", [(58, 1.0), (300, 0.3), (1500, 0.1)])
def test_streamed_dscc_matches_scipy(n, density):
    from scipy.stats import spearmanr

    import hic_gnn_b200 as hg
    from hic_gnn_b200 import metrics
    from oracle import loss as oloss

    truth = wish_from_map(small_map(n, density, seed=n), 1.0)
    coords = _structured_coords(truth, n)
    dist_truth, dist_out = oloss.triu_pairs(truth, coords)
    want = float(spearmanr(dist_truth.numpy(), dist_out.numpy())[0])
    tgt = hg.WishTarget.from_dense(truth.cuda())
    got = metrics.dscc_streamed(coords.cuda(), tgt)
    assert abs(got - want) < 1e-5, (got, want)
    assert abs(metrics.dscc(coords.cuda(), tgt) - want) < 1e-6          # exact path below STREAM_FROM loci
    # row blocks: histograms and cross sums add up
    import hic_gnn_b200.ops as ops

    parts = []
    for r0, r1 in [(0, n // 3), (n // 3, n)]:
        parts.append(hg.WishTarget.from_dense(truth.cuda(), r0, r1))
    # emulate a 2-rank all-reduce on one GPU: both blocks' kernels add into shared accumulators
    from hic_gnn_b200 import _native as N

    c = coords.cuda().contiguous()
    span = (c.max(0).values - c.min(0).values).double()
    dmax = float(torch.sqrt((span * span).sum())) * (1 + 1e-6) + 1e-30
    nb = metrics.NBINS
    ds, ts = nb / dmax, nb / (float(truth.max()) * (1 + 1e-6))
    hist_d = torch.zeros(nb, dtype=torch.int64, device="cuda")
    hist_t = torch.zeros_like(hist_d)
    for p in parts:
        N.check(N.lib().hicgat_rank_histograms(c.data_ptr(), p.data.data_ptr(), p.pitch, n, p.r0, p.r1, ds, ts, nb, hist_d.data_ptr(), hist_t.data_ptr(), ops._stream()))
    assert int(hist_d.sum()) == n * (n - 1) // 2 and int(hist_t.sum()) == n * (n - 1) // 2
    rd, rt = metrics._midranks(hist_d), metrics._midranks(hist_t)
    cross = torch.zeros(1, dtype=torch.float64, device="cuda")
    for p in parts:
        N.check(N.lib().hicgat_rank_cross_sum(c.data_ptr(), p.data.data_ptr(), p.pitch, n, p.r0, p.r1, ds, ts, nb, rd.data_ptr(), rt.data_ptr(), cross.data_ptr(), ops._stream()))
    assert abs(metrics._spearman_from_tables(hist_d, hist_t, cross, n * (n - 1) / 2) - want) < 1e-5


@pytest.mark.parametrize("n,density", [(300, 0.3), (2000, 0.05)])
def test_streamed_dscc_of_implicit_target_matches_dense(n, density):
    from scipy.stats import spearmanr

    from hic_gnn_b200 import metrics, utils
    from oracle import loss as oloss

    adj = small_map(n, density, seed=n + 1)
    truth = wish_from_map(adj, 1.0)
    coords = _structured_coords(truth, n)
    dist_truth, dist_out = oloss.triu_pairs(truth, coords)
    want = float(spearmanr(dist_truth.numpy(), dist_out.numpy())[0])
    data = utils.load_input(adj.numpy().copy(), np.zeros((n, 4), dtype=np.float32))
    sp = utils.sparse_wish_target(data, 1.0)
    got = metrics.dscc(coords.cuda(), sp)
    assert abs(got - want) < 1e-5, (got, want)


def test_streamed_dscc_at_10k_loci_matches_exact_gpu_sort():
    """BASELINE.json configs[3] size (9 970 loci, 5e7 pairs): streamed (what dscc() picks above STREAM_FROM loci) against the exact
    sort-based evaluation on the same GPU."""
    import hic_gnn_b200 as hg
    from hic_gnn_b200 import metrics, ops, synth

    n = 9970
    adj = synth.synthetic_map_chunked(n, 0.07, device="cuda")
    _, tgt = ops.cont2dist(adj, 1.0, want_f64=False, want_f32=True)
    coords = _structured_coords(torch.zeros(n, n), 5).cuda()
    got = metrics.dscc(coords, tgt)                      # streamed
    t, d = metrics.upper_pairs(coords, tgt)
    want = float(metrics._pearson(metrics.average_ranks(t.double()), metrics.average_ranks(d.double())))
    assert abs(got - want) < 1e-5, (got, want)
