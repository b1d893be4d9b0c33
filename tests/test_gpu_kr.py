"""KR normalisation on the GPU (row f-1) against the oracle's restatement of r_utils.R and the
reference's shipped known answer (chr19 1 Mb: KR -> cont2dist(0.4) -> dSCC of the shipped PDB)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["1mb", "500kb"])
def test_kr_matches_oracle_on_reference_maps(golden, tag):
    from hic_gnn_b200 import kr as gkr
    from oracle import kr as okr

    g, _ = golden
    m = g[f"{tag}_matrix"].copy()
    np.fill_diagonal(m, 0)
    want = okr.kr_norm(m)
    got = gkr.kr_norm(torch.tensor(m, device="cuda")).cpu().numpy()
    assert got.shape == want.shape
    # different summation order inside A @ x: entries may land on the neighbouring 1e-6 grid point
    # (the reference's own typo / no-typo variants differ the same way)
    assert np.abs(got - want).max() <= 1.0000001e-6
    assert (got != want).mean() < 0.01
    rows = got.sum(1)
    assert np.abs(rows - 1).max() < 2e-4  # balanced up to the rounding to 6 decimals


def test_kr_gemv_and_rounding_kernels():
    from hic_gnn_b200 import _native as N
    from hic_gnn_b200.ops import _stream

    gen = torch.Generator().manual_seed(0)
    for n in (1, 7, 58, 1001):
        A = torch.rand(n, n, generator=gen, dtype=torch.float64).cuda()
        x = torch.rand(n, generator=gen, dtype=torch.float64).cuda()
        y = torch.empty_like(x)
        N.check(N.lib().hicgat_gemv_f64(A.data_ptr(), A.stride(0), n, x.data_ptr(), y.data_ptr(), _stream()))
        assert torch.allclose(y, A @ x, rtol=1e-13, atol=0)
        out = torch.empty_like(A)
        N.check(N.lib().hicgat_kr_scale_round_f64(A.data_ptr(), A.stride(0), n, x.data_ptr(), out.data_ptr(), out.stride(0), 6, _stream()))
        want = np.round(((x.cpu().numpy()[:, None] * A.cpu().numpy()) * x.cpu().numpy()[None, :]), 6)
        assert np.array_equal(out.cpu().numpy(), want)
