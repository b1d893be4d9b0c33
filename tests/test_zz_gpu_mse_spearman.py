"""combined_loss_training.py:96-152 (GATNetHeadsChanged3LayersLeakyReLUv2, total = MSE + alpha (1 - dSCC), fixed epochs)
through ``train.step_loss`` / ``train.fit`` with ``mode="mse_spearman"``.

Every reference value is the f64 value of the reference expression evaluated at the oracle's (f32) coordinates: at
initialisation the structure is nearly collapsed and ATen's f32 matmul-form ``cdist`` moves the reference's OWN numbers
(probed on the CPU: MSE 1.1e-5 relative, Spearman -0.03764 vs -0.03692 in f64); the difference-form distances of the CUDA
path reproduce the f64 values.  Once the structure has spread out (40 epochs) the f32 reference agrees to 4e-7 / 3e-6.

(The file sorts last on purpose: it was added when the round's GPU budget was exhausted)
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_mse_spearman_loop_matches_oracle(golden):
    from hic_gnn_b200 import metrics as gmetrics, models as gmodels, ops as gops, train as gtrain, utils as gutils
    from oracle import graph as ograph, loop as oloop, loss as oloss, models as omodels, wish as owish

    g, _ = golden
    adj = g["1mb_kr_oracle"]
    n = adj.shape[0]
    x = 0.25 * torch.randn(n, 512, generator=torch.Generator().manual_seed(7))
    odata = ograph.load_input(adj.copy(), x.numpy())
    gdata = gutils.load_input(adj.copy(), x.numpy())
    truth = owish.cont2dist(odata.y.clone(), 1.0)
    target = gutils.wish_target(gdata.y, 1.0)
    alpha = 0.7
    torch.manual_seed(42)
    om = omodels.GATNetHeadsChanged3LayersLeakyReLUv2()
    gm = gmodels.GATNetHeadsChanged3LayersLeakyReLUv2().cuda()
    gm.load_state_dict(om.state_dict())
    idx = torch.triu_indices(n, n, 1)

    def oracle_values():
        """(MSE, Spearman, dRMSD) of the reference expression in f64 + the f32 reference's own Spearman / MSE."""
        with torch.no_grad():
            c = om.get_model(odata.x.float(), odata.edge_index)
            _, mse32, rho32, _ = oloss.mse_spearman_loss(c, truth, alpha)
            d = torch.cdist(c.double(), c.double())
            t = truth.float().double()
            mse = float(((d - t) ** 2).mean())
            drmsd = float(torch.sqrt(((t[idx[0], idx[1]] - d[idx[0], idx[1]]) ** 2).mean()))
            return mse, oloss.dscc(c.double(), truth), drmsd, float(mse32), rho32

    def check(tol_ref32_mse, tol_ref32_rho, tol_model):
        mse_o, rho_o, drmsd_o, mse32, rho32 = oracle_values()
        assert abs(mse32 - mse_o) <= tol_ref32_mse * mse_o and abs(rho32 - rho_o) <= tol_ref32_rho   # the reference's own f32 noise
        want = mse_o + alpha * (1.0 - rho_o)
        with torch.no_grad():
            # (i) the loss / rank kernels alone, at the ORACLE's coordinates: 1e-5
            c_o = om.get_model(odata.x.float(), odata.edge_index).cuda()
            loss_c, moments_c = gops.pairwise_loss(c_o, target, gtrain._KERNEL_MODE["mse_spearman"], None)
            assert abs(float(loss_c) - mse_o) <= 1e-5 * abs(mse_o)
            assert abs(float(loss_c) + alpha * (1.0 - gmetrics.dscc(c_o, target)) - want) <= 1e-5 * abs(want)
            assert abs(float(gtrain.drmsd_from_moments(moments_c, n)) - drmsd_o) <= 1e-5 * drmsd_o
            # (ii) through the CUDA model: its f32 coordinates differ from the CPU model's at rounding level (another GEMM
            # summation order; depends on the host's BLAS kernels), which the nearly collapsed initial structure amplifies --
            # 1.2e-5 on the MSE was seen on one host -- hence ``tol_model``
            loss, total, moments = gtrain.step_loss(gm, gdata.x.float(), gdata.edge_index, target, "mse_spearman", alpha=alpha)
        assert abs(float(loss) - mse_o) <= tol_model * abs(mse_o)
        assert abs(float(total) - want) <= tol_model * abs(want)
        assert abs(float(gtrain.drmsd_from_moments(moments, n)) - drmsd_o) <= tol_model * drmsd_o
        return want

    want0 = check(2e-4, 2e-2, 1e-4)
    # the loop itself: five epochs on the GPU from the shared start (value, MSE gradient, Adam)
    h_g = gtrain.fit(gm, gdata.x.float(), gdata.edge_index, target, mode="mse_spearman", lr=1e-3, thresh=0.0, max_steps=5, alpha=alpha)
    assert len(h_g) == 5 and all(np.isfinite(h_g)) and abs(h_g[0] - want0) <= 1e-4 * abs(want0)
    # teacher-forced again at a spread-out structure: 40 oracle epochs, then the same parameters on both sides (free runs
    # diverge at f32 rounding level: a 1e-6 input perturbation moves the oracle's own MSE by 7e-3 within 5 epochs)
    oloop.train(om, odata.x.float(), odata.edge_index, truth, mode="mse_spearman", lr=1e-3, thresh=0.0, max_steps=40, as_written=False, alpha=alpha)
    gm.load_state_dict(om.state_dict())
    check(5e-5, 1e-3, 3e-5)
    with pytest.raises(ValueError, match="CUDA graph"):
        gtrain.TrainStep(gm, gdata.x.float(), gdata.edge_index, target, mode="mse_spearman", use_cuda_graph=True)
