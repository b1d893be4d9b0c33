"""Fused relu(LayerNorm(x)) + residual (hicgat_ln_relu_add_fwd/bwd) against torch: the glue of the GAT net's
MLP head, models.py:670-690 `F.relu(self.norm_a(self.densea(x))) + x_initial`."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL_FWD = 2e-6   # max-norm relative, f32 kernel vs the f64 value of the reference formula
TOL_GRAD = 1e-5  # north_star tolerance for gradients


def rel_err(a, b):
    a, b = a.detach(), b.detach()
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-300))


def reference_f64(x, w, b, eps, res, gy, mask=None):
    """f64 evaluation of the reference expression and its autograd gradients.  `mask` (the kernel's own
    relu pattern) replaces the reference's where a pre-activation sits within f32 rounding of the kink."""
    x64 = x.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    b64 = b.double().requires_grad_(True)
    pre = F.layer_norm(x64, (x.shape[1],), w64, b64, eps)
    out = F.relu(pre) if mask is None else pre * mask.double()
    y = out + (res.double() if res is not None else 0.0)
    gx, gw, gb = torch.autograd.grad((y * gy.double()).sum(), (x64, w64, b64))
    return y.detach(), pre.detach(), gx, gw, gb


@pytest.mark.parametrize("c", [32, 64, 128, 256, 512])
@pytest.mark.parametrize("n", [1, 7, 58, 1001, 20000])
@pytest.mark.parametrize("with_res", [False, True])
def test_ln_relu_add_matches_torch(n, c, with_res):
    from hic_gnn_b200 import ops

    g = torch.Generator().manual_seed(1000 * c + n)
    x = (torch.randn(n, c, generator=g) * 1.7 + 0.3).cuda()
    res = torch.randn(n, c, generator=g).cuda() if with_res else None
    gy = torch.randn(n, c, generator=g).cuda()
    norm = torch.nn.LayerNorm(c).cuda()
    with torch.no_grad():
        norm.weight.copy_(1.0 + 0.2 * torch.randn(c, generator=g))
        norm.bias.copy_(0.1 * torch.randn(c, generator=g))
    xg = x.clone().requires_grad_(True)
    rg = res.clone().requires_grad_(True) if with_res else None
    y = ops.ln_relu_add(xg, norm, rg)
    grads = torch.autograd.grad((y * gy).sum(), (xg, norm.weight, norm.bias) + ((rg,) if with_res else ()))
    # the kernel's relu pattern, read off a residual-free forward of the same inputs
    with torch.no_grad():
        mask = ops.ln_relu_add(x, norm) > 0
    y_ref, pre_ref, gx_ref, gw_ref, gb_ref = reference_f64(x, norm.weight.detach(), norm.bias.detach(), norm.eps, res, gy, mask)
    # the pattern may differ from the f64 one only where the pre-activation is within f32 rounding of zero
    flips = mask != (pre_ref > 0)
    assert float(pre_ref[flips].abs().max()) < 1e-5 if flips.any() else True
    assert int(flips.sum()) <= max(2, n * c // 100000)
    assert rel_err(y, y_ref) < TOL_FWD
    assert rel_err(grads[0], gx_ref) < TOL_GRAD
    assert rel_err(grads[1], gw_ref) < TOL_GRAD
    assert rel_err(grads[2], gb_ref) < TOL_GRAD
    if with_res:
        assert torch.equal(grads[3], gy)
    # torch's own f32 CUDA path on the same inputs: same value within f32 rounding
    y_t = F.relu(norm(x)) + (res if with_res else 0.0)
    assert rel_err(y, y_t) < 5e-6


def test_ln_relu_add_is_bit_reproducible_and_rejects_bad_input():
    from hic_gnn_b200 import ops

    g = torch.Generator().manual_seed(3)
    x = torch.randn(5000, 256, generator=g).cuda().requires_grad_(True)
    gy = torch.randn(5000, 256, generator=g).cuda()
    norm = torch.nn.LayerNorm(256).cuda()
    outs = []
    for _ in range(3):
        y = ops.ln_relu_add(x, norm)
        outs.append(torch.autograd.grad((y * gy).sum(), (x, norm.weight, norm.bias)))
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))
    with pytest.raises(RuntimeError, match="unsupported width"):
        ops.ln_relu_add(torch.randn(4, 96).cuda(), torch.nn.LayerNorm(96).cuda())
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.ln_relu_add(torch.randn(4, 64), torch.nn.LayerNorm(64))


def test_gat_net_head_state_dict_and_eval_unchanged():
    """The fused glue keeps the reference's modules (state_dict keys) and computes what the unfused torch
    expression computes on the same CUDA tensors."""
    from hic_gnn_b200 import models

    torch.manual_seed(0)
    net = models.GATNetSelectiveResidualsUpdated().cuda()
    keys = set(net.state_dict())
    assert {"norm_a.weight", "norm_a.bias", "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias"} <= keys
    h = torch.randn(300, 512).cuda()
    with torch.no_grad():
        x0 = net.align_densea(h)
        want = F.relu(net.norm_a(net.densea(h))) + x0
        from hic_gnn_b200 import ops

        got = ops.ln_relu_add(net.densea(h), net.norm_a, x0)
    assert rel_err(got, want) < 5e-6
