"""Knight-Ruiz normalisation on the GPU ("next"): what the reference does
by shelling out to R (``os.system('Rscript normalize.R ...')``, HiC-GNN_main.py:85;
``normalize.R:1-11`` -> ``KRnorm``, ``r_utils.R:1-93``).

The iteration is the reference's inexact Newton-CG, statement for statement (including the
``Z``/``z`` typo at ``r_utils.R:60`` behind ``literal_typo``); the O(N^2) work -- every ``A %*% x``
and the final ``round(t(t(x*A)*x), 6)`` -- runs in the library's f64 kernels (csrc/kr.cu), the O(N)
vector algebra in torch on the device, and only scalars cross to the host.
"""
from __future__ import annotations

import torch

from . import _native as N
from .ops import _cuda, _stream


def _gemv(A: torch.Tensor, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    N.check(N.lib().hicgat_gemv_f64(A.data_ptr(), A.stride(0), A.shape[0], x.data_ptr(), out.data_ptr(), _stream()), "hicgat_gemv_f64")
    return out


def kr_norm(A: torch.Tensor, literal_typo: bool = True, decimals: int = 6) -> torch.Tensor:
    """``KRnorm`` (r_utils.R:1-93) of a square f64 CUDA matrix.  Returns the balanced matrix rounded to
    ``decimals`` (``round(., 6)``, r_utils.R:90); all-zero rows/columns are dropped first
    (r_utils.R:3-10) and NaNs zeroed (r_utils.R:14-15)."""
    _cuda(A)
    if A.dtype != torch.float64 or A.dim() != 2 or A.shape[0] != A.shape[1]:
        raise RuntimeError("kr_norm expects a square float64 CUDA matrix")
    A = torch.nan_to_num(A, nan=0.0)
    keep = A.sum(dim=0) != 0
    if not bool(keep.all()):
        A = A[keep][:, keep]
    A = A.contiguous()
    n = A.shape[0]
    dev = A.device
    tol, delta, Delta = 1e-6, 0.1, 3.0
    g, etamax = 0.9, 0.1
    eta, stop_tol = etamax, tol * 0.5
    e = torch.ones(n, dtype=torch.float64, device=dev)
    x = e.clone()
    rt = tol**2
    tmp = torch.empty_like(x)
    v = x * _gemv(A, x, tmp)
    rk = 1.0 - v
    rho_km1 = float(rk @ rk)
    rout = rold = rho_km2 = rho_km1
    z = rk.clone()
    p = z.clone()
    while rout > rt:  # outer iteration, r_utils.R:27
        k = 0
        y = e.clone()
        innertol = max(eta**2 * rout, rt)
        while rho_km1 > innertol:  # inner CG, r_utils.R:30
            k += 1
            if k == 1:
                z = rk / v
                p = z.clone()
                rho_km1 = float(rk @ z)
            else:
                beta = rho_km1 / rho_km2
                p = z + beta * p
            w = x * _gemv(A, (x * p).contiguous(), tmp) + v * p
            alpha = rho_km1 / float(p @ w)
            ap = alpha * p
            ynew = y + ap
            if float(ynew.min()) <= delta:  # r_utils.R:45-50
                if delta == 0:
                    break
                ind = ap < 0
                gamma = float(((delta - y[ind]) / ap[ind]).min())
                y = y + gamma * ap
                break
            if float(ynew.max()) >= Delta:  # r_utils.R:51-56
                ind = ynew > Delta
                gamma = float(((Delta - y[ind]) / ap[ind]).min())
                y = y + gamma * ap
                break
            y = ynew
            rk = rk - alpha * w
            rho_km2 = rho_km1
            if not literal_typo:
                z = rk / v
            rho_km1 = float(rk @ z)  # r_utils.R:60 (stale z when literal_typo)
        x = x * y
        v = x * _gemv(A, x.contiguous(), tmp)
        rk = 1.0 - v
        rho_km1 = float(rk @ rk)
        rout = rho_km1
        rat = rout / rold
        rold = rout
        res_norm = rout**0.5
        eta_o = eta
        eta = g * rat
        if g * eta_o**2 > 0.1:
            eta = max(eta, g * eta_o**2)
        eta = max(min(eta, etamax), stop_tol / res_norm)
    out = torch.empty_like(A)
    N.check(N.lib().hicgat_kr_scale_round_f64(A.data_ptr(), A.stride(0), n, x.contiguous().data_ptr(), out.data_ptr(), out.stride(0), decimals, _stream()),
            "hicgat_kr_scale_round_f64")
    return out
