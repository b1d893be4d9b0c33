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
    x = _kr_scaling(lambda v, out: _gemv(A, v, out), n, A.device, literal_typo)
    out = torch.empty_like(A)
    N.check(N.lib().hicgat_kr_scale_round_f64(A.data_ptr(), A.stride(0), n, x.contiguous().data_ptr(), out.data_ptr(), out.stride(0), decimals, _stream()),
            "hicgat_kr_scale_round_f64")
    return out


def kr_norm_csr(rowptr32: torch.Tensor, col32: torch.Tensor, val64: torch.Tensor, literal_typo: bool = True, decimals: int = 6) -> torch.Tensor:
    """``KRnorm`` on a symmetric CSR matrix with f64 values (the graph built straight from a contact list,
    ``utils.load_input_sparse``): same Newton-CG iteration, every ``A %*% x`` a CSR SpMV (``hicgat_spmv_csr_f64``), the final
    ``round(x_i a_ij x_j, 6)`` on the stored entries only.  Returns the balanced values f64[nnz] (entries that round to
    zero stay in the pattern with value 0: drop them with the caller's mask, as the dense path's ``load_input`` would)."""
    _cuda(rowptr32, col32, val64)
    n = rowptr32.numel() - 1
    if int((rowptr32[1:] == rowptr32[:-1]).sum()) or bool(torch.isnan(val64).any()):
        raise RuntimeError("kr_norm_csr: empty rows / NaN values must be removed by the caller (utils.load_input_sparse does)")
    lib = N.lib()

    def spmv(v, out):
        N.check(lib.hicgat_spmv_csr_f64(rowptr32.data_ptr(), col32.data_ptr(), val64.data_ptr(), n, v.data_ptr(), out.data_ptr(), _stream()), "hicgat_spmv_csr_f64")
        return out

    x = _kr_scaling(spmv, n, val64.device, literal_typo)
    row = torch.repeat_interleave(torch.arange(n, device=val64.device), (rowptr32[1:] - rowptr32[:-1]).long())
    v = (x[row] * val64) * x[col32.long()]                # t(t(x*A)*x), r_utils.R:75
    scale = 10.0 ** decimals
    return torch.round(v * scale) / scale                 # round(., 6): half-to-even on v * 1e6, like the dense kernel


def _kr_scaling(matvec, n: int, dev, literal_typo: bool) -> torch.Tensor:
    """The scaling vector x of KRnorm (r_utils.R:17-88) for a symmetric matrix given by ``matvec(v, out) -> out = A v``."""
    tol, delta, Delta = 1e-6, 0.1, 3.0
    g, etamax = 0.9, 0.1
    eta, stop_tol = etamax, tol * 0.5
    e = torch.ones(n, dtype=torch.float64, device=dev)
    x = e.clone()
    rt = tol**2
    tmp = torch.empty_like(x)
    v = x * matvec(x, tmp)
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
            w = x * matvec((x * p).contiguous(), tmp) + v * p
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
        v = x * matvec(x.contiguous(), tmp)
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
    return x
