"""Oracle (test infrastructure): Knight-Ruiz matrix balancing as the reference runs it.

Restates ``r_utils.R:1-93`` (called through ``normalize.R:1-11`` from e.g.
``HiC-GNN_main.py:85``) in numpy float64.  R is not available in this image, so this file
is pinned only end to end, through the reference's shipped structure/log known answers
(``Outputs/GM12878_1mb_chr19_list_{structure.pdb,log.txt}``; see tests).
"""
from __future__ import annotations

import numpy as np


def kr_norm(A: np.ndarray, literal_typo: bool = True) -> np.ndarray:
    """``KRnorm`` (``r_utils.R:1-93``): inexact Newton-CG balancing, then ``round(., 6)``.

    ``literal_typo=True`` keeps the reference's ``Z = rk/v`` (capital ``Z``,
    ``r_utils.R:60``): the preconditioned residual ``z`` is only set when ``k == 1`` and
    ``rho_km1`` is then computed with that stale ``z``.  ``False`` gives the textbook
    Knight-Ruiz update.  Both converge to the same balanced matrix up to +-1e-6 after the
    rounding.
    """
    A = np.array(A, dtype=np.float64, copy=True)
    # r_utils.R:3-10 -- remove all-zero columns/rows
    zeros = np.where(A.sum(axis=0) == 0)[0]
    if len(zeros) > 0:
        A = np.delete(np.delete(A, zeros, axis=0), zeros, axis=1)
    A[np.isnan(A)] = 0.0  # r_utils.R:14-15
    n = A.shape[0]
    tol, delta, Delta = 1e-6, 0.1, 3.0  # r_utils.R:12
    e = np.ones(n)
    g, etamax = 0.9, 0.1  # r_utils.R:21
    eta, stop_tol = etamax, tol * 0.5
    x = e.copy()
    rt = tol**2
    v = x * (A @ x)
    rk = 1.0 - v
    rho_km1 = float(rk @ rk)
    rout = rho_km1
    rold = rout
    rho_km2 = rho_km1
    z = rk.copy()
    p = z.copy()
    while rout > rt:  # outer iteration, r_utils.R:27
        k = 0
        y = e.copy()
        innertol = max(eta**2 * rout, rt)
        while rho_km1 > innertol:  # inner CG, r_utils.R:30
            k += 1
            if k == 1:
                z = rk / v
                p = z.copy()
                rho_km1 = float(rk @ z)
            else:
                beta = rho_km1 / rho_km2
                p = z + beta * p
            w = x * (A @ (x * p)) + v * p
            alpha = rho_km1 / float(p @ w)
            ap = alpha * p
            ynew = y + ap
            if ynew.min() <= delta:  # r_utils.R:45-50
                if delta == 0:
                    break
                ind = np.where(ap < 0)[0]
                gamma = np.min((delta - y[ind]) / ap[ind])
                y = y + gamma * ap
                break
            if ynew.max() >= Delta:  # r_utils.R:51-56
                ind = np.where(ynew > Delta)[0]
                gamma = np.min((Delta - y[ind]) / ap[ind])
                y = y + gamma * ap
                break
            y = ynew
            rk = rk - alpha * w
            rho_km2 = rho_km1
            if not literal_typo:
                z = rk / v
            rho_km1 = float(rk @ z)  # r_utils.R:60 (stale z when literal_typo)
        x = x * y
        v = x * (A @ x)
        rk = 1.0 - v
        rho_km1 = float(rk @ rk)
        rout = rho_km1
        rat = rout / rold
        rold = rout
        res_norm = np.sqrt(rout)
        eta_o = eta
        eta = g * rat
        if g * eta_o**2 > 0.1:
            eta = max(eta, g * eta_o**2)
        eta = max(min(eta, etamax), stop_tol / res_norm)
    result = (x[:, None] * A) * x[None, :]  # r_utils.R:75  t(t(x*A)*x)
    return np.round(result, 6)  # r_utils.R:90
