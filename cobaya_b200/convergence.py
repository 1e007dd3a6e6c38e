"""
Gelman-Rubin R-1 of means and the learned proposal covariance from the all-reduced
sufficient statistics (host side, D x D work once per checkpoint).

Restates the root-only block of ``MCMC.check_convergence_and_learn_proposal``
(cobaya/samplers/mcmc/mcmc.py:856-889) on sums instead of gathered per-chain arrays:
with M chains, N_c rows, means m_c and covariances C_c,

    W = sum N_c C_c / sum N_c                       (mcmc.py:856  mean_of_covs)
    B = 1/(M-1) sum (m_c - mbar)(m_c - mbar)^T      (mcmc.py:860  np.cov(means.T))
    R-1 = max |eig( L^-1 (B/dd^T) L^-T )|,  L = chol(W/dd^T), d = sqrt(diag B)   (:864-889)

Every rank holds identical sums after the all-reduce, so every rank computes the same
W / R-1 and no broadcast (mcmc.py:914,1005,1021) is needed.
"""

from __future__ import annotations

import numpy as np

from .flatmodel import inverse_cholesky


def unpack_sums(sums: np.ndarray, D: int):
    sums = np.asarray(sums, dtype=np.float64)
    DD = D * D
    M, S1, S2 = sums[0], sums[1], sums[2]
    Sm = sums[3: 3 + D]
    Smm = sums[3 + D: 3 + D + DD].reshape(D, D)
    SC = sums[3 + D + DD: 3 + D + 2 * DD].reshape(D, D)
    return M, S1, S2, Sm, Smm, SC


def rminus1_from_sums(sums: np.ndarray, D: int, shift: np.ndarray | None = None) -> dict:
    M, S1, S2, Sm, Smm, SC = unpack_sums(sums, D)
    shift = np.zeros(D) if shift is None else np.asarray(shift, dtype=np.float64)
    out = dict(M=int(round(M)), N=int(round(S1)), acceptance=S2 / S1,
               mean=Sm / M + shift, success=False, Rminus1=None)
    W = SC / S1
    W = (W + W.T) / 2
    out["W"] = W
    if M < 2:
        return out
    mbar = Sm / M
    B = (Smm - M * np.outer(mbar, mbar)) / (M - 1)
    B = (B + B.T) / 2
    out["B"] = B
    with np.errstate(all="ignore"):
        d = np.sqrt(np.diag(B))
        corr_of_means = (B / d).T / d
        norm_W = (W / d).T / d
        try:
            Linv = inverse_cholesky(norm_W)
            eigvals = np.linalg.eigvalsh(Linv.dot(corr_of_means).dot(Linv.T))
        except np.linalg.LinAlgError:
            return out
    if not np.all(np.isfinite(eigvals)):
        return out
    out["Rminus1"] = float(max(np.abs(eigvals)))
    out["success"] = True
    return out


def rminus1_cl_from_sums(bsums: np.ndarray, D: int, W: np.ndarray) -> float:
    """R-1 of the confidence bounds (mcmc.py:977-982): the standard deviation over chains
    (``np.std``, ddof=0) of each chain's lower/upper bound, in units of sqrt(diag W); the
    maximum over parameters and over the two bounds.  ``bsums`` = all-reduced output of
    ``cb2_bounds``: {M, sum b_low, sum b_low^2, sum b_up, sum b_up^2} (bounds shifted by a
    common vector, which leaves the standard deviation unchanged)."""
    bsums = np.asarray(bsums, dtype=np.float64)
    M = bsums[0]
    out = 0.0
    sd = np.sqrt(np.diag(W))
    for k in range(2):
        s1 = bsums[1 + 2 * D * k: 1 + 2 * D * k + D]
        s2 = bsums[1 + 2 * D * k + D: 1 + 2 * D * k + 2 * D]
        var = np.maximum(s2 / M - (s1 / M) ** 2, 0.0)
        out = max(out, float(np.max(np.sqrt(var) / sd)))
    return out
