"""fp64 restatement of the GP posterior record consumed by the ELBO inner loop
(host-side INPUT PREPARATION for tests and benchmarks; not on the product's compute path).

The arithmetic that produces ``gp.posteriors[s].{alpha, L, L_chol, sW}`` lives in
the third-party package ``gpyreg`` (``pyproject.toml:13`` of the reference:
``gpyreg >= 0.1.0``, a floor, CI installs acerbilab/gpyreg git HEAD).  gpyreg is
not vendored under ``/root/reference`` and is absent from this image, so this
file restates its published algorithm (Rasmussen & Williams Alg. 2.1 in the
gplite / gpyreg parameterisation):

    hyp = [ln ell (D), ln sigma_f, ln sigma_n (noise_N >= 1), m0, x_m (D), ln omega (D)]
    K    = sigma_f^2 exp(-1/2 sum_d ((x_d - x'_d)/ell_d)^2)              (SE-ARD)
    m(x) = m0 - 1/2 sum_d ((x_d - x_m,d)/omega_d)^2                      (negative quadratic)
    sn2_n   = exp(2 ln sigma_n) [+ s2_n for user-provided noise]
    sn2_div = min_n sn2_n ;  sn2_mat = diag(sn2_n / sn2_div)
    L    = chol(K / sn2_div + sn2_mat)  (upper),  L_chol = True          (sn2_div >= 1e-6)
    alpha = L \\ (L' \\ (y - m)) / sn2_div ,   sW = 1 / sqrt(sn2_div)
    low-noise branch (sn2_div < 1e-6): L = -(K + diag(sn2))^-1, L_chol = False

Parity anchor: the reference's own tests pin gpyreg-posterior o ``_gp_log_joint``
jointly against MATLAB goldens (``pyvbmc/testing/vbmc/test_variational_optimization.py:
120-211``); ``tests/test_oracle_golden.py`` shows this restatement reproduces them
(constant Gaussian noise, negative-quadratic mean, Cholesky branch).  The
heteroskedastic and low-noise branches are "parity unpinned" (no reference test
reaches the hot path with them); they only feed opaque ``alpha``/``L`` arrays, so
GPU-vs-oracle parity does not depend on them being gpyreg-exact.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla


def hyp_layout(D: int, noise_N: int = 1, mean_kind: str = "negquad"):
    """Index ranges into ``hyp`` (variational_optimization.py:1378-1392)."""
    cov_N = D + 1
    base = cov_N + noise_N
    n = {"zero": 0, "const": 1, "negquad": 1 + 2 * D}[mean_kind]
    return {"cov_N": cov_N, "noise_N": noise_N, "mean_start": base, "H": base + n}


def se_ard(X1, X2, ell, sf2):
    a = X1 / ell
    b = X2 / ell
    d2 = np.sum(a * a, 1)[:, None] + np.sum(b * b, 1)[None, :] - 2.0 * a @ b.T
    return sf2 * np.exp(-0.5 * np.maximum(d2, 0.0))


def mean_fn(X, hyp, D, noise_N=1, mean_kind="negquad"):
    lay = hyp_layout(D, noise_N, mean_kind)
    b = lay["mean_start"]
    if mean_kind == "zero":
        return np.zeros(X.shape[0])
    m0 = hyp[b]
    if mean_kind == "const":
        return np.full(X.shape[0], m0)
    xm = hyp[b + 1 : b + 1 + D]
    omega = np.exp(hyp[b + 1 + D : b + 1 + 2 * D])
    return m0 - 0.5 * np.sum(((X - xm) / omega) ** 2, axis=1)


def posterior(X, y, hyp, s2=None, noise_N=1, mean_kind="negquad", force_low_noise=False):
    """Posterior record for ONE hyper-parameter sample: dict(hyp, alpha, L, L_chol, sW)."""
    X = np.asarray(X, dtype=float)
    y = np.asarray(y, dtype=float).ravel()
    hyp = np.asarray(hyp, dtype=float).ravel()
    N, D = X.shape
    ell = np.exp(hyp[:D])
    sf2 = np.exp(2.0 * hyp[D])
    sn2 = np.full(N, np.exp(2.0 * hyp[D + 1]))
    if s2 is not None:
        sn2 = sn2 + np.asarray(s2, dtype=float).ravel()
    Kmat = se_ard(X, X, ell, sf2)
    # exact symmetric zero-distance diagonal
    Kmat[np.diag_indices(N)] = sf2
    m = mean_fn(X, hyp, D, noise_N, mean_kind)
    sn2_div = float(np.min(sn2))
    if sn2_div >= 1e-6 and not force_low_noise:
        A = Kmat / sn2_div + np.diag(sn2 / sn2_div)
        L = sla.cholesky(A, lower=False)
        alpha = sla.cho_solve((L, False), y - m) / sn2_div
        return {
            "hyp": hyp,
            "alpha": alpha,
            "L": L,
            "L_chol": True,
            "sW": np.ones(N) / np.sqrt(sn2_div),
        }
    A = Kmat + np.diag(sn2)
    Ainv = np.linalg.inv(A)
    return {
        "hyp": hyp,
        "alpha": Ainv @ (y - m),
        "L": -Ainv,
        "L_chol": False,
        "sW": np.ones(N) / np.sqrt(sn2_div),
    }


def posteriors(X, y, hyps, s2=None, noise_N=1, mean_kind="negquad", force_low_noise=False):
    hyps = np.atleast_2d(np.asarray(hyps, dtype=float))
    return [posterior(X, y, h, s2, noise_N, mean_kind, force_low_noise) for h in hyps]
