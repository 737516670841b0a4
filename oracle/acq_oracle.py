"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the acquisition-function ingredients.

gp_predict     gpyreg's ``GP.predict(x_star, separate_samples=...)`` for the posterior record the hot path consumes.
               gpyreg (``pyproject.toml:13``: ``gpyreg >= 0.1.0``, un-vendored, absent from this image) is restated from
               its published algorithm (Rasmussen & Williams eq. 2.25-2.26 in the gplite parameterisation, the same
               convention as ``workloads/gp_posterior.py``, which reproduces the reference's MATLAB goldens of
               ``_gp_log_joint``); the reference's own use of the same quantities pins the convention:
               ``acq_fcn_viqr.py:107-131`` (cross-covariance ``sf2 exp(-cdist(Xs/ell, X/ell)/2)``, Cholesky branch
               ``K_Xs_Xa - K_Xs_X @ C_tmp``, low-noise branch with the opposite sign).  PARITY UNPINNED against gpyreg
               itself; tests/test_oracle_golden.py checks it against an independent dense solve.
vp_pdf         ``VariationalPosterior.pdf`` for df = inf (variational_posterior.py:425-552), pinned to the unmodified
               reference by tests/golden/ref_acq.npz.
total_variance ``AbstractAcqFcn.__call__`` (abstract_acq_fcn.py:79-97);  acq_log  ``AcqFcnLog`` (acq_fcn_log.py:38-47).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import sys

import numpy as np
import scipy.linalg as sla

from . import gp_posterior as gpp


def gp_predict(X, posts, x_star, mean_kind="negquad", noise_N=1, separate_samples=False):
    """``(f_mu, f_s2)``: ``(Nx, S)`` each with ``separate_samples``, else ``(Nx, 1)`` mixture moments."""
    X = np.asarray(X, dtype=float)
    x_star = np.atleast_2d(np.asarray(x_star, dtype=float))
    N, D = X.shape
    Nx, S = x_star.shape[0], len(posts)
    f_mu = np.zeros((Nx, S))
    f_s2 = np.zeros((Nx, S))
    for s, p in enumerate(posts):
        hyp = np.asarray(p["hyp"], dtype=float).ravel()
        ell, sf2 = np.exp(hyp[:D]), np.exp(2.0 * hyp[D])
        a, b = x_star / ell, X / ell
        d2 = np.sum((a[:, None, :] - b[None, :, :]) ** 2, axis=2)  # direct differences
        Ks = sf2 * np.exp(-0.5 * d2)  # (Nx, N)
        f_mu[:, s] = gpp.mean_fn(x_star, hyp, D, noise_N, mean_kind) + Ks @ np.asarray(p["alpha"], dtype=float).ravel()
        L = np.asarray(p["L"], dtype=float)
        if p["L_chol"]:
            sW = float(np.asarray(p["sW"]).ravel()[0])
            V = sla.solve_triangular(L, sW * Ks.T, trans=1, lower=False, check_finite=False)
            f_s2[:, s] = sf2 - np.sum(V * V, axis=0)
        else:
            f_s2[:, s] = sf2 + np.sum(Ks.T * (L @ Ks.T), axis=0)
        f_s2[:, s] = np.maximum(f_s2[:, s], 0.0)  # remove numerical noise, i.e. negative variances
    if S > 1 and not separate_samples:
        fbar = np.sum(f_mu, axis=1, keepdims=True) / S
        vf = np.sum((f_mu - fbar) ** 2, axis=1, keepdims=True) / (S - 1)
        return fbar, np.sum(f_s2, axis=1, keepdims=True) / S + vf
    return f_mu, f_s2


def vp_pdf(vp, x, log_flag=False, grad_flag=False):
    """Transformed-space pdf of the mixture (orig_flag = False, df = inf): ``y (N, 1)`` [, ``dy (N, D)``]."""
    x = np.atleast_2d(np.asarray(x, dtype=float))
    N, D = x.shape
    mu, sigma, lambd, w = np.asarray(vp.mu), np.ravel(vp.sigma), np.ravel(vp.lambd), np.ravel(vp.w)
    nf = 1.0 / (2 * np.pi) ** (D / 2) / np.prod(lambd)
    y = np.zeros((N, 1))
    dy = np.zeros((N, D))
    for k in range(mu.shape[1]):
        d2 = np.sum(((x - mu[:, k]) / (sigma[k] * lambd)) ** 2, axis=1)
        nn = (nf * w[k] / sigma[k] ** D * np.exp(-0.5 * d2))[:, None]
        y += nn
        dy -= nn * (x - mu[:, k]) / (lambd**2 * sigma[k] ** 2)
    if log_flag:
        dy = dy / y
        with np.errstate(divide="ignore"):
            y = np.where(y == 0, -np.inf, np.log(np.where(y == 0, 1.0, y)))
    return (y, dy) if grad_flag else y


def total_variance(f_mu, f_s2):
    """``(f_bar, var_tot)`` of abstract_acq_fcn.py:82-97."""
    Ns = f_mu.shape[1]
    f_bar = np.sum(f_mu, axis=1, keepdims=True) / Ns
    var_bar = np.sum(f_s2, axis=1, keepdims=True) / Ns
    var_f = np.sum((f_mu - f_bar) ** 2, axis=1, keepdims=True) / (Ns - 1) if Ns > 1 else 0
    return np.ravel(f_bar), np.ravel(var_f + var_bar)


def acq_log(f_bar, var_tot, log_p, y_max):
    """acq_fcn_log.py:38-47 (log prospective uncertainty search)."""
    log_p = np.ravel(np.maximum(log_p, np.log(sys.float_info.min)))
    return -(np.log(var_tot) + f_bar - y_max + log_p)
