"""Parameter carrier mirroring the part of ``pyvbmc.variational_posterior.VariationalPosterior``
that sits on the ELBO path: construction defaults (variational_posterior.py:103-138), the
theta <-> (mu, sigma, lambda, w) packing (:623-759) and the soft bounds (:140-239).

It exists so that the B200 path can be used and tested without the reference installed; when
PyVBMC is installed its own ``VariationalPosterior`` works unchanged with every function of
this package (only the attributes ``D, K, mu, sigma, lambd, w, eta, optimize_*`` and the
methods ``set_parameters / get_parameters / get_bounds`` are used).  This is host-side
plumbing of O(D K) values; all arithmetic on the hot path is in the CUDA library.
"""
from __future__ import annotations

import numpy as np


class VariationalPosterior:
    def __init__(self, D: int, K: int = 2, x0=None):
        self.D, self.K = int(D), int(K)
        if x0 is None:
            centres = np.zeros((self.D, self.K))
        else:
            x0 = np.asarray(x0, dtype=float)
            if x0.size == self.D:
                centres = np.tile(x0.reshape(1, -1), (self.K, 1)).T
            else:
                centres = np.tile(x0.T, int(np.ceil(self.K / x0.T.shape[1])))[:, : self.K]
        self.w = np.full((1, self.K), 1.0 / self.K)
        self.eta = np.full((1, self.K), 1.0 / self.K)
        self.mu = centres + 1e-6 * np.random.randn(self.D, self.K)
        self.sigma = np.full((1, self.K), 1e-3)
        self.lambd = np.ones((self.D, 1))
        self.optimize_mu = self.optimize_sigma = self.optimize_lambd = self.optimize_weights = True
        self.bounds = None
        self.stats = None
        self._mode = None

    # ---- normalisation shared by get/set (:642-649, :749-756)
    def _normalise(self):
        scale = np.sqrt(np.sum(np.square(self.lambd)) / self.D)
        self.lambd = np.reshape(self.lambd, (-1, 1)) / scale
        self.sigma = np.reshape(self.sigma, (1, -1)) * scale
        if self.optimize_weights:
            self.w = np.reshape(self.w, (1, -1)) / np.sum(self.w)

    def get_parameters(self, raw_flag=True):
        self._normalise()
        blocks = []
        if self.optimize_sigma:
            blocks.append(np.ravel(self.sigma))
        if self.optimize_lambd:
            blocks.append(np.ravel(self.lambd))
        if self.optimize_weights:
            blocks.append(np.ravel(self.w))
        tail = np.concatenate(blocks) if blocks else np.zeros(0)
        if raw_flag:
            tail = np.log(tail)
        head = np.ravel(self.mu, order="F") if self.optimize_mu else np.zeros(0)
        return np.concatenate((head, tail))

    def set_parameters(self, theta, raw_flag=True):
        theta = np.array(theta, dtype=float, copy=True)
        D, K = self.D, self.K
        n_pos = (K if self.optimize_weights else 0) + (D if self.optimize_lambd else 0) + (K if self.optimize_sigma else 0)
        if not raw_flag and n_pos and np.any(theta[theta.size - n_pos :] < 0.0):
            raise ValueError("sigma, lambda and weights must be positive when raw_flag = False")
        at = 0
        if self.optimize_mu:
            self.mu = theta[: D * K].reshape((D, K), order="F")
            at = D * K
        if self.optimize_sigma:
            blk = theta[at : at + K]
            self.sigma = np.exp(blk) if raw_flag else blk
            at += K
        if self.optimize_lambd:
            blk = theta[at : at + D]
            self.lambd = np.exp(blk) if raw_flag else blk
        if self.optimize_weights:
            blk = theta[theta.size - K :]
            self.w = (np.exp(blk - np.max(blk)) if raw_flag else blk).reshape(1, K)
        self._normalise()
        self._mode = None

    def pdf(self, x, orig_flag=True, log_flag=False, grad_flag=False, df=np.inf):
        """Density of the mixture (variational_posterior.py:241-552) on the device; this mirror carries no parameter
        transformer, so only the transformed space (``orig_flag=False``) is available."""
        from .acquisition_functions import vp_pdf

        if orig_flag and getattr(self, "parameter_transformer", None) is None:
            raise NotImplementedError("pyvbmc_b200.VariationalPosterior has no parameter transformer: use orig_flag=False")
        return vp_pdf(self, x, orig_flag=orig_flag, log_flag=log_flag, grad_flag=grad_flag, df=df)

    def get_bounds(self, X, options, K=None):
        K = self.K if K is None else int(K)
        X = np.asarray(X, dtype=float)
        lo, hi = np.min(X, axis=0), np.max(X, axis=0)
        if self.bounds is None:
            inf = np.full((self.D,), np.inf)
            self.bounds = {"mu_lb": inf.copy(), "mu_ub": -inf, "lnscale_lb": inf.copy(), "lnscale_ub": -inf.copy()}
        b = self.bounds
        b["mu_lb"] = np.minimum(lo, b["mu_lb"])
        b["mu_ub"] = np.maximum(hi, b["mu_ub"])
        span = np.log(hi - lo)
        b["lnscale_lb"] = np.minimum(b["lnscale_lb"], span + np.log(options["tol_length"]))
        b["lnscale_ub"] = np.maximum(b["lnscale_ub"], span)
        if self.optimize_weights:
            b["eta_lb"] = -np.inf if options["tol_weight"] == 0 else np.log(0.5 * options["tol_weight"])
            b["eta_ub"] = 0
        groups = []
        if self.optimize_mu:
            groups.append("mu")
        if self.optimize_sigma or self.optimize_lambd:
            groups.append("lnscale")
        if self.optimize_weights:
            groups.append("eta")
        theta_bnd = {
            "lb": np.concatenate([np.tile(b[g + "_lb"], (K,)) for g in groups]),
            "ub": np.concatenate([np.tile(b[g + "_ub"], (K,)) for g in groups]),
            "tol_con": options["tol_con_loss"],
        }
        if self.optimize_weights:
            theta_bnd["weight_threshold"] = max(1 / (4 * K), options["tol_weight"])
            theta_bnd["weight_penalty"] = options["weight_penalty"]
        return theta_bnd
