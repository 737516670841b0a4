"""B200 replacements for the O(Nx) ingredients of PyVBMC's acquisition functions.

The reference evaluates, for a batch of search points ``Xs`` (up to ``search_cache`` = 8192 per active-sampling
step, option_configs/advanced_vbmc_options.ini:41),

    f_mu, f_s2 = gp.predict(x_star=Xs, separate_samples=True)        abstract_acq_fcn.py:79   (gpyreg)
    log_p      = vp.pdf(Xs, orig_flag=False, log_flag=True)          acq_fcn_log.py:38-42

and combines them on the host.  Here both run in the CUDA library (``csrc/gppred.cu``) for the GP that the ELBO path
already holds on the device; the combination stays NumPy, with the reference's signatures:

  gp_predict(gp, x_star, separate_samples=False)  -> (f_mu, f_s2)     gpyreg ``GP.predict`` (latent function)
  vp_pdf(vp, x, orig_flag, log_flag, grad_flag, df)                    ``VariationalPosterior.pdf``
  AcqFcnLog()(Xs, gp, vp, function_logger, optim_state)                abstract_acq_fcn.py:34-147 + acq_fcn_log.py

There is no CPU fallback: without the built library or a CUDA device these raise.
"""
from __future__ import annotations

import sys

import numpy as np

from .context import context_for_gp, entropy_context


def gp_predict(gp, x_star, separate_samples=False):
    """Predictive mean and variance of the latent function at ``x_star`` ``(Nx, D)``.

    ``separate_samples=True``: ``(Nx, S)`` arrays, one column per hyper-parameter sample (what the acquisition
    functions ask for); else the moments of the equal-weight mixture over samples, ``(Nx, 1)`` each."""
    x_star = np.asarray(x_star, dtype=float)
    if x_star.ndim == 1:
        x_star = x_star[None, :]
    ctx = context_for_gp(gp, need_L=True)
    f_mu, f_s2 = ctx.gp_predict(x_star)
    S = f_mu.shape[1]
    if S > 1 and not separate_samples:
        fbar = np.sum(f_mu, axis=1, keepdims=True) / S
        vf = np.sum((f_mu - fbar) ** 2, axis=1, keepdims=True) / (S - 1)
        return fbar, np.sum(f_s2, axis=1, keepdims=True) / S + vf
    return f_mu, f_s2


def vp_pdf(vp, x, orig_flag=True, log_flag=False, grad_flag=False, df=np.inf):
    """``VariationalPosterior.pdf`` (variational_posterior.py:241-552) for the Gaussian mixture (``df`` infinite or 0).

    Transformed space (``orig_flag=False``) needs nothing but the mixture arrays.  Original space follows the reference:
    points outside the original bounds get density 0 (``-inf`` with ``log_flag``), the others are mapped with
    ``vp.parameter_transformer`` and corrected by its log-Jacobian (:429-440, :524-552)."""
    if np.isfinite(df) and df != 0:
        raise NotImplementedError("heavy-tailed variants of the variational posterior pdf are not on the device path")
    if orig_flag and log_flag and grad_flag:
        raise NotImplementedError("vbmc_pdf:NoOriginalGrad: Gradient computation in original space not supported yet.")
    x = np.array(x, dtype=float, copy=True)
    if x.ndim == 1:
        x = x[None, :] if x.size == vp.D else x[:, None]
    N = x.shape[0]
    mask = np.full(N, True)
    if orig_flag:
        pt = vp.parameter_transformer
        mask = np.logical_and(np.all(x > pt.lb_orig, axis=1), np.all(x < pt.ub_orig, axis=1))
        x[mask] = pt(x[mask])
    y, dy = entropy_context().vp_pdf(vp, x, log_flag=log_flag, grad_flag=grad_flag)
    y = y[:, None]
    y[~mask] = -np.inf if log_flag else 0
    if orig_flag:
        ladj = vp.parameter_transformer.log_abs_det_jacobian(x[mask])[:, np.newaxis]
        if log_flag:
            y[mask] -= ladj
        else:
            y[mask] /= np.exp(ladj)
    if grad_flag:
        return y, dy
    return y


def total_variance(f_mu, f_s2):
    """``(f_bar, var_tot)``: mean over hyper-samples and total variance (abstract_acq_fcn.py:82-97)."""
    Ns = f_mu.shape[1]
    f_bar = np.sum(f_mu, axis=1, keepdims=True) / Ns
    var_bar = np.sum(f_s2, axis=1, keepdims=True) / Ns
    var_f = np.sum((f_mu - f_bar) ** 2, axis=1, keepdims=True) / (Ns - 1) if Ns > 1 else 0
    return np.ravel(f_bar), np.ravel(var_f + var_bar)


class AcqFcnLog:
    """Prospective uncertainty search, log-valued (acq_fcn_log.py:12-50), with the ``AbstractAcqFcn.__call__``
    protocol (abstract_acq_fcn.py:34-147): same arguments, same post-processing (variance regularisation, clamp at
    ``-realmax``, hard-bound masking), GP prediction and mixture density on the device."""

    def __init__(self):
        self.acq_info = {"compute_var_log_joint": False, "log_flag": True}

    def get_info(self):
        return self.acq_info

    def __call__(self, Xs, gp, vp, function_logger, optim_state):
        Xs = np.asarray(Xs, dtype=float)
        if Xs.ndim == 1:
            Xs = Xs[None, :]
        if np.any(optim_state.get("integer_vars")):
            raise NotImplementedError("integer-valued variables are not handled by the device acquisition path")
        f_mu, f_s2 = gp_predict(gp, Xs, separate_samples=True)
        f_bar, var_tot = total_variance(f_mu, f_s2)
        log_p = np.ravel(np.maximum(vp_pdf(vp, Xs, orig_flag=False, log_flag=True), np.log(sys.float_info.min)))
        acq = -(np.log(var_tot) + f_bar - function_logger.y_max + log_p)  # acq_fcn_log.py:45-47
        if optim_state.get("variance_regularized_acq_fcn"):
            tol_var = optim_state.get("tol_gp_var")
            low = var_tot < tol_var
            if np.any(low):
                acq[low] += tol_var / var_tot[low] - 1  # log-valued acquisition (:126-129)
        acq = np.maximum(acq, -sys.float_info.max)
        pt = getattr(vp, "parameter_transformer", None)
        if pt is not None and optim_state.get("lb_eps_orig") is not None:
            X_orig = pt.inverse(Xs)
            out = np.logical_or(np.any(X_orig < optim_state.get("lb_eps_orig"), axis=1),
                                np.any(X_orig > optim_state.get("ub_eps_orig"), axis=1))
            acq[out] = np.inf
        return acq.reshape(-1)
