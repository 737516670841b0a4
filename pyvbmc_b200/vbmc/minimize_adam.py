"""Device-resident counterpart of ``pyvbmc/vbmc/minimize_adam.py`` for the ELBO objective.

The reference hands ``minimize_adam`` a Python closure (variational_optimization.py:238-249)::

    def vb_train_mc_fun(theta_):
        res = _neg_elcbo(theta_, gp, vp0, elcbo_beta, ns_ent_K, compute_grad=True,
                         compute_var=compute_var, theta_bnd=theta_bnd)
        return res[0], res[1]

and pays a host round trip (pack theta, H2D, launch, D2H, NumPy update) per iteration.  Here the
objective is named by its ingredients, theta / the moment estimates / the iterate table stay in HBM,
and one iteration (theta -> parameters, evaluation, Adam update, clamp; minimize_adam.py:87-104) is a
CUDA graph replayed back to back (captured in iteration pairs: the entropy kernel alternates between two noise
buffers); batches of 20 iterations are issued one ahead of the host, which reads a batch back while the next one
runs.  The control flow around it -- batches of 20 iterations, the linear-fit early-stopping test, the returned
averages (:106-145) -- follows the reference's semantics; the fit is evaluated in closed form.
"""
from __future__ import annotations

import numpy as np

from ..context import context_for_gp
from ..entropy.entmc_vbmc import draw_seed
from .variational_optimization import _pack_params


def minimize_adam_elcbo(
    gp,
    vp,
    x0,
    Ns,
    theta_bnd=None,
    lb=None,
    ub=None,
    tol_fun: float = 0.001,
    max_iter: int = 2000,
    master_min: float = 0.001,
    master_max: float = 0.1,
    master_decay: float = 200,
    use_early_stopping: bool = True,
    *,
    seed=None,
    offset=0,
):
    """``minimize_adam(lambda t: _neg_elcbo(t, gp, vp, 0, Ns, True, False, theta_bnd)[:2], x0, lb, ub, ...)``.

    Returns ``(x, y, x_tab, y_tab, n_iter)`` exactly like the reference.  ``vp`` ends up holding the
    parameters of the last evaluated iterate, as it does after the reference loop.  The Monte-Carlo
    draws are device Philox draws keyed ``(seed, offset + iteration)``; ``seed`` defaults to one draw
    of the global NumPy stream, so ``np.random.seed`` makes the whole run reproducible."""
    x0 = np.array(x0, dtype=float, copy=True).ravel()
    D, K = vp.D, vp.K
    batch_size = 20
    tol_x = 0.001
    tol_x_max = 0.1
    tol_fun_max = tol_fun * 100
    min_iter = batch_size * 2
    n_vars = x0.size
    Ns_even = int(np.ceil(Ns / 2)) * 2
    if Ns_even <= 0:
        raise ValueError("minimize_adam_elcbo needs Ns > 0 (the deterministic entropy goes through SciPy)")

    vp.set_parameters(x0)
    if vp.optimize_weights:
        vp.eta = (x0[-K:] - np.amax(x0[-K:])).reshape(1, -1)
    optimize = (bool(vp.optimize_mu), bool(vp.optimize_sigma), bool(vp.optimize_lambd), bool(vp.optimize_weights))
    ctx = context_for_gp(gp, need_L=False)
    use_bounds = ctx.set_bounds(theta_bnd)
    prm = np.zeros(ctx.param_len(D, K))
    _pack_params(prm, vp, x0, optimize, use_bounds)
    if seed is None:
        seed = draw_seed()
    ctx.adam_init(D, K, prm, x0, optimize, Ns_even, use_bounds, seed, offset, lb, ub, max_iter, master_min, master_max,
                  master_decay)

    x_tab = np.zeros((n_vars, max_iter))
    y_tab = np.full((max_iter,), np.nan)
    i = -1
    issued = 0

    def issue():
        nonlocal issued
        n_ = min(batch_size, max_iter - issued)
        if n_ > 0:
            ctx.adam_enqueue(n_)
            issued += n_
        return n_

    # One batch is always in flight while the host digests the previous one (D2H of its values, the stopping rule):
    # the device never idles at a batch boundary.  If the rule fires, the speculative batch is simply not read.
    n = issue()
    while n > 0:
        n_next = issue()
        y, xs = ctx.adam_fetch(i + 1, n)
        y_tab[i + 1 : i + 1 + n] = y
        x_tab[:, i + 1 : i + 1 + n] = xs.T
        i += n
        if not np.all(np.isfinite(y)):
            raise FloatingPointError("non-finite objective in the device Adam loop (fall back to the host loop)")
        is_minibatch_end = (i + 1) % batch_size == 0
        if use_early_stopping and is_minibatch_end and i + 1 >= min_iter:
            # Stopping rule of the reference (minimize_adam.py:106-138), written out: least-squares line through the
            # last batch of objective values on a centred abscissa, its slope and the slope's variance
            # (RSS / (n - 2) / sum t^2), against the drift of the batch-averaged iterate.
            y_b = y_tab[i - batch_size + 1 : i + 1]
            t = np.arange(batch_size) - 0.5 * (batch_size - 1)
            tt = float(t @ t)
            slope = float(t @ y_b) / tt
            resid = y_b - (np.mean(y_b) + slope * t)
            slope_var = float(resid @ resid) / (batch_size - 2) / tt
            x_now = np.mean(x_tab[:, i - batch_size + 1 : i + 1], axis=1)
            x_prev = np.mean(x_tab[:, i - 2 * batch_size + 1 : i + 1 - batch_size], axis=1)
            dx = np.sqrt(np.sum((x_now - x_prev) ** 2) / batch_size)
            flat_strict = abs(slope) < np.sqrt(slope_var + tol_fun**2)
            flat_loose = abs(slope) < np.sqrt(slope_var + tol_fun_max**2)
            if (dx < tol_x and flat_loose) or (flat_strict and dx < tol_x_max):
                break
        n = n_next

    x = np.mean(x_tab[:, i - batch_size + 1 : i + 1], axis=1)
    y = np.mean(y_tab[i - batch_size + 1 : i + 1])
    # the reference's last f(x) call left vp at the iterate BEFORE the last update
    last = x_tab[:, i - 1] if i >= 1 else x0
    vp.set_parameters(last)
    if vp.optimize_weights:
        vp.eta = (last[-K:] - np.amax(last[-K:])).reshape(1, -1)
    return x, y, x_tab[:, 0 : i + 1], y_tab[0 : i + 1], i + 1
