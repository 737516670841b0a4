"""B200 replacements for the objective part of
``pyvbmc/vbmc/variational_optimization.py``: ``_neg_elcbo`` (:991-1235), ``_gp_log_joint``
(:1238-1606), ``_vp_bound_loss`` (:503-606) and ``_soft_bound_loss`` (:609-657).

Signatures, return tuples, exception types and the side effects on ``vp`` follow the
reference; the arithmetic runs in the CUDA library behind ``include/vbmc_b200.h``.
"""
from __future__ import annotations

import math

import numpy as np

from ..config import config
from ..context import context_for_gp
from ..entropy.entmc_vbmc import draw_eps_numpy, draw_seed


def _as_flags(grad_flags):
    if np.isscalar(grad_flags):
        return (bool(grad_flags),) * 4  # :1296-1300
    return tuple(bool(g) for g in grad_flags)


def _gp_log_joint(vp, gp, grad_flags, avg_flag=True, jacobian_flag=True, compute_var=False, separate_K=False):
    """Expected variational log joint via the GP surrogate.

    Returns ``(G, dG, varG, dvarG, var_ss)`` or, with ``separate_K``,
    ``(G, dG, varG, dvarG, var_ss, I_sk, J_sjk)`` exactly like the reference (scalars
    unwrapped when there is a single hyper-parameter sample, ``dG=None`` without flags,
    ``varG=None`` unless requested, ``dvarG`` always ``None``)."""
    g = _as_flags(grad_flags)
    compute_vargrad = compute_var and any(g)
    if compute_vargrad and compute_var != 2:
        raise NotImplementedError(
            "Computation of gradient of log joint variance is currently "
            "available only for diagonal approximation of the variance."
        )  # :1302-1307
    if compute_var == 2:
        raise NotImplementedError("Diagonal approximation of GP log-joint variance not implemented.")  # :1467-1471
    ctx = context_for_gp(gp, need_L=bool(compute_var))
    r = ctx.gplogjoint(vp, g, avg_flag, jacobian_flag, int(bool(compute_var)), separate_K)
    S = ctx.S
    G, dG = r["G"], r["dG"]
    varG = r["varG"] if compute_var else None
    var_ss = r["var_ss"] if (compute_var and S > 1 and avg_flag) else 0
    if not r["per_s"]:
        G = float(G[0])  # averaged (:1594) or S == 1 (:1598-1602)
        if varG is not None:
            varG = float(varG[0]) if (S > 1 and avg_flag) else (varG if S > 1 else varG[:1])
    if separate_K:
        return G, dG, varG, None, var_ss, r["I_sk"], r["J_sjk"]
    return G, dG, varG, None, var_ss


def _soft_bound_loss(x, slb, sub, tol_con=1e-3, compute_grad=False):
    """Quadratic penalty outside ``[slb, sub]`` (:609-657).  O(len(x)) host arithmetic on the
    caller's vectors; inside ``_neg_elcbo`` the same loss is evaluated by the CUDA finalize
    kernel, this stand-alone form exists for API parity with the reference's tests."""
    x = np.asarray(x, dtype=float)
    slb = np.asarray(slb, dtype=float)
    sub = np.asarray(sub, dtype=float)
    ell = (sub - slb) * tol_con
    y = 0.0
    dy = np.zeros(x.shape)
    for idx, edge in ((x < slb, slb), (x > sub, sub)):
        if np.any(idx):
            y += 0.5 * np.sum(((x[idx] - edge[idx]) / ell[idx]) ** 2)
            if compute_grad:
                dy[idx] = (x[idx] - edge[idx]) / ell[idx] ** 2
    if compute_grad:
        return y, dy
    return y


def _shift_eta(vp, theta, K):
    """``vp.eta = theta[-K:]; vp.eta -= amax(vp.eta); vp.eta = reshape(vp.eta, (1, -1))`` (:1082-1085).

    The slice is a view, so the subtraction lands in the caller's ``theta``: after ``_neg_elcbo`` returns, the
    caller's array has ``max(eta) == 0`` (``minimize_adam`` keeps iterating on that renormalised array,
    minimize_adam.py:87-98), ``vp.eta`` aliases it, and ``_vp_bound_loss`` (:1195-1209) reads the shifted eta."""
    eta = theta[-K:]
    eta -= np.amax(eta)
    vp.eta = np.reshape(eta, (1, -1))


def _bound_inputs(vp, theta):
    """The pieces of theta that ``_vp_bound_loss`` reads (:536-555)."""
    D, K = vp.D, vp.K
    pos = D * K if vp.optimize_mu else 0
    if vp.optimize_sigma:
        ln_sigma = theta[pos : pos + K]
        pos += K
    else:
        ln_sigma = np.log(np.ravel(vp.sigma))
    if vp.optimize_lambd:
        ln_lambd = theta[pos : pos + D]
    else:
        ln_lambd = np.log(np.ravel(vp.lambd))
    eta = theta[-K:] if vp.optimize_weights else None
    return ln_sigma, ln_lambd, eta


def _vp_bound_loss(vp, theta, theta_bnd, tol_con=1e-3, compute_grad=True):
    """Soft-bound loss on ``[mu | ln sigma_k + ln lambda_d | eta]`` (:503-606), including the
    reference's row-major reshape of the ln-scale gradient block (:584-586)."""
    D, K = vp.D, vp.K
    theta = np.asarray(theta, dtype=float)
    ln_sigma, ln_lambd, eta = _bound_inputs(vp, theta)
    mu = theta[: D * K] if vp.optimize_mu else np.asarray(vp.mu).ravel(order="F")
    ln_scale = np.reshape(ln_lambd, (-1, 1)) + np.reshape(ln_sigma, (1, -1))
    ext = []
    if vp.optimize_mu:
        ext.append(mu.ravel())
    ext.append(ln_scale.ravel(order="F"))
    if vp.optimize_weights:
        ext.append(np.ravel(eta))
    ext = np.concatenate(ext)
    lb, ub = np.ravel(theta_bnd["lb"]), np.ravel(theta_bnd["ub"])
    if not compute_grad:
        return _soft_bound_loss(ext, lb, ub, tol_con)
    L, dL = _soft_bound_loss(ext, lb, ub, tol_con, compute_grad=True)
    parts = []
    pos = 0
    if vp.optimize_mu:
        parts.append(dL[: D * K])
        pos = D * K
    dls = np.reshape(dL[pos : pos + D * K], (D, K))
    if vp.optimize_sigma:
        parts.append(np.sum(dls, axis=0))
    if vp.optimize_lambd:
        parts.append(np.sum(dls, axis=1))
    if vp.optimize_weights:
        parts.append(dL[-K:])
    return L, np.concatenate(parts)


def _pack_params(prm, vp, theta, optimize, use_bounds):
    """One parameter block of the C ABI (``ParamLayout`` in csrc/common.cuh) from a ``vp`` whose
    ``set_parameters(theta)`` has just run: ``[mu | sigma | lambda | w | eta | bound inputs]``."""
    D, K = vp.D, vp.K
    DK = D * K
    prm[:DK] = theta[:DK] if optimize[0] else np.ravel(vp.mu, order="F")
    prm[DK : DK + K] = np.ravel(vp.sigma)
    prm[DK + K : DK + K + D] = np.ravel(vp.lambd)
    prm[DK + K + D : DK + 2 * K + D] = np.ravel(vp.w)
    prm[DK + 2 * K + D : DK + 3 * K + D] = np.ravel(vp.eta)
    if use_bounds:
        ln_sigma_b, ln_lambd_b, eta_b = _bound_inputs(vp, theta)
        prm[DK + 3 * K + D : DK + 4 * K + D] = ln_sigma_b
        prm[DK + 4 * K + D : DK + 4 * K + 2 * D] = ln_lambd_b
        if eta_b is not None:
            prm[DK + 4 * K + 2 * D :] = eta_b


def neg_elcbo_batch(vp_vec, gp, theta_bnd=None, thetas=None):
    """Value-only negative ELCBO of MANY candidate variational posteriors in one launch.

    Replaces the loop of the reference's ``_sieve`` (variational_optimization.py:775-787)::

        for i, vp0 in enumerate(vp0_vec):
            theta = vp0.get_parameters()
            nelbo_tmp, _, _, _, varF_tmp = _neg_elcbo(theta, gp, vp0, 0, ns_ent_K_fast, 0, compute_var, theta_bnd)

    for the default ``ns_ent_K_fast == 0`` (deterministic entropy bound) and ``compute_var == 0``.
    Every ``vp0`` is mutated exactly as that loop does (``get_parameters`` renormalises it in place,
    ``_neg_elcbo`` calls ``set_parameters(theta)`` and shifts ``eta``, :1080-1085).  All candidates
    must share ``D``, ``K`` and the ``optimize_*`` flags.  Returns ``(F, G, H)`` arrays of length B."""
    vp_vec = list(vp_vec)
    B = len(vp_vec)
    if B == 0:
        z = np.zeros(0)
        return z, z.copy(), z.copy()
    vp0 = vp_vec[0]
    D, K = vp0.D, vp0.K
    optimize = (bool(vp0.optimize_mu), bool(vp0.optimize_sigma), bool(vp0.optimize_lambd), bool(vp0.optimize_weights))
    ctx = context_for_gp(gp, need_L=False)
    use_bounds = ctx.set_bounds(theta_bnd)
    n = ctx.param_len(D, K)
    prm = np.zeros((B, n))
    for b, vp in enumerate(vp_vec):
        if (vp.D, vp.K) != (D, K) or (bool(vp.optimize_mu), bool(vp.optimize_sigma), bool(vp.optimize_lambd),
                                      bool(vp.optimize_weights)) != optimize:
            raise ValueError("neg_elcbo_batch: all candidates must share D, K and the optimize_* flags")
        theta = np.asarray(vp.get_parameters() if thetas is None else thetas[b], dtype=float)
        vp.set_parameters(theta)  # :1080
        if vp.optimize_weights:
            _shift_eta(vp, theta, K)  # :1082-1085
        _pack_params(prm[b], vp, theta, optimize, use_bounds)
    out = ctx.negelcbo_batch(D, K, prm, optimize, use_bounds)
    return out[:, 0].copy(), out[:, 1].copy(), out[:, 2].copy()


def _neg_elcbo(
    theta,
    gp,
    vp,
    beta=0.0,
    Ns=0,
    compute_grad=True,
    compute_var=None,
    theta_bnd=None,
    _entropy_alpha=0.0,
    separate_K=False,
    *,
    eps=None,
    seed=None,
    offset=0,
):
    """Negative evidence lower confidence bound and its gradient.

    Drop-in for the reference function: returns ``(F, dF, G, H, varF)`` or the 11-tuple
    ``(F, dF, G, H, varF, dH, varG_ss, varG, varH, I_sk, J_sjk)`` with ``separate_K``.
    ``vp`` is mutated exactly like the reference does (``set_parameters(theta)`` and the
    shifted ``eta``, :1080-1085).  Keyword-only extensions: ``eps`` (explicit draws,
    shape ``(K, Ns/2, D)``), ``seed`` and ``offset`` (explicit Philox key)."""
    if not math.isfinite(beta):
        beta = 0
    if compute_var is None:
        compute_var = beta != 0
    if compute_grad and beta != 0 and compute_var != 2:
        raise NotImplementedError("Computation of the gradient of ELBO with full variance not supported")  # :1066-1070
    if separate_K and compute_grad:
        raise ValueError(
            "Computing the gradient of variational parameters and requesting per-component results at the same time."
        )  # :1114-1118
    if compute_var == 2:
        raise NotImplementedError("Diagonal approximation of GP log-joint variance not implemented.")

    theta = np.asarray(theta, dtype=float)
    K, D = vp.K, vp.D
    optimize = (bool(vp.optimize_mu), bool(vp.optimize_sigma), bool(vp.optimize_lambd), bool(vp.optimize_weights))
    Ns_even = math.ceil(Ns / 2) * 2 if Ns > 0 else 0  # (math, not NumPy: scalar ufunc calls cost ~1 us each on this path)

    if (not compute_var and not separate_K and eps is None and any(optimize) and theta.ndim == 1
            and (Ns_even == 0 or config.rng_mode != "numpy")):
        # Hot path (minimize_adam / BFGS / sieve objective): the raw theta goes to the device, where set_parameters
        # (:1080), the eta shift (:1082-1085) and the bound-loss inputs are evaluated; the host only moves O(P) bytes.
        P_in = (D * K if optimize[0] else 0) + (K if optimize[1] else 0) + (D if optimize[2] else 0) + (K if optimize[3] else 0)
        if theta.size != P_in:
            raise ValueError(f"theta has {theta.size} entries, the optimised groups of this posterior need {P_in}")
        ctx = context_for_gp(gp, need_L=False)
        if Ns_even > 0:
            if seed is None:
                seed = draw_seed()
            if config.host_noise_prefetch:  # theta-independent: overlaps the rest of this call's host work
                ctx.noise_prefetch(D, K, Ns_even, seed, offset)
        use_bounds = ctx.set_bounds(theta_bnd)
        out, vpo, tmpl, ptrs = ctx.theta_buffers(D, K)
        th = theta if theta.flags.c_contiguous and theta.flags.writeable else np.array(theta, dtype=float)
        DK = D * K
        if optimize == (True, True, True, True):
            tm = None
        else:  # the groups theta does not carry come from the posterior as it stands (variational_posterior.py:680-759)
            tm = tmpl
            if not optimize[0]:
                tm[:DK] = np.ravel(vp.mu, order="F")
            tm[DK : DK + K] = np.ravel(vp.sigma)
            tm[DK + K : DK + K + D] = np.ravel(vp.lambd)
            tm[DK + K + D : DK + 2 * K + D] = np.ravel(vp.w)
            tm[DK + 2 * K + D : DK + 3 * K + D] = np.ravel(vp.eta)
        ctx.negelcbo_theta(D, K, th, tm is not None, optimize, Ns_even, compute_grad, use_bounds, seed or 0, offset, None, ptrs)
        # side effects on vp and on the caller's theta, as the reference leaves them
        if optimize[0]:
            vp.mu = np.array(th[:DK]).reshape((D, K), order="F")
        v = vpo.copy()  # ONE copy of the pinned block; sigma / lambda / w are views of it (three copies cost 2 us more)
        vp.sigma = v[:K].reshape(1, K)
        vp.lambd = v[K : K + D].reshape(D, 1)
        if optimize[3]:
            vp.w = v[K + D :].reshape(1, K)
            if th is not theta:
                theta[-K:] = th[-K:]
            vp.eta = theta[-K:].reshape(1, -1)  # a view of the caller's array (:1082-1085)
        vp._mode = None
        dF = None
        if compute_grad:
            P = (DK if optimize[0] else 0) + (K if optimize[1] else 0) + (D if optimize[2] else 0) + (K if optimize[3] else 0)
            dF = out[8 : 8 + P].copy()
        return float(out[0]), dF, float(out[1]), float(out[2]), 0  # (beta != 0 implies compute_var: never here)

    vp.set_parameters(theta)  # :1080
    if vp.optimize_weights:
        _shift_eta(vp, theta, K)  # :1082-1085 (in place on the CALLER's theta, like the reference)

    ctx = context_for_gp(gp, need_L=bool(compute_var))
    use_bounds = ctx.set_bounds(theta_bnd)

    if Ns_even > 0 and eps is None and seed is None:
        if config.rng_mode == "numpy":
            eps = draw_eps_numpy(K, Ns_even, D)
        else:
            seed = draw_seed()
    if eps is not None:
        eps = np.ascontiguousarray(eps, dtype=float)
        if eps.size != K * (Ns_even // 2) * D:
            raise ValueError("eps must have shape (K, Ns/2, D)")

    if not compute_var and not separate_K:
        # hot path (minimize_adam / BFGS / sieve objective): one packed block in, one packed block out
        prm, out = ctx.flat_buffers(D, K)
        DK = D * K
        _pack_params(prm, vp, theta, optimize, use_bounds)
        ctx.negelcbo_flat(D, K, prm, optimize, Ns_even, compute_grad, use_bounds, eps, seed or 0, None, False, out,
                          offset=offset)
        F, G, H = float(out[0]), float(out[1]), float(out[2])
        dF = None
        if compute_grad:
            P = (DK if optimize[0] else 0) + (K if optimize[1] else 0) + (D if optimize[2] else 0) + (K if optimize[3] else 0)
            dF = out[8 : 8 + P].copy()
        return F, dF, G, H, 0  # (beta != 0 implies compute_var and never takes this branch)

    ln_sigma_b = ln_lambd_b = eta_b = None
    if use_bounds:
        ln_sigma_b, ln_lambd_b, eta_b = _bound_inputs(vp, theta)
    r = ctx.negelcbo(
        vp, optimize, Ns_even, compute_grad, bool(compute_var), separate_K, use_bounds,
        ln_sigma_b, ln_lambd_b, eta_b, eps=eps, seed=seed or 0,
    )
    F, G, H = r["F"], r["G"], r["H"]
    dF = r["dF"] if compute_grad else None
    dH = r["dH"] if compute_grad else None
    varH = 0  # :1179
    varG = r["varF"] if compute_var else 0
    if compute_var and ctx.S == 1:
        varG = np.array([varG])  # the reference only unwraps G and dG for a single hyper-sample (:1598-1602)
    varG_ss = r["varG_ss"] if compute_var else 0
    varF = varG + varH if compute_var else 0
    if beta != 0:  # dead in practice: elcbo_beta is hard-wired to 0 (:743); value-only branch
        F += beta * np.sqrt(varF)
    if separate_K:
        return F, dF, G, H, varF, dH, varG_ss, varG, varH, r["I_sk"], r["J_sjk"]
    return F, dF, G, H, varF
