"""fp64 NumPy restatement of PyVBMC's ELBO inner loop -- TEST INFRASTRUCTURE ONLY.

Each function cites the reference file:line it restates (paths relative to the
reference checkout, acerbilab/pyvbmc @ 15858017).  The restatement is written
from the mathematics, vectorised differently from the reference, and:

* takes the Monte-Carlo noise ``eps`` as an explicit argument (the reference
  draws it from the global NumPy RNG, ``pyvbmc/entropy/entmc_vbmc.py:64-68``);
  :func:`draw_eps_like_reference` reproduces the reference's draw order;
* evaluates mixture densities through log-sum-exp (the reference evaluates
  ``nconst / sigma**D * exp(-d2/2)`` directly, ``entmc_vbmc.py:77,87-88`` and
  ``entlb_vbmc.py:94-97``, which under/overflows for far-apart components);
  results agree wherever the reference is finite.

Parity status: PINNED by ``tests/test_oracle_golden.py`` (MATLAB known-answer
fixtures of the reference + outputs of the unmodified reference).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace

import numpy as np
import scipy.linalg as sla

LOG2PI = float(np.log(2.0 * np.pi))
EPS64 = float(np.spacing(1.0))


# --------------------------------------------------------------------------
# Variational-posterior carrier + theta packing
# (pyvbmc/variational_posterior/variational_posterior.py:103-138, 623-759)
# --------------------------------------------------------------------------
@dataclass
class OracleVP:
    D: int
    K: int
    mu: np.ndarray  # (D, K)
    sigma: np.ndarray  # (1, K)
    lambd: np.ndarray  # (D, 1)
    w: np.ndarray  # (1, K)
    eta: np.ndarray  # (1, K)
    optimize_mu: bool = True
    optimize_sigma: bool = True
    optimize_lambd: bool = True
    optimize_weights: bool = True
    bounds: dict = field(default=None)

    @classmethod
    def create(cls, D, K, mu, sigma, lambd, w, eta, optimize=(True,) * 4):
        return cls(
            int(D),
            int(K),
            np.array(mu, dtype=float).reshape(D, K),
            np.array(sigma, dtype=float).reshape(1, K),
            np.array(lambd, dtype=float).reshape(D, 1),
            np.array(w, dtype=float).reshape(1, K),
            np.array(eta, dtype=float).reshape(1, K),
            *[bool(o) for o in optimize],
        )

    def copy(self):
        return OracleVP.create(
            self.D,
            self.K,
            self.mu.copy(),
            self.sigma.copy(),
            self.lambd.copy(),
            self.w.copy(),
            self.eta.copy(),
            (
                self.optimize_mu,
                self.optimize_sigma,
                self.optimize_lambd,
                self.optimize_weights,
            ),
        )


def _renormalise(vp: OracleVP):
    # variational_posterior.py:642-649 and :749-756
    nl = np.sqrt(np.sum(vp.lambd**2) / vp.D)
    vp.lambd = vp.lambd.reshape(-1, 1) / nl
    vp.sigma = vp.sigma.reshape(1, -1) * nl
    if vp.optimize_weights:
        vp.w = vp.w.reshape(1, -1) / np.sum(vp.w)


def get_parameters(vp: OracleVP, raw_flag: bool = True) -> np.ndarray:
    """variational_posterior.py:623-678 (also renormalises ``vp`` in place)."""
    _renormalise(vp)
    head = vp.mu.ravel(order="F") if vp.optimize_mu else np.zeros(0)
    tail = []
    if vp.optimize_sigma:
        tail.append(vp.sigma.ravel())
    if vp.optimize_lambd:
        tail.append(vp.lambd.ravel())
    if vp.optimize_weights:
        tail.append(vp.w.ravel())
    tail = np.concatenate(tail) if tail else np.zeros(0)
    return np.concatenate([head, np.log(tail) if raw_flag else tail])


def set_parameters(vp: OracleVP, theta: np.ndarray, raw_flag: bool = True):
    """variational_posterior.py:680-759."""
    theta = np.array(theta, dtype=float)  # private copy (:703)
    D, K = vp.D, vp.K
    pos = 0
    if vp.optimize_mu:
        vp.mu = theta[: D * K].reshape((D, K), order="F")
        pos = D * K
    if vp.optimize_sigma:
        s = theta[pos : pos + K]
        vp.sigma = np.exp(s) if raw_flag else s
        pos += K
    if vp.optimize_lambd:
        l = theta[pos : pos + D]
        vp.lambd = np.exp(l) if raw_flag else l
    if vp.optimize_weights:
        eta = theta[-K:]
        if raw_flag:
            vp.w = np.exp(eta - np.max(eta)).reshape(1, K)
        else:
            vp.w = eta.reshape(1, K)
    _renormalise(vp)


def get_bounds(vp: OracleVP, X: np.ndarray, options: dict, K: int = None) -> dict:
    """Soft bounds from the training inputs (variational_posterior.py:140-239)."""
    K = vp.K if K is None else K
    D = vp.D
    if vp.bounds is None:
        vp.bounds = {
            "mu_lb": np.full(D, np.inf),
            "mu_ub": np.full(D, -np.inf),
            "lnscale_lb": np.full(D, np.inf),
            "lnscale_ub": np.full(D, -np.inf),
        }
    lo, hi = X.min(axis=0), X.max(axis=0)
    b = vp.bounds
    b["mu_lb"] = np.minimum(lo, b["mu_lb"])
    b["mu_ub"] = np.maximum(hi, b["mu_ub"])
    ln_range = np.log(hi - lo)
    b["lnscale_lb"] = np.minimum(b["lnscale_lb"], ln_range + np.log(options["tol_length"]))
    b["lnscale_ub"] = np.maximum(b["lnscale_ub"], ln_range)
    if vp.optimize_weights:
        b["eta_lb"] = -np.inf if options["tol_weight"] == 0 else np.log(0.5 * options["tol_weight"])
        b["eta_ub"] = 0
    lbs, ubs = [], []
    if vp.optimize_mu:
        lbs.append(np.tile(b["mu_lb"], K))
        ubs.append(np.tile(b["mu_ub"], K))
    if vp.optimize_sigma or vp.optimize_lambd:
        lbs.append(np.tile(b["lnscale_lb"], K))
        ubs.append(np.tile(b["lnscale_ub"], K))
    if vp.optimize_weights:
        lbs.append(np.tile(b["eta_lb"], K))
        ubs.append(np.tile(b["eta_ub"], K))
    out = {
        "lb": np.concatenate(lbs),
        "ub": np.concatenate(ubs),
        "tol_con": options["tol_con_loss"],
    }
    if vp.optimize_weights:
        out["weight_threshold"] = max(1 / (4 * K), options["tol_weight"])
        out["weight_penalty"] = options["weight_penalty"]
    return out


def _softmax_jacobian(eta: np.ndarray) -> np.ndarray:
    """dw/deta built from exp(eta) exactly like entmc_vbmc.py:122-129."""
    e = np.exp(np.ravel(eta))
    s = e.sum()
    return np.diag(e) / s - np.outer(e, e) / s**2


def _flat(vp):
    return (
        int(vp.D),
        int(vp.K),
        np.asarray(vp.mu, dtype=float).reshape(int(vp.D), int(vp.K)),
        np.asarray(vp.sigma, dtype=float).ravel(),
        np.asarray(vp.lambd, dtype=float).ravel(),
        np.asarray(vp.w, dtype=float).ravel(),
        np.asarray(vp.eta, dtype=float).ravel(),
    )


def _pack_grad(grad_flags, g_mu, g_sigma, g_lambd, g_w):
    parts = []
    if grad_flags[0]:
        parts.append(g_mu.ravel(order="F"))
    if grad_flags[1]:
        parts.append(g_sigma)
    if grad_flags[2]:
        parts.append(g_lambd)
    if grad_flags[3]:
        parts.append(g_w)
    return np.concatenate(parts) if parts else np.zeros(0)


# --------------------------------------------------------------------------
# Monte-Carlo entropy (pyvbmc/entropy/entmc_vbmc.py:39-134)
# --------------------------------------------------------------------------
def even_ns(Ns) -> int:
    """entmc_vbmc.py:61 -- draws per component, rounded up to even."""
    return int(np.ceil(Ns / 2)) * 2


def draw_eps_like_reference(K: int, Ns, D: int) -> np.ndarray:
    """Consume the global NumPy RNG in the reference's order (entmc_vbmc.py:64-67):
    one ``randn(Ns//2, D)`` block per component.  Returns ``(K, Ns//2, D)``."""
    half = even_ns(Ns) // 2
    return np.stack([np.random.randn(half, D) for _ in range(K)], axis=0)


def entmc(vp, eps_half: np.ndarray, grad_flags=(True,) * 4, jacobian_flag=True, chunk=4096):
    """Monte-Carlo mixture entropy and gradient; ``eps_half`` is ``(K, Ns/2, D)``
    and is mirrored antithetically (entmc_vbmc.py:67-68)."""
    D, K, mu, sigma, lambd, w, eta = _flat(vp)
    eps_half = np.asarray(eps_half, dtype=float).reshape(K, -1, D)
    Ns = 2 * eps_half.shape[1]
    any_grad = any(grad_flags)

    g_mu = np.zeros((D, K))
    g_sigma = np.zeros(K)
    g_lambd = np.zeros(D)
    g_w = np.zeros(K)

    with np.errstate(divide="ignore"):
        logw = np.log(w)
    # log of  nconst / sigma_k^D   (entmc_vbmc.py:53-56,77)
    lognorm = -0.5 * D * LOG2PI - np.sum(np.log(lambd)) - D * np.log(sigma)
    inv_s2 = 1.0 / sigma**2
    H = 0.0

    for j in range(K):
        acc_logq = 0.0  # sum_i log q(x_i)
        acc_ls = np.zeros(D)  # sum_i lsum(x_i) / q(x_i)
        acc_els = np.zeros(D)  # sum_i eps_i * lsum(x_i) / q(x_i)
        acc_nq = np.zeros(K)  # sum_i N_k(x_i) / q(x_i)
        for sign in (1.0, -1.0):
            for a in range(0, Ns // 2, chunk):
                e = sign * eps_half[j, a : a + chunk]  # (n, D)
                x = e * (lambd * sigma[j]) + mu[:, j]  # (n, D)   :70
                t = (x[:, :, None] - mu[None, :, :]) / lambd[None, :, None]  # (n, D, K)
                d2 = np.sum(t * t, axis=1) * inv_s2  # (n, K)   :75-76
                logn = lognorm - 0.5 * d2  # log N_k(x)
                la = logw + logn
                m = la.max(axis=1, keepdims=True)
                logq = (m + np.log(np.exp(la - m).sum(axis=1, keepdims=True)))[:, 0]  # :78-80
                acc_logq += logq.sum()
                if any_grad:
                    nq = np.exp(logn - logq[:, None])  # N_k / q   (n, K)   :85-90
                    # lsum/q, lsum_d = sum_k w_k N_k (x_d - mu_dk) / (sigma_k lambda_d)^2   :93-95
                    ls = np.einsum("ndk,nk->nd", t, nq * (w * inv_s2)) / lambd
                    acc_ls += ls.sum(axis=0)
                    acc_els += (e * ls).sum(axis=0)
                    acc_nq += nq.sum(axis=0)
        H -= w[j] * acc_logq / Ns  # :80
        if grad_flags[0]:
            g_mu[:, j] = w[j] * acc_ls / Ns  # :98
        if grad_flags[1]:
            g_sigma[j] = w[j] * np.sum(acc_els * lambd) / Ns  # :102-103
        if grad_flags[2]:
            g_lambd += w[j] * sigma[j] * acc_els / Ns  # :106-108
        if grad_flags[3]:
            g_w[j] -= acc_logq / Ns  # :111
            g_w -= w[j] * acc_nq / Ns  # :112

    if jacobian_flag and grad_flags[1]:
        g_sigma = g_sigma * sigma  # :115-116
    if jacobian_flag and grad_flags[2]:
        g_lambd = g_lambd * lambd  # :119-120
    if jacobian_flag and grad_flags[3]:
        g_w = _softmax_jacobian(eta) @ g_w  # :123-130
    return float(H), _pack_grad(grad_flags, g_mu, g_sigma, g_lambd, g_w)


# --------------------------------------------------------------------------
# Entropy lower bound (pyvbmc/entropy/entlb_vbmc.py:43-180)
# --------------------------------------------------------------------------
def entlb(vp, grad_flags=(True,) * 4, jacobian_flag=True):
    D, K, mu, sigma, lambd, w, eta = _flat(vp)
    g_mu = np.zeros((D, K))
    g_sigma = np.zeros(K)
    g_lambd = np.zeros(D)
    g_w = np.zeros(K)

    if K == 1:
        # exact single-Gaussian entropy, entlb_vbmc.py:60-78
        H = 0.5 * D * (1 + LOG2PI) + D * np.sum(np.log(sigma)) + np.sum(np.log(lambd))
        g_sigma = D / sigma
        g_lambd = 1.0 / lambd
    else:
        s2 = sigma[:, None] ** 2 + sigma[None, :] ** 2  # (K, K)   :84
        diff = (mu.T[:, None, :] - mu.T[None, :, :]) / lambd  # (K, K, D): (mu_i - mu_j)/lambda
        r2 = np.sum(diff * diff, axis=2)
        # log gamma_ij,  gamma = nconst / s^D exp(-r2/(2 s2))   :87-94
        lg = -0.5 * D * LOG2PI - np.sum(np.log(lambd)) - 0.5 * D * np.log(s2) - 0.5 * r2 / s2
        with np.errstate(divide="ignore"):
            la = lg + np.log(w)[None, :]
        m = la.max(axis=1, keepdims=True)
        lgs = (m + np.log(np.exp(la - m).sum(axis=1, keepdims=True)))[:, 0]  # log gammasum_i
        H = -np.sum(w * lgs)  # :97
        if any(grad_flags):
            gam_i = np.exp(lg - lgs[:, None])  # gamma_ij / gammasum_i
            gam_j = np.exp(lg - lgs[None, :])  # gamma_ij / gammasum_j   (:100-105)
            # symmetric pair weight  w_i w_j gamma_ij (1/gammasum_i + 1/gammasum_j)
            pair = (w[:, None] * w[None, :]) * (gam_i + gam_j)
            if grad_flags[0]:
                # :107-110,121-131   mu_grad[:, j] = -sum_i pair_ij (mu_i - mu_j)/(s2_ij lambda^2)
                g_mu = -np.einsum("ij,ijd->dj", pair / s2, diff) / lambd[:, None]
            if grad_flags[1]:
                # :112-115,133-139
                dsig = -D / s2 + r2 / s2**2
                g_sigma = -sigma * np.sum(pair * dsig, axis=0)
            if grad_flags[2]:
                # :141-156
                inner = np.einsum("ij,ijd->jd", w[:, None] * gam_j, diff * diff / s2[:, :, None] - 1.0)
                g_lambd = -np.sum(w[:, None] * inner, axis=0) / lambd
            if grad_flags[3]:
                # :158-159
                g_w = -lgs - np.sum(w[None, :] * gam_j, axis=1)
    if jacobian_flag and grad_flags[1]:
        g_sigma = g_sigma * sigma  # :162-163
    if jacobian_flag and grad_flags[2]:
        g_lambd = g_lambd * lambd  # :166-167
    if jacobian_flag and grad_flags[3]:
        g_w = _softmax_jacobian(eta) @ g_w  # :170-177
    return float(H), _pack_grad(grad_flags, g_mu, np.ravel(g_sigma), np.ravel(g_lambd), np.ravel(g_w))


# --------------------------------------------------------------------------
# GP-surrogate expected log joint (variational_optimization.py:1238-1606)
# --------------------------------------------------------------------------
def make_gp(X, posts, mean_kind="negquad", noise_N=1, y=None):
    """Plain carrier for what ``_gp_log_joint`` reads from a gpyreg GP
    (variational_optimization.py:1311,1367-1398).  ``posts``: list of dicts with
    ``hyp, alpha, L, L_chol, sW``."""
    X = np.asarray(X, dtype=float)
    return SimpleNamespace(
        X=X,
        y=y,
        D=X.shape[1],
        cov_N=X.shape[1] + 1,
        noise_N=int(noise_N),
        mean_kind=mean_kind,
        posteriors=[
            SimpleNamespace(
                hyp=np.asarray(p["hyp"], dtype=float).ravel(),
                alpha=np.asarray(p["alpha"], dtype=float).ravel(),
                L=None if p.get("L") is None else np.asarray(p["L"], dtype=float),
                L_chol=bool(p["L_chol"]),
                sW=np.asarray(p["sW"], dtype=float).ravel(),
            )
            for p in posts
        ],
    )


def gp_log_joint(
    vp,
    gp,
    grad_flags,
    avg_flag=True,
    jacobian_flag=True,
    compute_var=False,
    separate_K=False,
):
    """Returns ``(G, dG, varG, dvarG, var_ss[, I_sk, J_sjk])`` with the reference's
    conventions (scalars unwrapped when S == 1, ``dG=None`` without flags)."""
    if np.isscalar(grad_flags):
        grad_flags = (bool(grad_flags),) * 4  # :1296-1300
    any_grad = bool(np.any(grad_flags))
    if compute_var and any_grad and compute_var != 2:
        raise NotImplementedError("gradient of the log-joint variance")  # :1302-1307
    if compute_var == 2:
        raise NotImplementedError("diagonal approximation of the variance")  # :1467-1471

    D, K, mu, sigma, lambd, w, eta = _flat(vp)
    X = gp.X
    N = X.shape[0]
    S = len(gp.posteriors)
    quad = gp.mean_kind == "negquad"
    zero = gp.mean_kind == "zero"

    G = np.zeros(S)
    g_mu = np.zeros((D, K, S))
    g_sigma = np.zeros((K, S))
    g_lambd = np.zeros((D, S))
    g_w = np.zeros((K, S))
    varG = np.zeros(S)
    I_sk = np.zeros((S, K))
    J_sjk = np.zeros((S, K, K))

    dX = mu.T[:, :, None] - X.T[None, :, :]  # (K, D, N)   :1362-1364
    sl2 = (sigma[:, None] * lambd[None, :]) ** 2  # (K, D)  (sigma_k lambda_d)^2

    for s, post in enumerate(gp.posteriors):
        hyp = post.hyp
        ell = np.exp(hyp[:D])
        ln_sf2 = 2.0 * hyp[D]
        sum_lnell = np.sum(hyp[:D])
        base = gp.cov_N + gp.noise_N
        m0 = 0.0 if zero else hyp[base]  # :1383-1386
        alpha = post.alpha
        sn2_eff = 1.0 / post.sW[0] ** 2  # :1398

        tau = np.sqrt(sl2 + ell[None, :] ** 2)  # (K, D)   :1401
        lnnf = ln_sf2 + sum_lnell - np.sum(np.log(tau), axis=1)  # (K,)  :1402-1404
        delta = dX / tau[:, :, None]  # (K, D, N)   :1405
        z = np.exp(lnnf[:, None] - 0.5 * np.sum(delta**2, axis=1))  # (K, N)   :1406
        I = z @ alpha + m0  # :1407
        if quad:
            xm = hyp[base + 1 : base + 1 + D]
            omega = np.exp(hyp[base + 1 + D : base + 1 + 2 * D])
            # :1409-1424
            nu = -0.5 * np.sum((mu.T**2 + sl2 - 2.0 * mu.T * xm + xm**2) / omega**2, axis=1)
            I = I + nu
        G[s] = np.sum(w * I)  # :1425
        I_sk[s] = I

        if any_grad:
            za = z * alpha[None, :]  # (K, N)
            if grad_flags[0]:
                # :1430-1436
                gm = -np.einsum("kdn,kn->kd", delta, za) / tau
                if quad:
                    gm = gm - (mu.T - xm) / omega**2
                g_mu[:, :, s] = (w[:, None] * gm).T
            q = np.einsum("kdn,kn->kd", delta**2 - 1.0, za)  # sum_n alpha_n z_kn (delta^2 - 1)
            if grad_flags[1]:
                # :1438-1450
                gs = sigma * np.sum((lambd[None, :] / tau) ** 2 * q, axis=1)
                if quad:
                    gs = gs - sigma * np.sum(lambd**2 / omega**2)
                g_sigma[:, s] = w * gs
            if grad_flags[2]:
                # :1452-1462
                gl = (sigma[:, None] / tau) ** 2 * q * lambd[None, :]
                if quad:
                    gl = gl - (sigma[:, None] ** 2) * lambd[None, :] / omega**2
                g_lambd[:, s] = np.sum(w[:, None] * gl, axis=0)
            if grad_flags[3]:
                g_w[:, s] = I  # :1464-1465

        if compute_var:
            # :1472-1514
            s2sum = sigma[:, None] ** 2 + sigma[None, :] ** 2  # (K, K)
            tau_jk = np.sqrt(s2sum[:, :, None] * lambd[None, None, :] ** 2 + ell**2)  # (K, K, D)
            lnnf_jk = ln_sf2 + sum_lnell - np.sum(np.log(tau_jk), axis=2)
            d_jk = (mu.T[:, None, :] - mu.T[None, :, :]) / tau_jk
            J = np.exp(lnnf_jk - 0.5 * np.sum(d_jk**2, axis=2))
            if post.L_chol:
                # z_k' (L'L)^-1 z_j / sn2_eff via the two triangular solves of :1490-1501
                v = sla.solve_triangular(post.L, z.T, trans=1, check_finite=False)
                u = sla.solve_triangular(post.L, v, trans=0, check_finite=False)
                J = J - (z @ u) / sn2_eff
            else:
                J = J + z @ (post.L @ z.T)  # :1502-1503
            # the reference evaluates only j <= k, as  J_jk = prior_jk - z_k . solve(z_j)  (:1473-1503);
            # in this layout that is the lower triangle J[k, j] -- use exactly those entries.
            Jl = np.tril(J, -1)
            Jd = np.diag(J)
            varG[s] = np.sum(w**2 * np.maximum(EPS64, Jd)) + 2.0 * np.sum(np.outer(w, w) * Jl)  # :1505-1514
            J_sjk[s] = Jl + Jl.T + np.diag(Jd)

    if compute_var:
        varG = np.maximum(varG, EPS64)  # :1517-1518
    else:
        varG = None

    if any_grad:
        parts = []
        if grad_flags[0]:
            parts.append(g_mu.reshape((D * K, S), order="F"))  # :1525
        if jacobian_flag and grad_flags[1]:
            parts.append(g_sigma * sigma[:, None])  # :1529-1531
        if jacobian_flag and grad_flags[2]:
            parts.append(g_lambd * lambd[:, None])  # :1534-1536
        if jacobian_flag and grad_flags[3]:
            parts.append(_softmax_jacobian(eta) @ g_w)  # :1539-1546
        dG = np.concatenate(parts, axis=0)
    else:
        dG = None
    dvarG = None

    var_ss = 0
    if S > 1 and avg_flag:  # :1578-1596
        G_bar = np.sum(G) / S
        if compute_var:
            varG_ss = np.sum((G - G_bar) ** 2) / (S - 1)
            var_ss = varG_ss + np.std(varG, ddof=1)  # sic: a std added to a variance (:1586)
            varG = np.sum(varG) / S + varG_ss
        G = G_bar
        if any_grad:
            dG = np.sum(dG, axis=1) / S
    if S == 1:  # :1598-1602
        G = G[0]
        if any_grad:
            dG = dG[:, 0]
    if separate_K:
        return G, dG, varG, dvarG, var_ss, I_sk, (J_sjk if compute_var else None)
    return G, dG, varG, dvarG, var_ss


# --------------------------------------------------------------------------
# Soft bounds (variational_optimization.py:503-657)
# --------------------------------------------------------------------------
def soft_bound_loss(x, slb, sub, tol_con=1e-3, compute_grad=False):
    x = np.asarray(x, dtype=float)
    slb = np.asarray(slb, dtype=float)
    sub = np.asarray(sub, dtype=float)
    ell = (sub - slb) * tol_con  # :639
    below = x < slb
    above = x > sub
    with np.errstate(invalid="ignore", divide="ignore"):
        viol = np.where(below, x - slb, np.where(above, x - sub, 0.0))
        y = 0.5 * float(np.sum(np.where(below | above, (viol / ell) ** 2, 0.0)))
        if not compute_grad:
            return y
        dy = np.where(below | above, viol / ell**2, 0.0)
    return y, dy


def vp_bound_loss(vp, theta, theta_bnd, tol_con=1e-3, compute_grad=True):
    D, K = vp.D, vp.K
    theta = np.asarray(theta, dtype=float)
    pos = 0
    if vp.optimize_mu:
        mu = theta[: D * K]
        pos = D * K
    else:
        mu = np.asarray(vp.mu).ravel(order="F")
    if vp.optimize_sigma:
        ln_sigma = theta[pos : pos + K]
        pos += K
    else:
        ln_sigma = np.log(np.ravel(vp.sigma))
    if vp.optimize_lambd:
        ln_lambd = theta[pos : pos + D]
    else:
        ln_lambd = np.log(np.ravel(vp.lambd))
    ln_scale = ln_lambd[:, None] + ln_sigma[None, :]  # (D, K)   :557
    ext = []
    if vp.optimize_mu:
        ext.append(mu)
    # the reference tests ``vp.optimize_sigma or vp.optimize_lambda`` (sic, :561); sigma is
    # always optimised so the block is always present
    ext.append(ln_scale.ravel(order="F"))
    if vp.optimize_weights:
        ext.append(theta[-K:])
    ext = np.concatenate(ext)
    lb = np.ravel(theta_bnd["lb"])
    ub = np.ravel(theta_bnd["ub"])
    if not compute_grad:
        return soft_bound_loss(ext, lb, ub, tol_con)
    L, dL = soft_bound_loss(ext, lb, ub, tol_con, compute_grad=True)
    out = []
    pos = 0
    if vp.optimize_mu:
        out.append(dL[: D * K])
        pos = D * K
    # NOTE the reference reshapes the (column-major) ln-scale block row-major here
    # (np.reshape(..., (D, K)) at :584-586) -- replicated literally.
    dls = np.reshape(dL[pos : pos + D * K], (D, K))
    if vp.optimize_sigma:
        out.append(dls.sum(axis=0))
    if vp.optimize_lambd:
        out.append(dls.sum(axis=1))
    if vp.optimize_weights:
        out.append(dL[-K:])
    return L, np.concatenate(out)


# --------------------------------------------------------------------------
# Negative ELCBO (variational_optimization.py:991-1235)
# --------------------------------------------------------------------------
def neg_elcbo(
    theta,
    gp,
    vp: OracleVP,
    beta=0.0,
    Ns=0,
    compute_grad=True,
    compute_var=None,
    theta_bnd=None,
    entropy_alpha=0.0,
    separate_K=False,
    eps_half=None,
):
    """``eps_half`` (K, Ns/2, D) replaces the reference's global-RNG draws; when it
    is None and Ns > 0 the draws are taken from ``np.random`` in the reference's order."""
    if not np.isfinite(beta):
        beta = 0
    if compute_var is None:
        compute_var = beta != 0
    if compute_grad and beta != 0 and compute_var != 2:
        raise NotImplementedError("gradient of ELBO with full variance")  # :1066-1070
    K = vp.K
    theta = np.asarray(theta, dtype=float)
    set_parameters(vp, theta)  # :1080
    if vp.optimize_weights:
        # :1082-1085 -- `vp.eta = theta[-K:]` is a VIEW of the caller's array and `vp.eta -= amax(vp.eta)` shifts it
        # in place: the caller's theta leaves this function with max(eta) == 0 (minimize_adam's iterate is
        # renormalised on every call) and the soft-bound loss below reads the SHIFTED eta.
        eta = theta[-K:]
        eta -= np.max(eta)
        vp.eta = eta.reshape(1, -1)
    if compute_grad:
        grad_flags = (vp.optimize_mu, vp.optimize_sigma, vp.optimize_lambd, vp.optimize_weights)
    else:
        grad_flags = (False,) * 4

    I_sk = J_sjk = None
    varG = varG_ss = 0
    dG = None
    if separate_K:
        if compute_grad:
            raise ValueError("gradient and per-component results requested together")  # :1114-1118
        if compute_var:
            G, _, varG, _, varG_ss, I_sk, J_sjk = gp_log_joint(vp, gp, grad_flags, 1, 1, compute_var, True)
        else:
            G, dG, _, _, _, I_sk, _ = gp_log_joint(vp, gp, grad_flags, 1, 1, 0, True)
    else:
        if compute_var:
            G, dG, varG, dvarG, varG_ss = gp_log_joint(vp, gp, grad_flags, 1, 1, compute_var)
        else:
            G, dG, _, _, _ = gp_log_joint(vp, gp, grad_flags, 1, 1, 0)

    if Ns > 0:  # :1163-1168
        if eps_half is None:
            eps_half = draw_eps_like_reference(K, Ns, vp.D)
        H, dH = entmc(vp, eps_half, grad_flags, 1)
    else:
        H, dH = entlb(vp, grad_flags, 1)

    F = -G - H
    if compute_grad:
        dF = -dG - dH
    else:
        dF = None
        dH = None
    varH = 0
    varF = varG + varH if compute_var else 0
    if beta != 0:  # dead in practice (elcbo_beta = 0), value-only branch kept (:1186-1189)
        F += beta * np.sqrt(varF)

    if theta_bnd is not None:  # :1195-1229
        if compute_grad:
            L, dL = vp_bound_loss(vp, theta, theta_bnd, tol_con=theta_bnd["tol_con"])
            dF = dF + dL
        else:
            L = vp_bound_loss(vp, theta, theta_bnd, tol_con=theta_bnd["tol_con"], compute_grad=False)
        F += L
        if vp.optimize_weights:
            thresh = theta_bnd["weight_threshold"]
            pen = theta_bnd["weight_penalty"]
            wv = np.ravel(vp.w)
            F += np.sum(np.where(wv < thresh, wv, thresh)) * pen  # :1213-1219
            if compute_grad:
                gw = _softmax_jacobian(vp.eta) @ (pen * (wv < thresh))  # :1221-1226
                dF = dF.copy()
                dF[-K:] += gw
    if separate_K:
        return F, dF, G, H, varF, dH, varG_ss, varG, varH, I_sk, J_sjk
    return F, dF, G, H, varF
