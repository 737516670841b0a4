"""Seeded synthetic workloads of SURVEY.md section 8(d) (input preparation only).

Builds, for each BASELINE.json config, the inputs of one ``_neg_elcbo`` call:
training set ``X, y``, ``S`` hyper-parameter samples with their GP posterior
records (``workloads.gp_posterior``), the variational-posterior arrays, ``theta`` and
the soft bounds.  Everything is fp64 host data; nothing here touches the GPU, the
oracle or the CUDA library.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
from scipy.special import gammaln, logsumexp

from . import gp_posterior as gpp

OPTIONS = {"tol_con_loss": 0.01, "tol_weight": 1e-2, "weight_penalty": 0.1, "tol_length": 1e-6}

# name -> (D, N, K, S, total draws N_s, target)
CONFIGS = {
    "C1": dict(D=2, N=50, K=2, S=8, Ns_total=160, target="rosenbrock2"),
    "C2": dict(D=10, N=200, K=20, S=4, Ns_total=100_000, target="mvt"),
    "C3": dict(D=20, N=400, K=50, S=8, Ns_total=400_000, target="lumpy"),
    "C4": dict(D=6, N=200, K=30, S=6, Ns_total=200_000, target="rosenbrock6_noisy"),
    "C5": dict(D=20, N=400, K=50, S=32, Ns_total=3_200_000, target="lumpy"),
}


def ns_per_component(Ns_total: int, K: int) -> int:
    """BASELINE's N_s is the TOTAL number of draws; ``entmc_vbmc`` takes draws per
    component rounded up to even (entmc_vbmc.py:61; callers pass ceil(ns/K),
    variational_optimization.py:728)."""
    return 2 * int(np.ceil(np.ceil(Ns_total / K) / 2))


def _target(name, X, rng):
    N, D = X.shape
    s2 = None
    if name == "mvt":
        nu = 5.0
        sc = np.linspace(0.5, 1.5, D)
        q = np.sum((X / sc) ** 2, axis=1)
        f = (
            gammaln((nu + D) / 2)
            - gammaln(nu / 2)
            - 0.5 * D * np.log(nu * np.pi)
            - np.sum(np.log(sc))
            - 0.5 * (nu + D) * np.log1p(q / nu)
        )
    elif name == "lumpy":
        r = np.random.default_rng(7)
        M = 12
        means = r.uniform(-1, 1, size=(M, D))
        stds = 0.4 * (1 + r.uniform(size=(M, D)))
        wts = r.dirichlet(np.ones(M))
        lp = (
            -0.5 * np.sum(((X[:, None, :] - means) / stds) ** 2, axis=2)
            - np.sum(np.log(stds), axis=1)
            - 0.5 * D * np.log(2 * np.pi)
        )
        f = logsumexp(lp + np.log(wts), axis=1)
    elif name == "rosenbrock2":
        # examples/scripts/pyvbmc_example_1_full_code.py:9-31 (likelihood + N(0,3^2) prior)
        f = -np.sum((X[:, :-1] ** 2 - X[:, 1:]) ** 2 + (X[:, :-1] - 1) ** 2 / 100, axis=1)
        f = f - 0.5 * np.sum((X / 3.0) ** 2, axis=1) - D * np.log(3.0 * np.sqrt(2 * np.pi))
    elif name == "rosenbrock6_noisy":
        # examples/scripts/pyvbmc_example_6_full_code.py:29 -- per-point noise variance
        f = -np.sum((X[:, :-1] ** 2 - X[:, 1:]) ** 2 + (X[:, :-1] - 1) ** 2 / 100, axis=1)
        f = f - 0.5 * np.sum((X / 3.0) ** 2, axis=1) - D * np.log(3.0 * np.sqrt(2 * np.pi))
        s2 = 1.0 + 0.5 * np.sum(X**2, axis=1)
    else:
        raise ValueError(name)
    return f, s2


def make_problem(name="C3", ill_conditioned=False, optimize=(True,) * 4, S=None, K=None, N=None, mean_kind="negquad"):
    cfg = dict(CONFIGS[name])
    if S is not None:
        cfg["S"] = S
    if K is not None:
        cfg["K"] = K
    if N is not None:
        cfg["N"] = N
    D, N, K, S = cfg["D"], cfg["N"], cfg["K"], cfg["S"]
    rng = np.random.default_rng(0)
    if cfg["target"] == "rosenbrock2":
        X = rng.uniform(-3, 3, size=(N, D))
    else:
        X = rng.normal(size=(N, D))
    f, s2 = _target(cfg["target"], X, rng)
    noise_sd = 0.1 if s2 is None else np.sqrt(s2)
    y = f + noise_sd * rng.normal(size=N)

    hrng = np.random.default_rng(11)
    lay = gpp.hyp_layout(D, 1, mean_kind)
    hyps = np.zeros((S, lay["H"]))
    for s in range(S):
        h = hyps[s]
        h[:D] = hrng.normal(0.0, 0.2, size=D)
        if ill_conditioned:
            h[D] = 4.7 + 0.05 * hrng.normal()
            h[D + 1] = -5.4 + 0.3 * hrng.normal()
        else:
            h[D] = np.log(np.std(y)) + 0.1 * hrng.normal()
            h[D + 1] = np.log(1e-2)
        b = lay["mean_start"]
        if mean_kind != "zero":
            h[b] = np.max(y) + 0.1 * hrng.normal()
        if mean_kind == "negquad":
            h[b + 1 : b + 1 + D] = 0.1 * hrng.normal(size=D)
            h[b + 1 + D : b + 1 + 2 * D] = np.log(2.0) + 0.1 * hrng.normal(size=D)
    posts = gpp.posteriors(X, y, hyps, s2=s2, noise_N=1, mean_kind=mean_kind)
    gp = make_gp(X, posts, mean_kind=mean_kind, noise_N=1, y=y)

    vrng = np.random.default_rng(1)
    mu = 0.5 * vrng.normal(size=(D, K))
    sigma = 0.5 * np.exp(0.1 * vrng.normal(size=K))
    lambd = np.ones(D)
    eta = 0.3 * vrng.normal(size=K)
    eta = eta - eta.max()
    w = np.exp(eta)
    w = w / w.sum()
    # theta = vp.get_parameters() for lambda already normalised (mean lambda^2 = 1)
    parts = []
    if optimize[0]:
        parts.append(mu.ravel(order="F"))
    if optimize[1]:
        parts.append(np.log(sigma))
    if optimize[2]:
        parts.append(np.log(lambd))
    if optimize[3]:
        parts.append(np.log(w))
    theta = np.concatenate(parts)
    theta_bnd = soft_bounds(X, K, optimize, OPTIONS)
    return SimpleNamespace(
        name=name,
        D=D,
        N=N,
        K=K,
        S=S,
        Ns_total=cfg["Ns_total"],
        Ns_K=ns_per_component(cfg["Ns_total"], K),
        X=X,
        y=y,
        s2=s2,
        hyps=hyps,
        posts=posts,
        gp=gp,
        mu=mu,
        sigma=sigma,
        lambd=lambd,
        w=w,
        eta=eta,
        optimize=tuple(bool(o) for o in optimize),
        theta=theta,
        theta_bnd=theta_bnd,
        mean_kind=mean_kind,
    )


def make_gp(X, posts, mean_kind="negquad", noise_N=1, y=None):
    """Duck-typed stand-in for a trained ``gpyreg.GP``: exactly the attributes that
    ``_gp_log_joint`` reads (pyvbmc/vbmc/variational_optimization.py:1311,1367-1398)."""
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


def soft_bounds(X, K, optimize, options):
    """theta_bnd as ``VariationalPosterior.get_bounds`` builds it from the training inputs
    (pyvbmc/variational_posterior/variational_posterior.py:140-239) for a fresh posterior."""
    lo, hi = X.min(axis=0), X.max(axis=0)
    ln_range = np.log(hi - lo)
    lbs, ubs = [], []
    if optimize[0]:
        lbs.append(np.tile(lo, K))
        ubs.append(np.tile(hi, K))
    lbs.append(np.tile(ln_range + np.log(options["tol_length"]), K))
    ubs.append(np.tile(ln_range, K))
    out = {}
    if optimize[3]:
        eta_lb = -np.inf if options["tol_weight"] == 0 else np.log(0.5 * options["tol_weight"])
        lbs.append(np.full(K, eta_lb))
        ubs.append(np.zeros(K))
        out["weight_threshold"] = max(1 / (4 * K), options["tol_weight"])
        out["weight_penalty"] = options["weight_penalty"]
    out.update(lb=np.concatenate(lbs), ub=np.concatenate(ubs), tol_con=options["tol_con_loss"])
    return out


def draw_eps(K, Ns_K, D, seed=0):
    """Parity-mode noise in the reference's draw order (per component ``randn(Ns/2, D)``,
    entmc_vbmc.py:64-67), from a private MT19937 stream seeded like ``np.random.seed``."""
    rs = np.random.RandomState(seed)
    return np.stack([rs.randn(Ns_K // 2, D) for _ in range(K)], axis=0)
