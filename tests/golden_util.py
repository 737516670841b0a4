"""Load ``tests/golden/*.npz`` into oracle-side objects (shared by CPU and GPU tests)."""
import os
from types import SimpleNamespace

import numpy as np

from oracle import elbo_oracle as eo
from oracle import gp_posterior as gpp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

REF_CASES = [
    "c1",
    "c2",
    "c2_noweights",
    "c2_ill",
    "c2_const",
    "c2_zero",
    "c2_var",
    "c2_var_s1",
    "c2_lownoise",
    "c3",
    "c4",
    "c4_var",
]
VAR_CASES = ["c1", "c2_var", "c2_var_s1", "c2_lownoise", "c4_var"]


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def relmax(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _scalar(a):
    return float(np.asarray(a, dtype=float).reshape(-1)[0])


def relerr(a, b):
    a, b = _scalar(a), _scalar(b)
    return abs(a - b) / max(abs(b), 1e-300)


def load_case(stem):
    """-> namespace(g=npz, gp=oracle gp, vp()=fresh OracleVP factory, theta_bnd=dict, opt=tuple)."""
    g = load_npz("ref_" + stem)
    D, K, S, N = int(g["D"]), int(g["K"]), int(g["S"]), int(g["N"])
    mean_kind = str(g["mean_kind"])
    posts = []
    for s in range(S):
        L = g["L"][s] if "L" in g.files else None
        posts.append(
            dict(hyp=g["hyps"][s], alpha=g["alpha"][s], L=L, L_chol=bool(g["L_chol"][s]), sW=np.full(N, g["sW0"][s]))
        )
    gp = eo.make_gp(g["X"], posts, mean_kind=mean_kind, noise_N=1, y=g["y"])
    opt = tuple(bool(o) for o in g["optimize"])

    def fresh_vp():
        return eo.OracleVP.create(D, K, g["vp_mu"], g["vp_sigma"], g["vp_lambd"], g["vp_w"], g["vp_eta"], opt)

    def standalone_vp():
        return eo.OracleVP.create(D, K, g["sa_mu"], g["sa_sigma"], g["sa_lambd"], g["sa_w"], g["sa_eta"], opt)

    bnd = {"lb": g["bnd_lb"], "ub": g["bnd_ub"], "tol_con": float(g["bnd_tol_con"])}
    if opt[3]:
        bnd["weight_threshold"] = float(g["bnd_weight_threshold"])
        bnd["weight_penalty"] = float(g["bnd_weight_penalty"])
    return SimpleNamespace(
        g=g, D=D, K=K, S=S, N=N, Ns_K=int(g["Ns_K"]), gp=gp, vp=fresh_vp, sa_vp=standalone_vp, theta_bnd=bnd, opt=opt,
        mean_kind=mean_kind, posts=posts,
    )


def eps_for(seed, K, Ns_K, D):
    """The reference's draws after ``np.random.seed(seed)`` (entmc_vbmc.py:64-67)."""
    rs = np.random.RandomState(int(seed))
    return np.stack([rs.randn(Ns_K // 2, D) for _ in range(K)], axis=0)
