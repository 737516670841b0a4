"""Generates tests/golden/ref_adam.npz from the UNMODIFIED reference minimize_adam (needs /root/reference;
run in the build container:  python -m oracle.make_golden_adam)."""
import os

import numpy as np

from . import ref_loader
from .minimize_adam_oracle import noisy_quadratic


def main():
    ref_loader.load()
    from pyvbmc.vbmc.minimize_adam import minimize_adam as ref_adam

    out = {}
    for name, kw in (("default", {}), ("box", {"lb": np.full(6, -0.5), "ub": np.full(6, 0.7), "max_iter": 130}),
                     ("noearly", {"use_early_stopping": False, "max_iter": 75, "master_max": 0.05})):
        f, x0 = noisy_quadratic()
        x, y, x_tab, y_tab, n = ref_adam(f, x0.copy(), **kw)
        out[name + "_x"], out[name + "_y"], out[name + "_xtab"], out[name + "_ytab"], out[name + "_n"] = x, y, x_tab, y_tab, n
    # the real objective: the closure of variational_optimization.py:238-249 around the unmodified _neg_elcbo.  The
    # reference shifts the eta block of the iterate in place on every call (:1082-1085), so x_tab's eta rows are
    # renormalised -- a quadratic objective cannot pin that.
    from . import synthetic as syn

    ref = ref_loader.load()
    pr = syn.make_problem("C2", N=60)
    opt = (True, True, True, True)
    rgp = ref_loader.make_ref_gp(pr.X, pr.y, pr.posts, pr.mean_kind)
    vp = ref_loader.make_ref_vp(pr.D, pr.K, pr.vp.mu, pr.vp.sigma, pr.vp.lambd, pr.vp.w, pr.vp.eta, opt)
    theta0 = pr.theta.copy()
    theta0[-pr.K:] += 6.0
    Ns_K = 10

    def f(theta_):
        r = ref._neg_elcbo(theta_, rgp, vp, 0.0, Ns_K, compute_grad=True, compute_var=False, theta_bnd=pr.theta_bnd)
        return r[0], r[1]

    np.random.seed(11)
    kw = dict(max_iter=60, master_max=0.05, use_early_stopping=True)
    x, y, x_tab, y_tab, n = ref_adam(f, theta0.copy(), **kw)
    out.update(elbo_theta0=theta0, elbo_x=x, elbo_y=y, elbo_xtab=x_tab, elbo_ytab=y_tab, elbo_n=n, elbo_Ns_K=Ns_K,
               elbo_seed=11, elbo_N=60)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_adam.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
