"""Development aid (GPU box): 200 Adam iterations on the ELBO objective at a config -- device-resident loop
(pyvbmc_b200.minimize_adam_elcbo) vs the reference-shaped host loop over the drop-in _neg_elcbo."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv
from oracle.minimize_adam_oracle import minimize_adam as adam_host
from workloads import synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
pr = syn.make_problem(cfg)
def fresh():
    vp = pv.VariationalPosterior(pr.D, pr.K)
    vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy(), pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()
    return vp
for Ns_K in (pr.Ns_K, 28):
    kw = dict(max_iter=iters, use_early_stopping=False, master_max=0.01)
    vp = fresh()
    th0 = np.asarray(vp.get_parameters(), dtype=float)
    pv.minimize_adam_elcbo(pr.gp, fresh(), th0, Ns_K, pr.theta_bnd, seed=1, max_iter=40, use_early_stopping=False)  # warm-up
    t0 = time.perf_counter()
    x, y, xt, yt, n = pv.minimize_adam_elcbo(pr.gp, vp, th0, Ns_K, pr.theta_bnd, seed=1, **kw)
    t_dev = time.perf_counter() - t0
    vp2 = fresh()
    k = {"i": 0}
    def f(t):
        F, dF, *_ = pv._neg_elcbo(t, pr.gp, vp2, 0.0, Ns_K, True, False, pr.theta_bnd, seed=1, offset=k["i"])
        k["i"] += 1
        return F, dF
    adam_host(f, th0.copy(), max_iter=40, use_early_stopping=False)
    k["i"] = 0
    t0 = time.perf_counter()
    xh, yh, xth, yth, nh = adam_host(f, th0.copy(), **kw)
    t_host = time.perf_counter() - t0
    print(f"{cfg} Ns_K={Ns_K}: device loop {t_dev/n*1e6:.1f} us/iteration ({n/t_dev:.0f} it/s), host loop over the drop-in "
          f"{t_host/nh*1e6:.1f} us/iteration ({nh/t_host:.0f} it/s); max |x_dev - x_host| {np.max(np.abs(xt - xth)):.2e}, y {yt[0]:.6f} -> {yt[-1]:.6f}")
