"""Development aid (GPU box): where the host-buffer (e2e) call of C3 spends its time.  Wall clock per piece of
pyvbmc_b200._neg_elcbo: the whole call, the C-ABI calls inside it, and (by difference) the NumPy / Python side."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyvbmc_b200 as pv
from pyvbmc_b200 import context as ctxmod
from workloads import synthetic as syn

pv.config.host_noise_prefetch = os.environ.get("VBMC_HOST_PREFETCH", "0") == "1"
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
pr = syn.make_problem(cfg)
Ns = int(sys.argv[2]) if len(sys.argv) > 2 else pr.Ns_K
vp = pv.VariationalPosterior(pr.D, pr.K)
vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy(), pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()
theta = pr.theta.copy()
call = lambda: pv._neg_elcbo(theta, pr.gp, vp, 0.0, Ns, True, False, pr.theta_bnd)
for _ in range(50):
    call()
ctx = pv.context_for_gp(pr.gp)
acc = {}


def wrap(obj, name):
    f = getattr(obj, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return r

    setattr(obj, name, g)


n = 2000
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    print(cfg, Ns, "whole call %.1f us" % (1e6 * (time.perf_counter() - t0) / n))
for name in ("noise_prefetch", "set_bounds", "theta_buffers", "negelcbo_theta"):
    wrap(ctx, name)
t0 = time.perf_counter()
for _ in range(n):
    call()
tot = time.perf_counter() - t0
print("instrumented whole call %.1f us" % (1e6 * tot / n))
for k, v in acc.items():
    print("  %-16s %.1f us" % (k, 1e6 * v / n))
print("  %-16s %.1f us" % ("python/numpy rest", 1e6 * (tot - sum(acc.values())) / n))
