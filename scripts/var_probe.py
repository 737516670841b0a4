"""Development aid (GPU box): the variance path (_eval_full_elcbo's call) end to end, C3 and C5-sized S."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyvbmc_b200 as pv
from workloads import synthetic as syn
for cfg, S in (("C3", 8), ("C3", 32), ("C2", 4)):
    pr = syn.make_problem(cfg, S=S)
    vp = pv.VariationalPosterior(pr.D, pr.K)
    vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy(), pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()
    call = lambda: pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, 0, False, True, None, 0.0, True)
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    for _ in range(20):
        r = call()
    print(cfg, "S", S, "variance path us/call %.1f" % (1e6 * (time.perf_counter() - t0) / 20), "varF", r[4])
