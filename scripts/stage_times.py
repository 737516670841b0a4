"""Development aid: per-stage device timeline of one synchronous _neg_elcbo evaluation (GPU box)."""
import os
import sys

os.environ["VBMC_STAGE_TIMING"] = "1"
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv
from workloads import synthetic as syn

for cfg in sys.argv[1:] or ["C3", "C2"]:
    pr = syn.make_problem(cfg)
    vp = pv.VariationalPosterior(pr.D, pr.K)
    vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu, pr.sigma.reshape(1, -1), pr.lambd.reshape(-1, 1), pr.w.reshape(1, -1), pr.eta.reshape(1, -1)
    ctx = pv.context_for_gp(pr.gp)
    for Ns in (pr.Ns_K, 2, 0):
        acc = None
        for it in range(60):
            pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, Ns, True, False, pr.theta_bnd)
            if it >= 10:
                t = ctx.stage_times()
                acc = t if acc is None else {k: acc[k] + v for k, v in t.items()}
        print(cfg, "Ns_K=%d" % Ns, {k: round(v / 50, 1) for k, v in acc.items()})
