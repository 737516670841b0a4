"""Development aid (GPU box): a short run through the paths touched last (small-draw-count entropy kernel, root-forked
generator + 5-CTA parameter kernel, split-phase device Adam) for compute-sanitizer:
  compute-sanitizer --tool memcheck python scripts/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyvbmc_b200 as pv
from workloads import synthetic as syn

for cfg, Ns in (("C2", 28), ("C2", 5000), ("C3", 28), ("C3", 8000)):
    pr = syn.make_problem(cfg)
    vp = pv.VariationalPosterior(pr.D, pr.K)
    vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy(), pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()
    for it in range(4):  # eager, capture, replay, replay
        F, dF, *_ = pv._neg_elcbo(pr.theta.copy(), pr.gp, vp, 0.0, Ns, True, False, pr.theta_bnd, seed=5 + it)
    x, y, xt, yt, n = pv.minimize_adam_elcbo(pr.gp, vp, pr.theta.copy(), Ns, pr.theta_bnd, seed=3, max_iter=46, use_early_stopping=False)
    print(cfg, Ns, "F", F, "adam y", yt[0], "->", yt[-1], "variant", pv.context_for_gp(pr.gp).entmc_variant_used(), flush=True)
pv.clear_caches()
print("SANITIZE_RUN_DONE")
