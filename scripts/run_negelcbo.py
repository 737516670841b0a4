"""Development aid (GPU box): a few eager (graph off) _neg_elcbo evaluations -- the command ncu wraps to
capture gplj / entmc / tail kernels.  VBMC_GRAPH=0 python scripts/run_negelcbo.py [C3] [n]"""
import os
import sys

os.environ.setdefault("VBMC_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv
from workloads import synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
pr = syn.make_problem(cfg)
vp = pv.VariationalPosterior(pr.D, pr.K)
vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu, pr.sigma.reshape(1, -1), pr.lambd.reshape(-1, 1), pr.w.reshape(1, -1), pr.eta.reshape(1, -1)
for i in range(n):
    F, dF, *_ = pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, pr.Ns_K, True, False, pr.theta_bnd)
print(cfg, "F", F)
