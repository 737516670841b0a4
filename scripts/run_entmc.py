"""Development aid (GPU box): launch the default entmc kernel on a config a few times -- the command
ncu wraps.  python scripts/run_entmc.py [C3] [n_launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv
from workloads import synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
pr = syn.make_problem(cfg)
vp = pv.VariationalPosterior(pr.D, pr.K)
vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu, pr.sigma.reshape(1, -1), pr.lambd.reshape(-1, 1), pr.w.reshape(1, -1), pr.eta.reshape(1, -1)
ctx = pv.Context(0)
for i in range(n):
    H, dH = ctx.entmc(vp, pr.Ns_K, (True,) * 4, True, seed=7 + i)
print(cfg, "H", H)
ctx.close()
