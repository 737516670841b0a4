import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1]
if "torch" in mode:
    import torch
    torch.cuda.set_device(0)
import pyvbmc_b200 as pv
from workloads import synthetic as syn
pr = syn.make_problem("C3", S=8)
vp = pv.VariationalPosterior(pr.D, pr.K)
vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy(), pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()
if "ev" in mode:
    from pyvbmc_b200.distributed import ShardedNegElcbo
    ev = ShardedNegElcbo(pr.gp, device=0, seed=1234)
    ev(pr.theta, vp, pr.Ns_K, pr.theta_bnd)
if "flush" in mode:
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
def step():
    return pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, pr.Ns_K, True, False, pr.theta_bnd)
for _ in range(5):
    step()
ts = []
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(50):
        step()
    ts.append((time.perf_counter() - t0) / 50 * 1e6)
print(mode, ["%.1f" % t for t in ts])
