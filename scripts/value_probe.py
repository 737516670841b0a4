"""Development aid (GPU box): device-resident step time of C3 through vbmc_negelcbo_enqueue, back to back, no L2 flush."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyvbmc_b200 as pv
from workloads import synthetic as syn
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
pr = syn.make_problem(cfg)
Ns = int(sys.argv[2]) if len(sys.argv) > 2 else pr.Ns_K
vp = pv.VariationalPosterior(pr.D, pr.K)
vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy(), pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()
for _ in range(5):
    pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, Ns, True, False, pr.theta_bnd)
ctx = pv.context_for_gp(pr.gp)
st = torch.cuda.ExternalStream(ctx.stream)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(4):
    for _ in range(200):
        ctx.enqueue()
    ctx.synchronize()
    a.record(st)
    n = 500
    l0 = ctx.launch_count
    for _ in range(n):
        ctx.enqueue()
    b.record(st)
    b.synchronize()
    print(cfg, Ns, "device us/eval %.2f" % (1e3 * a.elapsed_time(b) / n), "launches/eval", (ctx.launch_count - l0) / n)
