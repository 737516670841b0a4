"""Development aid (GPU box): device timeline (CUPTI, through torch.profiler) of one C3 evaluation on each path:
host-buffer call, device-resident enqueue, device Adam iteration.  Prints kernel start / end relative to the first
kernel of the window, in time order.  Tracing adds host overhead: read the STRUCTURE (gaps, overlaps), not the totals."""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import pyvbmc_b200 as pv
from workloads import synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
pr = syn.make_problem(cfg)
Ns = int(sys.argv[2]) if len(sys.argv) > 2 else pr.Ns_K
vp = pv.VariationalPosterior(pr.D, pr.K)
vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy(), pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()
theta = pr.theta.copy()
call = lambda: pv._neg_elcbo(theta, pr.gp, vp, 0.0, Ns, True, False, pr.theta_bnd)
for _ in range(30):
    call()
ctx = pv.context_for_gp(pr.gp)


def show(tag, fn, n_show):
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    ev = ev[-n_show:]
    t0 = ev[0].time_range.start
    print("==", tag)
    for e in ev:
        m = re.search(r"(\w+)\s*(<[^(]*)?\(", e.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
        name = (m.group(1) if m else e.name)[:38]
        print("  %-38s start %8.1f  end %8.1f  dur %6.1f" % (name, e.time_range.start - t0, e.time_range.end - t0, e.time_range.end - e.time_range.start))


def host_calls():
    for _ in range(6):
        call()


def enq():
    for _ in range(6):
        ctx.enqueue()
    ctx.synchronize()


show("host-buffer call (last 2 of 6)", host_calls, 14)
call()
show("device-resident enqueue (last 2 of 6)", enq, 10)
vp2 = pv.VariationalPosterior(pr.D, pr.K)
vp2.mu, vp2.sigma, vp2.lambd, vp2.w, vp2.eta = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy(), pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()


def adam():
    pv.minimize_adam_elcbo(pr.gp, vp2, pr.theta.copy(), Ns, theta_bnd=pr.theta_bnd, max_iter=20, use_early_stopping=False, seed=3)


adam()
show("device Adam (last 4 iterations of 20)", adam, 27)
