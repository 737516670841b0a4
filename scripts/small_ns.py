"""Development aid (GPU box): drop-in _neg_elcbo latency vs draw count, graph replay on/off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv
from workloads import synthetic as syn
pr = syn.make_problem(sys.argv[1] if len(sys.argv) > 1 else "C3")
vp = pv.VariationalPosterior(pr.D, pr.K)
vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu, pr.sigma.reshape(1, -1), pr.lambd.reshape(-1, 1), pr.w.reshape(1, -1), pr.eta.reshape(1, -1)
for Ns in (0, 2, 28, 200, 2000, pr.Ns_K):
    for _ in range(20):
        pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, Ns, True, False, pr.theta_bnd)
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, Ns, True, False, pr.theta_bnd)
    print(f"graph={os.environ.get('VBMC_GRAPH','1')} Ns_K={Ns:6d}: {(time.perf_counter()-t0)/n*1e6:7.1f} us/eval")
