"""Development aid: where does the host time of one drop-in _neg_elcbo call go?  (GPU box)"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv
from workloads import synthetic as syn

pr = syn.make_problem(sys.argv[1] if len(sys.argv) > 1 else "C3")
vp = pv.VariationalPosterior(pr.D, pr.K)
vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu, pr.sigma.reshape(1, -1), pr.lambd.reshape(-1, 1), pr.w.reshape(1, -1), pr.eta.reshape(1, -1)


def call():
    return pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, pr.Ns_K, True, False, pr.theta_bnd)


for _ in range(20):
    call()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    call()
dt = (time.perf_counter() - t0) / n
print(f"e2e {dt*1e6:.1f} us/eval  ({1/dt:.0f} evals/s)")
# tiny draw count: host overhead + fixed GPU latency only
t0 = time.perf_counter()
for _ in range(n):
    pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, 2, True, False, pr.theta_bnd)
dt2 = (time.perf_counter() - t0) / n
print(f"e2e with Ns=2 {dt2*1e6:.1f} us/eval")
t0 = time.perf_counter()
for _ in range(n):
    pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, 0, True, False, pr.theta_bnd)
dt3 = (time.perf_counter() - t0) / n
print(f"e2e with entlb (Ns=0) {dt3*1e6:.1f} us/eval")
prof = cProfile.Profile()
prof.enable()
for _ in range(n):
    call()
prof.disable()
st = pstats.Stats(prof)
st.sort_stats("tottime").print_stats(14)
