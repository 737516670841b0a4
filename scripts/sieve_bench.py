"""Development aid (GPU box): the sieve loop of variational_optimization.py:775-787 -- B candidates, value-only,
deterministic entropy -- one at a time through the drop-in _neg_elcbo vs one batched launch vs the oracle port."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv
from oracle import elbo_oracle as eo
from workloads import synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
pr = syn.make_problem(cfg)
rng = np.random.default_rng(0)
vps, thetas = [], []
for b in range(B):
    vp = pv.VariationalPosterior(pr.D, pr.K)
    vp.mu = pr.mu + 0.3 * rng.normal(size=pr.mu.shape)
    vp.sigma = pr.sigma.reshape(1, -1) * np.exp(0.2 * rng.normal(size=(1, pr.K)))
    vp.lambd = pr.lambd.reshape(-1, 1).copy()
    vp.w = pr.w.reshape(1, -1).copy()
    vp.eta = pr.eta.reshape(1, -1).copy()
    vps.append(vp)
pv.neg_elcbo_batch(vps[:4], pr.gp, pr.theta_bnd)  # warm-up (context, GP upload)
t0 = time.perf_counter()
F, G, H = pv.neg_elcbo_batch(vps, pr.gp, pr.theta_bnd)
t_batch = time.perf_counter() - t0
ctx = pv.context_for_gp(pr.gp)
prm = np.zeros((B, ctx.param_len(pr.D, pr.K)))
t0 = time.perf_counter()
F2, G2, H2 = pv.neg_elcbo_batch(vps, pr.gp, pr.theta_bnd)
t_batch2 = time.perf_counter() - t0
n1 = min(B, 200)
t0 = time.perf_counter()
F1 = np.array([pv._neg_elcbo(vps[b].get_parameters(), pr.gp, vps[b], 0.0, 0, False, False, pr.theta_bnd)[0] for b in range(n1)])
t_single = (time.perf_counter() - t0) / n1
no = min(B, 3)
t0 = time.perf_counter()
for b in range(no):
    vo = eo.OracleVP.create(pr.D, pr.K, vps[b].mu, vps[b].sigma, vps[b].lambd, vps[b].w, vps[b].eta, (True,) * 4)
    Fo = eo.neg_elcbo(eo.get_parameters(vo), pr.gp, vo, 0.0, 0, False, False, pr.theta_bnd)[0]
t_oracle = (time.perf_counter() - t0) / no
print(f"{cfg} B={B}: batched {t_batch*1e3:.1f} ms ({t_batch/B*1e6:.1f} us/candidate; 2nd call {t_batch2*1e3:.1f} ms), "
      f"one at a time {t_single*1e6:.1f} us/candidate, oracle port {t_oracle*1e3:.1f} ms/candidate; "
      f"max |F_batch - F_single| / |F| = {np.max(np.abs(F[:n1] - F1) / np.abs(F1)):.2e}, oracle relerr {abs(F[no-1]-Fo)/abs(Fo):.2e}")
