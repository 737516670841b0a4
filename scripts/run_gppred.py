"""Development aid (GPU box): a few GP-prediction launches at the search-cache size -- the command ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyvbmc_b200 as pv
from workloads import synthetic as syn
pr = syn.make_problem("C3")
rng = np.random.default_rng(0)
Xs = pr.X[rng.integers(0, pr.N, size=8192)] + 0.5 * rng.normal(size=(8192, pr.D))
for _ in range(4):
    f_mu, f_s2 = pv.gp_predict(pr.gp, Xs, separate_samples=True)
print("gp_predict", f_mu.shape, float(f_s2.mean()))
