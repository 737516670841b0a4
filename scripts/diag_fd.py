"""Development aid (GPU box): finite-difference diagnostics of _neg_elcbo with frozen Monte-Carlo noise."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pyvbmc_b200 as pv
from golden_util import load_case
from test_gpu_parity import case_vp

c = load_case("c2")
theta0 = c.g["theta2"].copy()
Ns = 400
for prec in ("f64", "f32"):
    pv.config.precision = prec
    for graphs in ("1", "0"):
        os.environ["VBMC_GRAPH"] = graphs
        pv.clear_caches()
        Fg, dF, *_ = pv._neg_elcbo(theta0, c.gp, case_vp(pv, c), 0.0, Ns, True, False, c.theta_bnd, seed=21)
        Fv = [pv._neg_elcbo(theta0, c.gp, case_vp(pv, c), 0.0, Ns, False, False, c.theta_bnd, seed=21)[0] for _ in range(4)]
        Fg2 = [pv._neg_elcbo(theta0, c.gp, case_vp(pv, c), 0.0, Ns, True, False, c.theta_bnd, seed=21)[0] for _ in range(4)]
        print(prec, "graphs", graphs, "F(grad)", Fg, "F(value-only) x4", Fv, "F(grad) x4", Fg2)
        for i in (191, 3, 205, 225, 240):
            h = 1e-5 * max(1.0, abs(theta0[i]))
            tp, tm = theta0.copy(), theta0.copy()
            tp[i] += h
            tm[i] -= h
            fv = [pv._neg_elcbo(t, c.gp, case_vp(pv, c), 0.0, Ns, False, False, c.theta_bnd, seed=21)[0] for t in (tp, tm)]
            fg = [pv._neg_elcbo(t, c.gp, case_vp(pv, c), 0.0, Ns, True, False, c.theta_bnd, seed=21)[0] for t in (tp, tm)]
            print("   i", i, "dF", dF[i], "fd(value-only)", (fv[0] - fv[1]) / (2 * h), "fd(grad path)", (fg[0] - fg[1]) / (2 * h))
