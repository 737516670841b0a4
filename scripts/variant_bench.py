"""Development aid: time the fp32 entmc kernel variants on config C3 and compare them with the
all-fp64 kernel on identical Philox draws.  Run on the GPU box:  python scripts/variant_bench.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv
from workloads import synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
pr = syn.make_problem(cfg)
vp = pv.VariationalPosterior(pr.D, pr.K)
vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu, pr.sigma.reshape(1, -1), pr.lambd.reshape(-1, 1), pr.w.reshape(1, -1), pr.eta.reshape(1, -1)
names = {0: "fast(expanded)", 1: "dsplit", 2: "packed", 3: "scalar", 4: "warp-autonomous", 5: "tensor-core"}
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else (3, 2, 1, 0, 4, 5)
ref = None
for variant in variants:
    os.environ["VBMC_ENTMC_VARIANT"] = str(variant)
    ctx = pv.Context(0)
    if ref is None:
        ref = ctx.entmc(vp, pr.Ns_K, (True,) * 4, True, seed=7, precision="f64")
    H, dH = ctx.entmc(vp, pr.Ns_K, (True,) * 4, True, seed=7)
    ctx.set_kernel_timing(True)
    for _ in range(3):
        ctx.entmc(vp, pr.Ns_K, (True,) * 4, True, seed=7)
    ctx.entmc_kernel_ms()
    for _ in range(20):
        ctx.entmc(vp, pr.Ns_K, (True,) * 4, True, seed=7)
    ms, n = ctx.entmc_kernel_ms()
    relH = abs(H - ref[0]) / abs(ref[0])
    relg = np.abs(dH - ref[1]).max() / np.abs(ref[1]).max()
    print(f"{cfg} variant {variant} {names[variant]:15s} kernel {ms*1e3:8.1f} us  relH {relH:.2e}  rel_dH {relg:.2e}")
    ctx.close()
