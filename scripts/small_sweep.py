"""Development aid (GPU box): the fp32 entropy kernels at SMALL draw counts (the reference's defaults are tens of draws per
component): one thread per pair (variant 0) against eight lanes per pair (variant 6), CUDA events around the launch of one
stand-alone entropy evaluation, checked against the all-fp64 kernel on identical Philox draws."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv

for D, K in ((20, 50), (10, 20), (6, 30), (2, 4), (32, 64)):
    rng = np.random.default_rng(K * 100 + D)
    vp = pv.VariationalPosterior(D, K)
    vp.mu = 0.5 * rng.normal(size=(D, K))
    vp.sigma = (0.5 * np.exp(0.1 * rng.normal(size=K))).reshape(1, -1)
    vp.lambd = np.ones((D, 1))
    eta = 0.3 * rng.normal(size=K)
    vp.eta = (eta - eta.max()).reshape(1, -1)
    vp.w = (np.exp(vp.eta) / np.exp(vp.eta).sum()).reshape(1, -1)
    for Ns_K in (28, 64, 128, 256, 512, 1024):
        ref = None
        line = f"D={D:2d} K={K:2d} Ns_K={Ns_K:5d}:"
        for variant in (0, 6):
            os.environ["VBMC_ENTMC_VARIANT"] = str(variant)
            ctx = pv.Context(0)
            try:
                if ref is None:
                    ref = ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=7, precision="f64")
                H, dH = ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=7)
                ctx.set_kernel_timing(True)
                for _ in range(3):
                    ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=7)
                ctx.entmc_kernel_ms()
                for _ in range(15):
                    ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=7)
                ms, n = ctx.entmc_kernel_ms()
                relg = np.abs(dH - ref[1]).max() / np.abs(ref[1]).max()
                line += f"  v{variant}->{ctx.entmc_variant_used()} {ms*1e3:6.1f} us relH {abs(H-ref[0])/max(abs(ref[0]),1):.1e} rel_dH {relg:.1e};"
            except Exception as exc:
                line += f"  v{variant} {type(exc).__name__}: {exc};"
            finally:
                ctx.close()
        print(line, flush=True)
