"""Development aid (GPU box): which fp32 entmc kernel variant is fastest where?  For a grid of (D, K) at ~400k total
draws: warp-autonomous (4) and expanded (0) CUDA-core kernels against the tensor-core kernel (5; forced), timed with
CUDA events around the launch(es) of one stand-alone entropy evaluation (variant 5: table + noise generator + main
kernel, and the main kernel alone), and checked against the all-fp64 kernel on identical Philox draws."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv

names = {0: "expanded", 4: "warp-autonomous", 5: "tensor-core"}
total = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
for D, K in ((20, 16), (20, 24), (20, 32), (20, 40), (20, 48), (20, 50), (20, 64), (10, 20), (10, 32), (10, 48), (10, 64),
             (6, 30), (6, 48), (32, 50)):
    rng = np.random.default_rng(K * 100 + D)
    vp = pv.VariationalPosterior(D, K)
    vp.mu = 0.5 * rng.normal(size=(D, K))
    vp.sigma = (0.5 * np.exp(0.1 * rng.normal(size=K))).reshape(1, -1)
    vp.lambd = np.ones((D, 1))
    eta = 0.3 * rng.normal(size=K)
    vp.eta = (eta - eta.max()).reshape(1, -1)
    vp.w = (np.exp(vp.eta) / np.exp(vp.eta).sum()).reshape(1, -1)
    Ns_K = 2 * int(np.ceil(total / K / 2))
    ref = None
    line = f"D={D:2d} K={K:2d} Ns_K={Ns_K:6d}:"
    for variant in (0, 4, 5):
        os.environ["VBMC_ENTMC_VARIANT"] = str(variant)
        ctx = pv.Context(0)
        try:
            if ref is None:
                ref = ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=7, precision="f64")
            H, dH = ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=7)
            used = ctx.entmc_variant_used()
            ctx.set_kernel_timing(True)
            for _ in range(3):
                ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=7)
            ctx.entmc_kernel_ms()
            for _ in range(15):
                ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=7)
            main_ms = ctx.entmc_main_kernel_ms()
            ms, n = ctx.entmc_kernel_ms()
            relg = np.abs(dH - ref[1]).max() / np.abs(ref[1]).max()
            extra = f" (main {main_ms*1e3:5.1f})" if used == 5 else ""
            line += f"  v{variant}->{used} {ms*1e3:6.1f} us{extra} rel_dH {relg:.1e};"
        except Exception as exc:
            line += f"  v{variant} {type(exc).__name__};"
        finally:
            ctx.close()
    print(line, flush=True)
