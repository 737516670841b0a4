"""Generates tests/golden/ref_adam.npz from the UNMODIFIED reference minimize_adam (needs /root/reference;
run in the build container:  python -m oracle.make_golden_adam)."""
import os

import numpy as np

from . import ref_loader
from .minimize_adam_oracle import noisy_quadratic


def main():
    ref_loader.load()
    from pyvbmc.vbmc.minimize_adam import minimize_adam as ref_adam

    out = {}
    for name, kw in (("default", {}), ("box", {"lb": np.full(6, -0.5), "ub": np.full(6, 0.7), "max_iter": 130}),
                     ("noearly", {"use_early_stopping": False, "max_iter": 75, "master_max": 0.05})):
        f, x0 = noisy_quadratic()
        x, y, x_tab, y_tab, n = ref_adam(f, x0.copy(), **kw)
        out[name + "_x"], out[name + "_y"], out[name + "_xtab"], out[name + "_ytab"], out[name + "_n"] = x, y, x_tab, y_tab, n
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_adam.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
