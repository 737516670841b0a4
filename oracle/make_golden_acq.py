"""Generates tests/golden/ref_acq.npz from the UNMODIFIED reference (needs /root/reference or oracle/_ref; run in the
build container:  python -m oracle.make_golden_acq) -- TEST INFRASTRUCTURE ONLY.

What the unmodified reference produces here:
  * ``VariationalPosterior.pdf(x, orig_flag=False, log_flag, grad_flag)`` (variational_posterior.py:241-552);
  * ``AcqFcnLog.__call__`` (abstract_acq_fcn.py:34-147 + acq_fcn_log.py:21-50) on a duck-typed GP whose ``predict`` is
    the oracle's restatement of gpyreg's (gpyreg itself is absent): this pins the aggregation over hyper-samples, the
    acquisition formula, the variance regularisation and the hard-bound masking -- everything around ``gp.predict``.
"""
import os
from types import SimpleNamespace

import numpy as np

from . import acq_oracle as ao
from . import ref_loader
from . import synthetic as syn


def main():
    ref = ref_loader.load()
    from pyvbmc.acquisition_functions import AcqFcnLog

    out = {}
    rng = np.random.default_rng(23)
    for tag, cfg, kw, Nx in (("c2", "C2", dict(N=64), 96), ("c4", "C4", dict(N=60), 80), ("c1", "C1", {}, 50)):
        pr = syn.make_problem(cfg, **kw)
        D, K = pr.D, pr.K
        vp = ref_loader.make_ref_vp(D, K, pr.vp.mu, pr.vp.sigma, pr.vp.lambd, pr.vp.w, pr.vp.eta)
        # search points: half around the mixture centres, half around the training inputs, a few far away
        comp = rng.integers(0, K, size=Nx)
        Xs = pr.vp.mu[:, comp].T + 1.5 * np.ravel(pr.vp.sigma)[comp][:, None] * np.ravel(pr.vp.lambd) * rng.normal(size=(Nx, D))
        Xs[Nx // 2 :] = pr.X[rng.integers(0, pr.N, size=Nx - Nx // 2)] + 0.05 * rng.normal(size=(Nx - Nx // 2, D))
        Xs[:3] += 40.0  # pdf underflows to 0: log pdf = -inf, clamped to log(realmin) by the acquisition function
        out[f"{tag}_Xs"] = Xs
        y = vp.pdf(Xs, orig_flag=False)
        ly, dly = vp.pdf(Xs, orig_flag=False, log_flag=True, grad_flag=True)
        _, dyy = vp.pdf(Xs, orig_flag=False, log_flag=False, grad_flag=True)
        out.update({f"{tag}_pdf": y, f"{tag}_logpdf": ly, f"{tag}_dlogpdf": dly, f"{tag}_dpdf": dyy})

        f_mu, f_s2 = ao.gp_predict(pr.X, pr.posts, Xs, pr.mean_kind, separate_samples=True)
        gp = SimpleNamespace(predict=lambda x_star, separate_samples=True, f=(f_mu, f_s2): f, D=D)
        flog = SimpleNamespace(y_max=float(np.max(pr.y)))
        optim_state = {"integer_vars": None, "variance_regularized_acq_fcn": True, "tol_gp_var": 1e-4,
                       "lb_eps_orig": np.full((1, D), -30.0), "ub_eps_orig": np.full((1, D), 30.0)}
        acq = AcqFcnLog()(Xs.copy(), gp, vp, flog, optim_state)
        out.update({f"{tag}_f_mu": f_mu, f"{tag}_f_s2": f_s2, f"{tag}_acq": acq, f"{tag}_y_max": flog.y_max,
                    f"{tag}_N": pr.N})
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_acq.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
