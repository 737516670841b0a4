"""Oracle-side view of :mod:`workloads.synthetic` -- TEST INFRASTRUCTURE ONLY.
Adds an :class:`oracle.elbo_oracle.OracleVP` to the plain-array problem description."""
from workloads.synthetic import CONFIGS, OPTIONS, draw_eps, ns_per_component  # noqa: F401
from workloads.synthetic import make_problem as _make_problem

from . import elbo_oracle as eo


def make_problem(*args, **kwargs):
    pr = _make_problem(*args, **kwargs)
    pr.vp = eo.OracleVP.create(pr.D, pr.K, pr.mu, pr.sigma, pr.lambd, pr.w, pr.eta, pr.optimize)
    theta = eo.get_parameters(pr.vp)
    assert abs(theta - pr.theta).max() < 1e-12
    bnd = eo.get_bounds(pr.vp, pr.X, OPTIONS, pr.K)
    assert all((bnd[k] == pr.theta_bnd[k]).all() if hasattr(bnd[k], "shape") else bnd[k] == pr.theta_bnd[k] for k in bnd)
    return pr
