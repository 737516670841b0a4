"""Re-export of :mod:`workloads.gp_posterior` (the gpyreg-posterior restatement that the
oracle tests pin against the reference's MATLAB fixtures) -- TEST INFRASTRUCTURE ONLY."""
from workloads.gp_posterior import *  # noqa: F401,F403
from workloads.gp_posterior import hyp_layout, mean_fn, posterior, posteriors, se_ard  # noqa: F401
