"""Seeded synthetic inputs for tests and benchmarks (host-side input preparation only).

``gp_posterior``  restates what the third-party ``gpyreg`` package produces for a trained GP
                  (``alpha, L, L_chol, sW``) -- the hot path consumes these as opaque arrays.
``synthetic``     the workloads of SURVEY.md section 8(d) / BASELINE.json ``configs``.

Nothing here is on the product's compute path and nothing here imports ``oracle`` or the
CUDA library.
"""
