"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the PyVBMC ELBO inner loop.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker (or as the CPU arm being
timed), never as the implementation of :mod:`pyvbmc_b200`.

Contents
--------
``elbo_oracle``   fp64 NumPy restatement of the reference's hot path
                  (``entmc_vbmc``, ``entlb_vbmc``, ``_gp_log_joint``, ``_neg_elcbo``,
                  soft-bound loss, theta packing, ``get_bounds``), with the
                  Monte-Carlo noise ``eps`` as an explicit argument.
``gp_posterior``  fp64 restatement of the GP posterior record the path consumes
                  (``alpha, L, L_chol, sW``) -- the arithmetic lives in the
                  un-vendored third-party package ``gpyreg`` (pyproject floor
                  ``gpyreg >= 0.1.0``, absent from this image).
``ref_loader``    imports the UNMODIFIED reference from ``/root/reference`` behind
                  ``sys.modules`` stubs for its missing third-party imports.  Only
                  usable in the build container (the GPU box has no reference tree).
``make_golden``   regenerates ``tests/golden/*.npz`` from ``ref_loader``.
``synthetic``     the seeded synthetic workloads of SURVEY.md section 8(d).

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks the restatement
against the reference's own MATLAB known-answer fixtures (re-exported into
``tests/golden/matlab_*.npz``) and against outputs of the unmodified reference
run in the build container (``tests/golden/ref_*.npz``).
"""
