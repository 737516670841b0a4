"""The drop-in claim on hardware: the UNMODIFIED reference ``optimize_vp`` (variational_optimization.py:90-391: sieve ->
Adam / BFGS on the best candidates -> full ELCBO -> pruning) running on top of ``pyvbmc_b200.install()``, against the
same call on the unpatched reference (CPU), on the problems of the reference's own end-to-end tests
(pyvbmc/testing/vbmc/test_variational_optimization.py:244-413) with the reference's own tolerances.

The reference package comes from ``oracle/_ref`` on the GPU box (archive packed by ``oracle/build_ref.py``); gpyreg is
absent everywhere, so the GP is the restated posterior record (fixed, plausible hyper-parameters instead of ``gp.fit``).
"""
import os
import time

import numpy as np
import pytest

from oracle import gp_posterior as gpp
from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference package not available")]


def _options(D, user=None):
    ref_loader.load()
    import pyvbmc.vbmc as vpk
    from pyvbmc.vbmc.options import Options

    base = os.path.join(os.path.dirname(vpk.__file__), "option_configs")
    opts = Options(os.path.join(base, "basic_vbmc_options.ini"), evaluation_parameters={"D": D}, user_options=user)
    opts.load_options_file(os.path.join(base, "advanced_vbmc_options.ini"), evaluation_parameters={"D": D})
    return opts


def _mixture_gp(D, S=3, seed=0):
    """GP surrogate of the log density of 0.5 N([-m, 0..], I) + 0.5 N([m, 0..], I) (the reference tests' target)."""
    from scipy.stats import multivariate_normal as mvn

    rng = np.random.default_rng(seed)
    if D == 1:
        m = 2.0
        X = np.linspace(-5, 5, 200).reshape(-1, 1)
    else:
        m = 1.5
        g = np.linspace(-3, 3, 9)
        X = np.stack(np.meshgrid(g, g), axis=-1).reshape(-1, 2)
    c1, c2 = np.zeros(D), np.zeros(D)
    c1[0], c2[0] = m, -m
    y = np.log(0.5 * mvn.pdf(X, mean=c1, cov=1.0) + 0.5 * mvn.pdf(X, mean=c2, cov=1.0))
    lay = gpp.hyp_layout(D, 1, "negquad")
    hyps = np.zeros((S, lay["H"]))
    for s in range(S):
        h = hyps[s]
        h[:D] = np.log(1.5) + 0.05 * rng.normal(size=D)
        h[D], h[D + 1] = np.log(2.0) + 0.05 * rng.normal(), np.log(1e-3)
        b = lay["mean_start"]
        h[b] = y.max()
        h[b + 1 : b + 1 + D] = 0.0
        h[b + 1 + D : b + 1 + 2 * D] = np.log(2.5)
    gp = ref_loader.make_ref_gp(X, y.reshape(-1, 1), gpp.posteriors(X, y, hyps))
    return gp, m


def _run(D, fast, slow, optim_state, user=None, seed=11):
    ref = ref_loader.load()
    from pyvbmc.vbmc import variational_optimization as vo

    gp, m = _mixture_gp(D)
    opts = _options(D, user)
    np.random.seed(seed)
    vp = ref.VariationalPosterior(D=D, K=2)
    t0 = time.perf_counter()
    vp, var_ss, pruned = vo.optimize_vp(opts, dict(optim_state), vp, gp, fast, slow)
    return vp, time.perf_counter() - t0, m


def _truth_moments(D, m):
    mu = np.zeros(D)
    cov = np.eye(D)
    cov[0, 0] += m * m
    return mu, cov


def _kl(mu1, S1, mu2, S2):
    ref_loader.load()
    from pyvbmc.stats import kl_div_mvn

    return np.abs(kl_div_mvn(np.atleast_2d(mu1), np.atleast_2d(S1), np.atleast_2d(mu2), np.atleast_2d(S2)))


@pytest.mark.parametrize("D,fast,slow,entropy_switch,elbo_tol,kl_tol", [
    (2, 100, 2, False, 0.1, 2e-2),    # test_vp_optimize_2D_g_mixture (:301-360)
    (1, 100, 2, False, 0.05, 2e-2),   # test_vp_optimize_1D_g_mixture (:244-298; KL on analytic moments instead of 1e7 samples)
    (1, 10, 1, True, 0.25, 2e-2),     # test_vp_optimize_deterministic_entropy_approximation (:363-413): BFGS + entlb
])
def test_unmodified_optimize_vp_on_the_device(D, fast, slow, entropy_switch, elbo_tol, kl_tol):
    import pyvbmc_b200 as pv

    ref_loader.load()  # puts the (stubbed) reference package on sys.path
    from pyvbmc.vbmc import variational_optimization as vo

    state = {"warmup": True, "entropy_switch": entropy_switch}
    vp_cpu, t_cpu, m = _run(D, fast, slow, state)
    before = {k: getattr(vo, k) for k in ("_neg_elcbo", "_gp_log_joint", "entmc_vbmc", "entlb_vbmc", "minimize_adam", "_sieve")}
    sites = pv.install(device_adam=True, batched_sieve=True)
    try:
        assert vo._neg_elcbo is pv._neg_elcbo and len(sites) >= 12
        launches0 = sum(c.launch_count for c, _ in pv.context._gp_ctx.values())
        vp_gpu, t_gpu, _ = _run(D, fast, slow, state)
        launches = sum(c.launch_count for c, _ in pv.context._gp_ctx.values()) - launches0
    finally:
        pv.uninstall()
    for k, f in before.items():
        assert getattr(vo, k) is f, k  # uninstall() restored every site
    assert launches > 100  # the patched run really went through the CUDA library
    print(f"\noptimize_vp D={D} entropy_switch={entropy_switch}: unpatched {t_cpu:.2f} s, patched {t_gpu:.2f} s; "
          f"elbo cpu {vp_cpu.stats['elbo']:.4f} gpu {vp_gpu.stats['elbo']:.4f}; {launches} kernel launches")
    # the reference's own acceptance criteria, for both runs
    mu_t, S_t = _truth_moments(D, m)
    for vp in (vp_cpu, vp_gpu):
        assert np.abs(vp.stats["elbo"]) < elbo_tol
        mu, S = vp.moments(orig_flag=False, cov_flag=True)
        assert np.all(_kl(mu_t, S_t, mu, S) < kl_tol)
        assert np.isfinite(vp.stats["elbo_sd"]) and vp.stats["I_sk"].shape[1] == vp.K
    # and the two runs agree with each other well inside those tolerances (different RNG streams: statistical)
    assert abs(vp_cpu.stats["elbo"] - vp_gpu.stats["elbo"]) < elbo_tol
    mu_c, S_c = vp_cpu.moments(orig_flag=False, cov_flag=True)
    mu_g, S_g = vp_gpu.moments(orig_flag=False, cov_flag=True)
    assert np.all(_kl(mu_c, S_c, mu_g, S_g) < kl_tol)


def test_device_pdf_inside_the_reference_class():
    """install(device_pdf=True): the reference's own VariationalPosterior.pdf -> device, same values, restored after."""
    import pyvbmc_b200 as pv

    ref = ref_loader.load()
    rng = np.random.default_rng(1)
    D, K = 3, 4
    vp = ref_loader.make_ref_vp(D, K, rng.normal(size=(D, K)), np.exp(0.2 * rng.normal(size=K)), np.exp(0.2 * rng.normal(size=D)),
                                np.ones(K) / K, np.zeros(K))
    x = rng.normal(size=(50, D))
    orig = ref.VariationalPosterior.pdf
    want = vp.pdf(x, orig_flag=False, log_flag=True)
    want_o = vp.pdf(x)  # original space: identity transform of an unbounded problem, Jacobian 1
    want_t = vp.pdf(x, orig_flag=False, df=4.0)
    pv.install(device_pdf=True)
    try:
        assert ref.VariationalPosterior.pdf is not orig
        got = vp.pdf(x, orig_flag=False, log_flag=True)
        got_o = vp.pdf(x)
        got_t = vp.pdf(x, orig_flag=False, df=4.0)  # heavy-tailed variant: the reference's own code
    finally:
        pv.uninstall()
    assert ref.VariationalPosterior.pdf is orig
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12) and np.allclose(got_o, want_o, rtol=1e-12)
    assert np.array_equal(got_t, want_t)
