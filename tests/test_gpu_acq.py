"""Acquisition-function ingredients on the device against the oracle and the goldens produced by the
unmodified reference (tests/golden/ref_acq.npz, oracle/make_golden_acq.py).  Every call goes through the C ABI
(vbmc_gp_predict / vbmc_vp_pdf)."""
from types import SimpleNamespace

import numpy as np
import pytest

from golden_util import load_npz
from oracle import acq_oracle as ao
from oracle import elbo_oracle as eo
from oracle import gp_posterior as gpp
from oracle import synthetic as syn

pytestmark = pytest.mark.gpu

ACQ_CASES = [("c2", "C2", dict(N=64)), ("c4", "C4", dict(N=60)), ("c1", "C1", {})]
TOL_MU = 1e-11   # f_mu = m + k*.alpha: plain fp64 dot products, relative to max |f_mu|
TOL_S2 = 1e-9    # f_s2 = sf2 - |L^-T k*|^2 is a cancellation: absolute error relative to sf2 (both sides solve with
                 # the same factor; the device multiplies by the explicit triangular inverse)


@pytest.fixture(scope="module")
def pv():
    import pyvbmc_b200 as pv

    yield pv
    pv.clear_caches()


def _vp(pv, pr):
    vp = pv.VariationalPosterior(pr.D, pr.K)
    vp.mu, vp.sigma, vp.lambd = pr.vp.mu.copy(), pr.vp.sigma.copy(), pr.vp.lambd.copy()
    vp.w, vp.eta = pr.vp.w.copy(), pr.vp.eta.copy()
    return vp


def _check_predict(pv, pr, Xs):
    f_mu, f_s2 = pv.gp_predict(pr.gp, Xs, separate_samples=True)
    o_mu, o_s2 = ao.gp_predict(pr.X, pr.posts, Xs, pr.mean_kind, separate_samples=True)
    assert f_mu.shape == o_mu.shape == (Xs.shape[0], pr.S)
    sf2 = np.exp(2.0 * pr.hyps[:, pr.D])
    e_mu = np.abs(f_mu - o_mu).max() / np.abs(o_mu).max()
    e_s2 = (np.abs(f_s2 - o_s2) / sf2[None, :]).max()
    assert e_mu < TOL_MU and e_s2 < TOL_S2, (e_mu, e_s2)
    assert np.all(f_s2 >= 0.0)
    # mixture moments (separate_samples = False)
    m1, v1 = pv.gp_predict(pr.gp, Xs)
    om, ov = ao.gp_predict(pr.X, pr.posts, Xs, pr.mean_kind, separate_samples=False)
    assert m1.shape == om.shape and np.abs(m1 - om).max() <= TOL_MU * np.abs(om).max()
    assert np.abs(v1 - ov).max() <= 1e-9 * max(np.abs(ov).max(), sf2.max())
    return f_mu, f_s2


@pytest.mark.parametrize("tag,cfg,kw", ACQ_CASES)
def test_gp_predict_golden_points(pv, tag, cfg, kw):
    g = load_npz("ref_acq")
    pr = syn.make_problem(cfg, **kw)
    f_mu, f_s2 = _check_predict(pv, pr, g[f"{tag}_Xs"])
    sf2 = np.exp(2.0 * pr.hyps[:, pr.D])
    assert np.abs(f_mu - g[f"{tag}_f_mu"]).max() <= TOL_MU * np.abs(g[f"{tag}_f_mu"]).max()
    assert (np.abs(f_s2 - g[f"{tag}_f_s2"]) / sf2[None, :]).max() < TOL_S2


@pytest.mark.parametrize("cfg,kw,Nx", [("C2", {}, 333), ("C2", dict(ill_conditioned=True), 100), ("C2", dict(mean_kind="const"), 64),
                                       ("C2", dict(mean_kind="zero"), 31), ("C4", {}, 257), ("C3", {}, 1000),
                                       ("C2", dict(N=33, S=1), 5), ("C2", dict(N=7, S=3), 1)])
def test_gp_predict_shapes_and_means(pv, cfg, kw, Nx):
    """Ragged sizes (Nx and N not multiples of the 32-point / 8-row tiles), every mean function, heteroskedastic noise
    (C4), a single hyper-sample, the ill-conditioned hyper-parameter set, the headline N = 400."""
    pr = syn.make_problem(cfg, **kw)
    rng = np.random.default_rng(Nx)
    Xs = pr.X[rng.integers(0, pr.N, size=Nx)] + 0.3 * rng.normal(size=(Nx, pr.D))
    Xs[0] = pr.X[0]  # exactly on a training point: the variance collapses towards the noise floor
    _check_predict(pv, pr, Xs)


def test_gp_predict_low_noise_branch(pv):
    """L_chol = False: L = -(K + sn2 I)^-1, f_s2 = sf2 + k* . (L k*) (dense product instead of the triangular one)."""
    pr = syn.make_problem("C2", N=48)
    pr.posts = gpp.posteriors(pr.X, pr.y, pr.hyps, s2=pr.s2, mean_kind=pr.mean_kind, force_low_noise=True)
    pr.gp = eo.make_gp(pr.X, pr.posts, mean_kind=pr.mean_kind, y=pr.y)
    assert not any(p["L_chol"] for p in pr.posts)
    rng = np.random.default_rng(3)
    Xs = pr.X[rng.integers(0, pr.N, size=70)] + 0.3 * rng.normal(size=(70, pr.D))
    _check_predict(pv, pr, Xs)


def test_gp_predict_full_batch_is_consistent(pv):
    """The search-cache size of the reference (8192 points, advanced_vbmc_options.ini:41) at the headline GP: the batch
    result equals the same points predicted in small batches (bitwise: a point's result does not depend on its tile),
    a 256-point sample agrees with the oracle, far-away points return the prior."""
    pr = syn.make_problem("C3")
    rng = np.random.default_rng(0)
    Nx = 8192
    Xs = pr.X[rng.integers(0, pr.N, size=Nx)] + 0.5 * rng.normal(size=(Nx, pr.D))
    Xs[-4:] += 1e3
    f_mu, f_s2 = pv.gp_predict(pr.gp, Xs, separate_samples=True)
    for lo, hi in ((0, 1), (5, 70), (4000, 4100), (8100, 8192)):
        a, b = pv.gp_predict(pr.gp, Xs[lo:hi], separate_samples=True)
        assert np.array_equal(a, f_mu[lo:hi]) and np.array_equal(b, f_s2[lo:hi])
    idx = rng.choice(Nx, size=256, replace=False)
    o_mu, o_s2 = ao.gp_predict(pr.X, pr.posts, Xs[idx], pr.mean_kind, separate_samples=True)
    sf2 = np.exp(2.0 * pr.hyps[:, pr.D])
    assert np.abs(f_mu[idx] - o_mu).max() <= TOL_MU * np.abs(o_mu).max()
    assert (np.abs(f_s2[idx] - o_s2) / sf2[None, :]).max() < TOL_S2
    assert np.allclose(f_s2[-4:], sf2[None, :], rtol=1e-14)  # k* = 0: prior variance, prior mean
    with pytest.raises(ValueError):
        pv.gp_predict(pr.gp, Xs[:, :-1])
    e_mu, e_s2 = pv.gp_predict(pr.gp, np.zeros((0, pr.D)), separate_samples=True)
    assert e_mu.shape == (0, pr.S) and e_s2.shape == (0, pr.S)


@pytest.mark.parametrize("tag,cfg,kw", ACQ_CASES)
def test_vp_pdf_golden(pv, tag, cfg, kw):
    """VariationalPosterior.pdf(x, orig_flag=False, ...) against the unmodified reference."""
    g = load_npz("ref_acq")
    pr = syn.make_problem(cfg, **kw)
    vp = _vp(pv, pr)
    Xs = g[f"{tag}_Xs"]
    y = vp.pdf(Xs, orig_flag=False)
    assert y.shape == (Xs.shape[0], 1)
    assert np.allclose(y, g[f"{tag}_pdf"], rtol=1e-12, atol=0)
    ly, dly = pv.vp_pdf(vp, Xs, orig_flag=False, log_flag=True, grad_flag=True)
    assert np.array_equal(np.isneginf(ly), np.isneginf(g[f"{tag}_logpdf"])) and np.isneginf(ly[:3]).all()
    fin = np.isfinite(g[f"{tag}_logpdf"]).ravel()
    assert np.allclose(ly[fin], g[f"{tag}_logpdf"][fin], rtol=1e-12, atol=1e-12)
    assert np.allclose(dly[fin], g[f"{tag}_dlogpdf"][fin], rtol=1e-10, atol=1e-12)
    assert np.isnan(dly[~fin]).all() and np.isnan(g[f"{tag}_dlogpdf"][~fin]).all()  # 0 / 0, as in the reference
    _, dy = pv.vp_pdf(vp, Xs, orig_flag=False, grad_flag=True)
    assert np.allclose(dy, g[f"{tag}_dpdf"], rtol=1e-10, atol=1e-300)
    with pytest.raises(NotImplementedError):
        pv.vp_pdf(vp, Xs, orig_flag=False, df=3.0)
    with pytest.raises(NotImplementedError):
        vp.pdf(Xs)  # original space needs a parameter transformer
    # 1-D input (the reference's handle_0D_1D_input decorator): one point of D coordinates
    y1 = pv.vp_pdf(vp, Xs[5], orig_flag=False)
    assert y1.shape == (1, 1) and y1[0, 0] == y[5, 0]


@pytest.mark.parametrize("tag,cfg,kw", ACQ_CASES)
def test_acq_fcn_log_golden(pv, tag, cfg, kw):
    """AcqFcnLog.__call__ (abstract_acq_fcn.py:34-147 + acq_fcn_log.py) against the unmodified reference run on the same
    points: aggregation over hyper-samples, log prospective uncertainty, variance regularisation, hard bounds."""
    g = load_npz("ref_acq")
    pr = syn.make_problem(cfg, **kw)
    vp = _vp(pv, pr)
    vp.parameter_transformer = SimpleNamespace(inverse=lambda x: x)  # unbounded problem: identity
    D = pr.D
    optim_state = {"integer_vars": None, "variance_regularized_acq_fcn": True, "tol_gp_var": 1e-4,
                   "lb_eps_orig": np.full((1, D), -30.0), "ub_eps_orig": np.full((1, D), 30.0)}
    flog = SimpleNamespace(y_max=float(g[f"{tag}_y_max"]))
    acq = pv.AcqFcnLog()(g[f"{tag}_Xs"], pr.gp, vp, flog, optim_state)
    ref = g[f"{tag}_acq"]
    assert acq.shape == ref.shape and np.array_equal(np.isinf(acq), np.isinf(ref)) and np.isinf(ref[:3]).all()
    fin = np.isfinite(ref)
    # log(var_tot) amplifies the variance's cancellation error where the GP is certain: compare on the acquisition scale
    assert np.abs(acq[fin] - ref[fin]).max() <= 1e-6 * max(1.0, np.abs(ref[fin]).max())
