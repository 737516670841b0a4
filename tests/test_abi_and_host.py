"""CPU: the C-ABI library loads and exports every symbol include/vbmc_b200.h declares (no compute calls
without a GPU), the product refuses to run without a CUDA device, and the host-side plumbing (theta
packing, soft bounds, name rebinding) matches the oracle / the reference's MATLAB fixtures."""
import os
import re
import sys
import types

import numpy as np
import pytest

from golden_util import load_npz
from oracle import elbo_oracle as eo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_available():
    from oracle import ref_loader

    return ref_loader.available()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vbmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vbmc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from pyvbmc_b200 import _capi

    lib = _capi.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vbmc_b200.h but not exported by libvbmc_b200.so"
        assert n in _capi.PROTOTYPES, f"{n} has no ctypes prototype in pyvbmc_b200/_capi.py"
    assert lib.vbmc_abi_version() == 1
    assert isinstance(lib.vbmc_device_count(), int)


def test_no_cpu_fallback():
    import pyvbmc_b200 as pv

    if pv._capi.load().vbmc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    vp = pv.VariationalPosterior(2, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pv.entlb_vbmc(vp)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pv.entmc_vbmc(vp, 10)


def test_product_does_not_import_oracle():
    """Nothing under pyvbmc_b200/ may import oracle/ (the checker) or the reference."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pyvbmc_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+(oracle|pyvbmc\b(?!_))", txt, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def _pair(D, K, seed, opt=(True,) * 4):
    import pyvbmc_b200 as pv

    rng = np.random.default_rng(seed)
    mu, sigma, lambd = rng.normal(size=(D, K)), np.exp(rng.normal(size=K)), np.exp(rng.normal(size=D))
    w = rng.dirichlet(np.ones(K))
    a = pv.VariationalPosterior(D, K)
    a.mu, a.sigma, a.lambd, a.w, a.eta = mu.copy(), sigma.reshape(1, -1).copy(), lambd.reshape(-1, 1).copy(), w.reshape(1, -1).copy(), np.log(w).reshape(1, -1)
    a.optimize_mu, a.optimize_sigma, a.optimize_lambd, a.optimize_weights = opt
    b = eo.OracleVP.create(D, K, mu, sigma, lambd, w, np.log(w), opt)
    return a, b


@pytest.mark.parametrize("opt", [(True,) * 4, (True, True, True, False), (True, True, False, True), (False, True, True, True)])
def test_theta_packing_matches_oracle(opt):
    a, b = _pair(3, 4, 0, opt)
    ta, tb = a.get_parameters(), eo.get_parameters(b)
    assert np.array_equal(ta, tb)
    t2 = ta + 0.1 * np.random.default_rng(1).normal(size=ta.size)
    a.set_parameters(t2)
    eo.set_parameters(b, t2)
    for f in ("mu", "sigma", "lambd", "w"):
        assert np.array_equal(np.asarray(getattr(a, f)), np.asarray(getattr(b, f))), f
    assert a.sigma.shape == (1, 4) and a.lambd.shape == (3, 1) and a.w.shape == (1, 4) and a.mu.shape == (3, 4)
    # raw_flag=False round trip and its positivity check (variational_posterior.py:706-718)
    tn = a.get_parameters(raw_flag=False)
    a.set_parameters(tn, raw_flag=False)
    assert np.allclose(a.get_parameters(raw_flag=False), tn)
    bad = tn.copy()
    bad[-1] = -1.0
    with pytest.raises(ValueError):
        a.set_parameters(bad, raw_flag=False)
    # theta is copied, never aliased (reference regression test test_variational_posterior.py:501-513)
    t3 = a.get_parameters()
    a.set_parameters(t3)
    t3[:] = 0
    assert not np.all(np.ravel(a.mu) == 0) or not opt[0]


def test_get_bounds_matlab_fixture():
    # pyvbmc/testing/variational_posterior/test_variational_posterior.py:761-791
    import pyvbmc_b200 as pv

    m = load_npz("matlab_bounds")
    vp = pv.VariationalPosterior(2, 2)
    vp.mu = m["mu"]
    options = {"tol_con_loss": 0.01, "tol_weight": 1e-2, "weight_penalty": 0.1, "tol_length": 1e-6}
    bnd = vp.get_bounds(m["X"], options)
    assert np.allclose(bnd["lb"], m["lb"]) and np.allclose(bnd["ub"], m["ub"])
    assert bnd["tol_con"] == 0.01 and bnd["weight_threshold"] == 0.125 and bnd["weight_penalty"] == 0.1
    # accumulating bounds over calls + tol_weight == 0 (final boost, :207-208)
    bnd2 = vp.get_bounds(m["X"] * 2.0, dict(options, tol_weight=0), K=3)
    assert bnd2["lb"].size == 2 * 3 + 2 * 3 + 3 and np.all(np.isneginf(bnd2["lb"][-3:]))
    assert np.all(bnd2["lb"][:2] <= bnd["lb"][:2])


def test_bound_loss_host_mirror_matches_oracle_and_matlab():
    import pyvbmc_b200 as pv

    m = load_npz("matlab_vbmc")
    vp = pv.VariationalPosterior(2, 2)
    vp.mu = m["mu"]
    options = {"tol_con_loss": 0.01, "tol_weight": 1e-2, "weight_penalty": 0.1, "tol_length": 1e-6}
    theta = vp.get_parameters()
    bnd = vp.get_bounds(m["X"], options, 2)
    L, dL = pv._vp_bound_loss(vp, theta, bnd)
    assert L == 0.0 and np.all(dL == 0.0)
    theta[-1] = 1.0
    L, dL = pv._vp_bound_loss(vp, theta, bnd, tol_con=0.01)
    assert np.isclose(L, m["bound_L"]) and np.isclose(dL[-1], m["bound_dL_last"]) and np.all(dL[:-1] == 0.0)
    x = np.zeros(3)
    x[0], x[1] = 15.0, -20.0
    L, dL = pv._soft_bound_loss(x, np.full(3, -10.0), np.full(3, 10.0), compute_grad=True)
    assert np.isclose(L, 156250.0) and np.isclose(dL[0], 12500.0) and np.isclose(dL[1], -25000.0)
    # D != K with violations in every block: must equal the oracle (which equals the reference incl. its reshape quirk)
    a, b = _pair(3, 5, 3)
    X = np.random.default_rng(0).normal(size=(30, 3))
    bnd = a.get_bounds(X, options)
    th = a.get_parameters()
    th[1] += 9.0
    th[3 * 5 + 2] += 7.0
    th[3 * 5 + 5 + 1] -= 30.0
    th[-1] -= 12.0
    La, dLa = pv._vp_bound_loss(a, th, bnd, tol_con=0.01)
    Lb, dLb = eo.vp_bound_loss(b, th, bnd, tol_con=0.01)
    assert np.isclose(La, Lb, rtol=1e-14) and np.allclose(dLa, dLb, rtol=1e-14, atol=0)


def test_install_rebinds_every_site(monkeypatch):
    import pyvbmc_b200 as pv
    import importlib

    inst = importlib.import_module("pyvbmc_b200.install")

    def mk(name, attrs):
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, object())
        return m

    fake = {
        "pyvbmc": mk("pyvbmc", []),
        "pyvbmc.entropy": mk("pyvbmc.entropy", ["entmc_vbmc", "entlb_vbmc"]),
        "pyvbmc.vbmc": mk("pyvbmc.vbmc", []),
        "pyvbmc.vbmc.variational_optimization": mk(
            "pyvbmc.vbmc.variational_optimization",
            ["entmc_vbmc", "entlb_vbmc", "_gp_log_joint", "_neg_elcbo", "_vp_bound_loss", "_soft_bound_loss"],
        ),
        "pyvbmc.vbmc.active_sample": mk("pyvbmc.vbmc.active_sample", ["_gp_log_joint", "_neg_elcbo"]),
    }
    for k, v in fake.items():
        monkeypatch.setitem(sys.modules, k, v)
    before = fake["pyvbmc.vbmc.active_sample"]._neg_elcbo
    done = inst.install()
    assert len(done) == 10
    assert fake["pyvbmc.entropy"].entmc_vbmc is pv.entmc_vbmc
    assert fake["pyvbmc.vbmc.variational_optimization"]._neg_elcbo is pv._neg_elcbo
    assert fake["pyvbmc.vbmc.active_sample"]._gp_log_joint is pv._gp_log_joint
    inst.uninstall()
    assert fake["pyvbmc.vbmc.active_sample"]._neg_elcbo is before


def test_neg_elcbo_argument_errors_without_gpu():
    """Argument validation happens before any device work (reference :1066-1070, :1114-1118)."""
    import pyvbmc_b200 as pv

    vp = pv.VariationalPosterior(2, 2)
    theta = vp.get_parameters()
    with pytest.raises(NotImplementedError):
        pv._neg_elcbo(theta, None, vp, 1.0, 0, True, None, None)
    with pytest.raises(ValueError):
        pv._neg_elcbo(theta, None, vp, 0.0, 0, True, False, None, 0.0, True)
    with pytest.raises(NotImplementedError):
        pv._gp_log_joint(vp, None, True, True, True, True)
    # a theta that does not match the optimised groups never reaches the device (the C entry point trusts its length)
    with pytest.raises(ValueError):
        pv._neg_elcbo(theta[:-1], None, vp, 0.0, 10, True, False, None)
    vp.optimize_weights = False
    with pytest.raises(ValueError):
        pv._neg_elcbo(theta, None, vp, 0.0, 0, True, False, None)


def _make_closure(gp="GP", vp0="VP", elcbo_beta=0, ns_ent_K=100, compute_var=False, theta_bnd=None, call_name="_neg_elcbo"):
    """A closure shaped like the reference's ``vb_train_mc_fun`` (variational_optimization.py:238-249)."""
    ns = {}
    src = (
        "def outer(gp, vp0, elcbo_beta, ns_ent_K, compute_var, theta_bnd, %s):\n"
        "    def vb_train_mc_fun(theta_):\n"
        "        res = %s(theta_, gp, vp0, elcbo_beta, ns_ent_K, compute_grad=True, compute_var=compute_var, theta_bnd=theta_bnd)\n"
        "        return res[0], res[1]\n"
        "    return vb_train_mc_fun\n" % (call_name, call_name)
    )
    # `_neg_elcbo` must be a GLOBAL of the closure (as in the reference), not a free variable
    src = src.replace(", %s):" % call_name, "):")
    exec(src, ns)
    ns[call_name] = lambda *a, **k: (1.5, np.zeros(3))
    return ns["outer"](gp, vp0, elcbo_beta, ns_ent_K, compute_var, theta_bnd)


def test_minimize_adam_wrapper_recognises_the_elcbo_closure_only():
    from pyvbmc_b200.install import elcbo_closure_ingredients, make_minimize_adam

    bnd = {"lb": 1}
    assert elcbo_closure_ingredients(_make_closure(theta_bnd=bnd)) == ("GP", "VP", 100, bnd)
    assert elcbo_closure_ingredients(_make_closure(elcbo_beta=0.5)) is None        # variance-weighted objective
    assert elcbo_closure_ingredients(_make_closure(compute_var=True)) is None
    assert elcbo_closure_ingredients(_make_closure(ns_ent_K=0)) is None            # deterministic entropy
    assert elcbo_closure_ingredients(_make_closure(call_name="other_fun")) is None
    assert elcbo_closure_ingredients(lambda x: (0.0, x)) is None                   # any other objective
    assert elcbo_closure_ingredients(np.sin) is None

    calls = []

    def ref_adam(f, x0, *a):
        calls.append(("ref", a))
        return "ref-result"

    def dev_loop(gp, vp0, x0, Ns, theta_bnd, *a):
        calls.append(("dev", gp, vp0, Ns, theta_bnd, a))
        if Ns == 7:
            raise FloatingPointError
        return "dev-result"

    wrapped = make_minimize_adam(ref_adam, dev_loop)
    x0 = np.zeros(3)
    assert wrapped(_make_closure(theta_bnd=bnd), x0, None, None, 1e-3, 50, 0.001, 0.05, 200) == "dev-result"
    assert calls[-1] == ("dev", "GP", "VP", 100, bnd, (None, None, 1e-3, 50, 0.001, 0.05, 200, True))
    assert wrapped(lambda x: (0.0, x), x0, max_iter=5) == "ref-result" and calls[-1][0] == "ref"
    assert calls[-1][1] == (None, None, 0.001, 5, 0.001, 0.1, 200, True)          # reference defaults forwarded
    assert wrapped(_make_closure(ns_ent_K=7), x0) == "ref-result"                   # non-finite -> host loop
    assert [c[0] for c in calls[-2:]] == ["dev", "ref"]


@pytest.mark.skipif(not _ref_available(), reason="needs the reference package (checkout or oracle/_ref archive)")
def test_minimize_adam_wrapper_matches_the_reference_closure_shape():
    """The free variables and globals the wrapper keys on are those of the reference's own nested function."""
    import types

    from oracle import ref_loader
    from pyvbmc_b200.install import _ELCBO_FREEVARS

    ref_loader.load()
    from pyvbmc.vbmc import variational_optimization as vo

    def nested(code):
        for c in code.co_consts:
            if isinstance(c, types.CodeType):
                yield c
                yield from nested(c)

    inner = [c for c in nested(vo.optimize_vp.__code__) if c.co_name == "vb_train_mc_fun"]
    assert len(inner) == 1
    assert set(_ELCBO_FREEVARS) <= set(inner[0].co_freevars) and "_neg_elcbo" in inner[0].co_names
    import inspect

    assert list(inspect.signature(vo.minimize_adam).parameters) == [
        "f", "x0", "lb", "ub", "tol_fun", "max_iter", "master_min", "master_max", "master_decay", "use_early_stopping"]


def test_sieve_wrapper_batches_the_candidate_loop_and_keeps_the_reference_order():
    """make_sieve around a stand-in with the reference's loop shape (variational_optimization.py:775-800)."""
    from pyvbmc_b200.install import make_sieve

    mod = types.SimpleNamespace()
    truth = {"a": 3.0, "b": -1.0, "c": 2.0, "d": 0.5}
    real_calls = []

    def real_neg_elcbo(theta, gp, vp, beta=0.0, Ns=0, compute_grad=True, compute_var=None, theta_bnd=None):
        real_calls.append(vp)
        return truth[vp] + (100.0 if Ns else 0.0), None, 0.0, 0.0, 0.0

    mod._neg_elcbo = real_neg_elcbo

    def ref_sieve(init_N, ns_fast=0):
        # (shape of the reference: evaluate every candidate through the MODULE's _neg_elcbo, argsort, reorder)
        vp0_vec = np.array(list("abcd")[:init_N], dtype=object)
        vp0_type = np.arange(init_N) + 10
        if init_N == 0:
            return ("vp-copy", 1, 0, False, 7, ns_fast)
        fill = np.zeros(init_N)
        for i, vp0 in enumerate(vp0_vec):
            fill[i] = mod._neg_elcbo(np.full(2, float(i)), "GP", vp0, 0, ns_fast, 0, False, {"lb": 0})[0]
        order = np.argsort(fill)
        return (vp0_vec[order], vp0_type[order], 0, False, 7, ns_fast)

    batch_calls = []

    def batch_fn(vps, gp, theta_bnd, thetas=None):
        batch_calls.append((list(vps), gp, theta_bnd, [t.tolist() for t in thetas]))
        F = np.array([truth[v] for v in vps])
        return F, F, F

    sieve = make_sieve(ref_sieve, mod, batch_fn)
    out = sieve(4)
    assert list(out[0]) == ["b", "d", "c", "a"] and list(out[1]) == [11, 13, 12, 10] and out[2:] == (0, False, 7, 0)
    assert len(batch_calls) == 1 and real_calls == []                       # ONE batched call, no single calls
    assert batch_calls[0][0] == list("abcd") and batch_calls[0][1] == "GP" and batch_calls[0][2] == {"lb": 0}
    assert batch_calls[0][3] == [[0.0, 0.0], [1.0, 1.0], [2.0, 2.0], [3.0, 3.0]]
    assert mod._neg_elcbo is real_neg_elcbo                                 # restored
    # stochastic fast entropy (ns_ent_K_fast > 0): untouched reference flow through the real function
    out = sieve(3, ns_fast=5)
    assert list(out[0]) == ["b", "c", "a"] and len(batch_calls) == 1 and real_calls == list("abc")
    # no candidates: the reference's early return passes through
    assert sieve(0) == ("vp-copy", 1, 0, False, 7, 0)
    # the module function is restored even if the reference raises
    def boom(*a, **k):
        raise RuntimeError("x")
    with pytest.raises(RuntimeError):
        make_sieve(boom, mod, batch_fn)()
    assert mod._neg_elcbo is real_neg_elcbo


@pytest.mark.skipif(not _ref_available(), reason="needs the reference package (checkout or oracle/_ref archive)")
def test_sieve_wrapper_matches_the_reference_loop_shape():
    """The reference's _sieve calls the module-level _neg_elcbo positionally as
    (theta, gp, vp0, 0, ns_ent_K_fast, 0, compute_var, theta_bnd) and returns a 6-tuple led by (vp0_vec, vp0_type)."""
    import inspect

    from oracle import ref_loader

    ref_loader.load()
    from pyvbmc.vbmc import variational_optimization as vo

    src = inspect.getsource(vo._sieve)
    assert "_neg_elcbo(" in src and "ns_ent_K_fast" in src and "np.argsort(nelcbo_fill)" in src
    assert "_neg_elcbo" in vo._sieve.__code__.co_names  # looked up in the module at call time => rebinding works
    call = src[src.index("_neg_elcbo("):]
    args = [a.strip() for a in call[call.index("(") + 1 : call.index(")")].split(",") if a.strip()]
    assert args == ["theta", "gp", "vp0", "0", "ns_ent_K_fast", "0", "compute_var", "theta_bnd"]


@pytest.mark.skipif(not _ref_available(), reason="needs the reference package (checkout or oracle/_ref archive)")
@pytest.mark.parametrize("best_N", [1, 5])
def test_sieve_wrapper_end_to_end_on_the_unmodified_reference(best_N):
    """make_sieve around the REAL pyvbmc _sieve (options, get_hpd, _vb_init, soft bounds all run as shipped), with
    the reference's own _neg_elcbo as the stand-in for the batched kernel: same candidates, same order, same types
    and same returned settings as the unwrapped reference with the same NumPy seed."""
    from oracle import gp_posterior as gpp
    from oracle import ref_loader
    from pyvbmc_b200.install import make_sieve

    ref_loader.load()
    import pyvbmc.vbmc as vpk
    from pyvbmc.vbmc import variational_optimization as vo
    from pyvbmc.vbmc.options import Options

    D, K, N, S = 3, 4, 60, 2
    rng = np.random.default_rng(5)
    X = rng.normal(size=(N, D))
    y = -0.5 * np.sum(X**2, axis=1) + 0.1 * rng.normal(size=N)
    lay = gpp.hyp_layout(D, 1, "negquad")
    hyps = np.zeros((S, lay["H"]))
    for s in range(S):
        h = hyps[s]
        h[:D] = rng.normal(0.0, 0.3, size=D)
        h[D], h[D + 1] = rng.normal(0.5, 0.2), np.log(1e-2)
        b = lay["mean_start"]
        h[b] = y.max()
        h[b + 1 : b + 1 + D] = 0.1 * rng.normal(size=D)
        h[b + 1 + D : b + 1 + 2 * D] = np.log(2.0)
    gp = ref_loader.make_ref_gp(X, y.reshape(-1, 1), gpp.posteriors(X, y, hyps))
    base = os.path.join(os.path.dirname(vpk.__file__), "option_configs")
    opts = Options(os.path.join(base, "basic_vbmc_options.ini"), evaluation_parameters={"D": D}, user_options=None)
    opts.load_options_file(os.path.join(base, "advanced_vbmc_options.ini"), evaluation_parameters={"D": D})
    optim_state = {"entropy_switch": False}

    def fresh_vp():
        eta = np.zeros(K)
        return ref_loader.make_ref_vp(D, K, 0.3 * np.arange(D * K).reshape(D, K) / (D * K), 0.4 * np.ones(K), np.ones(D),
                                      np.ones(K) / K, eta)

    def describe(out):
        vec, typ = out[0], out[1]
        return [np.concatenate([v.get_parameters(), [t]]) for v, t in zip(vec, typ)], out[2:]

    real_neg_elcbo = vo._neg_elcbo
    batch_sizes = []

    def batch_fn(vps, gp_, theta_bnd, thetas=None):  # CPU stand-in for neg_elcbo_batch: the reference, one at a time
        batch_sizes.append(len(vps))
        F = np.array([real_neg_elcbo(t, gp_, v, 0, 0, 0, False, theta_bnd)[0] for v, t in zip(vps, thetas)])
        return F, F, F

    np.random.seed(123)
    want, want_rest = describe(vo._sieve(opts, optim_state, fresh_vp(), gp, init_N=12, best_N=best_N, K=K))
    np.random.seed(123)
    got, got_rest = describe(make_sieve(vo._sieve, vo, batch_fn)(opts, optim_state, fresh_vp(), gp, init_N=12, best_N=best_N, K=K))
    assert batch_sizes == [12] and vo._neg_elcbo is real_neg_elcbo
    assert got_rest == want_rest and len(got) == len(want) == 12
    for a, b in zip(got, want):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)


@pytest.mark.skipif(not _ref_available(), reason="needs the reference package (checkout or oracle/_ref archive)")
def test_both_wrappers_inside_the_unmodified_optimize_vp():
    """The real pyvbmc optimize_vp (sieve -> Adam on the best candidates -> full ELCBO -> pruning) with BOTH optional
    wrappers installed and the reference's own functions as CPU stand-ins for the two device entry points: the
    closure is recognised at the real call site, every argument is forwarded, and the result is bit-identical to the
    unwrapped run with the same NumPy seed."""
    from oracle import gp_posterior as gpp
    from oracle import ref_loader
    from pyvbmc_b200.install import make_minimize_adam, make_sieve

    ref_loader.load()
    import pyvbmc.vbmc as vpk
    from pyvbmc.vbmc import variational_optimization as vo
    from pyvbmc.vbmc.options import Options

    D, K, N, S = 3, 3, 50, 2
    rng = np.random.default_rng(5)
    X = rng.normal(size=(N, D))
    y = -0.5 * np.sum(X**2, axis=1) + 0.1 * rng.normal(size=N)
    lay = gpp.hyp_layout(D, 1, "negquad")
    hyps = np.zeros((S, lay["H"]))
    for s in range(S):
        h = hyps[s]
        h[:D] = rng.normal(0.0, 0.3, size=D)
        h[D], h[D + 1] = rng.normal(0.5, 0.2), np.log(1e-2)
        b = lay["mean_start"]
        h[b] = y.max()
        h[b + 1 : b + 1 + D] = 0.1 * rng.normal(size=D)
        h[b + 1 + D : b + 1 + 2 * D] = np.log(2.0)
    gp = ref_loader.make_ref_gp(X, y.reshape(-1, 1), gpp.posteriors(X, y, hyps))
    base = os.path.join(os.path.dirname(vpk.__file__), "option_configs")
    opts = Options(os.path.join(base, "basic_vbmc_options.ini"), evaluation_parameters={"D": D},
                   user_options={"max_iter_stochastic": 60})
    opts.load_options_file(os.path.join(base, "advanced_vbmc_options.ini"), evaluation_parameters={"D": D})

    def run():
        vp = ref_loader.make_ref_vp(D, K, 0.3 * np.arange(D * K).reshape(D, K) / (D * K), 0.4 * np.ones(K), np.ones(D),
                                    np.ones(K) / K, np.zeros(K))
        np.random.seed(3)
        vp2, var_ss, pruned = vo.optimize_vp(opts, {"entropy_switch": False, "warmup": False}, vp, gp, 4, 1, K)
        return np.concatenate([vp2.get_parameters(), [vp2.stats["elbo"], vp2.stats["elbo_sd"], var_ss, pruned]])

    want = run()
    ref_adam, ref_sieve, real_neg_elcbo = vo.minimize_adam, vo._sieve, vo._neg_elcbo
    seen = {"adam": [], "sieve": []}

    def dev_loop(gp_, vp0, x0, Ns, theta_bnd, lb, ub, tol_fun, max_iter, master_min, master_max, master_decay, early):
        seen["adam"].append((Ns, max_iter, master_min, master_max, master_decay, tol_fun, early))
        f = lambda t: real_neg_elcbo(t, gp_, vp0, 0, Ns, compute_grad=True, compute_var=False, theta_bnd=theta_bnd)[:2]
        return ref_adam(f, x0, lb, ub, tol_fun, max_iter, master_min, master_max, master_decay, early)

    def batch_fn(vps, gp_, theta_bnd, thetas=None):
        seen["sieve"].append(len(vps))
        F = np.array([real_neg_elcbo(t, gp_, v, 0, 0, 0, False, theta_bnd)[0] for v, t in zip(vps, thetas)])
        return F, F, F

    vo.minimize_adam = make_minimize_adam(ref_adam, dev_loop)
    vo._sieve = make_sieve(ref_sieve, vo, batch_fn)
    try:
        got = run()
    finally:
        vo.minimize_adam, vo._sieve = ref_adam, ref_sieve
    assert vo._neg_elcbo is real_neg_elcbo
    assert len(seen["sieve"]) == 1 and seen["sieve"][0] > 0      # the candidate loop became one batch
    assert len(seen["adam"]) >= 1 and all(a[1] == 60 and a[0] > 0 for a in seen["adam"])  # every Adam run was routed
    np.testing.assert_array_equal(got, want)


class _AdamStandIn:
    """Host stand-in for the device side of minimize_adam_elcbo: the same enqueue / fetch protocol, the Adam update of
    the reference (minimize_adam.py:87-99) on a seeded noisy quadratic.  Issues are recorded to check the pipelining."""

    def __init__(self, f):
        self.f = f
        self.log = []

    def set_bounds(self, theta_bnd):
        return False

    def param_len(self, D, K):
        return D * K + 5 * K + 2 * D

    def adam_init(self, D, K, prm, x0, optimize, Ns, use_bounds, seed, offset, lb, ub, max_iter, mmin, mmax, mdecay):
        self.x = np.array(x0, dtype=float)
        self.m = self.v = 0.0
        self.i = 0
        self.cfg = (mmin, mmax, mdecay)
        self.y_tab, self.x_tab = [], []

    def adam_enqueue(self, n):
        self.log.append(("enqueue", self.i, n))
        mmin, mmax, mdecay = self.cfg
        for _ in range(n):
            y, g = self.f(self.x)
            self.m = 0.9 * self.m + 0.1 * g
            self.v = 0.999 * self.v + 0.001 * g**2
            i = self.i
            step = mmin + (mmax - mmin) * np.exp(-(i + 1) / mdecay)
            self.x = self.x - step * (self.m / (1 - 0.9 ** (i + 1))) / (np.sqrt(self.v / (1 - 0.999 ** (i + 1))) + np.sqrt(np.spacing(1)))
            self.y_tab.append(y)
            self.x_tab.append(self.x.copy())
            self.i += 1

    def adam_fetch(self, i0, n):
        self.log.append(("fetch", i0, n))
        assert i0 + n <= self.i
        return np.array(self.y_tab[i0 : i0 + n]), np.array(self.x_tab[i0 : i0 + n])


@pytest.mark.parametrize("early,max_iter", [(True, 400), (False, 70), (True, 30)])
def test_minimize_adam_elcbo_host_loop_matches_the_reference_loop(monkeypatch, early, max_iter):
    """Control flow of pyvbmc_b200.minimize_adam_elcbo (batches issued one ahead of the host, closed-form stopping rule,
    returned averages) against the oracle's restatement of minimize_adam.py:61-145 (itself pinned to the unmodified
    reference): same iterates, same stopping iteration."""
    import pyvbmc_b200 as pv
    from oracle.minimize_adam_oracle import minimize_adam, noisy_quadratic
    from pyvbmc_b200.vbmc import minimize_adam as mod

    D, K = 1, 2  # theta = [mu (2) | ln sigma (2) | ln lambda (1) | eta (2)]: 7 entries; a 7-variable quadratic
    f1, _ = noisy_quadratic(n=7, seed=3)
    f2, _ = noisy_quadratic(n=7, seed=3)
    x0 = np.full(7, 2.0)
    x0[-2:] = [-0.5, 0.0]  # max(eta) == 0 already: the in-place shift of the real objective is a no-op here
    kw = dict(max_iter=max_iter, master_max=0.05, use_early_stopping=early)
    ref = minimize_adam(f1, x0.copy(), **kw)
    fake = _AdamStandIn(f2)
    monkeypatch.setattr(mod, "context_for_gp", lambda gp, need_L=False: fake)
    vp = pv.VariationalPosterior(D, K)
    got = pv.minimize_adam_elcbo(object(), vp, x0.copy(), 10, None, seed=1, **kw)
    assert got[4] == ref[4]
    assert np.allclose(got[0], ref[0], rtol=0, atol=1e-12) and abs(got[1] - ref[1]) < 1e-12
    assert np.allclose(got[2], ref[2], rtol=0, atol=1e-12) and np.allclose(got[3], ref[3], rtol=0, atol=1e-12)
    # pipelining: every fetch of a batch is preceded by the enqueue of the NEXT one (while iterations remain)
    enq = [e for e in fake.log if e[0] == "enqueue"]
    for idx, e in enumerate(fake.log):
        if e[0] == "fetch":
            issued_before = sum(n for k, _, n in fake.log[:idx] if k == "enqueue")
            assert issued_before >= min(e[1] + e[2] + 20, max_iter) or issued_before == max_iter
    assert sum(n for _, _, n in enq) <= max_iter
    if early and max_iter == 400:
        assert ref[4] < max_iter  # the rule fired (otherwise this case tests nothing)
