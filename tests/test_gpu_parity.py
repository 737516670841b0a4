"""GPU parity: the CUDA path (through the C ABI, via the reference-shaped Python functions of
``pyvbmc_b200``) against the committed golden vectors of the unmodified reference and against
the fp64 oracle on identical inputs (theta, GP arrays, eps).

Tolerances (BASELINE.json north_star): |F - F_ref| / |F_ref| <= 1e-5 and
max|dF - dF_ref| / max|dF_ref| <= 1e-4 for the default fp32-compute/fp64-accumulate entropy
kernel; the all-fp64 kernel and the fp64 log-joint / lower-bound kernels are held to 1e-9.
"""
import numpy as np
import pytest

from golden_util import REF_CASES, VAR_CASES, eps_for, load_case, load_npz, relerr, relmax
from oracle import elbo_oracle as eo
from oracle import gp_posterior as gpp
from oracle import synthetic as syn

pytestmark = pytest.mark.gpu

TOL_F32_VAL, TOL_F32_GRAD = 1e-5, 1e-4
TOL_F64 = 1e-9


@pytest.fixture(scope="module")
def pv():
    import pyvbmc_b200

    return pyvbmc_b200


def make_vp(pv, D, K, mu, sigma, lambd, w, eta, opt=(True,) * 4):
    state = np.random.get_state()
    vp = pv.VariationalPosterior(D, K)
    np.random.set_state(state)
    vp.mu = np.array(mu, dtype=float).reshape(D, K)
    vp.sigma = np.array(sigma, dtype=float).reshape(1, K)
    vp.lambd = np.array(lambd, dtype=float).reshape(D, 1)
    vp.w = np.array(w, dtype=float).reshape(1, K)
    vp.eta = np.array(eta, dtype=float).reshape(1, K)
    vp.optimize_mu, vp.optimize_sigma, vp.optimize_lambd, vp.optimize_weights = [bool(o) for o in opt]
    return vp


def case_vp(pv, c, which="vp"):
    g = c.g
    p = "vp_" if which == "vp" else "sa_"
    return make_vp(pv, c.D, c.K, g[p + "mu"], g[p + "sigma"], g[p + "lambd"], g[p + "w"], g[p + "eta"], c.opt)


# ------------------------------------------------------------------ entropy kernels
@pytest.mark.parametrize("stem", REF_CASES)
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_entmc_golden(pv, stem, prec):
    c = load_case(stem)
    g = c.g
    ctx = pv.entropy_context()
    tv, tg = (TOL_F32_VAL, TOL_F32_GRAD) if prec == "f32" else (TOL_F64, TOL_F64)
    vp = case_vp(pv, c, "sa")
    for key, seed, flags, jac in [
        ("ent", g["ent_seed"], c.opt, True),
        ("entnj", g["entnj_seed"], (True,) * 4, False),
        ("entw", g["entnj_seed"], (False, False, False, True), True),
    ]:
        eps = eps_for(seed, c.K, c.Ns_K, c.D)
        H, dH = ctx.entmc(vp, c.Ns_K, flags, jac, eps=eps, precision=prec)
        assert dH.shape == g[key + "_dH"].shape
        assert relerr(H, g[key + "_H"]) < tv, (key, H, g[key + "_H"])
        assert relmax(dH, g[key + "_dH"]) < tg, key


@pytest.mark.parametrize("stem", REF_CASES)
def test_entlb_golden(pv, stem):
    c = load_case(stem)
    g = c.g
    vp = case_vp(pv, c, "sa")
    H, dH = pv.entlb_vbmc(vp, c.opt, True)
    assert relerr(H, g["elb_H"]) < TOL_F64 and relmax(dH, g["elb_dH"]) < TOL_F64
    H, dH = pv.entlb_vbmc(vp, (True,) * 4, False)
    assert relerr(H, g["elbnj_H"]) < TOL_F64 and relmax(dH, g["elbnj_dH"]) < TOL_F64


def test_entropy_edge_cases(pv):
    g = load_npz("ref_entropy_edge")
    for i, (D, K, Ns) in enumerate(g["specs"]):
        vp = make_vp(pv, D, K, g[f"mu{i}"], g[f"sigma{i}"], g[f"lambd{i}"], g[f"w{i}"], g[f"eta{i}"])
        eps = eps_for(100 + i, K, eo.even_ns(Ns), D)
        for prec, tv, tg in (("f32", TOL_F32_VAL, TOL_F32_GRAD), ("f64", TOL_F64, TOL_F64)):
            H, dH = pv.entropy_context().entmc(vp, Ns, (True,) * 4, True, eps=eps, precision=prec)
            assert relerr(H, g[f"H{i}"]) < tv and relmax(dH, g[f"dH{i}"]) < tg, (i, prec)
        Hl, dHl = pv.entlb_vbmc(vp)
        assert relerr(Hl, g[f"Hl{i}"]) < TOL_F64 and relmax(dHl, g[f"dHl{i}"]) < TOL_F64


def test_entropy_grad_flag_shapes(pv):
    # pyvbmc/testing/entropy/test_entmc_vbmc.py:174-183, test_entlb_vbmc.py:122-131
    D, K = 4, 3
    vp = pv.VariationalPosterior(D, K)
    _, dH = pv.entmc_vbmc(vp, Ns=1e3, grad_flags=(False,) * 4)
    assert dH.shape == (0,)
    _, dH = pv.entmc_vbmc(vp, Ns=1e3, grad_flags=(False, False, False, True))
    assert dH.shape == (K,)
    _, dH = pv.entlb_vbmc(vp, grad_flags=(False,) * 4)
    assert dH.shape == (0,)
    _, dH = pv.entlb_vbmc(vp, grad_flags=(False, False, False, True))
    assert dH.shape == (K,)


def test_matlab_entropy(pv):
    # test_entlb_vbmc.py:101-119 (exact), test_entmc_vbmc.py:140-171 (1 %)
    m = load_npz("matlab_entropy")
    D, K, Ns = int(m["D"]), int(m["K"]), int(m["Ns"])
    vp = make_vp(pv, D, K, m["mu"], m["sigma"], m["lambd"], m["w"], m["eta"])
    Hl, dHl = pv.entlb_vbmc(vp, jacobian_flag=int(m["jacobian_flag"]))
    assert np.isclose(Hl, m["Hl"]) and np.allclose(dHl, m["dHl"])
    np.random.seed(42)
    H, dH = pv.entmc_vbmc(vp, Ns, grad_flags=(True,) * 4, jacobian_flag=int(m["jacobian_flag"]))
    assert np.isclose(H, m["H"], rtol=0.01)
    assert np.allclose(dH, m["dH"], rtol=0.01, atol=0.01)


def test_entmc_single_gaussian_closed_form(pv):
    # test_entmc_vbmc.py:52-69
    D, K = 3, 1
    vp = make_vp(pv, D, K, np.ones((D, K)), np.ones(K), np.ones(D), np.ones(K), np.ones(K))
    H_exact = 0.5 * D * (1 + np.log(2 * np.pi))
    dH_exact = np.concatenate([np.zeros(D), [D], np.ones(D), [H_exact - 1]])
    np.random.seed(1)
    H, dH = pv.entmc_vbmc(vp, 1e5, jacobian_flag=False)
    assert np.isclose(H, H_exact, rtol=0.01, atol=0.01)
    assert np.allclose(dH, dH_exact, rtol=0.01, atol=0.01)


def test_philox_mode_matches_oracle_on_dumped_draws(pv):
    """Production RNG: the kernel's Philox draws, dumped through vbmc_philox_normals and fed to
    the oracle, must reproduce the kernel's result; draws must look standard normal."""
    c = load_case("c2")
    vp = case_vp(pv, c, "sa")
    ctx = pv.entropy_context()
    Ns = 200
    eps = ctx.philox_normals(c.D, c.K, Ns, seed=1234, offset=7)
    assert eps.shape == (c.K, Ns // 2, c.D)
    assert abs(eps.mean()) < 0.02 and abs(eps.std() - 1) < 0.02
    H, dH = ctx.entmc(vp, Ns, (True,) * 4, True, eps=None, seed=1234, offset=7)
    Ho, dHo = eo.entmc(c.sa_vp(), eps, (True,) * 4, True)
    assert relerr(H, Ho) < TOL_F32_VAL and relmax(dH, dHo) < TOL_F32_GRAD
    H2, dH2 = ctx.entmc(vp, Ns, (True,) * 4, True, eps=None, seed=1234, offset=7)
    assert H2 == H and np.array_equal(dH, dH2)  # bitwise run-to-run determinism
    H3, _ = ctx.entmc(vp, Ns, (True,) * 4, True, eps=None, seed=1235, offset=7)
    assert H3 != H
    big = ctx.philox_normals(8, 4, 200000, seed=5)
    assert abs(big.mean()) < 5e-3 and abs(big.std() - 1) < 5e-3
    assert abs(np.mean(big**4) - 3.0) < 0.05


def test_philox_full_size_dump_and_replay_c3(pv):
    """The headline workload in production mode, end to end against the oracle: the device's Philox draws for C3
    (8000 per component, 50 components, D = 20) are dumped and replayed through the NumPy restatement of entmc_vbmc;
    the tensor-core kernel (automatic choice at this size) must reproduce H and dH on exactly those draws."""
    pr = syn.make_problem("C3")
    vp = make_vp(pv, pr.D, pr.K, pr.mu, pr.sigma, pr.lambd, pr.w, pr.eta)
    ctx = pv.Context(0)
    try:
        eps = ctx.philox_normals(pr.D, pr.K, pr.Ns_K, seed=4321, offset=3)
        assert eps.shape == (pr.K, pr.Ns_K // 2, pr.D)
        H, dH = ctx.entmc(vp, pr.Ns_K, (True,) * 4, True, seed=4321, offset=3)
        assert ctx.entmc_variant_used() == 5
    finally:
        ctx.close()
    Ho, dHo = eo.entmc(pr.vp.copy(), eps, (True,) * 4, True)
    assert relerr(H, Ho) < TOL_F32_VAL and relmax(dH, dHo) < TOL_F32_GRAD


# ------------------------------------------------------------------ GP log joint
@pytest.mark.parametrize("stem", REF_CASES)
def test_gplogjoint_golden(pv, stem):
    c = load_case(stem)
    g = c.g
    vp = case_vp(pv, c, "sa")
    G, dG, varG, dvarG, var_ss = pv._gp_log_joint(vp, c.gp, c.opt, True, True, False)
    assert relerr(G, g["gp_G"]) < TOL_F64 and relmax(dG, g["gp_dG"]) < TOL_F64
    assert varG is None and dvarG is None and var_ss == 0
    if c.S > 1:
        G, dG, *_ = pv._gp_log_joint(vp, c.gp, c.opt, False, True, False)
        assert G.shape == (c.S,) and dG.shape == g["gp_dG_noavg"].shape
        assert relmax(G, g["gp_G_noavg"]) < TOL_F64 and relmax(dG, g["gp_dG_noavg"]) < TOL_F64
    G, dG, *_ = pv._gp_log_joint(vp, c.gp, False, True, True, False)
    assert dG is None and relerr(G, g["gp_G"]) < TOL_F64


def test_matlab_gp_log_joint_and_neg_elcbo(pv):
    # pyvbmc/testing/vbmc/test_variational_optimization.py:120-211 (gradient / value parts)
    m = load_npz("matlab_vbmc")
    D = K = 2
    posts = gpp.posteriors(m["X"], m["y"], m["hyp"])
    gp = eo.make_gp(m["X"], posts)
    vp = make_vp(pv, D, K, m["mu"], 1e-3 * np.ones(K), np.ones(D), np.ones(K) / K, np.ones(K) / K)
    G, dG, varG, dvarG, var_ss = pv._gp_log_joint(vp, gp, True, True, True, False, False)
    assert np.allclose(dG, m["dG"]) and np.isclose(G, m["G"])
    theta = vp.get_parameters()
    F, dF, G, H, varF = pv._neg_elcbo(theta, gp, vp, 0.0, 0, True, False, None, 0.0, False)
    assert np.isclose(F, m["F"]) and np.isclose(G, m["G"]) and np.isclose(H, m["H"])
    assert np.allclose(dF, m["dF"])


# ------------------------------------------------------------------ negative ELCBO
@pytest.mark.parametrize("stem", REF_CASES)
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_negelcbo_mc_golden(pv, stem, prec):
    c = load_case(stem)
    g = c.g
    pv.config.precision = prec
    try:
        vp = case_vp(pv, c)
        eps = eps_for(0, c.K, c.Ns_K, c.D)
        F, dF, G, H, varF = pv._neg_elcbo(g["theta"], c.gp, vp, 0.0, c.Ns_K, True, False, c.theta_bnd, eps=eps)
    finally:
        pv.config.precision = "f32"
    tv, tg = (TOL_F32_VAL, TOL_F32_GRAD) if prec == "f32" else (TOL_F64, TOL_F64)
    assert relerr(F, g["mc_F"]) < tv and relerr(G, g["mc_G"]) < TOL_F64 and relerr(H, g["mc_H"]) < tv
    assert dF.shape == g["mc_dF"].shape and relmax(dF, g["mc_dF"]) < tg
    assert varF == 0
    for a, b in [(vp.mu, "post_mu"), (vp.sigma, "post_sigma"), (vp.lambd, "post_lambd"), (vp.w, "post_w"), (vp.eta, "post_eta")]:
        assert np.asarray(a).shape == g[b].shape and np.allclose(a, g[b], rtol=1e-13, atol=0)


@pytest.mark.parametrize("stem", REF_CASES)
def test_negelcbo_lb_golden(pv, stem):
    c = load_case(stem)
    g = c.g
    F, dF, G, H, varF = pv._neg_elcbo(g["theta2"], c.gp, case_vp(pv, c), 0.0, 0, True, False, c.theta_bnd)
    assert relerr(F, g["lb_F"]) < TOL_F64 and relerr(G, g["lb_G"]) < TOL_F64 and relerr(H, g["lb_H"]) < TOL_F64
    assert relmax(dF, g["lb_dF"]) < TOL_F64
    F, dF, *_ = pv._neg_elcbo(g["theta2"], c.gp, case_vp(pv, c), 0.0, 0, False, False, c.theta_bnd)
    assert dF is None and relerr(F, g["lbv_F"]) < TOL_F64


@pytest.mark.parametrize("stem", ["c2", "c3", "c4", "c2_noweights"])
def test_negelcbo_shifts_eta_of_the_callers_theta_in_place(pv, stem):
    """variational_optimization.py:1082-1085 mutates the CALLER's theta (eta block shifted to max == 0, vp.eta a
    view of it) and the bound loss reads the shifted eta: theta2 carries max(eta) ~ 5.5, so an unshifted bound loss
    would be off by orders of magnitude (lb_F is the unmodified reference's value)."""
    c = load_case(stem)
    g = c.g
    theta = g["theta2"].copy()
    vp = case_vp(pv, c)
    F, dF, *_ = pv._neg_elcbo(theta, c.gp, vp, 0.0, 0, True, False, c.theta_bnd)
    assert relerr(F, g["lb_F"]) < TOL_F64 and relmax(dF, g["lb_dF"]) < TOL_F64
    assert np.array_equal(theta, g["theta2_after"])
    if c.opt[3]:
        assert theta[-c.K:].max() == 0.0 and np.shares_memory(vp.eta, theta)
    theta = g["theta"].copy()
    pv._neg_elcbo(theta, c.gp, case_vp(pv, c), 0.0, c.Ns_K, True, False, c.theta_bnd, eps=eps_for(0, c.K, c.Ns_K, c.D))
    assert np.array_equal(theta, g["theta_after"])


def test_negelcbo_errors(pv):
    c = load_case("c1")
    with pytest.raises(ValueError):
        pv._neg_elcbo(c.g["theta2"], c.gp, case_vp(pv, c), 0.0, 0, True, False, None, 0.0, True)
    with pytest.raises(NotImplementedError):
        pv._neg_elcbo(c.g["theta2"], c.gp, case_vp(pv, c), 1.0, 0, True, None, None)
    with pytest.raises(NotImplementedError):
        pv._gp_log_joint(case_vp(pv, c), c.gp, True, True, True, True)
    with pytest.raises(NotImplementedError):
        pv._gp_log_joint(case_vp(pv, c), c.gp, False, True, True, 2)


def test_graph_replay_is_transparent(pv):
    """The hot-loop entry runs eagerly, then captures a CUDA graph, then replays it: all three must give
    bitwise the same answer for the same (theta, seed), follow theta and the seed, and survive a change of
    bounds / GP / draw count in between."""
    c = load_case("c2")
    g = c.g
    res = []
    for it in range(5):
        F, dF, G, H, _ = pv._neg_elcbo(g["theta"], c.gp, case_vp(pv, c), 0.0, 500, True, False, c.theta_bnd, seed=11)
        res.append((F, dF))
    for F, dF in res[1:]:
        assert F == res[0][0] and np.array_equal(dF, res[0][1])
    F2, dF2, *_ = pv._neg_elcbo(g["theta"], c.gp, case_vp(pv, c), 0.0, 500, True, False, c.theta_bnd, seed=12)
    assert F2 != res[0][0]
    th = g["theta"] + 1e-3
    F3, dF3, *_ = pv._neg_elcbo(th, c.gp, case_vp(pv, c), 0.0, 500, True, False, c.theta_bnd, seed=11)
    eps = pv.context_for_gp(c.gp).philox_normals(c.D, c.K, 500, seed=11, offset=0)
    Fo, dFo, *_ = eo.neg_elcbo(th, c.gp, c.vp(), 0.0, 500, True, False, c.theta_bnd, eps_half=eps)
    assert relerr(F3, Fo) < TOL_F32_VAL and relmax(dF3, dFo) < TOL_F32_GRAD
    # change the signature (no bounds, other Ns), come back: still the first answer
    pv._neg_elcbo(g["theta"], c.gp, case_vp(pv, c), 0.0, 300, True, False, None, seed=11)
    pv._neg_elcbo(g["theta"], c.gp, case_vp(pv, c), 0.0, 0, False, False, c.theta_bnd)
    for it in range(3):
        F, dF, *_ = pv._neg_elcbo(g["theta"], c.gp, case_vp(pv, c), 0.0, 500, True, False, c.theta_bnd, seed=11)
        assert F == res[0][0] and np.array_equal(dF, res[0][1])
    # another GP on the same shapes (fresh posterior arrays => repack => graphs invalidated)
    c2 = load_case("c2_ill")
    Fa, dFa, *_ = pv._neg_elcbo(g["theta"], c2.gp, case_vp(pv, c2), 0.0, 500, True, False, c2.theta_bnd, seed=11)
    for it in range(3):
        Fb, dFb, *_ = pv._neg_elcbo(g["theta"], c2.gp, case_vp(pv, c2), 0.0, 500, True, False, c2.theta_bnd, seed=11)
        assert Fb == Fa and np.array_equal(dFa, dFb)
    assert Fa != res[0][0]


def test_graph_staleness_is_per_context(pv):
    """Two GP contexts alive at once (the LRU holds 4): context A captures a graph, then A's buffers move (a much
    larger draw count through the eager path), then context B runs first.  A's captured graph must NOT be replayed
    with the freed pointers -- the reallocation notice cannot be swallowed by another context."""
    ca, cb = load_case("c2"), load_case("c2_ill")
    g = ca.g

    def eval_a(Ns, seed=11):
        return pv._neg_elcbo(g["theta"].copy(), ca.gp, case_vp(pv, ca), 0.0, Ns, True, False, ca.theta_bnd, seed=seed)

    def eval_b():
        return pv._neg_elcbo(g["theta"].copy(), cb.gp, case_vp(pv, cb), 0.0, 300, True, False, cb.theta_bnd, seed=11)

    first = [eval_a(200) for _ in range(4)]  # eager, capture, replay, replay
    eval_b(), eval_b(), eval_b()
    eval_a(60000)  # A's per-CTA record / tile buffers grow: every pointer captured above is stale
    Fb = eval_b()  # B looks at the reallocation notice first
    again = eval_a(200)
    assert again[0] == first[0][0] and np.array_equal(again[1], first[0][1])
    for _ in range(3):
        nxt = eval_a(200)
        assert nxt[0] == first[0][0] and np.array_equal(nxt[1], first[0][1])
    Fb2 = eval_b()
    assert Fb2[0] == Fb[0] and np.array_equal(Fb2[1], Fb[1])


# ------------------------------------------------------------------ variance path (full ELCBO evaluation)
@pytest.mark.parametrize("stem", VAR_CASES)
def test_variance_path_golden(pv, stem):
    """_eval_full_elcbo's call: value + variance + per-component terms (variational_optimization.py:474-485).
    varG = prior - explained is a cancellation; 1e-7 relative is the agreed fp64 noise floor between two
    orderings of the same triangular solves."""
    c = load_case(stem)
    g = c.g
    eps = eps_for(g["var_seed"], c.K, c.Ns_K, c.D)
    r = pv._neg_elcbo(g["theta2"], c.gp, case_vp(pv, c), 0.0, c.Ns_K, False, True, None, 0.0, True, eps=eps)
    F, dF, G, H, varF, dH, varG_ss, varG, varH, I_sk, J_sjk = r
    assert dF is None and dH is None and varH == 0
    assert relerr(F, g["var_F"]) < TOL_F32_VAL and relerr(G, g["var_G"]) < TOL_F64 and relerr(H, g["var_H"]) < TOL_F32_VAL
    assert relerr(varF, g["var_varF"]) < 1e-7 and relerr(varG, g["var_varG"]) < 1e-7
    if c.S > 1:
        assert relerr(varG_ss, g["var_varG_ss"]) < 1e-7
    else:
        assert varG_ss == 0
    assert I_sk.shape == (c.S, c.K) and relmax(I_sk, g["var_I_sk"]) < TOL_F64
    assert J_sjk.shape == (c.S, c.K, c.K) and relmax(J_sjk, g["var_J_sjk"]) < 1e-7
    G, dG, varG, dvarG, var_ss = pv._gp_log_joint(case_vp(pv, c, "sa"), c.gp, False, True, True, True)
    assert dG is None and dvarG is None
    assert relerr(G, g["gpv_G"]) < TOL_F64 and relerr(varG, g["gpv_varG"]) < 1e-7
    if c.S > 1:
        assert relerr(var_ss, g["gpv_var_ss"]) < 1e-7


def test_matlab_variance(pv):
    # pyvbmc/testing/vbmc/test_variational_optimization.py:120-149 and :165-199 (variance parts)
    m = load_npz("matlab_vbmc")
    D = K = 2
    posts = gpp.posteriors(m["X"], m["y"], m["hyp"])
    gp = eo.make_gp(m["X"], posts)
    vp = make_vp(pv, D, K, m["mu"], 1e-3 * np.ones(K), np.ones(D), np.ones(K) / K, np.ones(K) / K)
    G, dG, varG, dvarG, var_ss, I_sk, J_sjk = pv._gp_log_joint(vp, gp, False, True, True, True, True)
    assert np.isclose(G, m["G"]) and dG is None and dvarG is None
    assert np.isclose(varG, m["varG"]) and np.isclose(var_ss, m["var_ss"])
    assert I_sk.shape == (8, 2) and J_sjk.shape == (8, 2, 2)
    theta = vp.get_parameters()
    F, dF, G, H, varF, dH, varG_ss, varG, varH, I_sk, J_sjk = pv._neg_elcbo(theta, gp, vp, 0.0, 0, False, True, None, 0.0, True)
    assert np.isclose(F, m["F"]) and dF is None and np.isclose(G, m["G"]) and np.isclose(H, m["H"])
    assert np.isclose(varF, m["varG"]) and dH is None and np.isclose(varG, m["varG"]) and varH == 0.0


def test_variance_full_size_c3(pv):
    """N = 400, K = 50, S = 8 (10 MB of Cholesky factors): CUDA vs the oracle's triangular solves."""
    pr = syn.make_problem("C3")
    vp = make_vp(pv, pr.D, pr.K, pr.vp.mu, pr.vp.sigma, pr.vp.lambd, pr.vp.w, pr.vp.eta)
    G, dG, varG, dvarG, var_ss, I_sk, J_sjk = pv._gp_log_joint(vp, pr.gp, False, True, True, True, True)
    Go, _, varGo, _, var_sso, I_o, J_o = eo.gp_log_joint(pr.vp.copy(), pr.gp, False, True, True, True, True)
    assert relerr(G, Go) < TOL_F64 and relerr(varG, varGo) < 1e-7 and relerr(var_ss, var_sso) < 1e-7
    assert relmax(I_sk, I_o) < TOL_F64 and relmax(J_sjk, J_o) < 1e-7


# ------------------------------------------------------------------ full-size workloads vs the oracle
@pytest.mark.parametrize("name", ["C2", "C4", "C3"])
def test_full_size_against_oracle(pv, name):
    """BASELINE.json configs at their full draw counts: CUDA vs oracle on the same eps."""
    pr = syn.make_problem(name)
    eps = syn.draw_eps(pr.K, pr.Ns_K, pr.D, seed=0)
    Fo, dFo, Go, Ho, _ = eo.neg_elcbo(pr.theta, pr.gp, pr.vp.copy(), 0.0, pr.Ns_K, True, False, pr.theta_bnd, eps_half=eps)
    vp = make_vp(pv, pr.D, pr.K, pr.vp.mu, pr.vp.sigma, pr.vp.lambd, pr.vp.w, pr.vp.eta)
    F, dF, G, H, _ = pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, pr.Ns_K, True, False, pr.theta_bnd, eps=eps)
    assert relerr(F, Fo) < TOL_F32_VAL and relerr(H, Ho) < TOL_F32_VAL and relerr(G, Go) < TOL_F64
    assert relmax(dF, dFo) < TOL_F32_GRAD
    # per-block gradient parity (per-element relative error is ill-posed)
    D, K = pr.D, pr.K
    for lo, hi in [(0, D * K), (D * K, D * K + K), (D * K + K, D * K + K + D), (D * K + K + D, D * K + 2 * K + D)]:
        assert relmax(dF[lo:hi], dFo[lo:hi]) < TOL_F32_GRAD
    # bitwise determinism of the whole evaluation
    F2, dF2, *_ = pv._neg_elcbo(pr.theta, pr.gp, vp, 0.0, pr.Ns_K, True, False, pr.theta_bnd, eps=eps)
    assert F2 == F and np.array_equal(dF, dF2)


# ------------------------------------------------------------------ randomized shapes / edge cases vs the oracle
def _random_problem(rng, D, K, N, S, mean_kind="negquad", scale_spread=0.3):
    X = rng.normal(size=(N, D))
    y = -0.5 * np.sum(X**2, axis=1) + 0.1 * rng.normal(size=N)
    lay = gpp.hyp_layout(D, 1, mean_kind)
    hyps = np.zeros((S, lay["H"]))
    for s in range(S):
        h = hyps[s]
        h[:D] = rng.normal(0.0, 0.3, size=D)
        h[D] = rng.normal(0.5, 0.2)
        h[D + 1] = np.log(1e-2)
        b = lay["mean_start"]
        if mean_kind != "zero":
            h[b] = y.max()
        if mean_kind == "negquad":
            h[b + 1 : b + 1 + D] = 0.1 * rng.normal(size=D)
            h[b + 1 + D : b + 1 + 2 * D] = np.log(2.0) + 0.1 * rng.normal(size=D)
    posts = gpp.posteriors(X, y, hyps, mean_kind=mean_kind)
    gp = eo.make_gp(X, posts, mean_kind=mean_kind)
    mu = rng.normal(size=(D, K))
    sigma = np.exp(scale_spread * rng.normal(size=K))
    lambd = np.exp(0.3 * rng.normal(size=D))
    eta = rng.normal(size=K)
    w = np.exp(eta - eta.max())
    w /= w.sum()
    return gp, X, (mu, sigma, lambd, w, eta - eta.max())


SHAPES = [  # D, K, N, S, Ns_K
    (1, 1, 5, 1, 6), (1, 3, 9, 2, 7), (2, 2, 10, 8, 80), (3, 1, 12, 3, 33), (4, 7, 20, 2, 64), (5, 33, 40, 3, 20),
    (7, 5, 31, 1, 130), (12, 64, 50, 2, 10), (13, 17, 64, 5, 258), (20, 50, 97, 4, 100), (21, 9, 40, 2, 50),
    (24, 40, 33, 1, 40), (29, 3, 35, 2, 66), (32, 12, 64, 3, 34), (8, 130, 30, 1, 8), (6, 200, 25, 2, 4),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_random_shapes_against_oracle(pv, shape):
    D, K, N, S, Ns_K = shape
    rng = np.random.default_rng(1000 + D * 31 + K)
    for mean_kind, opt in (("negquad", (True,) * 4), ("const", (True, True, True, False)), ("zero", (True, True, False, True))):
        gp, X, (mu, sigma, lambd, w, eta) = _random_problem(rng, D, K, N, S, mean_kind)
        vo = eo.OracleVP.create(D, K, mu, sigma, lambd, w, eta, opt)
        theta = eo.get_parameters(vo)  # renormalises vo in place (lambda to unit RMS, sigma rescaled) ...
        mu, sigma, lambd, w, eta = vo.mu, vo.sigma, vo.lambd, vo.w, vo.eta  # ... so both sides start from that state
        bnd = eo.get_bounds(vo, X, syn.OPTIONS, K)
        theta = theta + 0.05 * rng.normal(size=theta.size)
        Ns_even = eo.even_ns(Ns_K)
        eps = rng.normal(size=(K, Ns_even // 2, D))
        Fo, dFo, Go, Ho, _ = eo.neg_elcbo(theta, gp, vo.copy(), 0.0, Ns_K, True, False, bnd, eps_half=eps)
        for prec, tv, tg in (("f32", TOL_F32_VAL, TOL_F32_GRAD), ("f64", TOL_F64, TOL_F64)):
            pv.config.precision = prec
            try:
                vp = make_vp(pv, D, K, mu, sigma, lambd, w, eta, opt)
                F, dF, G, H, _ = pv._neg_elcbo(theta, gp, vp, 0.0, Ns_K, True, False, bnd, eps=eps)
            finally:
                pv.config.precision = "f32"
            assert dF.shape == dFo.shape
            assert relerr(G, Go) < TOL_F64, (shape, mean_kind)
            assert abs(H - Ho) <= tv * max(abs(Ho), 1.0), (shape, mean_kind, prec, H, Ho)
            assert abs(F - Fo) <= tv * max(abs(Fo), 1.0), (shape, mean_kind, prec)
            assert relmax(dF, dFo) < tg, (shape, mean_kind, prec)
        # deterministic entropy + value-only + Philox smoke on the same shapes
        Fl, dFl, Gl, Hl, _ = eo.neg_elcbo(theta, gp, vo.copy(), 0.0, 0, True, False, bnd)
        F, dF, G, H, _ = pv._neg_elcbo(theta, gp, make_vp(pv, D, K, mu, sigma, lambd, w, eta, opt), 0.0, 0, True, False, bnd)
        assert relerr(F, Fl) < TOL_F64 and relmax(dF, dFl) < TOL_F64, (shape, mean_kind)
        F, dF, *_ = pv._neg_elcbo(theta, gp, make_vp(pv, D, K, mu, sigma, lambd, w, eta, opt), 0.0, Ns_K, False, False, bnd, eps=eps)
        assert dF is None and abs(F - Fo) <= TOL_F32_VAL * max(abs(Fo), 1.0)
        F, dF, *_ = pv._neg_elcbo(theta, gp, make_vp(pv, D, K, mu, sigma, lambd, w, eta, opt), 0.0, Ns_K, True, False, bnd, seed=5)
        assert np.isfinite(F) and np.all(np.isfinite(dF))


def test_disparate_scales_take_the_direct_path(pv):
    """Narrow components inside wide ones: the expanded distance would cancel, the conditioning guard must
    route those components through the direct-difference path and keep fp32 parity."""
    rng = np.random.default_rng(7)
    D, K = 6, 12
    mu = 0.3 * rng.normal(size=(D, K))
    sigma = np.array([1.0, 0.8, 1.2, 1e-2, 2e-2, 5e-3, 0.5, 1e-3, 0.3, 2.0, 1e-2, 0.7])
    lambd = np.exp(0.2 * rng.normal(size=D))
    eta = rng.normal(size=K)
    w = np.exp(eta - eta.max())
    w /= w.sum()
    vo = eo.OracleVP.create(D, K, mu, sigma, lambd, w, eta - eta.max())
    vp = make_vp(pv, D, K, mu, sigma, lambd, w, eta - eta.max())
    for Ns in (400, 70000):  # both fp32 kernels (small: fast, large: warp-autonomous)
        eps = np.random.default_rng(3).normal(size=(K, Ns // 2, D))
        Ho, dHo = eo.entmc(vo, eps, (True,) * 4, True)
        H, dH = pv.entmc_vbmc(vp, Ns, eps=eps)
        assert relerr(H, Ho) < TOL_F32_VAL and relmax(dH, dHo) < TOL_F32_GRAD, Ns
        H64, dH64 = pv.entropy_context().entmc(vp, Ns, eps=eps, precision="f64")
        assert relerr(H64, Ho) < TOL_F64 and relmax(dH64, dHo) < TOL_F64


def test_finite_difference_gradients(pv):
    """The reference's second kind of test (pyvbmc/testing/_check_grad.py): analytic gradient vs central
    differences of the value.

    * deterministic objective (Ns = 0: lower-bound entropy + log joint + soft bounds + weight penalty): every
      entry of dF is the exact derivative of F;
    * Monte-Carlo objective with the noise frozen (fixed Philox seed): only the WEIGHT block of dF is the exact
      derivative of the frozen-noise estimate.  For mu / sigma / lambda the reference's estimator
      (entmc_vbmc.py:84-108) keeps the reparameterisation path and drops the score term, whose expectation is
      zero but whose frozen-noise sample value is O(1/sqrt(Ns)) -- so those entries are compared with a
      statistical tolerance only."""
    c = load_case("c2")
    g = c.g
    theta0 = g["theta2"].copy()
    D, K = c.D, c.K
    rng = np.random.default_rng(0)

    def fd(f, i):
        h = 1e-5 * max(1.0, abs(theta0[i]))
        tp, tm = theta0.copy(), theta0.copy()
        tp[i] += h
        tm[i] -= h
        return (f(tp) - f(tm)) / (2 * h)

    pv.config.precision = "f64"
    try:
        # deterministic objective
        def f0(th):
            return pv._neg_elcbo(th, c.gp, case_vp(pv, c), 0.0, 0, False, False, c.theta_bnd)[0]

        F, dF, *_ = pv._neg_elcbo(theta0, c.gp, case_vp(pv, c), 0.0, 0, True, False, c.theta_bnd)
        for i in rng.choice(theta0.size, size=25, replace=False):
            d = fd(f0, i)
            assert abs(d - dF[i]) <= 1e-5 * max(1.0, abs(dF[i])) + 1e-6 * np.abs(dF).max(), (i, d, dF[i])

        # Monte-Carlo objective, frozen noise
        Ns = 400

        def f1(th):
            return pv._neg_elcbo(th, c.gp, case_vp(pv, c), 0.0, Ns, False, False, c.theta_bnd, seed=21)[0]

        F, dF, *_ = pv._neg_elcbo(theta0, c.gp, case_vp(pv, c), 0.0, Ns, True, False, c.theta_bnd, seed=21)
        n_w = theta0.size - K
        for i in rng.choice(np.arange(n_w, theta0.size), size=8, replace=False):  # weights: exact
            d = fd(f1, i)
            assert abs(d - dF[i]) <= 1e-5 * max(1.0, abs(dF[i])) + 1e-6 * np.abs(dF).max(), (i, d, dF[i])
        idx = rng.choice(n_w, size=12, replace=False)  # mu / sigma / lambda: up to the dropped score term
        err = np.array([fd(f1, i) - dF[i] for i in idx])
        assert np.abs(err).max() <= 8.0 / np.sqrt(Ns * K) * max(1.0, np.abs(dF).max()), err
    finally:
        pv.config.precision = "f32"


# ---------------------------------------------------------------------------------------------------
# every fp32 entmc kernel variant (the automatic choice only exercises one per problem size)
@pytest.mark.parametrize("variant", [0, 2, 4, 5, 6])
@pytest.mark.parametrize("stem", ["c2", "c4", "c1", "c2_ill"])
def test_entmc_every_variant_against_oracle_and_f64(pv, variant, stem, monkeypatch):
    """Forced kernel variant (VBMC_ENTMC_VARIANT is read when a context is created): eps-input parity with
    the reference's golden values, and Philox-mode agreement with the all-fp64 kernel on identical draws.
    Variant 5 is the tcgen05 / TMEM kernel (it falls back to 4 for K > 64), variant 6 the eight-lanes-per-pair kernel the
    automatic choice takes at the reference's default draw counts."""
    c = load_case(stem)
    g = c.g
    vp = case_vp(pv, c, "sa")
    monkeypatch.setenv("VBMC_ENTMC_VARIANT", str(variant))
    ctx = pv.Context(0)
    try:
        eps = eps_for(g["ent_seed"], c.K, c.Ns_K, c.D)
        H, dH = ctx.entmc(vp, c.Ns_K, c.opt, True, eps=eps)
        assert relerr(H, g["ent_H"]) < TOL_F32_VAL and relmax(dH, g["ent_dH"]) < TOL_F32_GRAD
        assert ctx.entmc_variant_used() in ((variant, 4) if variant == 5 and c.K > 64 else (variant,))
        # production RNG: same Philox key -> same draws in every kernel
        for Ns in (2, 258, 4000):
            Hd, dHd = ctx.entmc(vp, Ns, (True,) * 4, True, seed=11, offset=3, precision="f64")
            Hs, dHs = ctx.entmc(vp, Ns, (True,) * 4, True, seed=11, offset=3)
            assert abs(Hs - Hd) <= TOL_F32_VAL * max(abs(Hd), 1.0), (variant, stem, Ns)
            assert relmax(dHs, dHd) < TOL_F32_GRAD, (variant, stem, Ns)
            # value-only instantiation (no gradient groups)
            H0, d0 = ctx.entmc(vp, Ns, (False,) * 4, True, seed=11, offset=3)
            assert d0.shape == (0,) and abs(H0 - Hd) <= TOL_F32_VAL * max(abs(Hd), 1.0)
    finally:
        ctx.close()


# ---------------------------------------------------------------------------------------------------
# batched sieve evaluation: one launch for all candidates of variational_optimization.py:775-787
@pytest.mark.parametrize("shape", [(3, 2, 40, 2, "negquad"), (6, 30, 200, 6, "negquad"), (10, 20, 200, 4, "const"),
                                   (4, 1, 60, 3, "zero"), (20, 50, 400, 8, "negquad")])
def test_sieve_batch_matches_oracle_and_single_calls(pv, shape):
    D, K, N, S, mean_kind = shape
    rng = np.random.default_rng(D * 1000 + K)
    gp, X = _random_problem(rng, D, K, N, S, mean_kind)[:2]
    B = 7
    cands, oracle_F = [], []
    bnd = None
    for b in range(B):
        mu = X[rng.integers(0, N, size=K)].T + 0.3 * rng.normal(size=(D, K))
        sigma = 0.3 * np.exp(0.5 * rng.normal(size=K))
        lambd = np.exp(0.3 * rng.normal(size=D))
        eta = 0.5 * rng.normal(size=K)
        w = np.exp(eta - eta.max())
        w /= w.sum()
        vo = eo.OracleVP.create(D, K, mu, sigma, lambd, w, eta, (True,) * 4)
        theta = eo.get_parameters(vo)
        if bnd is None:
            bnd = eo.get_bounds(vo, X, syn.OPTIONS, K)
        if b % 2:  # push some candidates outside the soft bounds
            theta = theta + 2.0 * rng.normal(size=theta.size)
            eo.set_parameters(vo, theta)
        Fo, _, Go, Ho, _ = eo.neg_elcbo(theta, gp, vo.copy(), 0.0, 0, False, False, bnd)
        oracle_F.append((Fo, Go, Ho))
        vp = make_vp(pv, D, K, vo.mu, vo.sigma, vo.lambd, vo.w, vo.eta)
        cands.append((vp, theta))
    F, G, H = pv.neg_elcbo_batch([c[0] for c in cands], gp, bnd, thetas=[c[1] for c in cands])
    assert F.shape == (B,)
    for b in range(B):
        Fo, Go, Ho = oracle_F[b]
        assert relerr(G[b], Go) < TOL_F64 and relerr(H[b], Ho) < TOL_F64, (shape, b)
        assert abs(F[b] - Fo) <= TOL_F64 * max(abs(Fo), 1.0), (shape, b, F[b], Fo)
        # ... and the one-at-a-time drop-in path (what the reference's loop would call)
        vp1 = make_vp(pv, D, K, cands[b][0].mu, cands[b][0].sigma, cands[b][0].lambd, cands[b][0].w, cands[b][0].eta)
        F1, dF1, G1, H1, _ = pv._neg_elcbo(cands[b][1], gp, vp1, 0.0, 0, False, False, bnd)
        assert dF1 is None and abs(F[b] - F1) <= TOL_F64 * max(abs(F1), 1.0)
    # empty batch, and candidates that disagree on K are rejected
    F0, _, _ = pv.neg_elcbo_batch([], gp, bnd)
    assert F0.shape == (0,)
    if K > 1:
        other = pv.VariationalPosterior(D, K - 1)
        with pytest.raises(ValueError):
            pv.neg_elcbo_batch([cands[0][0], other], gp, bnd)


# ---------------------------------------------------------------------------------------------------
# device-resident Adam: minimize_adam.py:61-145 around the ELBO objective
@pytest.mark.parametrize("case", ["c2_all", "c2_noweights", "c4_box", "c2_etashift"])
def test_device_adam_matches_host_loop(pv, case):
    """The CUDA-graph Adam loop against the oracle's restatement of the reference loop driving the
    parity-pinned one-at-a-time evaluation with the same Philox key (seed, offset0 + iteration)."""
    from oracle.minimize_adam_oracle import minimize_adam as adam_oracle

    pr = syn.make_problem("C4" if case.startswith("c4") else "C2")
    opt = (True, True, True, case != "c2_noweights")
    Ns_K, seed, off0 = 200, 77, 5

    def fresh_vp():
        vp = make_vp(pv, pr.D, pr.K, pr.mu, pr.sigma, pr.lambd, pr.w, pr.eta, opt)
        return vp

    vp_h = fresh_vp()
    theta0 = np.asarray(vp_h.get_parameters(), dtype=float)
    if case == "c2_etashift":
        # max(eta) far from 0: every objective call renormalises the iterate's eta block in place
        # (variational_optimization.py:1082-1085) and the bound loss reads the shifted eta
        theta0[-pr.K:] += 6.0
    bnd = vp_h.get_bounds(pr.gp.X, syn.OPTIONS, pr.K)
    kw = dict(tol_fun=1e-3, max_iter=90, master_min=0.001, master_max=0.05, master_decay=200)
    if case == "c4_box":
        kw.update(lb=theta0 - 0.02, ub=theta0 + 0.03, use_early_stopping=False, max_iter=45)
    it = {"i": 0}

    def f(theta):
        F, dF, *_ = pv._neg_elcbo(theta, pr.gp, vp_h, 0.0, Ns_K, True, False, bnd, seed=seed, offset=off0 + it["i"])
        it["i"] += 1
        return F, dF

    xo, yo, xto, yto, no = adam_oracle(f, theta0.copy(), **kw)
    vp_d = fresh_vp()
    x, y, x_tab, y_tab, n = pv.minimize_adam_elcbo(pr.gp, vp_d, theta0.copy(), Ns_K, bnd, seed=seed, offset=off0, **kw)
    assert n == no and x_tab.shape == xto.shape and y_tab.shape == yto.shape
    assert np.max(np.abs(y_tab - yto)) <= 1e-7 * max(1.0, np.max(np.abs(yto)))
    assert np.max(np.abs(x_tab - xto)) <= 1e-7 * max(1.0, np.max(np.abs(xto)))
    assert np.max(np.abs(x - xo)) <= 1e-7 * max(1.0, np.max(np.abs(xo))) and abs(y - yo) <= 1e-7 * max(1.0, abs(yo))
    if case == "c4_box":
        assert np.all(x_tab <= (theta0 + 0.03)[:, None] + 1e-15) and np.all(x_tab >= (theta0 - 0.02)[:, None] - 1e-15)
    if case == "c2_etashift":
        assert np.all(np.abs(x_tab[-pr.K:, :].max(axis=0)) < 0.2)
    # vp is left at the last evaluated iterate, as after the reference loop
    np.testing.assert_allclose(np.ravel(vp_d.sigma), np.ravel(vp_h.sigma), rtol=1e-6)
    np.testing.assert_allclose(vp_d.mu, vp_h.mu, rtol=1e-6, atol=1e-9)


# ---------------------------------------------------------------------------------------------------
# sizes beyond the headline config (kept LAST in this file: it is the only test that was written after the
# round's GPU budget ran out, so it has not been seen green on hardware yet)
@pytest.mark.parametrize("K,Ns_K", [(50, 16000), (64, 12000), (33, 10000)])
def test_entmc_large_draw_counts_tensor_core_vs_f64(pv, K, Ns_K, monkeypatch):
    """Twice the headline work per GPU (what ONE GPU evaluates when bench.py --gpus 2 is cross-checked) and the
    K limits of the tensor-core kernel (forced: the automatic choice only takes it for 49 <= K <= 64): 11+ tiles
    per CTA, chunks that straddle components, against the all-fp64 kernel on identical Philox draws."""
    D = 20
    rng = np.random.default_rng(K)
    mu = 0.5 * rng.normal(size=(D, K))
    sigma = 0.5 * np.exp(0.1 * rng.normal(size=K))
    lambd = np.ones(D)
    eta = 0.3 * rng.normal(size=K)
    w = np.exp(eta - eta.max())
    w /= w.sum()
    vp = make_vp(pv, D, K, mu, sigma, lambd, w, eta - eta.max())
    monkeypatch.setenv("VBMC_ENTMC_VARIANT", "5")
    ctx = pv.Context(0)
    try:
        Hd, dHd = ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=21, offset=2, precision="f64")
        Hs, dHs = ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=21, offset=2)
        assert ctx.entmc_variant_used() == 5
        assert abs(Hs - Hd) <= TOL_F32_VAL * max(abs(Hd), 1.0)
        assert relmax(dHs, dHd) < TOL_F32_GRAD
    finally:
        ctx.close()


@pytest.mark.parametrize("D,K,Ns_K,want", [(20, 32, 12500, 5), (20, 24, 16668, 5), (10, 32, 12500, 5), (6, 48, 8334, 5),
                                           (10, 20, 20000, 4), (6, 30, 13334, 4), (20, 16, 25000, 4),
                                           (20, 50, 28, 6), (20, 50, 128, 6), (20, 50, 130, 0), (10, 20, 22, 6), (2, 3, 2, 6),
                                           (32, 64, 100, 6)])
def test_entmc_auto_selection_matches_the_measured_crossover(pv, D, K, Ns_K, want, monkeypatch):
    """Automatic choice of the fp32 entropy kernel at ~400k draws (scripts/variant_sweep.py): the tensor-core kernel from
    K >= 32, and from K >= 24 at D >= 16; the CUDA-core kernels below; and at the reference's default draw counts (up to 128
    draws per component, scripts/small_sweep.py) the eight-lanes-per-pair kernel -- each against the all-fp64 kernel on
    identical Philox draws."""
    monkeypatch.delenv("VBMC_ENTMC_VARIANT", raising=False)
    rng = np.random.default_rng(K * 100 + D)
    mu = 0.5 * rng.normal(size=(D, K))
    sigma = 0.5 * np.exp(0.1 * rng.normal(size=K))
    eta = 0.3 * rng.normal(size=K)
    w = np.exp(eta - eta.max())
    w /= w.sum()
    vp = make_vp(pv, D, K, mu, sigma, np.ones(D), w, eta - eta.max())
    ctx = pv.Context(0)
    try:
        Hd, dHd = ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=9, offset=1, precision="f64")
        Hs, dHs = ctx.entmc(vp, Ns_K, (True,) * 4, True, seed=9, offset=1)
        assert ctx.entmc_variant_used() == want
        assert abs(Hs - Hd) <= TOL_F32_VAL * max(abs(Hd), 1.0)
        assert relmax(dHs, dHd) < TOL_F32_GRAD
    finally:
        ctx.close()


def test_noise_prefetch_is_only_used_for_its_own_key(pv):
    """vbmc_noise_prefetch generates the draws of ONE key ahead of the call.  A call with another key (or another draw
    count) must ignore them and give exactly what it gives without any prefetch; the stale side-stream work must not leak
    into a later graph capture (drop-in call, device Adam)."""
    pr = syn.make_problem("C3")
    Ns_K = 6000  # smallest draw count that takes the tensor-core kernel (the only one with separate noise tiles)

    def fresh_vp():
        return make_vp(pv, pr.D, pr.K, pr.mu, pr.sigma, pr.lambd, pr.w, pr.eta)

    def call(seed, Ns=Ns_K):
        return pv._neg_elcbo(pr.theta.copy(), pr.gp, fresh_vp(), 0.0, Ns, True, False, pr.theta_bnd, seed=seed)

    base = [call(5) for _ in range(3)]  # eager, capture, replay: each prefetches its own key
    ctx = pv.context_for_gp(pr.gp)
    assert ctx.entmc_variant_used() == 5
    for F, dF, *_ in base[1:]:
        assert F == base[0][0] and np.array_equal(dF, base[0][1])
    ctx.noise_prefetch(pr.D, pr.K, Ns_K, 777, 0)       # wrong key
    again = call(5)
    assert again[0] == base[0][0] and np.array_equal(again[1], base[0][1])
    ctx.noise_prefetch(pr.D, pr.K, Ns_K // 2, 5, 0)    # right key, wrong draw count
    again = call(5)
    assert again[0] == base[0][0] and np.array_equal(again[1], base[0][1])
    other = call(6)
    assert other[0] != base[0][0]
    # a stale prefetch in front of the device Adam loop (captures a graph of iteration pairs)
    ctx.noise_prefetch(pr.D, pr.K, Ns_K, 999, 0)
    vp_a = fresh_vp()
    th0 = np.asarray(vp_a.get_parameters(), dtype=float)
    kw = dict(seed=3, max_iter=8, use_early_stopping=False, master_max=0.01)
    x1, y1, xt1, yt1, n1 = pv.minimize_adam_elcbo(pr.gp, vp_a, th0.copy(), Ns_K, pr.theta_bnd, **kw)
    x2, y2, xt2, yt2, n2 = pv.minimize_adam_elcbo(pr.gp, fresh_vp(), th0.copy(), Ns_K, pr.theta_bnd, **kw)
    assert n1 == n2 == 8 and np.array_equal(yt1, yt2) and np.array_equal(xt1, xt2)


def test_device_adam_split_phase_equals_one_shot(pv):
    """vbmc_adam_enqueue / vbmc_adam_fetch (one batch in flight ahead of the host) return bit for bit what
    vbmc_adam_steps returns; a speculative batch issued past the stopping point changes nothing that is read; the root-forked
    generator of a host-buffer call in between does not disturb the loop's look-ahead bookkeeping."""
    pr = syn.make_problem("C2")
    Ns_K = 400

    def fresh_vp():
        return make_vp(pv, pr.D, pr.K, pr.mu, pr.sigma, pr.lambd, pr.w, pr.eta)

    from pyvbmc_b200.vbmc.variational_optimization import _pack_params

    def init(ctx, max_iter):
        vp = fresh_vp()
        th0 = np.asarray(vp.get_parameters(), dtype=float)
        vp.set_parameters(th0)
        optimize = (True, True, True, True)
        ub = ctx.set_bounds(pr.theta_bnd)
        prm = np.zeros(ctx.param_len(pr.D, pr.K))
        _pack_params(prm, vp, th0, optimize, ub)
        ctx.adam_init(pr.D, pr.K, prm, th0, optimize, Ns_K, ub, 11, 0, None, None, max_iter, 0.001, 0.02, 200)

    ctx = pv.context_for_gp(pr.gp, need_L=False)
    init(ctx, 64)
    y_ref, x_ref = ctx.adam_steps(47)
    init(ctx, 64)
    ctx.adam_enqueue(20)
    ctx.adam_enqueue(20)                       # second batch in flight before the first one is read
    y0, x0 = ctx.adam_fetch(0, 20)
    ctx.adam_enqueue(7)
    y1, x1 = ctx.adam_fetch(20, 20)
    ctx.adam_enqueue(10)                       # speculative: never read
    y2, x2 = ctx.adam_fetch(40, 7)
    assert np.array_equal(np.concatenate([y0, y1, y2]), y_ref)
    assert np.array_equal(np.concatenate([x0, x1, x2]), x_ref)
    with pytest.raises(Exception):
        ctx.adam_fetch(50, 20)                 # not issued yet
    # a host-buffer evaluation behind the speculative batch is ordered after it and unaffected by it
    a = pv._neg_elcbo(pr.theta.copy(), pr.gp, fresh_vp(), 0.0, Ns_K, True, False, pr.theta_bnd, seed=5)
    b = pv._neg_elcbo(pr.theta.copy(), pr.gp, fresh_vp(), 0.0, Ns_K, True, False, pr.theta_bnd, seed=5)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
