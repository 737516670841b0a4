"""Regenerate ``tests/golden/*.npz`` -- TEST INFRASTRUCTURE ONLY.

Run in the build container (needs ``/root/reference``):

    python -m oracle.make_golden

Two kinds of fixtures are written:

``matlab_*.npz``  the reference's own MATLAB known-answer data for the hot path,
                  re-exported verbatim from ``pyvbmc/testing/{vbmc,entropy,
                  variational_posterior}`` together with the golden scalars that
                  the reference tests hold inline
                  (test_variational_optimization.py:120-211, test_entlb_vbmc.py:
                  101-119, test_entmc_vbmc.py:140-171, test_variational_posterior.py:
                  761-791).
``ref_*.npz``     inputs and outputs of the UNMODIFIED reference functions
                  (imported through ``oracle.ref_loader``) on seeded synthetic
                  problems (``oracle.synthetic``), reduced draw counts.
"""
from __future__ import annotations

import os

import numpy as np
from scipy.io import loadmat

from . import elbo_oracle as eo
from . import gp_posterior as gpp
from . import ref_loader
from . import synthetic as syn

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TESTING = os.path.join(ref_loader.REFERENCE_ROOT, "pyvbmc", "testing")


def _txt(*parts):
    return np.loadtxt(os.path.join(TESTING, *parts), delimiter=",")


def export_matlab():
    np.savez_compressed(
        os.path.join(OUT, "matlab_vbmc.npz"),
        X=_txt("vbmc", "X.txt"),
        y=_txt("vbmc", "y.txt"),
        hyp=_txt("vbmc", "hyp.txt"),
        mu=_txt("vbmc", "mu.txt"),
        dF=_txt("vbmc", "dF.txt"),
        dG=_txt("vbmc", "dG_gp_log_joint.txt"),
        # inline goldens of test_gp_log_joint / test_neg_elcbo
        G=-0.461812484952867,
        varG=6.598768992700180e-05,
        var_ss=1.031705745662353e-04,
        F=11.746298071422430,
        H=-11.284485586469563,
        # inline goldens of test_vp_bound_loss (theta[-1] = 1.0, tol_con = 0.01)
        bound_L=178.1123635679098,
        bound_dL_last=356.2247271358195,
    )
    mat = loadmat(os.path.join(TESTING, "entropy", "entropy-test.mat"))
    g = lambda n: mat["vp"][n].item().astype(float)  # noqa: E731
    np.savez_compressed(
        os.path.join(OUT, "matlab_entropy.npz"),
        D=mat["D"].item(),
        K=mat["K"].item(),
        Ns=mat["Ns"].item(),
        jacobian_flag=mat["jacobian_flag"].item(),
        mu=g("mu"),
        sigma=g("sigma"),
        lambd=g("lambda"),
        w=g("w"),
        eta=g("eta"),
        H=mat["H"].item(),
        dH=mat["dH"].squeeze(),
        Hl=mat["Hl"].item(),
        dHl=mat["dHl"].squeeze(),
    )
    np.savez_compressed(
        os.path.join(OUT, "matlab_bounds.npz"),
        X=_txt("variational_posterior", "X.txt"),
        mu=_txt("variational_posterior", "mu.txt"),
        lb=_txt("variational_posterior", "bnd_lb.txt"),
        ub=_txt("variational_posterior", "bnd_ub.txt"),
        tol_con=0.01,
        weight_threshold=0.125,
        weight_penalty=0.1,
    )


# (file stem, synthetic config, make_problem kwargs, draws per component, extras)
CASES = [
    ("c1", "C1", {}, 80, dict(var=True)),
    ("c2", "C2", {}, 40, dict()),
    ("c2_noweights", "C2", dict(optimize=(True, True, True, False)), 40, dict()),
    ("c2_ill", "C2", dict(ill_conditioned=True), 40, dict()),
    ("c2_const", "C2", dict(mean_kind="const"), 20, dict()),
    ("c2_zero", "C2", dict(mean_kind="zero"), 20, dict()),
    ("c2_var", "C2", dict(N=64), 20, dict(var=True)),
    ("c2_var_s1", "C2", dict(N=48, S=1), 20, dict(var=True)),
    ("c2_lownoise", "C2", dict(N=48), 20, dict(var=True, force_low_noise=True)),
    ("c3", "C3", {}, 40, dict()),
    ("c4", "C4", {}, 40, dict()),
    ("c4_var", "C4", dict(N=60), 20, dict(var=True)),
]


def _perturbed_theta(pr, seed=5, violate=True):
    rng = np.random.default_rng(seed)
    theta = pr.theta + 0.02 * rng.normal(size=pr.theta.size)
    if violate:
        D, K = pr.D, pr.K
        theta[1] += 12.0  # a mean outside its soft bound
        if pr.vp.optimize_sigma:
            theta[D * K + (3 % K)] += 6.0  # ln sigma above the ln-scale bound
        if pr.vp.optimize_lambd and pr.vp.optimize_sigma:
            theta[D * K + K + (1 % D)] -= 16.0  # ln lambda below the ln-scale bound
        if pr.vp.optimize_weights:
            theta[-2] -= 9.0  # eta below ln(tol_weight / 2)
    return theta


def export_case(stem, cfg, kw, Ns_K, extras):
    ref = ref_loader.load()
    pr = syn.make_problem(cfg, **kw)
    if extras.get("force_low_noise"):
        pr.posts = gpp.posteriors(pr.X, pr.y, pr.hyps, s2=pr.s2, mean_kind=pr.mean_kind, force_low_noise=True)
        pr.gp = eo.make_gp(pr.X, pr.posts, mean_kind=pr.mean_kind, y=pr.y)
    D, K, S, N = pr.D, pr.K, pr.S, pr.N
    opt = (pr.vp.optimize_mu, pr.vp.optimize_sigma, pr.vp.optimize_lambd, pr.vp.optimize_weights)
    rgp = ref_loader.make_ref_gp(pr.X, pr.y, pr.posts, pr.mean_kind)

    def fresh_vp():
        return ref_loader.make_ref_vp(D, K, pr.vp.mu, pr.vp.sigma, pr.vp.lambd, pr.vp.w, pr.vp.eta, opt)

    out = dict(
        cfg=cfg,
        D=D,
        K=K,
        S=S,
        N=N,
        Ns_K=Ns_K,
        mean_kind=pr.mean_kind,
        optimize=np.array(opt),
        X=pr.X,
        y=pr.y,
        hyps=pr.hyps,
        alpha=np.stack([p["alpha"] for p in pr.posts]),
        L_chol=np.array([p["L_chol"] for p in pr.posts]),
        sW0=np.array([p["sW"][0] for p in pr.posts]),
        vp_mu=pr.vp.mu,
        vp_sigma=pr.vp.sigma,
        vp_lambd=pr.vp.lambd,
        vp_w=pr.vp.w,
        vp_eta=pr.vp.eta,
        bnd_lb=pr.theta_bnd["lb"],
        bnd_ub=pr.theta_bnd["ub"],
        bnd_tol_con=pr.theta_bnd["tol_con"],
        bnd_weight_threshold=pr.theta_bnd.get("weight_threshold", np.nan),
        bnd_weight_penalty=pr.theta_bnd.get("weight_penalty", np.nan),
        eps_seed=0,
    )
    if extras.get("var"):
        out["L"] = np.stack([p["L"] for p in pr.posts])

    # (1) the Adam objective: value + gradient, MC entropy, soft bounds  (:238-249)
    theta = _perturbed_theta(pr)
    # a COPY: the reference shifts the eta block of its argument in place (:1082-1085), so the array handed in
    # comes back with max(eta) == 0; `theta` is the input as the caller built it, `theta_after` what the caller holds
    # afterwards
    out["theta"] = theta.copy()
    np.random.seed(0)
    vp = fresh_vp()
    F, dF, G, H, varF = ref._neg_elcbo(theta, rgp, vp, 0.0, Ns_K, True, False, pr.theta_bnd)
    out["theta_after"] = theta.copy()
    out.update(mc_F=F, mc_dF=dF, mc_G=G, mc_H=H, mc_varF=varF)
    out.update(post_mu=vp.mu, post_sigma=vp.sigma, post_lambd=vp.lambd, post_w=vp.w, post_eta=vp.eta)

    # (2) the BFGS / sieve objective: deterministic entropy, with and without gradient (:206-219, :777)
    theta2 = _perturbed_theta(pr, seed=6, violate=False)
    if opt[3]:
        theta2[-K:] += 5.5  # max(eta) far from 0: the bound loss must see the SHIFTED eta (:1082-1085, :1195-1209)
    out["theta2"] = theta2.copy()
    vp = fresh_vp()
    F, dF, G, H, varF = ref._neg_elcbo(theta2, rgp, vp, 0.0, 0, True, False, pr.theta_bnd)
    out["theta2_after"] = theta2.copy()
    out.update(lb_F=F, lb_dF=dF, lb_G=G, lb_H=H)
    vp = fresh_vp()
    F, dF, G, H, varF = ref._neg_elcbo(out["theta2"].copy(), rgp, vp, 0.0, 0, False, False, pr.theta_bnd)
    assert dF is None
    out.update(lbv_F=F)

    # (3) stand-alone pieces on the state left by (2)
    vp = fresh_vp()
    vp.set_parameters(theta2)
    if opt[3]:
        vp.eta = (theta2[-K:] - np.max(theta2[-K:])).reshape(1, -1)
    gf = opt
    G, dG, _, _, _ = ref._gp_log_joint(vp, rgp, gf, True, True, False)
    out.update(gp_G=G, gp_dG=dG)
    if S > 1:
        G, dG, _, _, _ = ref._gp_log_joint(vp, rgp, gf, False, True, False)
        out.update(gp_G_noavg=G, gp_dG_noavg=dG)
    np.random.seed(3)
    H, dH = ref.entmc_vbmc(vp, Ns_K, gf, True)
    out.update(ent_H=H, ent_dH=dH, ent_seed=3)
    np.random.seed(4)
    H, dH = ref.entmc_vbmc(vp, Ns_K, (True,) * 4, False)
    out.update(entnj_H=H, entnj_dH=dH, entnj_seed=4)
    np.random.seed(4)
    H, dH = ref.entmc_vbmc(vp, Ns_K, (False, False, False, True), True)
    out.update(entw_H=H, entw_dH=dH)
    H, dH = ref.entlb_vbmc(vp, gf, True)
    out.update(elb_H=H, elb_dH=dH)
    H, dH = ref.entlb_vbmc(vp, (True,) * 4, False)
    out.update(elbnj_H=H, elbnj_dH=dH)
    out.update(sa_mu=vp.mu, sa_sigma=vp.sigma, sa_lambd=vp.lambd, sa_w=vp.w, sa_eta=vp.eta)

    # (4) the full-ELCBO evaluation: variance + per-component terms (:474-485)
    if extras.get("var"):
        vp = fresh_vp()
        np.random.seed(1)
        r = ref._neg_elcbo(theta2, rgp, vp, 0.0, Ns_K, False, True, None, 0.0, True)
        F, dF, G, H, varF, dH, varG_ss, varG, varH, I_sk, J_sjk = r
        out.update(var_F=F, var_G=G, var_H=H, var_varF=varF, var_varG_ss=varG_ss, var_varG=varG, var_I_sk=I_sk, var_J_sjk=J_sjk, var_seed=1)
        vp = fresh_vp()
        vp.set_parameters(theta2)
        G, dG, varG, dvarG, var_ss = ref._gp_log_joint(vp, rgp, False, True, True, True)
        out.update(gpv_G=G, gpv_varG=varG, gpv_var_ss=var_ss)
    np.savez_compressed(os.path.join(OUT, f"ref_{stem}.npz"), **out)
    return out


def export_entropy_edge_cases():
    """Small stand-alone entropy cases: K = 1, 1-D inputs, odd Ns (rounded to even)."""
    ref = ref_loader.load()
    rng = np.random.default_rng(17)
    out = {}
    specs = [(3, 1, 9), (1, 2, 5), (2, 3, 7), (4, 3, 16), (7, 5, 33)]
    out["specs"] = np.array(specs)
    for i, (D, K, Ns) in enumerate(specs):
        mu = rng.normal(size=(D, K))
        sigma = np.exp(0.3 * rng.normal(size=K))
        lambd = np.exp(0.3 * rng.normal(size=D))
        eta = rng.normal(size=K)
        w = np.exp(eta) / np.exp(eta).sum()
        vp = ref_loader.make_ref_vp(D, K, mu, sigma, lambd, w, eta)
        np.random.seed(100 + i)
        H, dH = ref.entmc_vbmc(vp, Ns, (True,) * 4, True)
        Hl, dHl = ref.entlb_vbmc(vp, (True,) * 4, True)
        out.update({f"mu{i}": mu, f"sigma{i}": sigma, f"lambd{i}": lambd, f"w{i}": w, f"eta{i}": eta})
        out.update({f"H{i}": H, f"dH{i}": dH, f"Hl{i}": Hl, f"dHl{i}": dHl})
    np.savez_compressed(os.path.join(OUT, "ref_entropy_edge.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    export_matlab()
    export_entropy_edge_cases()
    for case in CASES:
        export_case(*case)
        print("wrote", case[0])


if __name__ == "__main__":
    main()
