"""Drop-in installation: rebind the reference's hot-path names to the B200 implementations.

The reference has no plugin registry; callers import the functions BY NAME, so every binding
site is patched (what the reference's own tests do with ``mocker.patch``,
pyvbmc/testing/vbmc/test_vbmc_finalboost.py:42-44):

  pyvbmc.entropy.{entmc_vbmc, entlb_vbmc}                       (pyvbmc/entropy/__init__.py:1-2)
  pyvbmc.vbmc.variational_optimization.{entmc_vbmc, entlb_vbmc, _gp_log_joint, _neg_elcbo,
                                        _vp_bound_loss, _soft_bound_loss}
                                                               (variational_optimization.py:10, uses
                                                                at :207,239,474,777,1121-1168)
  pyvbmc.vbmc.active_sample.{_gp_log_joint, _neg_elcbo}         (active_sample.py:18-22, :331,:637)

``optimize_vp``, ``_sieve``, ``_eval_full_elcbo``, ``minimize_adam`` and ``VBMC.optimize()``
then run unchanged on top.

``install(device_adam=True)`` additionally rebinds ``variational_optimization.minimize_adam`` to a wrapper
that recognises the reference's ELBO closure (``vb_train_mc_fun``, variational_optimization.py:238-249: free
variables ``gp, vp0, elcbo_beta, ns_ent_K, compute_var, theta_bnd`` around a call of ``_neg_elcbo``) and runs
the device-resident loop ``minimize_adam_elcbo`` for it; any other objective, a non-default configuration
(``elcbo_beta != 0``, ``compute_var``, deterministic entropy) or a non-finite objective goes to the
reference's own ``minimize_adam``.  ``install(batched_sieve=True)`` rebinds ``variational_optimization._sieve``
to a wrapper that lets the reference's ``_sieve`` do everything except its candidate loop (:775-787), which
becomes one call of ``neg_elcbo_batch`` (see :func:`make_sieve`).
"""
from __future__ import annotations

import importlib
import sys
import threading

from .entropy import entlb_vbmc, entmc_vbmc
from .vbmc.variational_optimization import _gp_log_joint, _neg_elcbo, _soft_bound_loss, _vp_bound_loss

_SITES = {
    "pyvbmc.entropy": {"entmc_vbmc": entmc_vbmc, "entlb_vbmc": entlb_vbmc},
    "pyvbmc.vbmc.variational_optimization": {
        "entmc_vbmc": entmc_vbmc,
        "entlb_vbmc": entlb_vbmc,
        "_gp_log_joint": _gp_log_joint,
        "_neg_elcbo": _neg_elcbo,
        "_vp_bound_loss": _vp_bound_loss,
        "_soft_bound_loss": _soft_bound_loss,
    },
    "pyvbmc.vbmc.active_sample": {"_gp_log_joint": _gp_log_joint, "_neg_elcbo": _neg_elcbo},
}
_saved = {}
# make_sieve swaps a module-level name while the reference's _sieve runs: one sieve at a time per process
_sieve_lock = threading.RLock()

_ELCBO_FREEVARS = ("gp", "vp0", "elcbo_beta", "ns_ent_K", "compute_var", "theta_bnd")


def elcbo_closure_ingredients(f):
    """``(gp, vp0, ns_ent_K, theta_bnd)`` if ``f`` is the reference's ``vb_train_mc_fun`` closure in its default
    configuration (stochastic entropy, no variance term), else ``None``."""
    code, cells = getattr(f, "__code__", None), getattr(f, "__closure__", None)
    if code is None or not cells or "_neg_elcbo" not in code.co_names:
        return None
    free = dict(zip(code.co_freevars, cells))
    if any(n not in free for n in _ELCBO_FREEVARS):
        return None
    try:
        v = {n: free[n].cell_contents for n in _ELCBO_FREEVARS}
    except ValueError:  # empty cell
        return None
    if v["elcbo_beta"] != 0 or v["compute_var"] or not (v["ns_ent_K"] and v["ns_ent_K"] > 0):
        return None
    return v["gp"], v["vp0"], v["ns_ent_K"], v["theta_bnd"]


def make_minimize_adam(reference_minimize_adam, device_loop=None):
    """The rebound ``minimize_adam``: device-resident loop for the ELBO closure, the reference's loop otherwise."""

    def minimize_adam(f, x0, lb=None, ub=None, tol_fun=0.001, max_iter=2000, master_min=0.001, master_max=0.1,
                      master_decay=200, use_early_stopping=True):
        ing = elcbo_closure_ingredients(f)
        if ing is not None:
            loop = device_loop
            if loop is None:
                from .vbmc.minimize_adam import minimize_adam_elcbo as loop
            gp, vp0, ns_ent_K, theta_bnd = ing
            try:
                return loop(gp, vp0, x0, ns_ent_K, theta_bnd, lb, ub, tol_fun, max_iter, master_min, master_max,
                            master_decay, use_early_stopping)
            except FloatingPointError:
                pass  # non-finite objective inside the graph: the host loop re-evaluates in fp64 where needed
        return reference_minimize_adam(f, x0, lb, ub, tol_fun, max_iter, master_min, master_max, master_decay,
                                       use_early_stopping)

    minimize_adam.__wrapped__ = reference_minimize_adam
    return minimize_adam


def make_sieve(reference_sieve, module, batch_fn=None):
    """The rebound ``_sieve`` (variational_optimization.py:660-809): the reference's own function does everything
    (options, soft bounds, candidate generation, return values); only its candidate loop (:775-787) is collapsed
    into ONE batched launch.  While the reference function runs, the module's ``_neg_elcbo`` is a recorder that
    notes ``(theta, vp0)`` and returns increasing placeholders, so the reference's ``argsort`` leaves the
    candidates in creation order; the batched values are then computed and the same ``argsort`` applied here.
    Any call that is not the default sieve evaluation (stochastic entropy, variance term, gradients) is passed
    to the real ``_neg_elcbo`` and nothing is re-ordered."""
    import numpy as np

    def _sieve(*args, **kwargs):
        real = None
        rec, mode = [], {"batch": None}

        def recorder(theta, gp, vp, beta=0.0, Ns=0, compute_grad=True, compute_var=None, theta_bnd=None, *a, **k):
            plain = Ns == 0 and not compute_grad and not compute_var and beta == 0 and not a and not k
            if mode["batch"] is None:
                mode["batch"] = bool(plain)
            if not (mode["batch"] and plain):
                mode["batch"] = False
                return real(theta, gp, vp, beta, Ns, compute_grad, compute_var, theta_bnd, *a, **k)
            rec.append((np.array(theta, dtype=float, copy=True), gp, vp, theta_bnd))
            return float(len(rec)), None, 0.0, 0.0, 0.0

        with _sieve_lock:  # (re-read under the lock: another thread's sieve may just have restored the name)
            real = module._neg_elcbo
            module._neg_elcbo = recorder
            try:
                out = reference_sieve(*args, **kwargs)
            finally:
                module._neg_elcbo = real
        if not rec:
            return out
        vp0_vec, vp0_type = out[0], out[1]
        fn = batch_fn
        if fn is None:
            from .vbmc.variational_optimization import neg_elcbo_batch as fn
        same = mode["batch"] and len(rec) == len(vp0_vec) and all(r[2] is v for r, v in zip(rec, vp0_vec))
        if same:
            F, _, _ = fn([r[2] for r in rec], rec[0][1], rec[0][3], thetas=[r[0] for r in rec])
        else:  # (never seen: the loop changed shape) evaluate one at a time, in the order the reference returned
            F = np.array([real(r[0], r[1], r[2], 0, 0, 0, False, r[3])[0] for r in rec])
            vp0_vec = np.array([r[2] for r in rec], dtype=object)
        order = np.argsort(np.asarray(F))  # :790
        return (vp0_vec[order], np.asarray(vp0_type)[order]) + tuple(out[2:])

    _sieve.__wrapped__ = reference_sieve
    return _sieve


def make_pdf(reference_pdf):
    """The rebound ``VariationalPosterior.pdf``: the Gaussian mixture (``df`` infinite or 0) is evaluated on the
    device (``vbmc_vp_pdf``), the heavy-tailed variants go to the reference's own method."""
    import numpy as np

    def pdf(self, x, orig_flag=True, log_flag=False, grad_flag=False, df=np.inf):
        if np.isfinite(df) and df != 0:
            return reference_pdf(self, x, orig_flag=orig_flag, log_flag=log_flag, grad_flag=grad_flag, df=df)
        from .acquisition_functions import vp_pdf

        return vp_pdf(self, x, orig_flag=orig_flag, log_flag=log_flag, grad_flag=grad_flag, df=df)

    pdf.__wrapped__ = reference_pdf
    return pdf


def install(device_adam=False, batched_sieve=False, device_pdf=False):
    """Patch an importable ``pyvbmc``; returns the list of ``module.name`` sites rebound."""
    done = []
    for modname, names in _SITES.items():
        mod = sys.modules.get(modname) or importlib.import_module(modname)
        for name, fn in names.items():
            if hasattr(mod, name):
                _saved.setdefault((modname, name), getattr(mod, name))
                setattr(mod, name, fn)
                done.append(f"{modname}.{name}")
    if device_adam:
        modname = "pyvbmc.vbmc.variational_optimization"
        mod = sys.modules.get(modname) or importlib.import_module(modname)
        if hasattr(mod, "minimize_adam") and (modname, "minimize_adam") not in _saved:
            _saved[(modname, "minimize_adam")] = mod.minimize_adam
            mod.minimize_adam = make_minimize_adam(mod.minimize_adam)
            done.append(f"{modname}.minimize_adam")
    if batched_sieve:
        modname = "pyvbmc.vbmc.variational_optimization"
        mod = sys.modules.get(modname) or importlib.import_module(modname)
        if hasattr(mod, "_sieve") and (modname, "_sieve") not in _saved:
            _saved[(modname, "_sieve")] = mod._sieve
            mod._sieve = make_sieve(mod._sieve, mod)
            done.append(f"{modname}._sieve")
    if device_pdf:  # acquisition functions call vp.pdf(Xs, orig_flag=False, log_flag=True) (acq_fcn_log.py:38-42)
        modname = "pyvbmc.variational_posterior.variational_posterior"
        mod = sys.modules.get(modname) or importlib.import_module(modname)
        cls = mod.VariationalPosterior
        if ("class:" + modname, "pdf") not in _saved_attrs:
            _saved_attrs[("class:" + modname, "pdf")] = (cls, cls.pdf)
            cls.pdf = make_pdf(cls.pdf)
            done.append(f"{modname}.VariationalPosterior.pdf")
    return done


_saved_attrs = {}


def uninstall():
    for (modname, name), fn in list(_saved.items()):
        setattr(sys.modules[modname], name, fn)
        del _saved[(modname, name)]
    for key, (owner, value) in list(_saved_attrs.items()):
        setattr(owner, key[1], value)
        del _saved_attrs[key]
