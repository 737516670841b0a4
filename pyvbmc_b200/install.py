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
reference's own ``minimize_adam``.
"""
from __future__ import annotations

import importlib
import sys

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


def install(device_adam=False):
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
    return done


def uninstall():
    for (modname, name), fn in list(_saved.items()):
        setattr(sys.modules[modname], name, fn)
        del _saved[(modname, name)]
