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


def install():
    """Patch an importable ``pyvbmc``; returns the list of ``module.name`` sites rebound."""
    done = []
    for modname, names in _SITES.items():
        mod = sys.modules.get(modname) or importlib.import_module(modname)
        for name, fn in names.items():
            if hasattr(mod, name):
                _saved.setdefault((modname, name), getattr(mod, name))
                setattr(mod, name, fn)
                done.append(f"{modname}.{name}")
    return done


def uninstall():
    for (modname, name), fn in list(_saved.items()):
        setattr(sys.modules[modname], name, fn)
        del _saved[(modname, name)]
