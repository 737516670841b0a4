"""pyvbmc_b200 -- B200-native (sm_100a) implementation of PyVBMC's ELBO inner loop.

Public surface (mirrors the reference's names for this path only):

  pyvbmc_b200.entropy.entmc_vbmc / entlb_vbmc
  pyvbmc_b200.vbmc.variational_optimization._neg_elcbo / _gp_log_joint / _vp_bound_loss / _soft_bound_loss
  pyvbmc_b200.variational_posterior.VariationalPosterior   (theta packing + soft bounds)
  pyvbmc_b200.install() / uninstall()                       (rebinds the names inside an installed pyvbmc)
  pyvbmc_b200.gp_predict / vp_pdf / AcqFcnLog                (acquisition-function ingredients)

All arithmetic runs in ``csrc/libvbmc_b200.so`` (C ABI: ``include/vbmc_b200.h``).  There is no
CPU fallback: without the built library or without a CUDA device, compute calls raise.
"""
from . import _capi
from .config import config
from .context import Context, clear_caches, context_for_gp, entropy_context
from .entropy import entlb_vbmc, entmc_vbmc
from .acquisition_functions import AcqFcnLog, gp_predict, total_variance, vp_pdf
from .install import install, uninstall
from .variational_posterior import VariationalPosterior
from .vbmc.minimize_adam import minimize_adam_elcbo
from .vbmc.variational_optimization import _gp_log_joint, _neg_elcbo, _soft_bound_loss, _vp_bound_loss, neg_elcbo_batch

__all__ = [
    "config",
    "Context",
    "context_for_gp",
    "entropy_context",
    "clear_caches",
    "entmc_vbmc",
    "entlb_vbmc",
    "_neg_elcbo",
    "neg_elcbo_batch",
    "minimize_adam_elcbo",
    "_gp_log_joint",
    "_vp_bound_loss",
    "_soft_bound_loss",
    "VariationalPosterior",
    "install",
    "uninstall",
    "gp_predict",
    "vp_pdf",
    "total_variance",
    "AcqFcnLog",
]
__version__ = "0.1.0"
