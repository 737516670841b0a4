"""Mirror of the hot-path part of ``pyvbmc.vbmc``."""
from .variational_optimization import _gp_log_joint, _neg_elcbo, _soft_bound_loss, _vp_bound_loss, neg_elcbo_batch

__all__ = ["_gp_log_joint", "_neg_elcbo", "_soft_bound_loss", "_vp_bound_loss", "neg_elcbo_batch"]
