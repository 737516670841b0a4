"""B200 replacement for ``pyvbmc.entropy.entlb_vbmc`` (pyvbmc/entropy/entlb_vbmc.py:6-180)."""
from __future__ import annotations

from ..context import entropy_context


def entlb_vbmc(vp, grad_flags=tuple([True] * 4), jacobian_flag=True, *, _ctx=None):
    """Entropy lower bound of the variational posterior (Jensen) and its gradient.

    Same signature and return convention as the reference: ``(H, dH)``."""
    ctx = _ctx if _ctx is not None else entropy_context()
    return ctx.entlb(vp, grad_flags, jacobian_flag)
