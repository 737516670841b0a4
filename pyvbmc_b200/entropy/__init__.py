"""Mirror of ``pyvbmc.entropy`` (the package attributes ARE the functions,
pyvbmc/entropy/__init__.py:1-2)."""
from .entlb_vbmc import entlb_vbmc
from .entmc_vbmc import entmc_vbmc

__all__ = ["entlb_vbmc", "entmc_vbmc"]
