"""Import the UNMODIFIED reference hot path from ``/root/reference`` (test infrastructure).

The reference (acerbilab/pyvbmc) is pure Python but imports several third-party
packages that are absent from this image (``gpyreg``, ``cma``, ``corner``,
``matplotlib``, ``imageio``, ``plotly``).  None of them is needed by the ELBO
inner loop: ``_gp_log_joint`` consumes the GP as opaque arrays plus three
``isinstance`` checks (``pyvbmc/vbmc/variational_optimization.py:1311-1398``).
We therefore pre-seed ``sys.modules`` with inert stubs, put the read-only
reference tree on ``sys.path`` and import the reference functions verbatim.

Where ``/root/reference`` does not exist (the GPU box) the same unmodified package is
taken from ``oracle/_ref`` (installed there by ``oracle/build_ref.py`` during
``__graft_entry__.build()``; git-ignored, shipped with the snapshot).  TEST / BENCH
INFRASTRUCTURE ONLY: the product never imports this module.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _is_ref(c):
    return bool(c) and os.path.isfile(os.path.join(c, "pyvbmc", "vbmc", "variational_optimization.py"))


def _find_root():
    """The reference checkout where it exists (build container); else the archive that ``oracle/build_ref.py`` packed
    into the git-ignored ``oracle/_ref`` (what travels to the GPU box), unpacked into a temporary directory."""
    for c in (os.environ.get("PYVBMC_REFERENCE_ROOT"), "/root/reference"):
        if _is_ref(c):
            return c
    archive = os.path.join(_HERE, "_ref", "pyvbmc_ref.zip")
    if os.path.isfile(archive) and os.environ.get("PYVBMC_REFERENCE_ROOT") != "none":
        import atexit
        import shutil
        import tempfile
        import zipfile

        tmp = tempfile.mkdtemp(prefix="pyvbmc_ref_")
        atexit.register(shutil.rmtree, tmp, ignore_errors=True)
        with zipfile.ZipFile(archive) as z:
            z.extractall(tmp)
        return tmp
    return "/root/reference"


REFERENCE_ROOT = _find_root()

_STUBBED = [
    "cma",
    "corner",
    "imageio",
    "matplotlib",
    "matplotlib.pyplot",
    "plotly",
    "plotly.graph_objects",
    "plotly.subplots",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyvbmc"))


def _install_stubs():
    for name in _STUBBED:
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    if "gpyreg" not in sys.modules:
        gpr = types.ModuleType("gpyreg")
        mf = types.ModuleType("gpyreg.mean_functions")

        class ZeroMean:  # noqa: D401 - marker classes for isinstance()
            pass

        class ConstantMean:
            pass

        class NegativeQuadratic:
            pass

        mf.ZeroMean, mf.ConstantMean, mf.NegativeQuadratic = (
            ZeroMean,
            ConstantMean,
            NegativeQuadratic,
        )
        gpr.mean_functions = mf
        gpr.GP = MagicMock
        gpr.__stub__ = True
        sys.modules["gpyreg"] = gpr
        sys.modules["gpyreg.mean_functions"] = mf


_CACHE = None


def load():
    """Return a namespace with the reference's own hot-path callables."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not available():
        raise RuntimeError(
            f"reference package not found at {REFERENCE_ROOT} nor under oracle/_ref "
            "(run `python -m oracle.build_ref` in the build container)"
        )
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import gpyreg  # the stub (or the real thing if ever installed)
    from pyvbmc.entropy import entlb_vbmc, entmc_vbmc
    from pyvbmc.variational_posterior import VariationalPosterior
    from pyvbmc.vbmc import variational_optimization as vo
    from pyvbmc.vbmc.minimize_adam import minimize_adam

    _CACHE = SimpleNamespace(
        gpyreg=gpyreg,
        entmc_vbmc=entmc_vbmc,
        entlb_vbmc=entlb_vbmc,
        VariationalPosterior=VariationalPosterior,
        _neg_elcbo=vo._neg_elcbo,
        _gp_log_joint=vo._gp_log_joint,
        _vp_bound_loss=vo._vp_bound_loss,
        _soft_bound_loss=vo._soft_bound_loss,
        variational_optimization=vo,
        minimize_adam=minimize_adam,
    )
    return _CACHE


class _Count:
    def __init__(self, n):
        self._n = n

    def hyperparameter_count(self, *_a):
        return self._n


def make_ref_gp(X, y, posts, mean_kind="negquad", noise_N=1):
    """Duck-typed ``gpyreg.GP`` for the reference's ``_gp_log_joint``.

    ``posts`` is a list of dicts with keys ``hyp, alpha, L, L_chol, sW`` (see
    ``oracle.gp_posterior``).  Field use: variational_optimization.py:1367-1398.
    """
    ref = load()
    mf = ref.gpyreg.mean_functions
    mean = {
        "negquad": mf.NegativeQuadratic,
        "const": mf.ConstantMean,
        "zero": mf.ZeroMean,
    }[mean_kind]()
    D = X.shape[1]
    return SimpleNamespace(
        X=X,
        y=y,
        D=D,
        covariance=_Count(D + 1),
        noise=_Count(noise_N),
        mean=mean,
        posteriors=[
            SimpleNamespace(
                hyp=np.asarray(p["hyp"], dtype=float),
                alpha=np.asarray(p["alpha"], dtype=float).reshape(-1, 1),
                L=np.asarray(p["L"], dtype=float),
                L_chol=bool(p["L_chol"]),
                sW=np.asarray(p["sW"], dtype=float).reshape(-1, 1),
            )
            for p in posts
        ],
    )


def make_ref_vp(D, K, mu, sigma, lambd, w, eta, optimize=(True, True, True, True)):
    ref = load()
    state = np.random.get_state()
    vp = ref.VariationalPosterior(D, K)  # ctor draws randn (variational_posterior.py:121)
    np.random.set_state(state)
    vp.mu = np.array(mu, dtype=float).reshape(D, K)
    vp.sigma = np.array(sigma, dtype=float).reshape(1, K)
    vp.lambd = np.array(lambd, dtype=float).reshape(D, 1)
    vp.w = np.array(w, dtype=float).reshape(1, K)
    vp.eta = np.array(eta, dtype=float).reshape(1, K)
    (
        vp.optimize_mu,
        vp.optimize_sigma,
        vp.optimize_lambd,
        vp.optimize_weights,
    ) = [bool(o) for o in optimize]
    return vp
