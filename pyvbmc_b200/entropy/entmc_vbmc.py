"""B200 replacement for ``pyvbmc.entropy.entmc_vbmc`` (pyvbmc/entropy/entmc_vbmc.py:6-134)."""
from __future__ import annotations

import numpy as np

from ..config import config
from ..context import entropy_context


def draw_seed() -> int:
    """53-bit Philox seed taken from the GLOBAL NumPy RNG, so that ``np.random.seed(s)``
    makes the estimate a deterministic function of the parameters, as in the reference
    (whose draws come from the same global state, entmc_vbmc.py:67)."""
    return int(np.random.random_sample() * 9007199254740992.0)  # 53 bits, one draw of the global stream


def draw_eps_numpy(K: int, Ns_even: int, D: int) -> np.ndarray:
    """Host draws in the reference's exact order: one ``randn(Ns//2, D)`` per component."""
    return np.stack([np.random.randn(Ns_even // 2, D) for _ in range(K)], axis=0)


def entmc_vbmc(vp, Ns, grad_flags=tuple([True] * 4), jacobian_flag=True, *, eps=None, seed=None, _ctx=None):
    """Monte Carlo estimate of the entropy of the variational posterior and its gradient.

    Same signature and return convention as the reference: ``(H: float, dH: ndarray)`` with
    ``dH = [mu-grad (column-major), sigma, lambda, w]`` restricted to ``grad_flags``
    (``dH.shape == (0,)`` when none).  ``Ns`` is the number of draws PER COMPONENT and is
    rounded up to even (antithetic pairs).

    Keyword-only extensions: ``eps`` -- explicit standard-normal draws of shape
    ``(K, Ns/2, D)`` (parity mode); ``seed`` -- explicit Philox seed.
    """
    D, K = int(vp.D), int(vp.K)
    Ns_even = int(np.ceil(Ns / 2)) * 2
    ctx = _ctx if _ctx is not None else entropy_context()
    if eps is None and seed is None:
        if config.rng_mode == "numpy":
            eps = draw_eps_numpy(K, Ns_even, D)
        else:
            seed = draw_seed()
    return ctx.entmc(vp, Ns_even, grad_flags, jacobian_flag, eps=eps, seed=seed or 0)
