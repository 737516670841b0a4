"""Run-time knobs of the B200 path (deliberately NOT part of the reference's .ini options)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class Config:
    #: CUDA device index used by the default context (``LOCAL_RANK`` under torchrun)
    device: int = 0
    #: "philox": device counter-based draws keyed by a seed taken from ``np.random`` (fast path);
    #: "numpy" : draw ``np.random.randn(Ns//2, D)`` per component on the host in the reference's
    #:           order (pyvbmc/entropy/entmc_vbmc.py:64-67) and upload -- bit-compatible RNG stream.
    rng_mode: str = "philox"
    #: "f32": fp32 compute / fp64 accumulation in the Monte-Carlo entropy kernel; "f64": all fp64
    precision: str = "f32"
    #: True: ``_neg_elcbo`` starts the draw generator with a separate ``vbmc_noise_prefetch`` call before it packs theta.
    #: False (default): the evaluation's own graph forks the generator at its root, beside the parameter kernel -- the
    #: same overlap on the device without the extra host call (measured 4 us of host time per evaluation)
    host_noise_prefetch: bool = False


config = Config()
