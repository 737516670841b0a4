"""ctypes binding of ``libvbmc_b200.so`` (C ABI declared in ``include/vbmc_b200.h``).

This is the ONLY compute path of the package: there is no CPU fallback.  If the
shared library has not been built, or no CUDA device is visible, every compute
entry point raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VBMC_LIB") or os.path.join(_HERE, "csrc", "libvbmc_b200.so")  # (VBMC_LIB: development builds)

OK, ERR_CUDA, ERR_ARG, ERR_UNSUPPORTED, ERR_STATE = 0, 1, 2, 3, 4
MEAN_ZERO, MEAN_CONST, MEAN_NEGQUAD = 0, 1, 2
PREC_F32, PREC_F64 = 0, 1
RNG_EPS, RNG_PHILOX = 0, 1

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class VP(C.Structure):
    _fields_ = [
        ("D", C.c_int),
        ("K", C.c_int),
        ("mu", c_double_p),
        ("sigma", c_double_p),
        ("lambd", c_double_p),
        ("w", c_double_p),
        ("eta", c_double_p),
    ]


class ElcboIn(C.Structure):
    _fields_ = [
        ("vp", VP),
        ("optimize", C.c_int * 4),
        ("ln_sigma_b", c_double_p),
        ("ln_lambd_b", c_double_p),
        ("eta_b", c_double_p),
        ("Ns", C.c_int64),
        ("compute_grad", C.c_int),
        ("compute_var", C.c_int),
        ("separate_K", C.c_int),
        ("use_bounds", C.c_int),
        ("rng_mode", C.c_int),
        ("eps", c_double_p),
        ("seed", C.c_uint64),
        ("offset", C.c_uint64),
        ("precision", C.c_int),
    ]


class AdamIn(C.Structure):
    _fields_ = [
        ("D", C.c_int),
        ("K", C.c_int),
        ("params", C.c_void_p),
        ("theta0", C.c_void_p),
        ("optimize", C.c_int * 4),
        ("Ns", C.c_int64),
        ("use_bounds", C.c_int),
        ("seed", C.c_uint64),
        ("offset", C.c_uint64),
        ("lb", C.c_void_p),
        ("ub", C.c_void_p),
        ("max_iter", C.c_int),
        ("master_min", C.c_double),
        ("master_max", C.c_double),
        ("master_decay", C.c_double),
        ("precision", C.c_int),
    ]


class ElcboOut(C.Structure):
    _fields_ = [
        ("F", C.c_double),
        ("G", C.c_double),
        ("H", C.c_double),
        ("varF", C.c_double),
        ("varG_ss", C.c_double),
        ("dF", c_double_p),
        ("dH", c_double_p),
        ("I_sk", c_double_p),
        ("J_sjk", c_double_p),
    ]


# name -> (restype, argtypes); mirrors include/vbmc_b200.h one to one
PROTOTYPES = {
    "vbmc_abi_version": (C.c_int, []),
    "vbmc_last_error": (C.c_char_p, []),
    "vbmc_device_count": (C.c_int, []),
    "vbmc_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "vbmc_ctx_destroy": (None, [C.c_void_p]),
    "vbmc_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "vbmc_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "vbmc_entmc": (
        C.c_int,
        [C.c_void_p, C.POINTER(VP), C.c_int64, c_int_p, C.c_int, C.c_int, c_double_p, C.c_uint64, C.c_uint64, C.c_int,
         c_double_p, c_double_p],
    ),
    "vbmc_entlb": (C.c_int, [C.c_void_p, C.POINTER(VP), c_int_p, C.c_int, c_double_p, c_double_p]),
    "vbmc_philox_normals": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_uint64, c_double_p]),
    "vbmc_gp_pack": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, C.c_int, c_double_p, c_double_p, c_int_p,
         c_double_p, C.c_int, C.c_int, C.c_int],
    ),
    "vbmc_gplogjoint": (
        C.c_int,
        [C.c_void_p, C.POINTER(VP), c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
         c_double_p, c_double_p, c_double_p],
    ),
    "vbmc_set_bounds": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_double, C.c_double, C.c_double]),
    "vbmc_negelcbo": (C.c_int, [C.c_void_p, C.POINTER(ElcboIn), C.POINTER(ElcboOut)]),
    "vbmc_negelcbo_flat": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_void_p, c_int_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
         C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p],
    ),
    "vbmc_negelcbo_theta": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, c_int_p, C.c_int64, C.c_int, C.c_int, C.c_uint64,
         C.c_uint64, C.c_int, C.c_void_p, C.c_void_p],
    ),
    "vbmc_noise_prefetch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_uint64]),
    "vbmc_raw_len": (C.c_size_t, [C.c_int, C.c_int]),
    "vbmc_out_len": (C.c_size_t, [C.c_int, C.c_int]),
    "vbmc_negelcbo_upload": (C.c_int, [C.c_void_p, C.POINTER(ElcboIn)]),
    "vbmc_negelcbo_partials_async": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vbmc_negelcbo_finalize_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vbmc_stream_synchronize": (C.c_int, [C.c_void_p]),
    "vbmc_negelcbo_enqueue": (C.c_int, [C.c_void_p]),
    "vbmc_gp_predict": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "vbmc_gp_predict_device_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p]),
    "vbmc_vp_pdf": (C.c_int, [C.c_void_p, C.POINTER(VP), C.c_int, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p]),
    "vbmc_p2p_export": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vbmc_p2p_open": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vbmc_p2p_unmap": (C.c_int, [C.c_void_p]),
    "vbmc_p2p_close": (C.c_int, [C.c_void_p]),
    "vbmc_read_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, c_double_p]),
    "vbmc_set_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "vbmc_entmc_kernel_ms": (C.c_int, [C.c_void_p, c_double_p, C.POINTER(C.c_int64)]),
    "vbmc_entmc_main_kernel_ms": (C.c_int, [C.c_void_p, c_double_p]),
    "vbmc_entmc_variant_used": (C.c_int, [C.c_void_p]),
    "vbmc_param_len": (C.c_size_t, [C.c_int, C.c_int]),
    "vbmc_adam_init": (C.c_int, [C.c_void_p, C.POINTER(AdamIn)]),
    "vbmc_adam_steps": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "vbmc_adam_enqueue": (C.c_int, [C.c_void_p, C.c_int]),
    "vbmc_adam_fetch": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "vbmc_negelcbo_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_void_p]),
    "vbmc_fma_peak": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "vbmc_stage_times": (C.c_int, [C.c_void_p, c_double_p]),
}

_lib = None


def load():
    """Load the shared library (works without a GPU: cudart is linked statically)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"pyvbmc_b200: CUDA extension not built ({LIB_PATH} missing). Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C pyvbmc_b200/csrc`. "
            "There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.vbmc_abi_version() != 1:
        raise RuntimeError("pyvbmc_b200: ABI version mismatch between _capi.py and libvbmc_b200.so")
    _lib = lib
    return lib


def last_error() -> str:
    return load().vbmc_last_error().decode("utf-8", "replace")


def check(rc: int):
    """Map a status code to the reference's exception types (SURVEY 8b error conventions)."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_ARG:
        raise ValueError(msg)
    raise RuntimeError(f"vbmc_b200 error {rc}: {msg}")
