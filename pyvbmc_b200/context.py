"""Device context: owns one packed GP plus the evaluation workspace on one B200.

Thin object wrapper over the C ABI (``include/vbmc_b200.h``).  NumPy fp64 in, NumPy
fp64 / Python floats out -- the ownership contract of the reference's functions
(SURVEY 8b): inputs are never kept, outputs are fresh host arrays, device buffers
never hang off ``vp`` / ``gp`` objects (those get pickled with dill by the reference).
"""
from __future__ import annotations

import atexit
import ctypes as C
import sys
import weakref
from collections import OrderedDict

import numpy as np

from . import _capi
from .config import config

_F64 = np.float64


def _arr(a, shape=None):
    out = np.ascontiguousarray(a, dtype=_F64)
    if shape is not None:
        out = out.reshape(shape)
    return out


def _ptr(a):
    return a.ctypes.data_as(_capi.c_double_p) if a is not None else None


def _flags(grad_flags):
    if np.isscalar(grad_flags):
        grad_flags = (bool(grad_flags),) * 4
    g = tuple(bool(x) for x in grad_flags)
    if len(g) != 4:
        raise ValueError("grad_flags must be a bool or a 4-tuple")
    return g


def packed_len(D, K, g):
    return (D * K if g[0] else 0) + (K if g[1] else 0) + (D if g[2] else 0) + (K if g[3] else 0)


class _VPView:
    """Flattened fp64 view of a (duck-typed) VariationalPosterior in the C layout."""

    def __init__(self, vp):
        self.D, self.K = int(vp.D), int(vp.K)
        D, K = self.D, self.K
        # vp.mu is (D, K); the C side wants component-major == ravel(order="F")
        self.mu = np.ascontiguousarray(np.asarray(vp.mu, dtype=_F64).reshape(D, K).T).reshape(-1)
        self.sigma = _arr(vp.sigma, (K,))
        self.lambd = _arr(vp.lambd, (D,))
        self.w = _arr(vp.w, (K,))
        self.eta = _arr(vp.eta, (K,))
        self.c = _capi.VP(D, K, _ptr(self.mu), _ptr(self.sigma), _ptr(self.lambd), _ptr(self.w), _ptr(self.eta))


def gp_mean_kind(gp) -> int:
    """variational_optimization.py:1328-1330,1383 -- decided there with isinstance on gpyreg classes."""
    kind = getattr(gp, "mean_kind", None)
    if kind is None:
        kind = type(gp.mean).__name__
    table = {
        "negquad": _capi.MEAN_NEGQUAD,
        "NegativeQuadratic": _capi.MEAN_NEGQUAD,
        "const": _capi.MEAN_CONST,
        "ConstantMean": _capi.MEAN_CONST,
        "zero": _capi.MEAN_ZERO,
        "ZeroMean": _capi.MEAN_ZERO,
    }
    if kind not in table:
        raise NotImplementedError(f"GP mean function {kind!r} is not supported on the log-joint path")
    return table[kind]


def gp_counts(gp, D):
    """(cov_N, noise_N) as read at variational_optimization.py:1367-1369."""
    if hasattr(gp, "covariance"):
        cov_N = int(gp.covariance.hyperparameter_count(D))
    else:
        cov_N = int(gp.cov_N)
    if hasattr(gp, "noise"):
        noise_N = int(gp.noise.hyperparameter_count())
    else:
        noise_N = int(gp.noise_N)
    return cov_N, noise_N


_live = weakref.WeakSet()


@atexit.register
def _close_all():  # runs before interpreter finalisation: contexts are released in order
    for ctx in list(_live):
        try:
            ctx.close()
        except Exception:
            pass


class Context:
    def __init__(self, device: int = None):
        lib = _capi.load()
        if lib.vbmc_device_count() <= 0:
            raise RuntimeError(
                "pyvbmc_b200: no CUDA device visible -- this package has no CPU fallback "
                "(use the reference NumPy path instead)"
            )
        self._lib = lib
        self.device = config.device if device is None else int(device)
        h = C.c_void_p()
        _capi.check(lib.vbmc_ctx_create(self.device, C.byref(h)))
        self._h = h
        _live.add(self)
        self._gp_token = None
        self._gp_has_L = False
        self._bnd_cache = None
        self._bnd_ids = None
        self._flat = {}
        self._opt_c = {}
        self.S = 0
        self.N = 0
        self.D = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vbmc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        # never touch the CUDA runtime while the interpreter is being torn down (other libraries
        # sharing the primary context may already have released it): leak at exit instead
        if sys.is_finalizing():
            return
        try:
            self.close()
        except Exception:
            pass

    def read_device(self, dev_ptr: int, n: int) -> np.ndarray:
        """Copy ``n`` doubles from a device pointer to a fresh host array, ordered after the
        work enqueued on this context's stream (synchronises)."""
        out = np.empty(int(n), dtype=_F64)
        _capi.check(self._lib.vbmc_read_device(self._h, C.c_void_p(int(dev_ptr)), int(n), _ptr(out)))
        return out

    # ------------------------------------------------------------------ bookkeeping
    @property
    def launch_count(self) -> int:
        return int(self._lib.vbmc_ctx_launch_count(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.vbmc_ctx_stream(self._h) or 0)

    def synchronize(self):
        _capi.check(self._lib.vbmc_stream_synchronize(self._h))

    def set_kernel_timing(self, on: bool):
        _capi.check(self._lib.vbmc_set_kernel_timing(self._h, int(bool(on))))

    def entmc_variant_used(self):
        """fp32 entmc kernel of the last staged evaluation (5 = tensor-core, 4 = warp-autonomous, 0 = expanded ...)."""
        return int(self._lib.vbmc_entmc_variant_used(self._h))

    def entmc_main_kernel_ms(self):
        """Average duration of the dominant entropy kernel alone (call before ``entmc_kernel_ms``, which resets)."""
        ms = C.c_double()
        _capi.check(self._lib.vbmc_entmc_main_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def entmc_kernel_ms(self):
        ms = C.c_double()
        n = C.c_int64()
        _capi.check(self._lib.vbmc_entmc_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def stage_times(self):
        """Device microseconds per stage of the last synchronous evaluation (needs VBMC_STAGE_TIMING=1)."""
        us = (C.c_double * 7)()
        _capi.check(self._lib.vbmc_stage_times(self._h, us))
        return dict(zip(("h2d", "entmc", "join_gplj", "reduce", "finalize", "d2h", "host_call"), [float(u) for u in us]))

    def fma_peak(self, kind=0) -> float:
        """Measured FMA issue peak of this device in TFLOP/s (bench.py's compute denominator);
        kind 0 = FFMA, 1 = DFMA, 2 = packed FFMA2."""
        tf = C.c_double()
        _capi.check(self._lib.vbmc_fma_peak(self._h, int(kind), C.byref(tf)))
        return float(tf.value)

    # ------------------------------------------------------------------ GP
    @staticmethod
    def gp_token(gp, need_L: bool):
        posts = gp.posteriors
        return (id(gp), len(posts), tuple((id(p.hyp), id(p.alpha), id(p.L)) for p in posts), bool(need_L))

    def pack_gp(self, gp, need_L: bool = False):
        """Upload the GP fields read by ``_gp_log_joint`` (variational_optimization.py:1311,1367-1398)."""
        X = _arr(gp.X)
        N, D = X.shape
        posts = gp.posteriors
        S = len(posts)
        cov_N, noise_N = gp_counts(gp, D)
        hyp = np.stack([_arr(p.hyp).reshape(-1) for p in posts])
        alpha = np.stack([_arr(p.alpha).reshape(-1) for p in posts])
        if alpha.shape != (S, N):
            raise ValueError("gp.posteriors[s].alpha must have N entries")
        L_chol = np.array([int(bool(p.L_chol)) for p in posts], dtype=np.int32)
        sn2 = np.array([1.0 / float(np.asarray(p.sW).reshape(-1)[0]) ** 2 for p in posts], dtype=_F64)
        L = None
        if need_L:
            L = np.stack([_arr(p.L).reshape(N, N) for p in posts])
        _capi.check(
            self._lib.vbmc_gp_pack(
                self._h, D, N, S, _ptr(X), _ptr(hyp), hyp.shape[1], _ptr(alpha), _ptr(L),
                L_chol.ctypes.data_as(_capi.c_int_p), _ptr(sn2), gp_mean_kind(gp), cov_N, noise_N,
            )
        )
        self.S, self.N, self.D = S, N, D
        self._gp_has_L = bool(need_L)

    # ------------------------------------------------------------------ bounds
    def set_bounds(self, theta_bnd):
        """Upload ``theta_bnd`` (variational_posterior.py:225-239) unless unchanged."""
        if theta_bnd is None:
            return False
        # fast path: the very same dict / arrays as last time (get_bounds builds fresh arrays per call,
        # variational_posterior.py:213-228, and nothing in the reference mutates them afterwards)
        # (held references compared with `is`: a bare id() is recycled once the object dies)
        held = self._bnd_ids
        if (
            held is not None
            and held[0] is theta_bnd
            and held[1] is theta_bnd["lb"]
            and held[2] is theta_bnd["ub"]
            and held[3:] == (theta_bnd["tol_con"], theta_bnd.get("weight_threshold"), theta_bnd.get("weight_penalty"))
        ):
            return True
        self._bnd_ids = (theta_bnd, theta_bnd["lb"], theta_bnd["ub"], theta_bnd["tol_con"],
                         theta_bnd.get("weight_threshold"), theta_bnd.get("weight_penalty"))
        lb = _arr(theta_bnd["lb"]).reshape(-1)
        ub = _arr(theta_bnd["ub"]).reshape(-1)
        tol = float(theta_bnd["tol_con"])
        thr = float(theta_bnd.get("weight_threshold", 0.0))
        pen = float(theta_bnd.get("weight_penalty", 0.0))
        c = self._bnd_cache
        if (
            c is not None
            and c[2:] == (tol, thr, pen)
            and c[0].shape == lb.shape
            and np.array_equal(c[0], lb)
            and np.array_equal(c[1], ub)
        ):
            return True
        _capi.check(self._lib.vbmc_set_bounds(self._h, lb.size, _ptr(lb), _ptr(ub), tol, thr, pen))
        self._bnd_cache = (lb.copy(), ub.copy(), tol, thr, pen)
        return True

    # ------------------------------------------------------------------ entropy
    def entmc(self, vp, Ns, grad_flags=(True,) * 4, jacobian_flag=True, eps=None, seed=0, offset=0, precision=None):
        g = _flags(grad_flags)
        v = _VPView(vp)
        Ns_even = int(np.ceil(Ns / 2)) * 2  # entmc_vbmc.py:61 (Ns may arrive as a float)
        if Ns_even <= 0:
            raise ValueError("Ns must be > 0")
        prec = _capi.PREC_F64 if (precision or config.precision) == "f64" else _capi.PREC_F32
        if eps is not None:
            eps = _arr(eps)
            if eps.size != v.K * (Ns_even // 2) * v.D:
                raise ValueError("eps must have shape (K, Ns/2, D)")
            mode = _capi.RNG_EPS
        else:
            mode = _capi.RNG_PHILOX
        H = C.c_double()
        dH = np.empty(packed_len(v.D, v.K, (True,) * 4), dtype=_F64)
        gf = (C.c_int * 4)(*[int(x) for x in g])
        _capi.check(
            self._lib.vbmc_entmc(
                self._h, C.byref(v.c), Ns_even, gf, int(bool(jacobian_flag)), mode, _ptr(eps), int(seed), int(offset),
                prec, C.byref(H), _ptr(dH),
            )
        )
        return float(H.value), dH[: packed_len(v.D, v.K, g)].copy()

    def entlb(self, vp, grad_flags=(True,) * 4, jacobian_flag=True):
        g = _flags(grad_flags)
        v = _VPView(vp)
        H = C.c_double()
        dH = np.empty(packed_len(v.D, v.K, (True,) * 4), dtype=_F64)
        gf = (C.c_int * 4)(*[int(x) for x in g])
        _capi.check(self._lib.vbmc_entlb(self._h, C.byref(v.c), gf, int(bool(jacobian_flag)), C.byref(H), _ptr(dH)))
        return float(H.value), dH[: packed_len(v.D, v.K, g)].copy()

    def philox_normals(self, D, K, Ns, seed=0, offset=0):
        Ns_even = int(np.ceil(Ns / 2)) * 2
        out = np.empty((K, Ns_even // 2, D), dtype=_F64)
        _capi.check(self._lib.vbmc_philox_normals(self._h, D, K, Ns_even, int(seed), int(offset), _ptr(out)))
        return out

    # ------------------------------------------------------------------ log joint
    def gplogjoint(self, vp, grad_flags, avg_flag=True, jacobian_flag=True, compute_var=False, separate_K=False):
        """Raw outputs of ``vbmc_gplogjoint``: dict(G, dG, varG, var_ss, I_sk, J_sjk)."""
        g = _flags(grad_flags)
        v = _VPView(vp)
        S, K, D = self.S, v.K, v.D
        jac = bool(jacobian_flag)
        g_eff = (g[0], g[1] and jac, g[2] and jac, g[3] and jac)
        P = packed_len(D, K, g_eff)
        per_s = S > 1 and not avg_flag
        G = np.zeros(S if per_s else 1, dtype=_F64)
        dG = np.zeros((P, S) if per_s else (P,), dtype=_F64)
        varG = np.zeros(S if per_s else 1, dtype=_F64)
        var_ss = np.zeros(1, dtype=_F64)
        I_sk = np.zeros((S, K), dtype=_F64) if separate_K else None
        J_sjk = np.zeros((S, K, K), dtype=_F64) if (separate_K and compute_var) else None
        gf = (C.c_int * 4)(*[int(x) for x in g])
        _capi.check(
            self._lib.vbmc_gplogjoint(
                self._h, C.byref(v.c), gf, int(bool(avg_flag)), int(jac), int(compute_var), int(bool(separate_K)),
                _ptr(G), _ptr(dG) if P else None, _ptr(varG), _ptr(var_ss), _ptr(I_sk), _ptr(J_sjk),
            )
        )
        return dict(G=G, dG=dG if any(g) else None, varG=varG, var_ss=float(var_ss[0]), I_sk=I_sk, J_sjk=J_sjk,
                    per_s=per_s)

    # ------------------------------------------------------------------ negative ELCBO
    def _elcbo_in(self, vp, optimize, ln_sigma_b, ln_lambd_b, eta_b, Ns, compute_grad, compute_var, separate_K,
                  use_bounds, eps, seed, offset, precision):
        v = _VPView(vp)
        keep = [v]
        inp = _capi.ElcboIn()
        inp.vp = v.c
        inp.optimize = (C.c_int * 4)(*[int(bool(o)) for o in optimize])
        for name, a, n in (("ln_sigma_b", ln_sigma_b, v.K), ("ln_lambd_b", ln_lambd_b, v.D), ("eta_b", eta_b, v.K)):
            if a is not None:
                a = _arr(a, (n,))
                keep.append(a)
                setattr(inp, name, _ptr(a))
        Ns_even = int(np.ceil(Ns / 2)) * 2 if Ns > 0 else 0
        inp.Ns = Ns_even
        inp.compute_grad = int(bool(compute_grad))
        inp.compute_var = int(bool(compute_var))
        inp.separate_K = int(bool(separate_K))
        inp.use_bounds = int(bool(use_bounds))
        if eps is not None and Ns_even > 0:
            eps = _arr(eps)
            if eps.size != v.K * (Ns_even // 2) * v.D:
                raise ValueError("eps must have shape (K, Ns/2, D)")
            keep.append(eps)
            inp.rng_mode = _capi.RNG_EPS
            inp.eps = _ptr(eps)
        else:
            inp.rng_mode = _capi.RNG_PHILOX
        inp.seed = int(seed)
        inp.offset = int(offset)
        inp.precision = _capi.PREC_F64 if (precision or config.precision) == "f64" else _capi.PREC_F32
        return inp, v, keep

    def negelcbo(self, vp, optimize, Ns, compute_grad, compute_var=False, separate_K=False, use_bounds=False,
                 ln_sigma_b=None, ln_lambd_b=None, eta_b=None, eps=None, seed=0, offset=0, precision=None):
        inp, v, keep = self._elcbo_in(vp, optimize, ln_sigma_b, ln_lambd_b, eta_b, Ns, compute_grad, compute_var,
                                      separate_K, use_bounds, eps, seed, offset, precision)
        D, K, S = v.D, v.K, self.S
        g = tuple(bool(o) for o in optimize) if compute_grad else (False,) * 4
        P = packed_len(D, K, g)
        out = _capi.ElcboOut()
        dF = np.empty(max(P, 1), dtype=_F64)
        dH = np.empty(max(P, 1), dtype=_F64)
        out.dF, out.dH = _ptr(dF), _ptr(dH)
        I_sk = J_sjk = None
        if separate_K:
            I_sk = np.zeros((S, K), dtype=_F64)
            out.I_sk = _ptr(I_sk)
            if compute_var:
                J_sjk = np.zeros((S, K, K), dtype=_F64)
                out.J_sjk = _ptr(J_sjk)
        _capi.check(self._lib.vbmc_negelcbo(self._h, C.byref(inp), C.byref(out)))
        del keep
        return dict(
            F=float(out.F), G=float(out.G), H=float(out.H), varF=float(out.varF), varG_ss=float(out.varG_ss),
            dF=dF[:P].copy() if compute_grad else None, dH=dH[:P].copy() if compute_grad else None,
            I_sk=I_sk, J_sjk=J_sjk,
        )

    # ------------------------------------------------------------------ low-overhead flat path
    def flat_buffers(self, D, K):
        """Preallocated host arrays of the flat entry point: (params, out)."""
        key = (D, K)
        buf = self._flat.get(key)
        if buf is None:
            n_in = D * K + 5 * K + 2 * D
            n_out = 8 + 2 * (D * K + 2 * K + D)
            buf = (np.zeros(n_in, dtype=_F64), np.zeros(n_out, dtype=_F64))
            self._flat[key] = buf
        return buf

    def negelcbo_flat(self, D, K, params, optimize, Ns_even, compute_grad, use_bounds, eps, seed, precision, want_dH,
                      out, offset=0):
        """``vbmc_negelcbo_flat``: one packed parameter block in, ``out`` filled in place."""
        og = self._opt_c.get(optimize)
        if og is None:
            og = self._opt_c[optimize] = (C.c_int * 4)(*[int(bool(o)) for o in optimize])
        if eps is not None:
            mode, eptr = _capi.RNG_EPS, eps.ctypes.data
        else:
            mode, eptr = _capi.RNG_PHILOX, None
        prec = _capi.PREC_F64 if (precision or config.precision) == "f64" else _capi.PREC_F32
        rc = self._lib.vbmc_negelcbo_flat(
            self._h, D, K, params.ctypes.data, og, Ns_even, int(compute_grad), int(use_bounds), mode, eptr, seed,
            int(offset), prec, int(want_dH), out.ctypes.data,
        )
        if rc:
            _capi.check(rc)

    # ------------------------------------------------------------------ theta-in fast path
    def theta_buffers(self, D, K):
        """Preallocated host arrays of ``vbmc_negelcbo_theta``: (out, vp_out, tmpl, their addresses)."""
        key = ("theta", D, K)
        buf = self._flat.get(key)
        if buf is None:
            P = D * K + 2 * K + D
            out, vpo, tmpl = np.zeros(8 + P, dtype=_F64), np.zeros(2 * K + D, dtype=_F64), np.zeros(self.param_len(D, K), dtype=_F64)
            buf = (out, vpo, tmpl, (out.ctypes.data, vpo.ctypes.data, tmpl.ctypes.data))
            self._flat[key] = buf
        return buf

    def noise_prefetch(self, D, K, Ns_even, seed, offset=0):
        """``vbmc_noise_prefetch``: start the generator of this evaluation's draws on the side stream."""
        rc = self._lib.vbmc_noise_prefetch(self._h, D, K, Ns_even, seed, int(offset))
        if rc:
            _capi.check(rc)

    def negelcbo_theta(self, D, K, theta, use_tmpl, optimize, Ns_even, compute_grad, use_bounds, seed, offset, precision, ptrs):
        """``vbmc_negelcbo_theta``: raw optimiser vector in (its eta block is shifted in place); ``ptrs`` = addresses of
        the ``(out, vp_out, tmpl)`` arrays of :meth:`theta_buffers`, filled in place."""
        og = self._opt_c.get(optimize)
        if og is None:
            og = self._opt_c[optimize] = (C.c_int * 4)(*[int(bool(o)) for o in optimize])
        prec = _capi.PREC_F64 if (precision or config.precision) == "f64" else _capi.PREC_F32
        rc = self._lib.vbmc_negelcbo_theta(
            self._h, D, K, theta.ctypes.data, ptrs[2] if use_tmpl else None, og, Ns_even,
            int(compute_grad), int(use_bounds), seed, int(offset), prec, ptrs[0], ptrs[1],
        )
        if rc:
            _capi.check(rc)

    def param_len(self, D, K):
        return int(self._lib.vbmc_param_len(int(D), int(K)))

    def negelcbo_batch(self, D, K, params, optimize, use_bounds):
        """``vbmc_negelcbo_batch``: ``params`` is ``(B, param_len)`` C-contiguous; returns ``(B, 4)`` = F, G, H, L."""
        params = np.ascontiguousarray(params, dtype=_F64)
        B = params.shape[0]
        if params.shape[1] != self.param_len(D, K):
            raise ValueError("negelcbo_batch: parameter blocks have the wrong length")
        out = np.empty((B, 4), dtype=_F64)
        og = (C.c_int * 4)(*[int(bool(o)) for o in optimize])
        _capi.check(self._lib.vbmc_negelcbo_batch(self._h, B, int(D), int(K), params.ctypes.data, og, int(bool(use_bounds)),
                                                  out.ctypes.data))
        return out

    # ------------------------------------------------------------------ device-resident Adam
    def adam_init(self, D, K, params, theta0, optimize, Ns_even, use_bounds, seed, offset, lb, ub, max_iter,
                  master_min, master_max, master_decay, precision=None):
        a = _capi.AdamIn()
        keep = [np.ascontiguousarray(params, dtype=_F64), np.ascontiguousarray(theta0, dtype=_F64)]
        a.D, a.K = int(D), int(K)
        a.params, a.theta0 = keep[0].ctypes.data, keep[1].ctypes.data
        for i in range(4):
            a.optimize[i] = int(bool(optimize[i]))
        a.Ns, a.use_bounds, a.seed, a.offset = int(Ns_even), int(bool(use_bounds)), int(seed), int(offset)
        for name, arr in (("lb", lb), ("ub", ub)):
            if arr is not None:
                arr = np.ascontiguousarray(arr, dtype=_F64)
                keep.append(arr)
                setattr(a, name, arr.ctypes.data)
        a.max_iter = int(max_iter)
        a.master_min, a.master_max, a.master_decay = float(master_min), float(master_max), float(master_decay)
        a.precision = _capi.PREC_F64 if (precision or config.precision) == "f64" else _capi.PREC_F32
        _capi.check(self._lib.vbmc_adam_init(self._h, C.byref(a)))
        self._adam_P = int(keep[1].size)

    def adam_steps(self, n):
        """Next ``n`` iterations: ``(y[n], x[n, P])`` (objective values seen, iterates after each update)."""
        y = np.empty(n, dtype=_F64)
        x = np.empty((n, self._adam_P), dtype=_F64)
        _capi.check(self._lib.vbmc_adam_steps(self._h, int(n), y.ctypes.data, x.ctypes.data))
        return y, x

    def adam_enqueue(self, n):
        """``vbmc_adam_enqueue``: issue the next ``n`` iterations, no synchronisation."""
        _capi.check(self._lib.vbmc_adam_enqueue(self._h, int(n)))

    def adam_fetch(self, i0, n):
        """``vbmc_adam_fetch``: wait for iterations ``[i0, i0 + n)`` only and return ``(y[n], x[n, P])``."""
        y = np.empty(n, dtype=_F64)
        x = np.empty((n, self._adam_P), dtype=_F64)
        _capi.check(self._lib.vbmc_adam_fetch(self._h, int(i0), int(n), y.ctypes.data, x.ctypes.data))
        return y, x

    # ------------------------------------------------------------------ peer-memory all-reduce (one node)
    def p2p_export(self, world, D, K) -> bytes:
        buf = (C.c_ubyte * 64)()
        _capi.check(self._lib.vbmc_p2p_export(self._h, int(world), int(D), int(K), buf))
        return bytes(buf)

    def p2p_open(self, rank, world, handles):
        blob = b"".join(handles)
        if len(blob) != 64 * world:
            raise ValueError("p2p_open: need one 64-byte handle per rank")
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        _capi.check(self._lib.vbmc_p2p_open(self._h, int(rank), int(world), arr))

    def p2p_unmap(self):
        _capi.check(self._lib.vbmc_p2p_unmap(self._h))

    def p2p_close(self):
        _capi.check(self._lib.vbmc_p2p_close(self._h))

    # split-phase API (multi-GPU / kernel-only timing); device pointers are plain ints
    def upload(self, vp, optimize, Ns, compute_grad=True, use_bounds=False, ln_sigma_b=None, ln_lambd_b=None,
               eta_b=None, eps=None, seed=0, offset=0, precision=None):
        inp, v, keep = self._elcbo_in(vp, optimize, ln_sigma_b, ln_lambd_b, eta_b, Ns, compute_grad, False, False,
                                      use_bounds, eps, seed, offset, precision)
        _capi.check(self._lib.vbmc_negelcbo_upload(self._h, C.byref(inp)))
        # the upload is asynchronous w.r.t. pageable eps memory only until the copy is enqueued;
        # pinned staging is owned by the library, so nothing needs to outlive this call.
        self.synchronize() if eps is not None else None
        return v.D, v.K

    def partials_async(self, rank, world, raw_dev_ptr):
        _capi.check(self._lib.vbmc_negelcbo_partials_async(self._h, int(rank), int(world), C.c_void_p(raw_dev_ptr)))

    def finalize_async(self, raw_dev_ptr, out_dev_ptr):
        _capi.check(self._lib.vbmc_negelcbo_finalize_async(self._h, C.c_void_p(raw_dev_ptr), C.c_void_p(out_dev_ptr)))

    # ------------------------------------------------------------------ acquisition-function ingredients (N4)
    def gp_predict(self, Xs):
        """``vbmc_gp_predict``: ``(f_mu, f_s2)``, each ``(Nx, S)``, for the packed GP (needs the factor L)."""
        Xs = _arr(Xs)
        if Xs.ndim != 2 or Xs.shape[1] != self.D:
            raise ValueError("gp_predict: x_star must have shape (Nx, D)")
        Nx = Xs.shape[0]
        f_mu = np.empty((Nx, self.S), dtype=_F64)
        f_s2 = np.empty((Nx, self.S), dtype=_F64)
        _capi.check(self._lib.vbmc_gp_predict(self._h, Nx, _ptr(Xs), _ptr(f_mu), _ptr(f_s2)))
        return f_mu, f_s2

    def gp_predict_device_ms(self, Nx, reps=10):
        ms = C.c_double()
        _capi.check(self._lib.vbmc_gp_predict_device_ms(self._h, int(Nx), int(reps), C.byref(ms)))
        return float(ms.value)

    def vp_pdf(self, vp, Xs, log_flag=False, grad_flag=False):
        """``vbmc_vp_pdf``: density of the mixture at ``Xs`` (transformed space): ``(y (Nx,), dy (Nx, D) | None)``."""
        v = _VPView(vp)
        Xs = _arr(Xs)
        if Xs.ndim != 2 or Xs.shape[1] != v.D:
            raise ValueError("vp_pdf: x must have shape (Nx, D)")
        Nx = Xs.shape[0]
        y = np.empty(Nx, dtype=_F64)
        dy = np.empty((Nx, v.D), dtype=_F64) if grad_flag else None
        _capi.check(self._lib.vbmc_vp_pdf(self._h, C.byref(v.c), Nx, _ptr(Xs), int(bool(log_flag)), int(bool(grad_flag)),
                                          _ptr(y), _ptr(dy)))
        return y, dy

    def enqueue(self):
        """``vbmc_negelcbo_enqueue``: the evaluation staged last, again, on this context's stream; no host sync."""
        rc = self._lib.vbmc_negelcbo_enqueue(self._h)
        if rc:
            _capi.check(rc)

    def raw_len(self, D, K):
        return int(self._lib.vbmc_raw_len(D, K))

    def out_len(self, D, K):
        return int(self._lib.vbmc_out_len(D, K))


# ---------------------------------------------------------------------- per-process caches
_entropy_ctx = {}
_last_gp = None
_gp_ctx = OrderedDict()
_GP_CTX_MAX = 4


def entropy_context(device=None) -> Context:
    """Context for the GP-free entry points (entmc / entlb)."""
    dev = config.device if device is None else int(device)
    if dev not in _entropy_ctx:
        _entropy_ctx[dev] = Context(dev)
    return _entropy_ctx[dev]


def _hold(gp):
    """Keeps a cached GP identifiable: a weak reference where the type allows one, else the object itself
    (a strong reference, so that its ``id`` cannot be recycled while the cache entry lives), plus the first
    and last posterior ``alpha`` arrays (small; their identity changes with every ``gp.update`` / ``gp.fit``)."""
    posts = gp.posteriors
    try:
        ref = weakref.ref(gp)
    except TypeError:
        ref = lambda gp=gp: gp  # noqa: E731
    return (ref, len(posts), posts[0].alpha, posts[-1].alpha)


def _held_is(hold, gp) -> bool:
    ref, n, a0, a1 = hold
    if ref() is not gp:
        return False
    posts = gp.posteriors
    return len(posts) == n and posts[0].alpha is a0 and posts[-1].alpha is a1


def context_for_gp(gp, need_L=False, device=None) -> Context:
    """Context holding ``gp`` on the device (packed once per trained GP, small LRU).

    Keyed by the identity of the GP object and of its posterior arrays: ``gp.update`` / ``gp.fit`` create
    new posterior records, which invalidates the entry.  Identity is checked against held references
    (never a bare ``id``, which CPython recycles once an object dies).  Nothing is attached to ``gp``."""
    dev = config.device if device is None else int(device)
    global _last_gp
    if _last_gp is not None and _last_gp[0] == dev and _held_is(_last_gp[2], gp):
        ctx = _last_gp[1]
        if ctx._h and (not need_L or ctx._gp_has_L):
            return ctx
    ctx = _context_for_gp_slow(gp, need_L, dev)
    _last_gp = (dev, ctx, _hold(gp))
    return ctx


def _context_for_gp_slow(gp, need_L, dev) -> Context:
    tok = (dev,) + Context.gp_token(gp, False)
    hit = _gp_ctx.get(tok)
    if hit is not None:
        ctx, hold = hit
        if ctx._h and _held_is(hold, gp):
            _gp_ctx.move_to_end(tok)
            if need_L and not ctx._gp_has_L:
                ctx.pack_gp(gp, need_L=True)
            return ctx
        del _gp_ctx[tok]
        ctx.close()
    ctx = Context(dev)
    ctx.pack_gp(gp, need_L=need_L)
    _gp_ctx[tok] = (ctx, _hold(gp))
    while len(_gp_ctx) > _GP_CTX_MAX:
        _, (old, _r) = _gp_ctx.popitem(last=False)
        old.close()
    return ctx


def clear_caches():
    global _last_gp
    _last_gp = None
    for ctx, _ in _gp_ctx.values():
        ctx.close()
    _gp_ctx.clear()
    for ctx in _entropy_ctx.values():
        ctx.close()
    _entropy_ctx.clear()
