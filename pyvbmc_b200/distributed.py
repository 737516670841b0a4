"""Multi-GPU evaluation of the negative ELCBO: one process per GPU (torchrun), the entropy
draws and the GP hyper-samples sharded over ranks, ONE all-reduce of the raw vector.

Sharding: for every mixture component, rank r owns the antithetic pairs
``[r*half/W, (r+1)*half/W)`` (a pair never straddles ranks); hyper-sample ``s`` belongs to rank
``s mod W``.  Each rank's kernels emit the raw (pre-Jacobian) sums already scaled by the GLOBAL
``1/Ns`` and ``1/S``; a SUM all-reduce of ``2 + 2P`` doubles (< 20 KB) yields the single-GPU
value on every rank, and every rank then runs the identical finalize kernel (Jacobians, soft
bounds, penalty), so theta updates stay replicated with no broadcast.  Philox draws are keyed
by the global (component, pair) index => the result does not depend on the number of ranks
beyond fp64 summation order.

``torch`` is used for the device tensors and ``torch.distributed`` (NCCL over NVLink/NVSwitch;
gloo in CPU tests of the host logic).
"""
from __future__ import annotations

import numpy as np


def pair_range(half_glob: int, rank: int, world: int):
    """Antithetic pairs of each component owned by ``rank`` (mirrors capi.cu:partials)."""
    return half_glob * rank // world, half_glob * (rank + 1) // world


def sample_indices(S: int, rank: int, world: int):
    """GP hyper-samples owned by ``rank``."""
    return list(range(rank, S, world))


def raw_layout(D: int, K: int):
    """Offsets inside the raw vector (mirrors RawLayout in csrc/common.cuh)."""
    block = K * D + 2 * K + D
    return {"H": 0, "G": 1, "ent": 4, "gp": 4 + block, "block": block, "total": 4 + 2 * block}


def negotiate_p2p(dist, group, rank, world, export_fn, open_fn, close_fn, unmap_fn=None):
    """Collective set-up of the peer-memory exchange (host logic, backend-agnostic: also runs over gloo).

    ``export_fn() -> bytes`` allocates the local exchange buffer and returns its IPC handle, ``open_fn(handles)``
    maps all ranks' buffers, ``close_fn()`` releases everything.  Every rank executes the SAME sequence of
    collectives whatever fails locally (two all-gathers and a barrier), and the path is only enabled if it works
    on ALL ranks."""
    ok = 1
    try:
        handle = export_fn()
    except Exception:
        handle, ok = b"\0" * 64, 0
    handles = [None] * world
    dist.all_gather_object(handles, (ok, handle), group=group)
    ok = int(all(h[0] for h in handles))
    if ok:
        try:
            open_fn([h[1] for h in handles])
        except Exception:
            ok = 0
    flags = [None] * world
    dist.all_gather_object(flags, ok, group=group)
    enabled = bool(all(flags))
    if not enabled and unmap_fn is not None:
        try:
            unmap_fn()  # drop the peers' mappings first ...
        except Exception:
            pass
    dist.barrier(group=group)  # nobody stores into a peer before every peer has zeroed its flags
    if not enabled:
        try:
            close_fn()  # ... and free the local buffer only when no peer can still map it
        except Exception:
            pass
    return enabled


class ShardedNegElcbo:
    """``_neg_elcbo(theta, gp, vp, 0, Ns, True, False, theta_bnd)`` evaluated by all ranks.

    Every rank must call :meth:`__call__` with the same ``theta``; every rank gets the same
    ``(F, dF, G, H, varF)``.
    """

    def __init__(self, gp, device=None, group=None, seed=0, single=False, ctx=None):
        """``ctx``: an already built device context (tests inject a host stand-in with the same methods; the
        raw / result vectors then live in host tensors and the collective runs over gloo)."""
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        use_dist = dist.is_initialized() and not single  # single=True: this rank evaluates everything itself
        self.rank = dist.get_rank(group) if use_dist else 0
        self.world = dist.get_world_size(group) if use_dist else 1
        if ctx is None:
            from .context import Context

            self.device = torch.cuda.current_device() if device is None else int(device)
            self.ctx = Context(self.device)
            self.ctx.pack_gp(gp)
            self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.device)
            self._tensor_device = torch.device("cuda", self.device)
        else:
            self.device, self.ctx, self.stream = device, ctx, None
            self._tensor_device = torch.device("cpu")
        self.seed = int(seed)
        self.step = 0
        self._raw = self._out = None
        self.p2p = False  # all-reduce over NVLink peer memory inside the tail kernel (else: NCCL all-reduce)

    def enable_p2p(self, D, K):
        """Map every rank's exchange buffer (CUDA IPC, one node) so that the raw vector is all-reduced over peer
        memory inside the tail kernel: one launch instead of raw kernel + NCCL all-reduce + final kernel.
        Collective: every rank must call it.  Returns whether the fused path is active on ALL ranks."""
        import os

        if self.world < 2 or self.world > 8 or os.environ.get("VBMC_P2P", "1") == "0":
            return False
        self.p2p = negotiate_p2p(
            self.dist, self.group, self.rank, self.world,
            lambda: self.ctx.p2p_export(self.world, D, K),
            lambda handles: self.ctx.p2p_open(self.rank, self.world, handles),
            self.ctx.p2p_close,
            self.ctx.p2p_unmap,
        )
        return self.p2p

    def _buffers(self, D, K):
        """Device-resident raw (pre-Jacobian, all-reduced) and result vectors, allocated on first use and
        re-allocated when the problem shape changes."""
        torch = self.torch
        n_raw, n_out = self.ctx.raw_len(D, K), self.ctx.out_len(D, K)
        if self._raw is None or self._raw.numel() != n_raw or self._out.numel() != n_out:
            self._raw = torch.zeros(n_raw, dtype=torch.float64, device=self._tensor_device)
            self._out = torch.zeros(n_out, dtype=torch.float64, device=self._tensor_device)
        return self._raw, self._out

    def p2p_self_check(self, F):
        """Collective: if ANY rank saw the peer exchange time out (poisoned result: F is NaN), every rank drops
        back to the NCCL all-reduce.  Call it once after the first evaluation."""
        if not self.p2p:
            return False
        flags = [None] * self.world
        self.dist.all_gather_object(flags, bool(np.isfinite(F)), group=self.group)
        if not all(flags):
            self.p2p = False
            self.ctx.synchronize()
            self.ctx.p2p_unmap()
            self.dist.barrier(group=self.group)  # every rank has dropped its mappings of the peers' buffers
            self.ctx.p2p_close()
        self.dist.barrier(group=self.group)
        return self.p2p

    def enqueue(self, D, K):
        """partials -> all-reduce -> finalize on the context stream (no host sync)."""
        raw, out = self._buffers(D, K)
        self.ctx.partials_async(self.rank, self.world, raw.data_ptr())
        if self.world > 1 and not self.p2p:
            if self.stream is None:  # host stand-in context
                self.dist.all_reduce(raw, op=self.dist.ReduceOp.SUM, group=self.group)
            else:
                with self.torch.cuda.stream(self.stream):
                    self.dist.all_reduce(raw, op=self.dist.ReduceOp.SUM, group=self.group)
        self.ctx.finalize_async(raw.data_ptr(), out.data_ptr())  # (with p2p: raw phases + peer all-reduce + final)
        return out

    def __call__(self, theta, vp, Ns, theta_bnd=None, eps=None):
        from .vbmc.variational_optimization import _bound_inputs, _shift_eta

        theta = np.asarray(theta, dtype=float)
        K, D = vp.K, vp.D
        vp.set_parameters(theta)
        if vp.optimize_weights:
            _shift_eta(vp, theta, K)  # in place on the caller's theta, like the reference (:1082-1085)
        optimize = (vp.optimize_mu, vp.optimize_sigma, vp.optimize_lambd, vp.optimize_weights)
        use_bounds = self.ctx.set_bounds(theta_bnd)
        b = _bound_inputs(vp, theta) if use_bounds else (None, None, None)
        self.step += 1
        self.ctx.upload(vp, optimize, Ns, True, use_bounds, b[0], b[1], b[2], eps=eps, seed=self.seed, offset=self.step)
        out = self.enqueue(D, K)
        P = sum(n for n, o in zip((D * K, K, D, K), optimize) if o)
        h = self.ctx.read_device(out.data_ptr(), 8 + P)  # D2H through the library's pinned buffer
        return float(h[0]), h[8 : 8 + P].copy(), float(h[1]), float(h[2]), 0

    def close(self):
        self.ctx.synchronize()
        if self.p2p:  # collective tear-down: unmap everywhere, then free (no buffer is freed while a peer maps it)
            self.p2p = False
            self.ctx.p2p_unmap()
            self.dist.barrier(group=self.group)
        self.ctx.close()
