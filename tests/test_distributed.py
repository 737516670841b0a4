"""Multi-GPU sharding.  CPU part: the partition arithmetic and, with a world_size-2 gloo group, the
linearity that the single all-reduce relies on (shard partial sums, scaled by the GLOBAL 1/Ns and 1/S,
add up to the full evaluation BEFORE the Jacobians).  GPU part: torchrun on 2 GPUs vs one GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import elbo_oracle as eo
from oracle import synthetic as syn
from pyvbmc_b200.distributed import pair_range, raw_layout, sample_indices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_everything_once():
    for half in (1, 7, 1001, 4000):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(half, dtype=int)
            for r in range(world):
                lo, hi = pair_range(half, r, world)
                assert 0 <= lo <= hi <= half
                seen[lo:hi] += 1
            assert np.all(seen == 1)
    for S in (1, 3, 8, 32):
        for world in (1, 2, 4, 8):
            got = sorted(s for r in range(world) for s in sample_indices(S, r, world))
            assert got == list(range(S))


def test_raw_layout_matches_abi():
    lay = raw_layout(20, 50)
    assert lay["block"] == 20 * 50 + 2 * 50 + 20 and lay["total"] == 4 + 2 * lay["block"]
    from pyvbmc_b200 import _capi

    lib = _capi.load()
    assert lib.vbmc_raw_len(20, 50) == lay["total"]
    assert lib.vbmc_out_len(20, 50) == 8 + 3 * lay["block"] + 20 * 50


def _shard_partial(rank, world, pr, eps):
    """What one rank contributes to the raw vector, computed with the oracle: pre-Jacobian sums scaled by
    the global 1/Ns and 1/S."""
    K, D, S = pr.K, pr.D, pr.S
    half = eps.shape[1]
    lo, hi = pair_range(half, rank, world)
    vp = pr.vp.copy()
    H_r, dH_r = eo.entmc(vp, eps[:, lo:hi], (True,) * 4, jacobian_flag=False)
    frac = (hi - lo) / half
    sidx = sample_indices(S, rank, world)
    gp_r = eo.make_gp(pr.X, [pr.posts[s] for s in sidx], mean_kind=pr.mean_kind)
    # jacobian_flag=True is needed to get all four gradient blocks; undo the (linear, shared) Jacobians after
    G_r, dG_r, *_ = eo.gp_log_joint(vp, gp_r, (True, False, False, False), True, True, False)
    return np.concatenate([[H_r * frac, G_r * len(sidx) / S], dH_r * frac, dG_r * len(sidx) / S])


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr = syn.make_problem("C2", S=4)
    eps = syn.draw_eps(pr.K, 202, pr.D, seed=1)
    part = torch.from_numpy(_shard_partial(rank, world, pr, eps))
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(part.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_allreduce_of_shard_partials_equals_full():
    import torch.multiprocessing as mp

    world, port = 2, 29533
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pr = syn.make_problem("C2", S=4)
    eps = syn.draw_eps(pr.K, 202, pr.D, seed=1)
    full = _shard_partial(0, 1, pr, eps)
    assert np.allclose(got, full, rtol=1e-12, atol=1e-13)


def _p2p_worker(rank, world, port, q, scenario):
    import torch.distributed as dist

    from pyvbmc_b200.distributed import negotiate_p2p

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    log = []

    def export_fn():
        if scenario == "export_fails_on_1" and rank == 1:
            raise RuntimeError("no IPC")
        return bytes([rank]) * 64

    def open_fn(handles):
        log.append([h[0] for h in handles])
        if scenario == "open_fails_on_0" and rank == 0:
            raise RuntimeError("no peer access")

    def close_fn():
        log.append("closed")

    def unmap_fn():
        log.append("unmapped")

    enabled = negotiate_p2p(dist, None, rank, world, export_fn, open_fn, close_fn, unmap_fn)
    q.put((rank, enabled, log))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scenario,port", [("ok", 29541), ("export_fails_on_1", 29542), ("open_fails_on_0", 29543)])
def test_gloo_p2p_negotiation_is_collective(scenario, port):
    """Host logic of ShardedNegElcbo.enable_p2p over a world_size-2 gloo group: the handles reach every rank in
    rank order, and a failure on ANY rank disables the peer-memory path on ALL ranks without a deadlock (every rank
    runs the same collectives)."""
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_p2p_worker, args=(r, world, port, q, scenario)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    enabled = [g[1] for g in got]
    assert enabled == ([True, True] if scenario == "ok" else [False, False])
    if scenario == "ok":
        assert all(g[2] == [[0, 1]] for g in got)  # every rank opened the handles of ranks 0, 1 in order
    else:  # tear-down order on every rank: drop the peers' mappings, (barrier), then free the local buffer
        assert all([x for x in g[2] if isinstance(x, str)] == ["unmapped", "closed"] for g in got)


@pytest.mark.gpu
def test_two_gpu_sharded_equals_single():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "scripts", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "DIST_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---------------------------------------------------------------------------------------------------------------
# ShardedNegElcbo end to end over a host stand-in for the device context (no GPU): every attribute the class
# touches must exist, the buffers must be allocated, and the call sequence upload -> partials -> all-reduce ->
# finalize -> read must produce the all-reduced result on every rank.
class _HostContext:
    """Same methods as pyvbmc_b200.context.Context for the split-phase path; the "kernels" are NumPy on host memory.
    partials: raw[0] = rank + 1, raw[4:4+P] = theta * (rank + 1); finalize: out[0] = raw[0], out[8:8+P] = raw[4:4+P]."""

    def __init__(self):
        self.calls = []
        self.theta = None
        self.closed = False

    @staticmethod
    def _view(ptr, n):
        import ctypes

        return np.ctypeslib.as_array((ctypes.c_double * n).from_address(int(ptr)))

    def raw_len(self, D, K):
        return raw_layout(D, K)["total"]

    def out_len(self, D, K):
        return 8 + 3 * raw_layout(D, K)["block"] + D * K

    def set_bounds(self, theta_bnd):
        self.calls.append("set_bounds")
        return theta_bnd is not None

    def upload(self, vp, optimize, Ns, compute_grad=True, use_bounds=False, ln_sigma_b=None, ln_lambd_b=None,
               eta_b=None, eps=None, seed=0, offset=0, precision=None):
        self.calls.append(("upload", int(offset)))
        self.D, self.K = vp.D, vp.K
        self.theta = np.concatenate([np.ravel(vp.mu, order="F"), np.log(np.ravel(vp.sigma)), np.log(np.ravel(vp.lambd)),
                                     np.ravel(vp.eta)])
        return vp.D, vp.K

    def partials_async(self, rank, world, raw_ptr):
        self.calls.append("partials")
        raw = self._view(raw_ptr, self.raw_len(self.D, self.K))
        raw[:] = 0.0
        raw[0] = rank + 1.0
        raw[4 : 4 + self.theta.size] = self.theta * (rank + 1.0)

    def finalize_async(self, raw_ptr, out_ptr):
        self.calls.append("finalize")
        raw = self._view(raw_ptr, self.raw_len(self.D, self.K))
        out = self._view(out_ptr, self.out_len(self.D, self.K))
        out[0] = raw[0]
        out[8 : 8 + self.theta.size] = raw[4 : 4 + self.theta.size]

    def read_device(self, ptr, n):
        self.calls.append("read")
        return self._view(ptr, n).copy()

    def synchronize(self):
        pass

    def close(self):
        self.closed = True


def _host_sharded(world_rank=None):
    import pyvbmc_b200 as pv
    from pyvbmc_b200.distributed import ShardedNegElcbo

    pr = syn.make_problem("C2", S=4)
    vp = pv.VariationalPosterior(pr.D, pr.K)
    vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy(), pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()
    ctx = _HostContext()
    ev = ShardedNegElcbo(pr.gp, seed=5, ctx=ctx)
    return pr, vp, ctx, ev


def test_sharded_negelcbo_host_standin_world1():
    """__call__, enqueue and close over a host stand-in context (regression: a deleted helper method made every
    use of the class raise AttributeError while the CPU suite stayed green)."""
    pr, vp, ctx, ev = _host_sharded()
    assert (ev.rank, ev.world, ev.p2p) == (0, 1, False)
    F, dF, G, H, varF = ev(pr.theta, vp, 200, pr.theta_bnd)
    P = pr.theta.size
    assert dF.shape == (P,) and F == 1.0 and varF == 0
    th = pr.theta.copy()
    th[-pr.K:] -= th[-pr.K:].max()
    # the stand-in echoes the parameters the evaluator staged: set_parameters renormalisation included
    assert np.allclose(dF[: pr.D * pr.K], th[: pr.D * pr.K], rtol=0, atol=1e-14)
    assert np.allclose(dF[-pr.K:], th[-pr.K:], rtol=0, atol=1e-14)
    assert ctx.calls == ["set_bounds", ("upload", 1), "partials", "finalize", "read"]
    out = ev.enqueue(pr.D, pr.K)  # device-resident step of bench.py
    assert out.numel() == ctx.out_len(pr.D, pr.K) and float(out[0]) == 1.0
    # a second call reuses the buffers, a new shape re-allocates them
    raw0 = ev._raw
    ev(pr.theta, vp, 200, pr.theta_bnd)
    assert ev._raw is raw0 and ctx.calls[-5:] == ["set_bounds", ("upload", 2), "partials", "finalize", "read"]
    r, o = ev._buffers(pr.D + 1, pr.K)
    assert r is not raw0 and r.numel() == ctx.raw_len(pr.D + 1, pr.K)
    assert ev.enable_p2p(pr.D, pr.K) is False and ev.p2p_self_check(F) is False  # world 1: never negotiated
    ev.close()
    assert ctx.closed


def _sharded_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr, vp, ctx, ev = _host_sharded()
    assert (ev.rank, ev.world) == (rank, world)
    F, dF, G, H, _ = ev(pr.theta, vp, 200, pr.theta_bnd)
    single_pr, single_vp, single_ctx, single = _host_sharded()
    single.world, single.rank = 1, 0  # what single=True does
    F1, dF1, *_ = single(pr.theta, single_vp, 200, pr.theta_bnd)
    q.put((rank, F, dF, F1, dF1, ctx.calls))
    dist.barrier()
    ev.close()
    dist.destroy_process_group()


def test_sharded_negelcbo_host_standin_gloo_world2():
    """World size 2 over gloo: the all-reduce sits between partials and finalize, every rank ends with the sum."""
    import torch.multiprocessing as mp

    world, port = 2, 29551
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(world)), key=lambda g: g[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, F, dF, F1, dF1, calls in got:
        assert F == 3.0 and F1 == 1.0  # (rank0: 1) + (rank1: 2)
        assert np.allclose(dF, 3.0 * dF1, rtol=1e-15, atol=0)
        assert calls == ["set_bounds", ("upload", 1), "partials", "finalize", "read"]
    assert np.array_equal(got[0][2], got[1][2])  # replicated bitwise


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_sharded_single_equals_plugin_path(cfg):
    """One GPU: ShardedNegElcbo(single=True) (split-phase entry points, what bench.py's device-resident loop and
    every multi-GPU rank run) against the drop-in ``_neg_elcbo`` on the same Philox key."""
    import pyvbmc_b200 as pv
    from pyvbmc_b200.distributed import ShardedNegElcbo

    pr = syn.make_problem(cfg)
    Ns_K = min(pr.Ns_K, 2000)

    def fresh_vp():
        vp = pv.VariationalPosterior(pr.D, pr.K)
        vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu.copy(), pr.sigma.reshape(1, -1).copy(), pr.lambd.reshape(-1, 1).copy(), pr.w.reshape(1, -1).copy(), pr.eta.reshape(1, -1).copy()
        return vp

    ev = ShardedNegElcbo(pr.gp, seed=77, single=True)
    try:
        for it in range(3):
            theta = pr.theta + 0.01 * it
            F, dF, G, H, _ = ev(theta.copy(), fresh_vp(), Ns_K, pr.theta_bnd)
            F1, dF1, G1, H1, _ = pv._neg_elcbo(theta.copy(), pr.gp, fresh_vp(), 0.0, Ns_K, True, False, pr.theta_bnd,
                                                seed=77, offset=it + 1)
            assert abs(F - F1) <= 1e-12 * abs(F1) and abs(G - G1) <= 1e-12 * abs(G1) and abs(H - H1) <= 1e-12 * abs(H1)
            assert np.abs(dF - dF1).max() <= 1e-9 * np.abs(dF1).max()
        out = ev.enqueue(pr.D, pr.K)
        ev.ctx.synchronize()
        assert np.isfinite(float(out[0]))
    finally:
        ev.close()


@pytest.mark.gpu
def test_bench_runs_and_prints_the_contract_line():
    """`python bench.py --gpus 1 --steps 2 --warmup 1` must exit 0 and print one JSON line with every contract key
    (the round-1 bench shipped broken because nothing ran it after the last refactor)."""
    import json

    env = dict(os.environ, VBMC_BENCH_QUICK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["gpu_launches"] > 0
    assert "C3" in line["config"]["workload"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in line["cpu_baseline"], key
