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
