"""torchrun entry: sharded evaluation (draws + hyper-samples split over ranks, one NCCL all-reduce)
must reproduce the single-GPU evaluation on the same Philox key.  Prints DIST_CHECK_OK on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyvbmc_b200 as pv
from pyvbmc_b200.distributed import ShardedNegElcbo
from workloads import synthetic as syn

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pv.config.device = local

ok = True
for cfg, S in (("C2", 6), ("C3", 8)):
    pr = syn.make_problem(cfg, S=S)
    vp = pv.VariationalPosterior(pr.D, pr.K)
    vp.mu, vp.sigma, vp.lambd, vp.w, vp.eta = pr.mu, pr.sigma.reshape(1, -1), pr.lambd.reshape(-1, 1), pr.w.reshape(1, -1), pr.eta.reshape(1, -1)
    Ns_K = 2002  # not divisible by 2 * world: uneven shards
    sharded = ShardedNegElcbo(pr.gp, device=local, seed=99)
    fused = ShardedNegElcbo(pr.gp, device=local, seed=99)
    p2p = fused.enable_p2p(pr.D, pr.K)  # all-reduce over NVLink peer memory inside the tail kernel
    if rank == 0:
        print(f"{cfg}: peer-memory all-reduce {'active' if p2p else 'NOT available (NCCL path only)'}")
    single = ShardedNegElcbo(pr.gp, device=local, seed=99, single=True)
    for it in range(6):
        theta = pr.theta + 0.01 * it
        ev_ = fused if (p2p and it >= 2) else sharded  # iterations 0-1: NCCL all-reduce, 2-5: peer-memory all-reduce
        F, dF, G, H, _ = ev_(theta, vp, Ns_K, pr.theta_bnd)
        ev_.step = it + 1  # (same Philox offset as `single`, whichever evaluator ran)
        sharded.step = fused.step = it + 1
        F1, dF1, G1, H1, _ = single(theta, vp, Ns_K, pr.theta_bnd)
        eF, eG, eH = abs(F - F1) / abs(F1), abs(G - G1) / abs(G1), abs(H - H1) / abs(H1)
        eg = np.abs(dF - dF1).max() / np.abs(dF1).max()
        # every rank must hold the identical result (replicated finalize, no broadcast)
        t = torch.tensor([F], dtype=torch.float64, device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = float(lo) == float(hi)
        if rank == 0:
            print(f"{cfg} world={world} it={it}: relF {eF:.2e} relG {eG:.2e} relH {eH:.2e} rel_dF {eg:.2e} replicated={same}")
        # the fp32 kernel sums each thread's few pairs in fp32 before the fp64 reduction, and the pair -> thread
        # map depends on the shard size: gradients agree to ~1e-9 (bitwise in fp64 mode), values to ~1e-15
        ok = ok and eF < 1e-12 and eG < 1e-12 and eH < 1e-12 and eg < 1e-8 and same
    sharded.close()
    fused.close()
    single.close()
dist.barrier()
if rank == 0:
    print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
