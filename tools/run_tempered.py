#!/usr/bin/env python
"""Parallel-tempered chains sharded across the GPUs of one box (BASELINE config 4 in miniature).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_tempered.py \
        --config etopo_fast --temps 8 --samples 20

Every rank builds the same mesh, owns a contiguous shard of the temperatures, evaluates its shard on its GPU
(`bayeslands_mcmc.likelihood_batch`), and the B log-likelihoods are all_gathered over NCCL; accept/reject and
replica-exchange decisions come from a host RNG every rank runs identically (bayeslands_b200/dist.py).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bayeslands_b200 import dist as bdist  # noqa: E402
from bayeslands_b200.mcmc import bayeslands_mcmc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="etopo_fast")
    ap.add_argument("--temps", type=int, default=8)
    ap.add_argument("--samples", type=int, default=20)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d, inp, m, simtime, rl, el, coords, likl_sed = bench.load_config(a.config)
    per_rank = -(-a.temps // world)
    mc = bayeslands_mcmc(True, simtime, a.samples, d["final_elev"], d["final_erdp"], d["final_erdp_pts"], coords, None, inp,
                         el, rl, (0.4, 0.6), (0.9, 1.1), "pt", likl_sed, batch=per_rank, device=local, mesh=m,
                         sealevel=d["sealevel"] if "sealevel" in d else None)
    ev = bdist.ShardedEvaluator(lambda r, e, s: mc.likelihood_batch(r, e, seeds=s), device=torch.device("cuda", local))
    betas = np.geomspace(1.0, 0.02, a.temps)
    t0 = time.perf_counter()
    out = bdist.tempered_chains(ev, a.samples, betas, rl, el, (rl[1] - rl[0]) * 0.03, (el[1] - el[0]) * 0.03,
                                seed=a.seed, swap_every=2)
    dt = time.perf_counter() - t0
    digest = float(np.sum(out["lik"]))
    if world > 1:
        t = torch.tensor([digest], dtype=torch.float64, device="cuda")
        lo = t.clone(); hi = t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert lo.item() == hi.item(), "ranks diverged"
    if rank == 0:
        print("config %s temps %d samples %d ranks %d: %.2f s, %.1f evals/s, swaps %d, cold chain: rain %.3f erod %.3e L %.2f, "
              "digest %.6f" % (a.config, a.temps, a.samples, world, dt, a.temps * a.samples / dt, out["swaps"],
                               out["rain"][-1, 0], out["erod"][-1, 0], out["lik"][-1, 0], digest))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
