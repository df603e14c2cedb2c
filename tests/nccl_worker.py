"""Worker of tests/test_gpu_nccl.py (one process per GPU under torch.distributed.run): ShardedEvaluator and tempered_chains on
the REAL engine over NCCL.  Every rank also evaluates the whole batch alone, so the sharded result can be compared bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bayeslands_b200 import dist as bdist  # noqa: E402
from bayeslands_b200.mcmc import bayeslands_mcmc  # noqa: E402


def main():
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d, inp, m, simtime, rl, el, coords, likl_sed = bench.load_config("crater_fast")
    n = 8

    def make(batch):
        return bayeslands_mcmc(True, simtime, 6, d["final_elev"], d["final_erdp"], d["final_erdp_pts"], coords, None, inp, el, rl,
                               (0.4, 0.6), (0.9, 1.1), "nccl", likl_sed, batch=batch, device=local, mesh=m)
    mc_shard, mc_full = make(-(-n // world)), make(n)
    ev = bdist.ShardedEvaluator(lambda r, e, s: mc_shard.likelihood_batch(r, e, seeds=s), device=torch.device("cuda", local))
    r, e, s = bench.proposals(3, 0, n, rl, el)
    got = ev(r, e, s)
    want, _ = mc_full.likelihood_batch(r, e, seeds=s)
    assert np.array_equal(got, want), (rank, got, want)
    # parallel-tempered chains: identical on every rank and identical to the single-GPU run of the same chains
    betas = np.geomspace(1.0, 0.05, n)
    args = (6, betas, rl, el, (rl[1] - rl[0]) * 0.03, (el[1] - el[0]) * 0.03)
    out = bdist.tempered_chains(ev, *args, seed=4, swap_every=2)
    alone = bdist.tempered_chains(lambda rr, ee, ss: mc_full.likelihood_batch(rr, ee, seeds=ss)[0], *args, seed=4, swap_every=2)
    for k in ("rain", "erod", "lik"):
        assert np.array_equal(out[k], alone[k]), (rank, k)
    t = torch.tensor([float(np.sum(out["lik"]))], dtype=torch.float64, device="cuda")
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert lo.item() == hi.item(), "ranks diverged"
    dist.barrier()
    if rank == 0:
        print("NCCL_WORKER_OK world=%d swaps=%d" % (world, out["swaps"]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
