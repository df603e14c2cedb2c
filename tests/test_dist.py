"""N>1 host logic on CPU: world_size-2 gloo processes exercising the sharding / gather / tempering code."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayeslands_b200 import dist as bdist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _fake_lik(rain, erod, seeds=None):
    # smooth surrogate with its maximum at (1.5, 5e-5)
    return -(np.square((np.asarray(rain) - 1.5) / 0.3) + np.square((np.asarray(erod) - 5e-5) / 1e-5)) * 50.0


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def local(r, e, s):
        calls.append(len(r))
        return _fake_lik(r, e)
    ev = bdist.ShardedEvaluator(local)
    rng = np.random.default_rng(0)
    rain, erod = rng.uniform(0, 3, 7), rng.uniform(3e-5, 7e-5, 7)          # ragged: 7 items over 2 ranks
    full = ev(rain, erod)
    ok = np.allclose(full, _fake_lik(rain, erod)) and calls == [4 if rank == 0 else 3]
    out = bdist.tempered_chains(ev, 40, [1.0, 0.5, 0.25, 0.1], (0, 3), (3e-5, 7e-5), 0.09, 1.2e-6, seed=3)
    q.put((rank, ok, out["lik"].sum(), out["swaps"], out["rain"][-1].tolist()))
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 255, 256):
        for w in (1, 2, 3, 8):
            spans = [bdist.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_gloo_world2_sharded_gather_and_tempering():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in ps]
    assert all(r[1] for r in res)
    # both ranks hold identical chains (shared host RNG + gathered likelihoods)
    assert res[0][2:] == res[1][2:]
    # and they equal the single-process run
    single = bdist.tempered_chains(lambda r, e, s: _fake_lik(r, e), 40, [1.0, 0.5, 0.25, 0.1], (0, 3), (3e-5, 7e-5),
                                   0.09, 1.2e-6, seed=3)
    assert np.isclose(single["lik"].sum(), res[0][2]) and single["swaps"] == res[0][3]


def test_replica_swap_rule():
    rs = np.random.RandomState(0)
    # a hotter chain holding a much better state always swaps down
    perm = bdist.replica_swap(np.array([-100.0, -10.0]), [1.0, 0.1], rs, even=True)
    assert perm.tolist() == [1, 0]
    # the reverse move is accepted with probability exp((b0-b1)(L1-L0)) ~ exp(-81): never
    perm = bdist.replica_swap(np.array([-10.0, -100.0]), [1.0, 0.1], rs, even=True)
    assert perm.tolist() == [0, 1]
    assert bdist.metropolis_accept(1e9, 0.0, 0.999) and not bdist.metropolis_accept(-50.0, 0.0, 0.5)
