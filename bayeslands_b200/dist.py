"""Multi-GPU layer: proposals / chains shard across the GPUs of one box.

Each proposal is a complete, independent forward model on a read-only mesh, so nothing is exchanged during
the time loop.  Per MCMC iteration the only collective is an all_gather of the B log-likelihood scalars
(`torch.distributed`, NCCL over NVLink on GPUs, gloo in the CPU tests); every rank then runs the same
host-side RNG and takes identical accept / swap decisions, so no second collective is needed.  The
reference has no parallel tempering: the swap rule here is the standard one.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int):
    """Contiguous split of n items over `world` ranks (first n % world ranks get one more)."""
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class ShardedEvaluator:
    """Evaluate log-likelihoods of a global batch with each rank computing its contiguous shard.

    `evaluate_local(rain, erod, seeds) -> lik[np.ndarray]` is the per-GPU hot path
    (`bayeslands_mcmc.likelihood_batch`); injectable so the sharding / gather logic is testable on CPU.
    """

    def __init__(self, evaluate_local: Callable, device: Optional[torch.device] = None):
        self.fn = evaluate_local
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = device or torch.device("cpu")

    def __call__(self, rain: np.ndarray, erod: np.ndarray, seeds: Optional[np.ndarray] = None) -> np.ndarray:
        n = len(rain)
        lo, hi = shard_range(n, self.world, self.rank)
        sd = None if seeds is None else np.asarray(seeds)[lo:hi]
        out = self.fn(np.asarray(rain)[lo:hi], np.asarray(erod)[lo:hi], sd)
        local = np.asarray(out[0] if isinstance(out, tuple) else out, dtype=np.float64)
        if self.world == 1:
            return local
        # equal-size buffers for all_gather_into_tensor: pad shards to the largest one
        m = -(-n // self.world)
        buf = torch.full((m,), float("nan"), dtype=torch.float64, device=self.device)
        buf[:hi - lo] = torch.as_tensor(local, dtype=torch.float64, device=self.device)
        allbuf = torch.empty(self.world * m, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(allbuf, buf)
        allv = allbuf.cpu().numpy().reshape(self.world, m)
        res = np.empty(n)
        for r in range(self.world):
            a, b = shard_range(n, self.world, r)
            res[a:b] = allv[r, :b - a]
        return res


def metropolis_accept(lik_new: float, lik_old: float, u: float, beta: float = 1.0) -> bool:
    """`mh_prob = min(1, exp(L' - L))` with the reference's overflow guard (bl_mcmc.py:693-698)."""
    try:
        mh = min(1.0, math.exp(beta * (lik_new - lik_old)))
    except OverflowError:
        mh = 1.0
    return u < mh


def replica_swap(lik: np.ndarray, betas: Sequence[float], rng: np.random.RandomState, even: bool):
    """One sweep of neighbour swaps between temperatures: pairs (0,1),(2,3).. or (1,2),(3,4)..; swap i<->j
    is accepted with probability min(1, exp((beta_i - beta_j) (L_j - L_i))).  Returns the permutation to
    apply to the per-temperature states."""
    n = len(lik)
    perm = np.arange(n)
    for i in range(0 if even else 1, n - 1, 2):
        j = i + 1
        a = (betas[i] - betas[j]) * (lik[j] - lik[i])
        u = rng.uniform()
        if a >= 0 or u < math.exp(a):
            perm[i], perm[j] = perm[j], perm[i]
    return perm


def tempered_chains(evaluate: Callable, samples: int, betas: Sequence[float], rainlimits, erodlimits,
                    step_rain: float, step_erod: float, seed: int = 0, swap_every: int = 5):
    """Parallel-tempered random-walk Metropolis: one chain per temperature, all chains advanced with ONE
    batched (and sharded) evaluation per iteration.  Every rank must call this with the same arguments; the
    host RNG is shared so all ranks stay in lock step.  Returns dict(rain, erod, lik [samples, nT], swaps)."""
    nT = len(betas)
    rs = np.random.RandomState(seed)
    rain = rs.uniform(rainlimits[0], rainlimits[1], nT)
    erod = rs.uniform(erodlimits[0], erodlimits[1], nT)
    lik = evaluate(rain, erod, np.arange(nT, dtype=np.int64) + seed * 7919)
    out_r = np.zeros((samples, nT)); out_e = np.zeros((samples, nT)); out_l = np.zeros((samples, nT))
    out_r[0], out_e[0], out_l[0] = rain, erod, lik
    nswap = 0
    for i in range(1, samples):
        pr = rain + rs.normal(0, step_rain, nT)
        pe = erod + rs.normal(0, step_erod, nT)
        bad = (pr < rainlimits[0]) | (pr > rainlimits[1]); pr[bad] = rain[bad]
        bad = (pe < erodlimits[0]) | (pe > erodlimits[1]); pe[bad] = erod[bad]
        lp = evaluate(pr, pe, np.arange(nT, dtype=np.int64) + (seed * 7919 + i) * nT)
        u = rs.uniform(size=nT)
        for c in range(nT):
            if metropolis_accept(lp[c], lik[c], u[c], betas[c]):
                rain[c], erod[c], lik[c] = pr[c], pe[c], lp[c]
        if swap_every and i % swap_every == 0:
            perm = replica_swap(lik, betas, rs, even=(i // swap_every) % 2 == 0)
            nswap += int((perm != np.arange(nT)).sum() // 2)
            rain, erod, lik = rain[perm], erod[perm], lik[perm]
        out_r[i], out_e[i], out_l[i] = rain, erod, lik
    return dict(rain=out_r, erod=out_e, lik=out_l, swaps=nswap)
