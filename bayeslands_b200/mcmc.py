"""`bl_mcmc.bayeslands_mcmc` look-alike on the batched B200 engine.

Same names, argument meaning and return shapes as the reference (`bl_mcmc.py:61-835`) for the path this
repo covers: `blackBox`, `interpolateArray` (folded into the engine), `likelihoodFunc`, the accept/reject
core of `sampler` - plus batched variants (`blackBox_batch`, `likelihood_batch`, `sampler_chains`) because
one GPU evaluates hundreds of proposals at once.  Plotting / file-dump helpers of the reference
(`viewMap` ... `viewCrossSection`, `bl_mcmc.py:217-455`) are out of scope.

Deviations, all documented in DESIGN.md:
  * `load_xml` is not repeated per proposal: the mesh is built once.
  * `m`, `n` arguments are accepted and ignored exactly like the reference's dead stores (F7).
  * `pos_likl = np.zeros(samples, likelihood)` (bl_mcmc.py:618, a TypeError) becomes `np.full` (F11).
  * RNG protocol (F6/H7): numpy global stream for proposals, stdlib `random` for `u`, explicit per-proposal
    seeds for the base shuffle.
"""
from __future__ import annotations

import collections
import math
import os
import random
import time
from typing import Optional

import numpy as np
import torch

from . import _lib, config, hostparams, mesh as meshmod
from .engine import Engine


def log_likelihood(sse_elev, G, pts_pred, pts_real, likl_sed, sed_weight=50.0):
    """`likelihoodFunc` arithmetic (bl_mcmc.py:488-508): tau^2 is the MSE, so the elevation term collapses to
    -G/2 (ln(2 pi tau^2) + 1).  Vectorised over a leading batch axis."""
    sse_elev = np.asarray(sse_elev, dtype=np.float64)
    tausq = sse_elev / G
    with np.errstate(divide="ignore", invalid="ignore"):
        lik = -0.5 * G * np.log(2 * math.pi * tausq) - 0.5 * sse_elev / tausq
        lik = np.where(np.isfinite(sse_elev), lik, -np.inf)        # failed evaluations (status bits) never win
        if likl_sed:
            d2 = np.square(pts_pred[..., 1:, :] - pts_real[1:, :])          # [B, nt-1, npts]
            npts = pts_real.shape[1]
            t2 = d2.sum(axis=-1) / npts                                      # [B, nt-1]
            lp = (-0.5 * npts * np.log(2 * math.pi * t2) - 0.5 * d2.sum(axis=-1) / t2).sum(axis=-1)
            lik = np.where(np.isfinite(lik), lik + lp * sed_weight, -np.inf)
    return lik, tausq


class bayeslands_mcmc(object):
    def __init__(self, muted, simtime, samples, real_elev, real_erdp, real_erdp_pts, erdp_coords, filename,
                 xmlinput, erodlimits, rainlimits, mlimit, nlimit, run_nb, likl_sed, batch: int = 1,
                 device: int = 0, dem_xyz: Optional[np.ndarray] = None, sealevel: Optional[np.ndarray] = None,
                 base_dir: Optional[str] = None, mesh=None, shuffle_seed: int = 0, engine=None):
        self.filename = filename
        self.input = xmlinput
        self.real_elev = np.asarray(real_elev, dtype=np.float64)
        self.real_erdp = real_erdp
        self.real_erdp_pts = np.asarray(real_erdp_pts, dtype=np.float64)
        self.erdp_coords = np.asarray(erdp_coords)
        self.likl_sed = likl_sed
        self.simtime = simtime
        self.samples = samples
        self.run_nb = run_nb
        self.muted = muted
        self.erodlimits = erodlimits
        self.rainlimits = rainlimits
        self.mlimit = mlimit
        self.nlimit = nlimit
        self.step_rain = (rainlimits[1] - rainlimits[0]) * 0.03        # bl_mcmc.py:89-92
        self.step_erod = (erodlimits[1] - erodlimits[0]) * 0.03
        self.step_m = (mlimit[1] - mlimit[0]) * 0.01
        self.step_n = (nlimit[1] - nlimit[0]) * 0.01
        self.sim_interval = hostparams.sim_interval(simtime)            # bl_mcmc.py:94
        self.burn_in = 0.05
        # ---- built once instead of per proposal
        self.inp = xmlinput if isinstance(xmlinput, config.Input) else config.parse_xml(xmlinput, base_dir=base_dir)
        if engine is not None:
            # a ready engine, e.g. a `grid_engine.GridEngine` for a regular lattice too large for one CTA; it must have been
            # built with the same real_elev / erdp_coords (it computes the SSE and the point series on the device)
            self.mesh = engine
            self.batch = int(engine.B)
            self.engine = engine
        else:
            if mesh is None:
                if dem_xyz is None:
                    dem_xyz = np.loadtxt(self.inp.demfile)
                mesh = meshmod.build_from_dem(dem_xyz, self.inp.btype, self.inp.Afactor)
            if sealevel is None and self.inp.seafile is not None:
                sealevel = np.loadtxt(self.inp.seafile)
            self.mesh = mesh
            self.batch = int(batch)
            self.engine = Engine(mesh, self.inp, self.batch, device=device, real_elev=self.real_elev,
                                 erdp_coords=self.erdp_coords, sealevel=sealevel)
        self.G = self.engine.G
        self.shuffle_seed = int(shuffle_seed)
        self._evals = 0
        # pinned staging for the end-to-end path
        self._h_in = torch.empty((2, self.batch), dtype=torch.float64).pin_memory()
        self._h_seed = torch.empty(self.batch, dtype=torch.int64).pin_memory()
        self._h_sse = torch.empty(self.batch, dtype=torch.float64).pin_memory()
        self._h_pts = torch.empty((self.batch, len(self.sim_interval), len(self.erdp_coords)),
                                  dtype=torch.float64).pin_memory()

    # ------------------------------------------------------------------------------------------------
    def _seeds(self, n):
        s = np.arange(self._evals, self._evals + n, dtype=np.int64) + self.shuffle_seed * 1000003
        self._evals += n
        return s

    def blackBox_batch(self, rain, erodibility, grids: bool = False, seeds=None):
        """`blackBox` for B proposals at once.  Returns (sse[B], pts[B, nt, npts], elev, erdp) as numpy; elev /
        erdp are [B, nt, ny, nx] when `grids`.  Host buffers in, host buffers out (H2D/D2H inside)."""
        rain = np.asarray(rain, dtype=np.float64); erodibility = np.asarray(erodibility, dtype=np.float64)
        n = len(rain)
        if n > self.batch:
            raise ValueError("batch of %d proposals exceeds engine batch %d" % (n, self.batch))
        if seeds is None:
            seeds = self._seeds(n)
        hin, hseed = self._h_in, self._h_seed
        hin[0, :n] = torch.from_numpy(rain); hin[1, :n] = torch.from_numpy(erodibility)
        hseed[:n] = torch.from_numpy(np.asarray(seeds, dtype=np.int64))
        if n < self.batch:
            hin[:, n:] = hin[:, :1]; hseed[n:] = hseed[:1]
        eng = self.engine
        dev_in = hin.to(eng.tdev, non_blocking=True)
        dev_seed = hseed.to(eng.tdev, non_blocking=True)
        eng.reset(dev_in[0], dev_in[1], dev_seed)
        pts, sse, elev, erdp = eng.run_schedule(self.sim_interval, grids=grids)
        self._h_sse.copy_(sse, non_blocking=True)
        self._h_pts.copy_(pts, non_blocking=True)
        torch.cuda.current_stream(eng.tdev).synchronize()
        # device status words: an evaluation whose pit cascade hit its iteration guard or that produced a non-finite
        # surface is wrong - its SSE becomes +inf, i.e. log-likelihood -inf, so Metropolis can never accept it
        # (a NaN likelihood would be accepted: min(1, exp(nan)) is 1)
        self.last_status = eng.status()[:n].copy() if hasattr(eng, "status") else np.zeros(n, dtype=np.uint32)
        failed = (self.last_status & _lib.ST_FATAL) != 0
        if failed.any():
            self._h_sse.numpy()[:n][failed] = np.inf
        out_e = out_d = None
        if grids:
            ny, nx = self.mesh.grid_ny, self.mesh.grid_nx
            out_e = elev[:n].cpu().numpy().reshape(n, len(self.sim_interval), ny, nx)
            out_d = erdp[:n].cpu().numpy().reshape(n, len(self.sim_interval), ny, nx)
        return self._h_sse.numpy()[:n].copy(), self._h_pts.numpy()[:n].copy(), out_e, out_d

    def likelihood_batch(self, rain, erodibility, seeds=None):
        """log-likelihood[B] (and tau^2_elev[B]) for B proposals - the end-to-end hot-path call."""
        sse, pts, _, _ = self.blackBox_batch(rain, erodibility, grids=False, seeds=seeds)
        return log_likelihood(sse, self.G, pts, self.real_erdp_pts, self.likl_sed)

    # ---- reference-shaped single-proposal API ------------------------------------------------------
    def blackBox(self, rain, erodibility, m, n):
        """bl_mcmc.py:97-165: three OrderedDicts keyed by simulation time."""
        sse, pts, elev, erdp = self.blackBox_batch([rain], [erodibility], grids=True)
        elev_vec, erdp_vec, erdp_pts_vec = collections.OrderedDict(), collections.OrderedDict(), collections.OrderedDict()
        for x, T in enumerate(self.sim_interval):
            self.simtime = T
            elev_vec[T] = elev[0, x]
            erdp_vec[T] = erdp[0, x]
            erdp_pts_vec[T] = pts[0, x]
        return elev_vec, erdp_vec, erdp_pts_vec

    def likelihoodFunc(self, input_vector, real_elev, real_erdp, real_erdp_pts, tausq_elev, tausq_erdp, tausq_erdp_pts):
        """bl_mcmc.py:482-513 (the tausq_* arguments are ignored there too)."""
        pred_elev_vec, pred_erdp_vec, pred_erdp_pts_vec = self.blackBox(input_vector[0], input_vector[1],
                                                                       input_vector[2], input_vector[3])
        tausq_elev = (np.sum(np.square(pred_elev_vec[self.simtime] - real_elev))) / real_elev.size
        tausq_pts = np.zeros(self.sim_interval.size)
        for i in range(self.sim_interval.size):
            tausq_pts[i] = np.sum(np.square(pred_erdp_pts_vec[self.sim_interval[i]] - self.real_erdp_pts[i])) / real_erdp_pts.shape[1]
        likelihood_elev = -0.5 * np.log(2 * math.pi * tausq_elev) - 0.5 * np.square(pred_elev_vec[self.simtime] - real_elev) / tausq_elev
        likelihood_erdp_pts = 0
        if self.likl_sed:
            for i in range(1, self.sim_interval.size):
                likelihood_erdp_pts += np.sum(-0.5 * np.log(2 * math.pi * tausq_pts[i]) - 0.5 * np.square(
                    pred_erdp_pts_vec[self.sim_interval[i]] - self.real_erdp_pts[i]) / tausq_pts[i])
            likelihood = np.sum(likelihood_elev) + (likelihood_erdp_pts * 50)
        else:
            likelihood = np.sum(likelihood_elev)
        return [likelihood, pred_elev_vec, pred_erdp_vec, pred_erdp_pts_vec]

    # ---- accept/reject core of sampler (bl_mcmc.py:633-704, :747-755) -------------------------------
    def sampler_chains(self, nchains: Optional[int] = None, np_seed: int = 0, py_seed: int = 0, log=None,
                       speculate: Optional[int] = None):
        """`nchains` independent random-walk Metropolis chains, batched forward evaluations.  Chain c draws from its own
        numpy RandomState(np_seed + c) and random.Random(py_seed + c) in the reference's order per iteration:
        normal(rain), normal(erod), 3 x normal(size=1) for the eta walk (consumed, unused - F8), then uniform for the
        accept test.  `speculate=K` (default batch // nchains) evaluates the next K proposals of every chain per launch
        assuming rejections, as `sampler` does; every evaluation keeps the base-shuffle seed of its (iteration, chain), so
        the chains are identical for any K.  `log(i, rain, erod, lik)` is called per iteration only when K = 1.
        Returns dict(pos_rain, pos_erod, pos_likl [samples, nchains], accept_ratio[nchains])."""
        C_ = int(nchains or self.batch)
        S_ = self.samples
        K = max(1, int(speculate if speculate is not None else self.batch // C_))
        K = max(1, min(K, self.batch // C_))
        rs = [np.random.RandomState(np_seed + c) for c in range(C_)]
        pr = [random.Random(py_seed + c) for c in range(C_)]
        rain = np.array([r.uniform(self.rainlimits[0], self.rainlimits[1]) for r in rs])
        erod = np.array([r.uniform(self.erodlimits[0], self.erodlimits[1]) for r in rs])
        lik, _ = self.likelihood_batch(rain, erod)      # the reference evaluates the start twice (:569, :610)
        seed0 = self._evals + self.shuffle_seed * 1000003          # seed of (iteration 0, chain 0); (i, c) -> seed0 + i*C + c
        pos_rain = np.zeros((S_, C_)); pos_erod = np.zeros((S_, C_)); pos_likl = np.zeros((S_, C_))
        pos_rain[0], pos_erod[0], pos_likl[0] = rain, erod, lik
        naccept = np.zeros(C_, dtype=np.int64)
        noise = [[] for _ in range(C_)]                             # per chain: (d_rain, d_erod, u) per iteration drawn so far
        it = np.zeros(C_, dtype=np.int64)                           # next iteration of every chain

        def draw_upto(c, n):
            while len(noise[c]) < n:
                dr = rs[c].normal(0, self.step_rain); de = rs[c].normal(0, self.step_erod)
                for _ in range(3):          # eta random walk draws (bl_mcmc.py:670-677): consumed, never used (F8)
                    rs[c].normal(0, 1.0, 1)
                noise[c].append((dr, de, pr[c].uniform(0, 1)))

        while np.any(it < S_ - 1):
            pr_, pe_, owner = [], [], []
            for c in range(C_):
                n = int(min(K, S_ - 1 - it[c]))
                draw_upto(c, int(it[c]) + n)
                for j in range(n):
                    dr, de, _ = noise[c][int(it[c]) + j]
                    v = rain[c] + dr
                    p_r = rain[c] if (v < self.rainlimits[0] or v > self.rainlimits[1]) else v
                    v = erod[c] + de
                    p_e = erod[c] if (v < self.erodlimits[0] or v > self.erodlimits[1]) else v
                    pr_.append(p_r); pe_.append(p_e); owner.append((c, int(it[c]) + j))
            seeds = np.array([seed0 + i * C_ + c for c, i in owner], dtype=np.int64)
            lik_p, _ = self.likelihood_batch(np.array(pr_), np.array(pe_), seeds=seeds)
            stopped = np.zeros(C_, dtype=bool)
            for q, (c, i) in enumerate(owner):
                if stopped[c]:
                    continue
                diff = lik_p[q] - lik[c]
                try:
                    mh = min(1, math.exp(diff))
                except OverflowError:
                    mh = 1
                if noise[c][i][2] < mh:
                    lik[c], rain[c], erod[c] = lik_p[q], pr_[q], pe_[q]
                    naccept[c] += 1
                    stopped[c] = True                      # the rest of this chain's look-ahead assumed a rejection
                pos_rain[i + 1, c], pos_erod[i + 1, c], pos_likl[i + 1, c] = rain[c], erod[c], lik[c]
                it[c] = i + 1
            if log is not None and K == 1:
                log(int(it.min()) - 1, rain, erod, lik)
        self._evals = seed0 - self.shuffle_seed * 1000003 + (S_ - 1) * C_
        return dict(pos_rain=pos_rain, pos_erod=pos_erod, pos_likl=pos_likl, accept_ratio=naccept / max(1, S_ - 1) * 100.0)

    def sampler(self, speculate: Optional[int] = None):
        """Single random-walk Metropolis chain with the reference's draw order and output files
        (`bl_mcmc.py:515-835`): numpy's global stream for the start point, the proposals and the three unused eta
        draws, stdlib `random` for the accept test; `description.txt`, `exp_data.txt` (one `rain erod likl` row per
        sample, `storeParams` :457-480), `prediction_data/mean_pred_{elev,erdp,erdp_pts}_<t>.txt` (post burn-in running
        means, :776-815), `experiment_stats.txt` (RMSEelev / RMSEerdp / RMSEerdp_pts, accept ratio, count list, :817-831)
        and `prediction_data/pred_xslc.txt` / `pred_yslc.txt` (mid-domain slices per sample, :832-833).  Plots are out of scope.
        Returns (pos_rain, pos_erod, pos_likl, accept_ratio).

        `speculate=K` (default: the engine's batch size; 1 = strictly sequential evaluation) evaluates the next K proposals
        of the chain in one batched launch, assuming they are all rejected
        (prefetching: a rejection leaves the state, and therefore the following proposal, unchanged); at the first
        acceptance the rest of the batch is discarded and the chain goes on from the new state with the noise already
        drawn for those iterations.  Every random draw is made in the sequential order and every evaluation uses the
        base-shuffle seed of its iteration, so the chain, the files and the RNG states afterwards are identical to the
        sequential run - only ~1 / acceptance-rate times fewer launches (one chain otherwise keeps one SM busy)."""
        start = time.time()
        samples = self.samples
        fn = self.filename
        if fn:
            os.makedirs(os.path.join(fn, "prediction_data"), exist_ok=True)
        pos_rain = np.zeros(samples); pos_erod = np.zeros(samples)
        rain = np.random.uniform(self.rainlimits[0], self.rainlimits[1])
        erod = np.random.uniform(self.erodlimits[0], self.erodlimits[1])
        m, n = 0.5, 1.0
        # first evaluation of the start point (:569): the eta step sizes written to description.txt come from it (:571-585)
        _, ipts, ielev, ierdp = self.blackBox_batch([rain], [erod], grids=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            eta_elev = np.log(np.var(ielev[0, -1] - self.real_elev))
            eta_erdp = np.log(np.var(ierdp[0, -1] - self.real_erdp)) if self.real_erdp is not None else float("nan")
            eta_erdp_pts = np.log(np.var(ipts[0, -1] - self.real_erdp_pts))
        step_eta_elev, step_eta_erdp, step_eta_erdp_pts = np.abs(eta_elev * 0.02), np.abs(eta_erdp * 0.02), np.abs(eta_erdp_pts * 0.02)
        if fn:
            with open(os.path.join(fn, "description.txt"), "a") as f:                         # :589-605
                for k, v in (("samples", samples), ("step_rain", self.step_rain), ("step_erod", self.step_erod),
                             ("step_m", self.step_m), ("step_n", self.step_n), ("step_eta_elev", step_eta_elev),
                             ("step_eta_erdp", step_eta_erdp), ("step_eta_erdp_pts", step_eta_erdp_pts),
                             ("Initial_proposed_rain", rain),
                             ("Initial_proposed_erod", erod), ("Initial_proposed_m", m), ("Initial_proposed_n", n),
                             ("erod_limits", self.erodlimits), ("rain_limits", self.rainlimits),
                             ("m_limit", self.mlimit), ("n_limit", self.nlimit)):
                    f.write("\n\t{0}: {1}".format(k, v))

        def evaluate(r, e):
            sse, pts, elev, erdp = self.blackBox_batch([r], [e], grids=True)
            lik, _ = log_likelihood(sse, self.G, pts, self.real_erdp_pts, self.likl_sed)
            return float(lik[0]), elev[0], erdp[0], pts[0]

        def store(r, e, l):
            if fn:
                with open(os.path.join(fn, "exp_data.txt"), "a") as f:
                    f.write("{0} {1} {2} \n".format(r, e, l))

        likelihood, elev, erdp, pts = evaluate(rain, erod)     # second evaluation of the start (:610)
        pos_likl = np.full(samples, likelihood)              # np.zeros(samples, likelihood) in the reference (F11)
        prev = (elev, erdp, pts)
        store(pos_rain[0], pos_erod[0], pos_likl[0])
        sum_elev, sum_erdp, sum_pts = elev.copy(), erdp.copy(), pts.copy()
        burnsamples = int(samples * self.burn_in)
        num_div = 0
        accept_counter = 0
        count_list = [0]                                     # sample numbers of the accepted proposals (:540, :631, :703)
        list_yslicepred = np.zeros((samples, self.real_elev.shape[0]))    # mid-domain slices of the accepted topography (:534-537)
        list_xslicepred = np.zeros((samples, self.real_elev.shape[1]))
        ymid, xmid = int(self.real_elev.shape[1] / 2), int(self.real_elev.shape[0] / 2)
        K = max(1, min(int(self.batch if speculate is None else speculate), self.batch))
        seed0 = self._evals + self.shuffle_seed * 1000003                       # base-shuffle seed of iteration 0
        nz_r, nz_e, us = [], [], []                                             # noise / uniforms drawn so far, per iteration

        def draw_upto(n):
            while len(us) < n:
                nz_r.append(np.random.normal(0, self.step_rain))
                nz_e.append(np.random.normal(0, self.step_erod))
                for _ in range(3):
                    np.random.normal(0, 1.0, 1)              # eta walk draws, consumed and unused (:670-677, F8)
                us.append(random.uniform(0, 1))

        def propose(j):
            p_rain = rain + nz_r[j]
            if p_rain < self.rainlimits[0] or p_rain > self.rainlimits[1]:
                p_rain = rain
            p_erod = erod + nz_e[j]
            if p_erod < self.erodlimits[0] or p_erod > self.erodlimits[1]:
                p_erod = erod
            return p_rain, p_erod

        i = 0
        while i < samples - 1:
            n = min(K, samples - 1 - i)
            if K == 1:
                # the sequential order of the reference: draw, evaluate, then the uniform
                nz_r.append(np.random.normal(0, self.step_rain)); nz_e.append(np.random.normal(0, self.step_erod))
                for _ in range(3):
                    np.random.normal(0, 1.0, 1)
                props = [propose(i)]
                sse, bpts, belev, berdp = self.blackBox_batch([props[0][0]], [props[0][1]], grids=True, seeds=[seed0 + i])
                us.append(random.uniform(0, 1))
            else:
                draw_upto(i + n)
                props = [propose(i + j) for j in range(n)]
                sse, bpts, belev, berdp = self.blackBox_batch([q[0] for q in props], [q[1] for q in props], grids=False,
                                                              seeds=seed0 + i + np.arange(n))
            liks, _ = log_likelihood(sse, self.G, bpts, self.real_erdp_pts, self.likl_sed)
            used = n
            for j in range(n):
                likelihood_proposal = float(liks[j])
                diff_likelihood = likelihood_proposal - likelihood
                try:
                    mh_prob = min(1, math.exp(diff_likelihood))
                except OverflowError:
                    mh_prob = 1
                it = i + j
                accepted = us[it] < mh_prob
                if accepted:
                    likelihood, rain, erod = likelihood_proposal, props[j][0], props[j][1]
                    pos_rain[it + 1], pos_erod[it + 1], pos_likl[it + 1] = rain, erod, likelihood
                    if K == 1:
                        prev = (belev[0], berdp[0], bpts[0])
                    else:                                    # prediction grids of the accepted proposal (same seed)
                        _, p1, e1, d1 = self.blackBox_batch([rain], [erod], grids=True, seeds=[seed0 + it])
                        prev = (e1[0], d1[0], p1[0])
                    accept_counter += 1
                    count_list.append(it)
                    final_predtopo = prev[0][-1]
                    list_yslicepred[it + 1, :] = final_predtopo[:, ymid]
                    list_xslicepred[it + 1, :] = final_predtopo[xmid, :]
                else:
                    pos_rain[it + 1], pos_erod[it + 1], pos_likl[it + 1] = pos_rain[it], pos_erod[it], pos_likl[it]
                    list_yslicepred[it + 1, :] = list_yslicepred[it, :]
                    list_xslicepred[it + 1, :] = list_xslicepred[it, :]
                store(pos_rain[it + 1], pos_erod[it + 1], pos_likl[it + 1])
                if it > burnsamples:
                    sum_elev += prev[0]; sum_erdp += prev[1]; sum_pts += prev[2]
                    num_div += 1
                if accepted:
                    used = j + 1                             # the rest of the batch assumed a rejection here
                    break
            i += used
        self._evals = seed0 - self.shuffle_seed * 1000003 + samples - 1
        accept_ratio = accept_counter / (samples * 1.0) * 100
        div = max(num_div, 1)
        mean_elev, mean_erdp, mean_pts = sum_elev / div, sum_erdp / div, sum_pts / div
        # RMSE of the mean predictions at the final time (:783, :791, :801) - same numpy broadcasting as the reference
        self.rmse_elev = float(np.sqrt(np.sum(np.square(mean_elev[-1] - self.real_elev)) / self.real_elev.size))
        self.rmse_erdp = (float(np.sqrt(np.sum(np.square(mean_erdp[-1] - self.real_erdp)) / np.size(self.real_erdp)))
                          if self.real_erdp is not None else float("nan"))
        self.rmse_erdp_pts = float(np.sqrt(np.sum(np.square(mean_pts[-1] - self.real_erdp_pts)) / self.real_erdp_pts.size))
        self.count_list = count_list
        if fn:
            for x, T in enumerate(self.sim_interval):
                np.savetxt(os.path.join(fn, "prediction_data", "mean_pred_elev_%s.txt" % T), mean_elev[x], fmt="%.5f")
                np.savetxt(os.path.join(fn, "prediction_data", "mean_pred_erdp_%s.txt" % T), mean_erdp[x], fmt="%.5f")
                np.savetxt(os.path.join(fn, "prediction_data", "mean_pred_erdp_pts_%s.txt" % T), mean_pts[x], fmt="%.5f")
            np.savetxt(os.path.join(fn, "prediction_data", "mean_pred_erdp_pts.txt"), mean_pts, fmt="%.5f")
            total_time = time.time() - start
            accepted_count = len(count_list)                 # the reference counts the start point too (:817-819)
            with open(os.path.join(fn, "experiment_stats.txt"), "w") as f:                    # :826-831
                f.write("RMSEelev: {0}\nRMSEerdp: {1}\nRMSEerdp_pts: {2}\nTime:(s) {3}\nTime:(mins) {4}\n".format(
                    self.rmse_elev, self.rmse_erdp, self.rmse_erdp_pts, total_time, total_time / 60.0))
                f.write("Accept ratio: {0} %\nSamples accepted : {1} out of {2}\n Count List : {3} ".format(
                    accepted_count / (samples * 1.0) * 100, accepted_count, samples, count_list))
                f.write("Time Elapsed: (s) {0} , (mins): {1}".format(total_time, total_time / 60.0))
            np.savetxt(os.path.join(fn, "prediction_data", "pred_xslc.txt"), list_xslicepred)     # :832-833
            np.savetxt(os.path.join(fn, "prediction_data", "pred_yslc.txt"), list_yslicepred)
        return pos_rain, pos_erod, pos_likl, accept_ratio
