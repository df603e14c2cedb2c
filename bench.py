#!/usr/bin/env python
"""bench.py - MCMC forward-model evaluations per second on the BASELINE.json workload.

A "step" is one pass of the hot path over one batch of B synthetic (rain, erodibility) proposals:
`blackBox` (full pyBadlands run to simtime with the 5 sim_interval grids) + `likelihoodFunc`.
Workload at N=1: BASELINE.json configs[1] = Examples/crater, full resolution (T = 50 000 a), batched
proposals on one B200.  Proposals shard across GPUs (weak scaling, B per GPU); the only exchange is the
NCCL all_gather of the B log-likelihoods per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the reference cannot run here:
Python 2 + gfortran + Triangle) on all host cores for the same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (npz, simtime, rain limits, erod limits, erdp coords, likl_sed)
    "crater": ("crater", 50000, (0.0, 3.0), (3.e-5, 7.e-5),
               [[60, 60], [52, 67], [74, 76], [62, 45], [72, 66], [85, 73], [90, 75], [44, 86], [100, 80], [88, 69]], True),
    "crater_fast": ("crater_fast", 15000, (0.0, 3.0), (3.e-5, 7.e-5),
                    [[60, 60], [72, 66], [85, 73], [90, 75], [44, 86], [100, 80], [88, 69], [79, 91], [96, 77], [42, 49]], True),
    "etopo_fast": ("etopo_fast", 500000, (0.0, 3.0), (3.e-6, 7.e-6),
                   [[42, 10], [39, 8], [75, 51], [59, 13], [40, 5], [6, 20], [14, 66], [4, 40], [68, 40], [72, 44]], True),
    "etopo": ("etopo", 1000000, (0.0, 3.0), (3.e-6, 7.e-6),
              [[42, 10], [39, 8], [75, 51], [59, 13], [40, 5], [6, 20], [14, 66], [4, 40], [72, 73], [46, 64]], True),
}
B_ALG_NODE_STEP = 340.0   # algorithmic bytes / node / time step / proposal
B_ALG_CELL = 160.0        # bytes / grid cell / evaluation


def load_config(name):
    from bayeslands_b200 import config, mesh
    npz, simtime, rl, el, coords, likl_sed = CONFIGS[name]
    # every array read now: a lazily read NpzFile would share one file descriptor (and its offset) between the forked pool
    # workers of run_reference, and concurrent reads then corrupt each other's zlib streams
    with np.load(os.path.join(ROOT, "data", npz + ".npz")) as f:
        d = {k: f[k] for k in f.files}
    inp = config.parse_xml(str(d["xml"]))
    m = mesh.build_from_dem(d["dem_xyz"], inp.btype, inp.Afactor)
    return d, inp, m, simtime, rl, el, np.array(coords), likl_sed


def proposals(step, rank, B, rl, el):
    rng = np.random.default_rng(5678 + 1000 * (step + 100000) + rank)
    return rng.uniform(rl[0], rl[1], B), rng.uniform(el[0], el[1], B), rng.integers(1, 2 ** 31, B).astype(np.int64)


# ---------------------------------------------------------------------------------------------------
# CPU oracle legs (the only places bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(cfg):
    from oracle import oracle
    d, inp, m, simtime, rl, el, coords, likl_sed = load_config(cfg)
    _W.update(om=oracle.OracleModel(m, inp, sealevel=d["sealevel"] if "sealevel" in d else None), d=d, simtime=simtime, coords=coords, likl_sed=likl_sed, oracle=oracle)


def _cpu_eval(args):
    rain, erod, seed = args
    om, d = _W["om"], _W["d"]
    ev, ed, pts = om.blackbox(rain, erod, _W["simtime"], _W["coords"], seed=int(seed))
    T = sorted(ev)[-1]
    lik, _ = _W["oracle"].likelihood(ev[T], d["final_elev"], np.array([pts[k] for k in sorted(pts)]),
                                     d["final_erdp_pts"], _W["likl_sed"])
    return lik, om.steps


def cpu_baseline(cfg, budget_s=12.0):
    """Oracle, one process, evaluations one after another (faithful to `python bl_mcmc.py`, 1 rank)."""
    _cpu_init(cfg)
    _, _, _, _, rl, el, _, _ = load_config(cfg)
    r, e, s = proposals(10 ** 6, 0, 64, rl, el)
    t0 = time.perf_counter()
    n = 0
    while n < len(r) and (n < 2 or time.perf_counter() - t0 < budget_s):
        _cpu_eval((r[n], e[n], s[n]))
        n += 1
    dt = time.perf_counter() - t0
    # context: the reference rebuilds the mesh in every blackBox call (bl_mcmc.py:126-129, F13); the oracle does not
    from bayeslands_b200 import mesh as meshmod
    d, inp = _W["d"], _W["om"].inp
    t1 = time.perf_counter()
    meshmod.build_from_dem(d["dem_xyz"], inp.btype, inp.Afactor)
    t_mesh = time.perf_counter() - t1
    return dict(value=n / dt, unit="evals/s", cores=1, kind="port",
                sample="%d %s evaluations, CPU oracle (C restatement, -O2), 1 process" % (n, cfg),
                mesh_rebuild_s=t_mesh, value_with_mesh_rebuild_per_eval=1.0 / (dt / n + t_mesh))


def run_reference(args):
    """--impl reference: the oracle port on every host core, bounded sample per step."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    ncores = os.cpu_count() or 1
    _, _, m, simtime, rl, el, _, _ = load_config(cfg)
    per_step = max(ncores, 8)
    # the oracle library is loaded (and exercised once) in THIS process too, so that the record of native objects the
    # driver keeps for the reference arm shows oracle/libbl_oracle.so; the pool workers are forks of it
    _cpu_init(cfg)
    r0, e0, s0 = proposals(-1000, 0, 1, rl, el)
    _cpu_eval((r0[0], e0[0], s0[0]))
    ctx = mp.get_context("fork")
    nsteps_seen = []
    with ctx.Pool(ncores) as pool:
        def one(step):
            r, e, s = proposals(step, 0, per_step, rl, el)
            return pool.map(_cpu_eval, list(zip(r, e, s)), chunksize=1)
        for w in range(args.warmup):
            one(-1 - w)
        t0 = time.perf_counter()
        for k in range(args.steps):
            nsteps_seen += [x[1] for x in one(k)]
        dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    mean_steps = float(np.mean(nsteps_seen)) if nsteps_seen else 0.0
    line = dict(impl="reference", metric="mcmc_forward_model_evals_per_sec", value=val, unit="evals/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=dt / args.steps * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="Examples/%s full blackBox+likelihood, T=%d a, N=%d TIN nodes, G=%d grid cells"
                                     % (cfg, simtime, m.N, m.grid_nx * m.grid_ny),
                            proposals_per_gpu_per_step=per_step, mean_time_steps_per_eval=mean_steps,
                            node_steps_per_s=val * mean_steps * m.N, l2_policy="n/a (host CPU arm)",
                            parallelism="%d host processes, one evaluation each" % ncores, status_bits=0),
                reference_is_oracle_port=True,
                cpu_baseline=dict(value=val, unit="evals/s", cores=ncores, kind="port",
                                  sample="%d evaluations per step over %d processes, CPU oracle = C restatement of the "
                                         "reference path, -O2, sqrt for Q**0.5 (the reference calls libm pow), one mesh build "
                                         "for all evaluations (the reference rebuilds it per proposal); the reference itself "
                                         "cannot run here: py2 + gfortran + Triangle" % (per_step, ncores)),
                e2e=dict(value=val, unit="evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bayeslands_b200.mcmc import bayeslands_mcmc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = args.config
    d, inp, m, simtime, rl, el, coords, likl_sed = load_config(cfg)
    B = args.batch
    mc = bayeslands_mcmc(True, simtime, 0, d["final_elev"], d["final_erdp"], d["final_erdp_pts"], coords, None, inp,
                         el, rl, (0.4, 0.6), (0.9, 1.1), "bench", likl_sed, batch=B, device=local, mesh=m,
                         sealevel=d["sealevel"] if "sealevel" in d else None)
    eng = mc.engine
    dev = eng.tdev
    gather = torch.empty(world * B, dtype=torch.float64, device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (value): inputs already in HBM when the timed region starts
    def dev_step(step, ev0=None, ev1=None):
        r, e, s = proposals(step, rank, B, rl, el)
        dr = torch.as_tensor(r, device=dev); de = torch.as_tensor(e, device=dev); ds = torch.as_tensor(s, device=dev)
        torch.cuda.synchronize(dev)
        return dr, de, ds

    staged = [dev_step(k) for k in range(args.steps)]
    for w in range(args.warmup):
        dr, de, ds = dev_step(-1 - w)
        eng.reset(dr, de, ds); eng.run_schedule(mc.sim_interval)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = eng.launches
    steps_total = 0
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        dr, de, ds = staged[k]
        eng.reset(dr, de, ds)
        kev[k][0].record()
        pts, sse, _, _ = eng.run_schedule(mc.sim_interval)
        kev[k][1].record()
        if world > 1:
            dist.all_gather_into_tensor(gather, sse)
    t_end.record()
    barrier()
    ms = t_start.elapsed_time(t_end)
    kernel_ms = [a.elapsed_time(b) for a, b in kev]
    launches = eng.launches - launches0
    sc = eng.scalars()
    steps_last = sc["steps"].astype(np.float64)           # time steps per evaluation of the last batch
    status = int(np.bitwise_or.reduce(sc["status"]))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end-to-end through the public API (host buffers in, host likelihoods out)
    for w in range(max(1, args.warmup // 2)):
        r, e, s = proposals(-100 - w, rank, B, rl, el)
        mc.likelihood_batch(r, e, seeds=s)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        r, e, s = proposals(500 + k, rank, B, rl, el)
        lik, _ = mc.likelihood_batch(r, e, seeds=s)
        if world > 1:
            dist.all_gather_into_tensor(gather, torch.as_tensor(lik, device=dev))
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if sampler else None
    e2e_val = world * B * args.steps / e2e_s
    h2d = B * (8 + 8 + 8)
    d2h = B * 8 + B * len(mc.sim_interval) * len(coords) * 8

    if rank == 0:
        # roofline of the dominant (only) kernel, bl_run_kernel: algorithmic bytes of the last launch / its duration
        alg_bytes = float(np.sum(steps_last) * m.N * B_ALG_NODE_STEP + B * B_ALG_CELL * eng.G)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_bytes / (kernel_ms[-1] * 1e-3) / 1e9
        # DRAM traffic of this launch, scaled from the committed ncu --set full capture (bytes per node-step there
        # times the node-steps of this launch); null when the capture is for another config
        traffic, traffic_source = None, "none: no ncu --set full capture for this config"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if cfg == "crater" and tj.get("N") == m.N:
                traffic = float(tj["dram_bytes_per_node_step"] * np.sum(steps_last) * m.N)
                traffic_source = ("not measured in this run: dram__bytes_read+write per node-step of the committed ncu --set full "
                                  "capture (profiles/ncu_traffic.json, %s) x the node-steps of this launch" % tj.get("capture", "?"))
        except Exception:
            pass
        cpu = cpu_baseline(cfg) if world == 1 and not args.no_cpu else None
        line = dict(metric="mcmc_forward_model_evals_per_sec", value=value, unit="evals/s", n_gpus=world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload="Examples/%s full blackBox+likelihood, T=%d a, N=%d TIN nodes, G=%d grid cells"
                                         % (cfg, simtime, m.N, eng.G),
                                proposals_per_gpu_per_step=B, mean_time_steps_per_eval=float(np.mean(steps_last)),
                                node_steps_per_s=value * float(np.mean(steps_last)) * m.N,
                                l2_policy="per-step working set %.0f MB > L2 (fresh proposals every step)"
                                          % (eng.workspace.numel() / 1e6),
                                parallelism="proposals sharded, %d per GPU" % B, status_bits=status,
                                launch=eng.info()),
                    e2e=dict(value=e2e_val, unit="evals/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=int(launches),
                    roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                                  traffic=traffic, traffic_source=traffic_source, kernel="bl_run_kernel",
                                  alg_bytes_per_node_step=B_ALG_NODE_STEP, kernel_ms=kernel_ms[-1],
                                  peak_source="MEASURED_PEAKS.json" if peaks else "fallback",
                                  share_of_step=sum(kernel_ms) / ms),
                    clocks=clocks)
        if cpu:
            line["cpu_baseline"] = cpu
        if cfg == "crater" and not args.no_other:
            # the other BASELINE configs, measured after the headline's timed region (short batches; same JSON line)
            del mc, eng
            torch.cuda.empty_cache()
            line["config"]["other_workloads"] = other_workloads(local, peak, not args.no_cpu and world == 1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _resident_rate(cfg, B, device, steps=2, warmup=1):
    """evals/s of `cfg` with inputs resident in HBM (same protocol as the headline), mean time steps, roofline fraction inputs."""
    import torch
    from bayeslands_b200.mcmc import bayeslands_mcmc
    d, inp, m, simtime, rl, el, coords, likl_sed = load_config(cfg)
    mc = bayeslands_mcmc(True, simtime, 0, d["final_elev"], d["final_erdp"], d["final_erdp_pts"], coords, None, inp,
                         el, rl, (0.4, 0.6), (0.9, 1.1), "bench", likl_sed, batch=B, device=device, mesh=m,
                         sealevel=d["sealevel"] if "sealevel" in d else None)
    eng = mc.engine
    dev = eng.tdev
    st = []
    for k in range(-warmup, steps):
        r, e, s = proposals(7000 + k, 0, B, rl, el)
        st.append((torch.as_tensor(r, device=dev), torch.as_tensor(e, device=dev), torch.as_tensor(s, device=dev)))
    torch.cuda.synchronize(dev)
    for k in range(warmup):
        eng.reset(*st[k]); eng.run_schedule(mc.sim_interval)
    torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(warmup, warmup + steps):
        eng.reset(*st[k]); eng.run_schedule(mc.sim_interval)
    t1.record()
    torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1)
    sc = eng.scalars()
    return mc, dict(value=B * steps / (ms * 1e-3), unit="evals/s", proposals_per_step=B, steps=steps,
                    mean_time_steps_per_eval=float(sc["steps"].mean()), N=m.N, G=eng.G,
                    status_bits=int(np.bitwise_or.reduce(sc["status"])),
                    alg_bytes_last=float(sc["steps"].sum() * m.N * B_ALG_NODE_STEP + B * B_ALG_CELL * eng.G), ms_last=ms / steps)


def other_workloads(device, peak, with_cpu):
    """Short measurements of the BASELINE configs the headline does not cover: etopo_fast (marine path), a crater_fast
    single chain with speculative batching, a small synthetic lattice.  Each: value, mean steps, roofline fraction, CPU leg."""
    import torch
    out = {}
    # ---- Examples/etopo_fast: sea-level curve, marine deposition + diffusion (BASELINE configs[2])
    mc, r = _resident_rate("etopo_fast", 888, device)
    r["roofline_frac"] = r.pop("alg_bytes_last") / (r.pop("ms_last") * 1e-3) / 1e9 / peak
    if with_cpu:
        r["cpu_baseline"] = cpu_baseline("etopo_fast", budget_s=3.0)
    out["etopo_fast"] = r
    del mc
    torch.cuda.empty_cache()
    # ---- Examples/crater_fast: ONE Metropolis chain, 300 samples, speculative batches of 148 proposals (configs[0])
    mc, r = _resident_rate("crater_fast", 148, device, steps=1, warmup=1)
    mc.samples = 300
    mc.filename = None
    np.random.seed(0)
    import random
    random.seed(0)
    l0 = mc.engine.launches
    t0 = time.perf_counter()
    pos_rain, pos_erod, pos_likl, acc = mc.sampler(speculate=148)
    dt = time.perf_counter() - t0
    r["roofline_frac"] = r.pop("alg_bytes_last") / (r.pop("ms_last") * 1e-3) / 1e9 / peak
    r["single_chain"] = dict(value=(mc.samples - 1) / dt, unit="chain iterations/s", samples=mc.samples, speculate=148,
                             accept_ratio_percent=float(acc), launches=int(mc.engine.launches - l0))
    if with_cpu:
        r["cpu_baseline"] = cpu_baseline("crater_fast", budget_s=3.0)
    out["crater_fast"] = r
    del mc
    torch.cuda.empty_cache()
    # ---- synthetic regular-grid landscape on the lattice engine, small lattice (configs[4] in miniature; the full
    #      2048 x 2048 x 256 run is `bench.py --config synthetic`)
    from bayeslands_b200 import synthetic
    from bayeslands_b200.grid_engine import GridEngine
    nx, T, B = 256, 50000.0, 16
    lat, inp = synthetic.build(nx, nx, T=T)
    eng = GridEngine(lat, inp, B, device=device, real_elev=lat.z0 + lat.disp * T)
    times = [T / 4, T / 2, 3 * T / 4, T]
    rr, ee = synthetic.proposals(B, offset=424242)
    dr, de = torch.as_tensor(rr, device=eng.tdev), torch.as_tensor(ee, device=eng.tdev)
    eng.reset(dr, de); eng.run_schedule(times)
    torch.cuda.synchronize(eng.tdev)
    l0 = eng.launches
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record()
    eng.reset(dr, de); eng.run_schedule(times)
    t1e.record()
    torch.cuda.synchronize(eng.tdev)
    ms = t0e.elapsed_time(t1e)
    sc = eng.scalars()
    N = (nx + 2) * (nx + 2)
    syn = dict(value=B / (ms * 1e-3), unit="evals/s", proposals_per_step=B, steps=1, lattice="%dx%d" % (nx, nx), N=N,
               mean_time_steps_per_eval=float(sc["steps"].mean()),
               mean_fill_sweeps_per_step=float(sc["fill_sweeps"].sum() / max(1.0, float(sc["steps"].sum()))),
               launches=int(eng.launches - l0),
               roofline_frac=float(sc["steps"].sum()) * N * 157.0 / (ms * 1e-3) / 1e9 / peak)
    if with_cpu:
        per_step = _synthetic_cpu(nx, nx, T)
        syn["cpu_baseline"] = dict(value=1.0 / (per_step * syn["mean_time_steps_per_eval"]), unit="evals/s", cores=1, kind="port",
                                   sample="2 time steps of one proposal on the same lattice, CPU oracle, scaled by the steps of an evaluation")
    out["synthetic_%d" % nx] = syn
    return out


# ---------------------------------------------------------------------------------------------------
# BASELINE config 5: synthetic regular-grid landscape on the lattice engine (optional: --config synthetic)
# ---------------------------------------------------------------------------------------------------
def _synthetic_cpu(nx, ny, T, nsteps=2):
    """Oracle on the same lattice: `nsteps` time steps of one proposal (a full evaluation takes hours on one core)."""
    from bayeslands_b200 import synthetic, mesh, forcing
    from oracle import oracle
    inp = synthetic.inputs(T)
    m = mesh.build_regular(nx, ny, 100.0, synthetic.surface(nx, ny), "fixed", order="colour")
    rate = np.zeros(m.N)
    rate[m.extra["regular"]["ext"][1:-1, 1:-1].ravel()] = synthetic.uplift_rate(nx, ny).ravel()
    disp = forcing.disp_border(np.where(np.arange(m.N) < m.boundsPt, -1.0e6, rate), m.FVmesh.neighbours,
                               m.FVmesh.edge_length, m.boundsPt)
    om = oracle.OracleModel(m, inp, seed=1, disp=disp)
    r, e = synthetic.proposals(1)
    om.reset(r[0], e[0], seed=1)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        om.streamflow(); om.sediment_flux(T)
    return (time.perf_counter() - t0) / nsteps


def run_synthetic(args):
    import torch
    import torch.distributed as dist
    from bayeslands_b200 import synthetic
    from bayeslands_b200.grid_engine import GridEngine
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = ny = args.nx
    T = float(args.simtime)
    B = args.batch
    lat, inp = synthetic.build(nx, ny, T=T)
    # observed topography = the initial surface lifted by the full uplift (any fixed grid does for the SSE reduction)
    real = lat.z0 + lat.disp * T
    eng = GridEngine(lat, inp, B, device=local, real_elev=real)
    dev = eng.tdev
    times = [T / 4, T / 2, 3 * T / 4, T]
    gather = torch.empty(world * B, dtype=torch.float64, device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one(step, host=False):
        r, e = synthetic.proposals(B, offset=1000 * (step + 100000) + rank)
        if not host:
            r, e = torch.as_tensor(r, device=dev), torch.as_tensor(e, device=dev)
        eng.reset(r, e)
        _, sse, _, _ = eng.run_schedule(times)
        if world > 1:
            dist.all_gather_into_tensor(gather, sse)
        return sse.cpu().numpy() if host else sse

    for w in range(args.warmup):
        one(-1 - w)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.launches
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record()
    for k in range(args.steps):
        one(k)
    t1e.record()
    barrier()
    ms = t0e.elapsed_time(t1e)
    launches = eng.launches - l0
    sc = eng.scalars()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)
    t0 = time.perf_counter()
    for k in range(args.steps):
        one(500 + k, host=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        N = (nx + 2) * (ny + 2)
        S = sc["steps"].astype(np.float64)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # erosive configuration: fill 16 + SFD 9 + donors 2 + discharge 24 + CFL 17 + erosion 41 + diffusion 48 B/node/step
        alg = 157.0
        achieved = float(np.sum(S)) * N * alg * args.steps / (ms * 1e-3) / 1e9
        line = dict(metric="mcmc_forward_model_evals_per_sec", value=value, unit="evals/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic",
                    config=dict(workload="synthetic %dx%d regular grid, uplift + rain, dep=0, T=%d a, N=%d nodes (lattice engine)"
                                         % (nx, ny, T, N),
                                proposals_per_gpu_per_step=B, mean_time_steps_per_eval=float(np.mean(S)),
                                mean_fill_sweeps_per_step=float(sc["fill_sweeps"].sum() / max(1.0, S.sum())),
                                l2_policy="per-proposal state %.0f MB x %d > L2" % (N * 43 / 1e6, B),
                                parallelism="proposals sharded, %d per GPU" % B),
                    e2e=dict(value=world * B * args.steps / e2e_s, unit="evals/s", h2d_bytes_per_step=B * 16, d2h_bytes_per_step=B * 8),
                    gpu_launches=int(launches),
                    roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None,
                                  kernel="all lattice kernels (k_fill_pass dominates)", alg_bytes_per_node_step=alg,
                                  peak_source="MEASURED_PEAKS.json" if peaks else "fallback"),
                    clocks=clocks)
        if world == 1 and not args.no_cpu:
            per_step = _synthetic_cpu(nx, ny, T)
            line["cpu_baseline"] = dict(value=1.0 / (per_step * float(np.mean(S))), unit="evals/s", cores=1, kind="port",
                                        sample="2 time steps of one proposal on the same lattice, CPU oracle, scaled by the "
                                               "%.0f steps of an evaluation" % float(np.mean(S)))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    # 592 = 4 landscapes per SM: two resident (compact layout), two queued behind them, so that an SM whose landscapes
    # needed fewer adaptive time steps takes the next one instead of idling until the slowest of a single wave is done
    # (measured: 296 -> 1 140, 592 -> 1 155, 888 -> 1 156 evals/s)
    ap.add_argument("--batch", type=int, default=592)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="crater", choices=sorted(CONFIGS) + ["synthetic"])
    ap.add_argument("--nx", type=int, default=2048, help="synthetic config: lattice size (nx = ny)")
    ap.add_argument("--simtime", type=float, default=50000.0, help="synthetic config: simulated years")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-other", action="store_true", help="skip the short measurements of the other BASELINE configs")
    args = ap.parse_args()
    if args.config == "synthetic":
        if args.impl == "reference":
            raise SystemExit("--impl reference is defined on the BASELINE metric config (crater); the synthetic config "
                             "reports its CPU oracle leg as cpu_baseline")
        run_synthetic(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
