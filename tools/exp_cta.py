"""Experiment (run on the GPU box): what does a second resident landscape per SM buy?

A crater mesh coarse enough (N ~ 5 600) for two CTAs' shared-memory working sets per SM, evaluated with
  1024 threads x 1 CTA/SM,  512 threads x 1 CTA/SM (dynamic shared memory forced to the maximum),  512 x 2,  256 x 2.
Same proposals, same results (bit-exact device code at every CTA size); only the residency changes.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayeslands_b200 import hostparams, mesh
from bayeslands_b200.engine import Engine

A = float(sys.argv[1]) if len(sys.argv) > 1 else 2.5
B = int(sys.argv[2]) if len(sys.argv) > 2 else 592
d, inp, m0, simtime, rl, el, coords, likl_sed = bench.load_config("crater")
m = mesh.build_from_dem(d["dem_xyz"], inp.btype, A)
r, e, s = bench.proposals(0, 0, B, rl, el)
times = hostparams.sim_interval(simtime)
out = dict(N=m.N, B=B, Afactor=A, runs=[])
ref = None
for tpb, dbg, label in ((1024, 0, "1024 x 1"), (512, 1, "512 x 1"), (512, 0, "512 x 2"), (256, 0, "256 x 2")):
    eng = Engine(m, inp, B, real_elev=d["final_elev"], erdp_coords=coords, tpb=tpb, debug=dbg)
    best = 1e30
    for it in range(3):
        eng.reset(r, e, s)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, sse, _, _ = eng.run_schedule(times)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    sse = sse.cpu().numpy()
    if ref is None:
        ref = sse
    pc = eng.phase_cycles().astype(np.float64)
    steps = eng.scalars()["steps"]
    out["runs"].append(dict(cfg=label, evals_per_s=B / best, ms=best * 1e3, same_sse=bool(np.array_equal(sse, ref)),
                            cycles_per_step=float((pc[:, :12].sum(1) / steps).mean()), mean_steps=float(steps.mean())))
    print(out["runs"][-1], flush=True)
    del eng
    torch.cuda.empty_cache()
print(json.dumps(out))
