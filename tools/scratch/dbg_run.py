"""Debug helper: a few time steps of one config (for compute-sanitizer)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, torch
from bayeslands_b200.engine import Engine
cfg = sys.argv[1]; B = int(sys.argv[2]); nsteps = int(sys.argv[3])
d, inp, m, simtime, rl, el, coords, likl_sed = bench.load_config(cfg)
eng = Engine(m, inp, B, real_elev=d["final_elev"], erdp_coords=coords, sealevel=d["sealevel"] if "sealevel" in d else None)
r, e, s = bench.proposals(0, 0, B, rl, el)
eng.reset(r, e, s)
for i in range(nsteps):
    eng.step(float(simtime))
    torch.cuda.synchronize()
    print("step", i, eng.scalars()["tNow"][:2], flush=True)
