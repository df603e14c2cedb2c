import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, torch
from bayeslands_b200 import hostparams
from bayeslands_b200.engine import Engine
cfg = sys.argv[1] if len(sys.argv) > 1 else "crater"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
d, inp, m, simtime, rl, el, coords, likl_sed = bench.load_config(cfg)
eng = Engine(m, inp, B, real_elev=d["final_elev"], erdp_coords=coords)
r, e, s = bench.proposals(0, 0, B, rl, el)
eng.reset(r, e, s)
eng.run_schedule(hostparams.sim_interval(simtime))
torch.cuda.synchronize()
pc = eng.phase_cycles().astype(np.float64)
steps = eng.scalars()["steps"].astype(np.float64)
per = lambda i: (pc[:, i] / steps).mean()
npass = per(14); items = per(15)
print("fill total cyc/step %.0f ; passes/step %.1f ; items/pass %.1f" % (per(0), npass, items / npass))
print("per pass: (a)+barrier %.0f ; (b)+barrier %.0f" % (per(12) / npass, per(13) / npass))
print("thread 0 per pass: search+select %.0f ; W,z test + ids load %.0f ; 8 W loads + min %.0f ; update+atomics %.0f" % (
    per(1) / npass, per(3) / npass, per(5) / npass, per(7) / npass))
