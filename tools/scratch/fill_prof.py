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
n = per(15)
print("fill total cyc/step %.0f ; passes(rounds)/step %.1f" % (per(0), n))
print("per round (thread 0, incl. barrier): a1 %.0f ; a2 %.0f ; b %.0f ; sum %.0f" % (per(12) / n, per(13) / n, per(14) / n, (per(12)+per(13)+per(14))/n))
print("per round max over threads: b own %.0f ; ids load %.0f ; W loads+min %.0f ; update+atomics %.0f" % (per(1) / n, per(3) / n, per(5) / n, per(7) / n))
