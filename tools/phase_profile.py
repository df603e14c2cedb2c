"""Per-phase cycle breakdown of the device time step (diagnostics; run on the GPU box)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayeslands_b200 import _lib
from bayeslands_b200.engine import Engine
import torch

cfg = sys.argv[1] if len(sys.argv) > 1 else "crater_fast"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
TPB = int(sys.argv[3]) if len(sys.argv) > 3 else 0
DBG = int(sys.argv[4]) if len(sys.argv) > 4 else 0
COMPACT = {"auto": None, "on": True, "off": False}[sys.argv[5] if len(sys.argv) > 5 else "auto"]
d, inp, m, simtime, rl, el, coords, likl_sed = bench.load_config(cfg)
eng = Engine(m, inp, B, real_elev=d["final_elev"], erdp_coords=coords, sealevel=d["sealevel"] if "sealevel" in d else None,
             tpb=TPB, debug=DBG, compact=COMPACT)
r, e, s = bench.proposals(0, 0, B, rl, el)
from bayeslands_b200 import hostparams
for it in range(2):
    eng.reset(r, e, s)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    eng.run_schedule(hostparams.sim_interval(simtime))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
pc = eng.phase_cycles().astype(np.float64)
steps = eng.scalars()["steps"]
tot = pc.sum(1)
print("config", cfg, "B", B, "tpb", TPB, "debug", DBG, "evals/s %.1f" % (B / dt), "wall ms", dt * 1e3, "mean steps", steps.mean())
print("mean cycles/step per phase:")
for i, nm in enumerate(_lib.PHASES):
    print("  %-12s %10.0f cyc  %5.1f%%" % (nm, (pc[:, i] / steps).mean(), 100 * pc[:, i].sum() / tot.sum()))
print("  diag streampower split (cyc/step): prepass+scan (compact layout: routing only) %.0f, event compaction %.0f, parallel replay %.0f, serial replay %.0f" % tuple((pc[:,i]/steps).mean() for i in (15,12,13,14)))
tot = pc[:, :12].sum(1)
print("  total        %10.0f cyc/step  (%.1f us at 1.965 GHz)" % ((tot / steps).mean(), (tot / steps).mean() / 1965.0))

names = {16: "erodep pass1+compact", 17: "erodep small basins", 18: "erodep large basins", 19: "erodep tail", 20: "#large basins",
         21: "#land pits"}
for i in sorted(names):
    print("  sub %-32s %10.1f /step" % (names[i], (pc[:, i] / steps).mean()))

info = eng.info()
print("  launch:", info)
if info["path"] == "compact":
    print("  sub discharge_c (cyc/step): depths %.0f, stable sort %.0f, records+init %.0f, level loop %.0f, Q scatter %.0f; stream power pre-pass %.0f"
          % tuple((pc[:, i] / steps).mean() for i in (22, 23, 24, 25, 26, 27)))
    print("  sub links_stack_c, both surfaces (cyc/step): links %.0f, bases+copy %.0f, walk A + splitter rank %.0f, walk B %.0f"
          % tuple((pc[:, i] / steps).mean() for i in (28, 29, 30, 31)))
    if DBG & 4:
        print("  fill diag, warp 0 (per step): rounds %.1f, evaluations of warp 0 %.0f, cycles scan %.0f, evaluation %.0f, barrier wait %.0f"
              % tuple((pc[:, i] / steps).mean() for i in (32, 33, 34, 35, 36)))
if DBG & 4 and info["path"] != "compact":
    print("  fill diag (warp 0, cyc/step): scan %.0f, eval %.0f, barrier %.0f; rounds/step %.1f, evals/step (CTA) %.0f, eval passes/step (warp 0) %.1f"
          % tuple((pc[:, i] / steps).mean() for i in (22, 23, 24, 25, 26, 27)))
