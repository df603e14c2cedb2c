"""Per-round trace of the depression filling of one CTA (diagnostics; needs the trace build:
   nvcc ... -DBL_FILL_TRACE, loaded through BAYESLANDS_B200_LIB).  Prints per-round cycles of the last time step."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayeslands_b200 import _lib
from bayeslands_b200.engine import Engine

cfg = sys.argv[1] if len(sys.argv) > 1 else "crater"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
DBG = int(sys.argv[3]) if len(sys.argv) > 3 else 0
d, inp, m, simtime, rl, el, coords, likl_sed = bench.load_config(cfg)
eng = Engine(m, inp, B, real_elev=d["final_elev"], erdp_coords=coords, debug=DBG)
r, e, s = bench.proposals(0, 0, B, rl, el)
eng.reset(r, e, s)
for _ in range(40):
    eng.step(float(simtime))
out = np.empty(eng.N, dtype=np.uint64)
for b in (0, 77):
    _lib.check(eng.L.bl_get_field(eng.ctx, 100, b, out.ctypes.data_as(C.c_void_p), eng._stream()))
    for wi, wname in ((0, "warp 0"), (1, "warp 5")):
        n = int(out[wi])
        if n == 0:
            continue
        rec = out[8 + 8000 * wi: 8 + 8000 * wi + n].astype(np.int64).reshape(-1, 5)
        T0, T1, T2, off, T3 = rec.T
        print("proposal %d %s: %d rounds, mean round %.0f cycles; list ready %.0f, eval %.0f, barrier wait %.0f; pending/round mean %.1f max %d"
              % (b, wname, len(rec), np.diff(T0).mean(), (T1 - T0).mean(), (T2 - T1).mean(), (T3 - T2).mean(), off.mean(), off.max()))
        if b == 0:
            for i in range(min(len(rec) - 1, 40)):
                print("  %3d round %6d | list ready %5d  evals done %5d  barrier passed %5d | pending %3d" % (
                    i, T0[i + 1] - T0[i], T1[i] - T0[i], T2[i] - T0[i], T3[i] - T0[i], off[i]))
