"""Run BASELINE config 5 (synthetic regular grid) on the lattice engine and report per-step cost.

  python tools/run_synthetic.py --nx 2048 --ny 2048 --batch 32 --T 50000 [--steps 10]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslands_b200 import synthetic                      # noqa: E402
from bayeslands_b200.grid_engine import GridEngine         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--ny", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--T", type=float, default=50000.0)
    ap.add_argument("--steps", type=int, default=0, help="stop after this many single steps (0: full run)")
    ap.add_argument("--group", type=int, default=0, help="proposals per fill group (0: default 64)")
    ap.add_argument("--bps", type=int, default=0, help="blocks per SM of the cooperative fill loop (0: occupancy limit)")
    a = ap.parse_args()
    t0 = time.time()
    lat, inp = synthetic.build(a.nx, a.ny, T=a.T)
    rain, erod = synthetic.proposals(a.batch)
    eng = GridEngine(lat, inp, a.batch, fill_group=a.group, fill_blocks_per_sm=a.bps)
    eng.reset(rain, erod)
    torch.cuda.synchronize()
    print("setup %.1f s, workspace %.2f GB" % (time.time() - t0, eng.workspace.numel() / 1e9), flush=True)
    N = (a.nx + 2) * (a.ny + 2)
    if a.steps:
        prev = np.zeros(a.batch, dtype=np.int64); pl = 0
        for s in range(a.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.step(a.T); e1.record(); torch.cuda.synchronize()
            sc = eng.scalars()
            ms = e0.elapsed_time(e1)
            print("step %3d  %8.2f ms  sweeps max %5d  launches %6d  dt %s  %.1f Gnode-steps/s" % (
                s, ms, int((sc["fill_sweeps"] - prev).max()), eng.launches - pl, sc["last_dt"][:3],
                a.batch * N / ms / 1e6), flush=True)
            prev = sc["fill_sweeps"].copy(); pl = eng.launches
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run_to_time(a.T); e1.record(); torch.cuda.synchronize()
        sc = eng.scalars()
        ms = e0.elapsed_time(e1)
        S = sc["steps"]
        print("run: %.1f ms, steps %d..%d, sweeps/step %.1f, launches %d, %.2f evals/s, %.2f Gnode-steps/s" % (
            ms, S.min(), S.max(), sc["fill_sweeps"].sum() / S.sum(), eng.launches, a.batch / ms * 1e3,
            S.sum() * N / ms / 1e6))


if __name__ == "__main__":
    main()
