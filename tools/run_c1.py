#!/usr/bin/env python
"""SURVEY.md 8d workload C1 on the GPU engine: Examples/crater_fast, (i) the 50 x 50 likelihood surface of
bl_surflikl.py:392-393 (2500 evaluations, 256 per launch) and (ii) a 1000-sample single random-walk Metropolis chain
(numpy seed 0 / random seed 0, reference step sizes = 3 % of the parameter ranges, bl_mcmc.py:89-90).

    python tools/run_c1.py [--samples 1000] [--out /tmp/c1]
"""
import argparse
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402
from bayeslands_b200.mcmc import bayeslands_mcmc                # noqa: E402
from bayeslands_b200.surflikl import likelihood_surface         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="crater_fast")
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--speculate", type=int, default=0, help="evaluate this many proposals ahead per launch (same chain)")
    a = ap.parse_args()
    d, inp, m, simtime, rl, el, coords, likl_sed = bench.load_config(a.config)

    def make(batch, samples, fn):
        return bayeslands_mcmc(True, simtime, samples, d["final_elev"], d["final_erdp"], d["final_erdp_pts"], coords, fn, inp,
                               el, rl, (0.4, 0.6), (0.9, 1.1), "c1", likl_sed, batch=batch, mesh=m)

    mc = make(256, 1, None)
    likelihood_surface(mc, rl, el, n=8)                          # warm-up
    t0 = time.perf_counter()
    rain, erod, lik, mse = likelihood_surface(mc, rl, el, n=50, out_dir=os.path.join(a.out, "liklsurf") if a.out else None)
    dt = time.perf_counter() - t0
    i, j = np.unravel_index(np.argmax(lik), lik.shape)
    print("likelihood surface 50x50 on %s: %.2f s = %.0f evals/s; argmax (rain %.2f, erod %.2e) L %.1f; min MSE %.4f at "
          "(rain %.2f, erod %.2e); K*sqrt(rain) at argmax %.3e (shipped truth 1.5, 5e-5 -> %.3e); rain = 0 row spread %.3g"
          % (a.config, dt, lik.size / dt, rain[i], erod[j], lik[i, j], mse.min(), rain[np.unravel_index(np.argmin(mse), mse.shape)[0]],
             erod[np.unravel_index(np.argmin(mse), mse.shape)[1]], erod[j] * np.sqrt(rain[i]), 5e-5 * np.sqrt(1.5),
             float(np.ptp(lik[0]))))
    del mc
    np.random.seed(0); random.seed(0)
    mc = make(max(1, a.speculate), a.samples, os.path.join(a.out, "chain") if a.out else None)
    t0 = time.perf_counter()
    pos_rain, pos_erod, pos_likl, acc = mc.sampler(speculate=a.speculate or 1)
    dt = time.perf_counter() - t0
    burn = int(0.05 * a.samples)
    print("single chain, %d samples, %s: %.2f s = %.1f chain iterations/s, %d launches; accepted %.1f %%; posterior mean rain %.3f "
          "erod %.3e; max L %.1f; chain digest %.10e" % (
              a.samples, ("speculating %d ahead" % a.speculate) if a.speculate else "sequential (one CTA busy)", dt,
              (a.samples + 1) / dt, mc.engine.launches, acc, pos_rain[burn:].mean(), pos_erod[burn:].mean(), pos_likl.max(),
              float(np.sum(pos_rain * 1e3 + pos_erod * 1e7 + pos_likl * 1e-3))))


if __name__ == "__main__":
    main()
