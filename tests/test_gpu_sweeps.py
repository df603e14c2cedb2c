"""Quantitative comparison with EVERY likelihood sweep the reference ships (`Examples/*/liklsurf_*/exp_data.txt`, imported
into data/*.npz as surf_false / surf_true): 2 500 forward models per sweep on the batched engine, per-cell relative
deviation of the elevation mean-squared-error column against the shipped value.

The shipped numbers come from the authors' Triangle mesh (not recorded anywhere: SURVEY.md 8c) and unseeded base
shuffles, ours from the hexagonal-lattice mesh - so the comparison is statistical, but it is a bound on how far the whole
restated path (oracle == device, bit for bit) can be from the reference's own runs: median and 95th percentile of
|mse_ours - mse_shipped| / mse_shipped over the 2 500 cells, asserted against the bounds stated below.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, ROOT)

# config -> (bound on the median, bound on the 95th percentile) of the per-cell relative deviation of the elevation MSE.
# Measured on the round-2 build (profiles/README.md): crater_fast 0.37 / 1.98, etopo_fast 0.25 / 1.62, crater 0.19 / 2.73,
# etopo 1.30 / 5.86 - large where the shipped MSE is at its noise floor (~1 m^2 at the truth: a 1 m mesh-dependent
# difference doubles it) and small in the log-likelihood (median 2 - 7 %).  The bounds are those values + 30 %: the test
# fails when the restated path moves AWAY from the reference's own runs.
BOUNDS = {"crater_fast": (0.49, 2.6), "crater": (0.25, 3.6), "etopo_fast": (0.33, 2.1), "etopo": (1.7, 7.7)}


@pytest.mark.parametrize("cfg", ["crater_fast", "etopo_fast", "crater", "etopo"])
def test_shipped_sweep_per_cell_deviation(cfg):
    import bench
    from bayeslands_b200.mcmc import bayeslands_mcmc
    from bayeslands_b200.surflikl import likelihood_surface
    d, inp, m, simtime, rl, el, coords, likl_sed = bench.load_config(cfg)
    mc = bayeslands_mcmc(True, simtime, 0, d["final_elev"], d["final_erdp"], d["final_erdp_pts"], coords, None, inp, el, rl,
                         (0.4, 0.6), (0.9, 1.1), "sweep", False, batch=500, mesh=m,
                         sealevel=d["sealevel"] if "sealevel" in d else None)
    rain, erod, lik, mse = likelihood_surface(mc, rl, el, n=50)
    assert ((mc.last_status & 9) == 0).all()          # no cascade-guard / NaN evaluation (bit 4 = drainage fallback: informational)
    ship = d["surf_false"]
    assert np.allclose(ship[:, 0].reshape(50, 50)[:, 0], rain) and np.allclose(ship[:, 1].reshape(50, 50)[0], erod)
    smse = ship[:, 4].reshape(50, 50)
    rel = np.abs(mse - smse) / smse
    med, p95 = float(np.median(rel)), float(np.percentile(rel, 95))
    # likelihood column: L = -G/2 (ln(2 pi mse) + 1) on both sides, so its deviation follows from the MSE's
    dl = np.abs(lik - ship[:, 2].reshape(50, 50)) / np.abs(ship[:, 2].reshape(50, 50))
    print("%s: elevation MSE vs shipped sweep: median rel dev %.4f, p95 %.4f, max %.4f; log-likelihood median rel dev %.5f; "
          "argmax ours (%.3g, %.3g) shipped (%.3g, %.3g)"
          % (cfg, med, p95, float(rel.max()), float(np.median(dl)), rain[np.unravel_index(np.argmax(lik), lik.shape)[0]],
             erod[np.unravel_index(np.argmax(lik), lik.shape)[1]], ship[np.argmax(ship[:, 2]), 0], ship[np.argmax(ship[:, 2]), 1]))
    drm = np.abs(np.sqrt(mse) - np.sqrt(smse))
    print("%s: |RMSE ours - RMSE shipped| median %.3f m, p95 %.3f m (relief of the observed surface %.0f m)"
          % (cfg, float(np.median(drm)), float(np.percentile(drm, 95)), float(np.ptp(d["final_elev"]))))
    bm, b95 = BOUNDS[cfg]
    assert med <= bm and p95 <= b95, (cfg, med, p95)
