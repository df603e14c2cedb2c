"""`bl_surflikl.likelihoodSurface` on the batched engine (reference: `bl_surflikl.py:373-460`).

The reference evaluates a 50 x 50 grid of (rain, erodibility) one model run at a time; here the 2500
evaluations are a handful of batched launches.  Output: the same `exp_data.txt` rows (`rain erod likl`,
`bl_surflikl.py:291-335`).  The sediment term has no x50 weight here, as in `bl_surflikl.py:360`.
"""
from __future__ import annotations

import os

import numpy as np

from .mcmc import bayeslands_mcmc, log_likelihood


def likelihood_surface(mc: bayeslands_mcmc, rainlimits, erodlimits, n: int = 50, evaluate=None, out_dir=None):
    """Returns (rain[n], erod[n], likl[n, n], mse[n, n]); likl[i, j] is for (rain[i], erod[j])."""
    rain = np.linspace(rainlimits[0], rainlimits[1], num=n, endpoint=False)     # bl_surflikl.py:392-393
    erod = np.linspace(erodlimits[0], erodlimits[1], num=n, endpoint=False)
    R, E = np.meshgrid(rain, erod, indexing="ij")
    r, e = R.ravel(), E.ravel()
    lik = np.empty(len(r)); mse = np.empty(len(r))
    B = mc.batch
    for lo in range(0, len(r), B):
        hi = min(len(r), lo + B)
        sse, pts, _, _ = mc.blackBox_batch(r[lo:hi], e[lo:hi], seeds=np.arange(lo, hi, dtype=np.int64))
        l, t2 = log_likelihood(sse, mc.G, pts, mc.real_erdp_pts, mc.likl_sed, sed_weight=1.0)
        lik[lo:hi], mse[lo:hi] = l, t2
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "exp_data.txt"), "w") as f:
            for a, b, c in zip(r, e, lik):
                f.write("{0}\t{1}\t{2}\n".format(a, b, c))
    return rain, erod, lik.reshape(n, n), mse.reshape(n, n)


def topo_generator(mc: bayeslands_mcmc, rain: float, erod: float, out_dir=None, final_noise: bool = True, seed=None):
    """`bl_topogenr.topoGenerator` (reference: `bl_topogenr.py:102-220`): run the model at the chosen "true"
    parameters, add Gaussian noise of variance `max * 0.01 + 0.5` to the final-time outputs (`:160-181`) and write
    `data/final_elev.txt`, `final_erdp.txt`, `final_erdp_pts.txt`, `initial_elev.txt` (`%.5f`, `:187-217`).
    Returns (final_elev, final_erdp, erdp_pts[nt, npts]) as written."""
    sse, pts, elev, erdp = mc.blackBox_batch([rain], [erod], grids=True)
    elev, erdp, pts = elev[0], erdp[0], pts[0].copy()
    rs = np.random.RandomState(seed)
    fe, fd = elev[-1].copy(), erdp[-1].copy()
    if final_noise:
        fe = fe + rs.normal(0, np.sqrt(abs(fe.max() * 0.01 + 0.5)), fe.size).reshape(fe.shape)
        fd = fd + rs.normal(0, np.sqrt(abs(fd.max() * 0.01 + 0.5)), fd.size).reshape(fd.shape)
        pts[-1] = pts[-1] + rs.normal(0, np.sqrt(abs(pts[-1].max() * 0.01 + 0.5)), pts[-1].size)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        np.savetxt(os.path.join(out_dir, "initial_elev.txt"), elev[0], fmt="%.5f")
        np.savetxt(os.path.join(out_dir, "final_elev.txt"), fe, fmt="%.5f")
        np.savetxt(os.path.join(out_dir, "final_erdp.txt"), fd, fmt="%.5f")
        np.savetxt(os.path.join(out_dir, "final_erdp_pts.txt"), pts, fmt="%.5f")
    return fe, fd, pts
