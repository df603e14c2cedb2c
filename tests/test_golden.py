"""Known answers the reference ships for this path - statistical, end to end.

The reference has no unit tests; what pins the path is data under Examples/ (imported into data/*.npz by
tools/import_reference_data.py): the noisy final topography at the true parameters and the 50x50
likelihood sweeps.  The authors' Triangle mesh is not reproducible, so agreement is to a few % in MSE.
"""
import math
import os

import numpy as np
import pytest

from conftest import CRATER_FAST_COORDS, ROOT
from oracle import oracle


def test_likelihood_formula_against_shipped_row(crater_fast):
    """L = -G/2 (ln(2 pi MSE) + 1); shipped row (1.5, 5e-05): MSE 1.08712.. -> -22099.0167."""
    d, _, _ = crater_fast
    surf = d["surf_false"]                       # rain erod likl sq_err ...
    row = surf[(np.isclose(surf[:, 0], 1.5)) & (np.isclose(surf[:, 1], 5e-5))][0]
    G = 123 * 123
    assert abs(-0.5 * G * (math.log(2 * math.pi * row[3]) + 1.0) - row[2]) < 1e-6 * abs(row[2])
    # rain = 0 rows are identical (no fluvial erosion): likelihood independent of erodibility
    r0 = surf[surf[:, 0] == 0.0]
    assert len(r0) == 50 and np.ptp(r0[:, 2]) == 0.0
    # argmax of the shipped surface is the true-parameter cell
    best = surf[np.argmax(surf[:, 2])]
    assert np.isclose(best[0], 1.5) and np.isclose(best[1], 5e-5)


def test_product_likelihood_helper_matches_reference_arithmetic():
    from bayeslands_b200.mcmc import log_likelihood
    rng = np.random.default_rng(0)
    pred = rng.normal(50, 10, (12, 13)); real = pred + rng.normal(0, 1, pred.shape)
    pts_p = rng.normal(0, 3, (5, 10)); pts_r = pts_p + rng.normal(0, 0.5, (5, 10)); pts_r[0] = 0; pts_p[0] = 0
    for sed in (False, True):
        want, tau = oracle.likelihood(pred, real, pts_p, pts_r, sed)
        sse = np.sum(np.square(pred - real))
        got, tau2 = log_likelihood(np.array([sse]), pred.size, pts_p[None], pts_r, sed)
        assert abs(got[0] - want) < 1e-9 * abs(want) and abs(tau2[0] - tau) < 1e-12 * tau


def test_rain_zero_is_independent_of_erodibility(crater_fast):
    d, inp, m = crater_fast
    om = oracle.OracleModel(m, inp)
    a, _, _ = om.blackbox(0.0, 3e-5, 15000, CRATER_FAST_COORDS, seed=1)
    b, _, _ = om.blackbox(0.0, 7e-5, 15000, CRATER_FAST_COORDS, seed=1)
    assert np.array_equal(a[15000], b[15000])


def test_mse_minimum_at_true_parameters(crater_fast):
    """Slices of the likelihood surface through the shipped truth (1.5, 5e-5): our restatement on our mesh
    must put the minimum MSE against the shipped final_elev.txt at the truth (+- one 50x50 grid cell), and the
    misfit there must be within a few m^2 of the noise floor 1.087 the authors' mesh gives."""
    d, inp, m = crater_fast
    om = oracle.OracleModel(m, inp)
    real = d["final_elev"]

    def mse(r, e):
        ev, _, _ = om.blackbox(r, e, 15000, CRATER_FAST_COORDS, seed=7)
        return float(np.mean((ev[15000] - real) ** 2))
    rains = [1.2, 1.38, 1.5, 1.62, 1.8]
    erods = [4.2e-5, 4.76e-5, 5e-5, 5.24e-5, 5.8e-5]
    mr = [mse(r, 5e-5) for r in rains]
    me = [mse(1.5, e) for e in erods]
    assert int(np.argmin(mr)) in (1, 2, 3) and int(np.argmin(me)) in (1, 2, 3), (mr, me)
    assert mr[2] < 6.0 and mr[0] > mr[2] and mr[-1] > mr[2]
    # the model explains > 90 % of the change from the initial topography
    init = float(np.mean((om.interpolate(m.elevation) - real) ** 2))
    assert mr[2] < 0.1 * init


def test_oracle_matches_reference_sfd_vectors(crater_fast):
    """tests/golden/sfd_reference_vectors.npz holds what the reference's own libUtils/sfd.c returns on the crater_fast
    mesh (tools/make_golden_sfd.py): the oracle's restatement must reproduce every array bit for bit - also where
    /root/reference and oracle/_ref do not exist."""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_sfd import surfaces
    from oracle import oracle
    d, inp, m = crater_fast
    g = np.load(os.path.join(ROOT, "tests", "golden", "sfd_reference_vectors.npz"))
    assert int(g["N"]) == m.N
    L = oracle.lib()
    om = oracle.OracleModel(m, inp)
    c_dp, c_ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    N = m.N
    for t, (z, fill) in enumerate(surfaces(m)):
        b = np.zeros(N, dtype=np.int32); r = np.zeros(N, dtype=np.int32)
        L.or_directions_base(C.byref(om.cm), z.ctypes.data_as(c_dp), b.ctypes.data_as(c_ip), r.ctypes.data_as(c_ip))
        assert np.array_equal(r, g["rcv1_%d" % t]) and np.array_equal(b, g["base1_%d" % t])
        mh = np.zeros(N); md = np.zeros(N)
        L.or_directions(C.byref(om.cm), fill.ctypes.data_as(c_dp), z.ctypes.data_as(c_dp), b.ctypes.data_as(c_ip),
                        r.ctypes.data_as(c_ip), mh.ctypes.data_as(c_dp), md.ctypes.data_as(c_dp))
        assert np.array_equal(r, g["rcv_%d" % t]) and np.array_equal(b, g["base_%d" % t])
        assert np.array_equal(mh, g["maxh_%d" % t]) and np.array_equal(md, g["maxdep_%d" % t])
        df = np.zeros(N)
        L.or_diffusion(C.byref(om.cm), z.ctypes.data_as(c_dp), df.ctypes.data_as(c_dp))
        assert np.array_equal(df, g["diff_%d" % t])
