"""GPU tests of the reference-shaped host API (bayeslands_mcmc / Model mirrors) against the oracle."""
import math
import random

import numpy as np
import pytest

from conftest import CRATER_FAST_COORDS

pytestmark = pytest.mark.gpu


def _mc(crater_fast, batch, likl_sed=True, samples=6):
    from bayeslands_b200.mcmc import bayeslands_mcmc
    d, inp, m = crater_fast
    return bayeslands_mcmc(True, 15000, samples, d["final_elev"], d["final_erdp"], d["final_erdp_pts"],
                           np.array(CRATER_FAST_COORDS), None, inp, [3.e-5, 7.e-5], [0.0, 3.0], [0.4, 0.6], [0.9, 1.1],
                           "t", likl_sed, batch=batch, mesh=m)


def _oracle_lik(crater_fast, om, rain, erod, seed, likl_sed):
    from oracle import oracle
    d, inp, m = crater_fast
    ev, ed, pts = om.blackbox(rain, erod, 15000, CRATER_FAST_COORDS, seed=int(seed))
    return oracle.likelihood(ev[15000], d["final_elev"], np.array([pts[k] for k in sorted(pts)]), d["final_erdp_pts"],
                             likl_sed)[0], ev, ed, pts


def test_blackbox_and_likelihoodfunc_match_reference_shapes_and_oracle(crater_fast):
    from oracle import oracle
    d, inp, m = crater_fast
    mc = _mc(crater_fast, 4)
    mc._evals = 0
    ev, ed, pts = mc.blackBox(1.5, 5e-5, 0.5, 1.0)
    assert list(ev.keys()) == [0, 3750, 7500, 11250, 15000] and ev[15000].shape == (123, 123) and pts[0].shape == (10,)
    om = oracle.OracleModel(m, inp)
    _, oev, oed, opts = _oracle_lik(crater_fast, om, 1.5, 5e-5, 0, True)
    for T in ev:
        assert np.array_equal(ev[T], oev[T]) and np.array_equal(ed[T], oed[T]) and np.array_equal(pts[T], opts[T])
    mc._evals = 0
    out = mc.likelihoodFunc([1.5, 5e-5, 0.5, 1.0], d["final_elev"], d["final_erdp"], d["final_erdp_pts"], 1.0, 1.0, 1.0)
    want = _oracle_lik(crater_fast, om, 1.5, 5e-5, 0, True)[0]
    assert abs(out[0] - want) <= 1e-9 * abs(want)


def test_likelihood_batch_matches_oracle(crater_fast):
    from oracle import oracle
    d, inp, m = crater_fast
    for sed in (False, True):
        mc = _mc(crater_fast, 5, likl_sed=sed)
        rain = np.array([1.5, 0.2, 2.8, 1.0, 0.0]); erod = np.array([5e-5, 6.5e-5, 3.2e-5, 4e-5, 5e-5])
        seeds = np.array([11, 12, 13, 14, 15], dtype=np.int64)
        lik, tau = mc.likelihood_batch(rain, erod, seeds=seeds)
        om = oracle.OracleModel(m, inp)
        for b in range(5):
            want = _oracle_lik(crater_fast, om, rain[b], erod[b], seeds[b], sed)[0]
            assert abs(lik[b] - want) <= 1e-9 * abs(want)


def test_sampler_chains_accept_reject_equal_oracle_chain(crater_fast):
    """Same RNG protocol driven by the oracle must take identical accept/reject decisions."""
    from oracle import oracle
    d, inp, m = crater_fast
    S, Cn = 5, 2
    mc = _mc(crater_fast, Cn, samples=S)
    mc._evals = 0
    out = mc.sampler_chains(Cn, np_seed=3, py_seed=9)
    om = oracle.OracleModel(m, inp)
    evals = 0
    # replay on the CPU: seeds are handed out in evaluation order, batch by batch (mcmc._seeds)
    rs = [np.random.RandomState(3 + c) for c in range(Cn)]
    pr = [random.Random(9 + c) for c in range(Cn)]
    rain = np.array([r.uniform(0.0, 3.0) for r in rs]); erod = np.array([r.uniform(3e-5, 7e-5) for r in rs])
    lik = np.array([_oracle_lik(crater_fast, om, rain[c], erod[c], evals + c, True)[0] for c in range(Cn)]); evals += Cn
    pos = [(rain.copy(), erod.copy())]
    for i in range(S - 1):
        p_r, p_e = rain.copy(), erod.copy()
        for c in range(Cn):
            v = rain[c] + rs[c].normal(0, mc.step_rain)
            if not (v < 0.0 or v > 3.0): p_r[c] = v
            v = erod[c] + rs[c].normal(0, mc.step_erod)
            if not (v < 3e-5 or v > 7e-5): p_e[c] = v
            for _ in range(3): rs[c].normal(0, 1.0, 1)
        lp = np.array([_oracle_lik(crater_fast, om, p_r[c], p_e[c], evals + c, True)[0] for c in range(Cn)]); evals += Cn
        for c in range(Cn):
            try: mh = min(1, math.exp(lp[c] - lik[c]))
            except OverflowError: mh = 1
            if pr[c].uniform(0, 1) < mh: lik[c], rain[c], erod[c] = lp[c], p_r[c], p_e[c]
        pos.append((rain.copy(), erod.copy()))
    for i in range(S):
        assert np.array_equal(out["pos_rain"][i], pos[i][0]) and np.array_equal(out["pos_erod"][i], pos[i][1])


def test_model_mirror_load_xml_run_to_time(crater_fast):
    from bayeslands_b200.model import Model
    from oracle import oracle
    d, inp, m = crater_fast
    model = Model()
    ie, ir = model.load_xml("0", str(d["xml"]), muted=True, mesh=m, seed=5)
    assert ie == [] and ir == []
    model.input.SPLero = 4e-5
    model.flow.erodibility.fill(4e-5)
    model.force.rainVal[:] = 2.0
    om = oracle.OracleModel(m, inp, seed=5)
    om.reset(2.0, 4e-5, seed=5)
    for T in (0, 3750, 7500):
        model.run_to_time(T, muted=True)
        om.run_to_time(float(T))
        assert model.tNow == om.tNow
        assert np.array_equal(model.elevation, om.z) and np.array_equal(model.cumdiff, om.cumdiff)
    model.run_to_time(1e9, muted=True)          # clamped to the XML end (model.py:240-243)
    assert model.tNow == 15000.0
    assert model.FVmesh.node_coords.shape == (m.N, 2) and model.flow.receivers.shape == (m.N,)


def test_likelihood_surface_argmax_matches_shipped_sweep(crater_fast, tmp_path):
    """the 50 x 50 sweep of bl_surflikl over (rain, erod) - 2500 forward models - must put its
    maximum on the shipped surface's ridge through (1.5, 5e-5) and follow the shipped surface's shape."""
    from bayeslands_b200.surflikl import likelihood_surface
    d, inp, m = crater_fast
    mc = _mc(crater_fast, 500, likl_sed=False)
    rain, erod, lik, mse = likelihood_surface(mc, [0.0, 3.0], [3.e-5, 7.e-5], n=50, out_dir=str(tmp_path))
    i, j = np.unravel_index(np.argmax(lik), lik.shape)
    ship = d["surf_false"]
    si = int(np.argmax(ship[:, 2]))
    ti, tj = divmod(si, 50)                              # shipped rows: rain outer loop, erod inner
    assert np.isclose(ship[si, 0], rain[ti]) and np.isclose(ship[si, 1], erod[tj])
    # The surface is a long ridge K * sqrt(rain) = const (incision ~ K (rain A)^m, m = 0.5): on a different mesh
    # the maximum slides along it, so the invariant is the ridge, and the 1-D slices through the shipped truth.
    ridge = erod[j] * np.sqrt(rain[i])
    assert abs(ridge / (erod[tj] * np.sqrt(rain[ti])) - 1.0) < 0.05, (rain[i], erod[j])
    assert abs(int(np.argmax(lik[ti, :])) - tj) <= 2 and abs(int(np.argmax(lik[:, tj])) - ti) <= 3
    # rain = 0 row: no fluvial erosion, likelihood independent of erodibility
    assert np.ptp(lik[0]) == 0.0
    # same shape as the shipped sweep (different mesh => compare ranks, not values)
    from scipy.stats import spearmanr
    rho = spearmanr(lik.ravel(), ship[:, 2]).correlation
    assert rho > 0.9, rho
    rows = np.loadtxt(tmp_path / "exp_data.txt")
    assert rows.shape == (2500, 3) and np.allclose(rows[:, 2], lik.ravel())


def test_topo_generator_writes_reference_files(crater_fast, tmp_path):
    from bayeslands_b200.surflikl import topo_generator
    mc = _mc(crater_fast, 2)
    fe, fd, pts = topo_generator(mc, 1.5, 5e-5, out_dir=str(tmp_path), seed=0)
    assert fe.shape == (123, 123) and pts.shape == (5, 10) and np.all(pts[0] == 0)
    back = np.loadtxt(tmp_path / "final_elev.txt")
    assert back.shape == (123, 123) and np.abs(back - fe).max() < 1e-5
    clean, _, _ = topo_generator(mc, 1.5, 5e-5, final_noise=False)
    noise_var = np.var(fe - clean)
    assert 0.8 * (clean.max() * 0.01 + 0.5) < noise_var < 1.2 * (clean.max() * 0.01 + 0.5)


def test_sampler_writes_reference_outputs_and_is_seeded(crater_fast, tmp_path):
    """bl_mcmc.sampler: global numpy stream + stdlib random => the chain is reproducible from the two seeds and
    leaves the reference's files behind (exp_data.txt rows, description.txt, stats, mean prediction grids)."""
    import os
    out = []
    for rep in range(2):
        d_ = tmp_path / ("run%d" % rep)
        mc = _mc(crater_fast, 1, samples=8)
        mc.filename = str(d_)
        np.random.seed(11); random.seed(12)
        out.append(mc.sampler())
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][2], out[1][2])
    rows = np.loadtxt(tmp_path / "run0" / "exp_data.txt")
    assert rows.shape == (8, 3) and np.allclose(rows[1:, 0], out[0][0][1:]) and np.allclose(rows[:, 2], out[0][2])
    assert os.path.isfile(tmp_path / "run0" / "description.txt") and os.path.isfile(tmp_path / "run0" / "experiment_stats.txt")
    g = np.loadtxt(tmp_path / "run0" / "prediction_data" / "mean_pred_elev_15000.txt")
    assert g.shape == (123, 123) and 20.0 < g.mean() < 120.0
    assert 0.0 <= out[0][3] <= 100.0
    # bl_mcmc.py:595-597, :776-833: eta step sizes, RMSE lines, count list, per-time point means, mid-domain slices
    desc = open(tmp_path / "run0" / "description.txt").read()
    assert all(k in desc for k in ("step_eta_elev:", "step_eta_erdp:", "step_eta_erdp_pts:", "Initial_proposed_rain:"))
    stats = open(tmp_path / "run0" / "experiment_stats.txt").read()
    assert stats.startswith("RMSEelev: ") and "RMSEerdp: " in stats and "RMSEerdp_pts: " in stats and "Count List : [0" in stats
    rm = float(stats.split("RMSEelev: ")[1].split()[0])
    d = crater_fast[0]
    assert abs(rm - np.sqrt(np.mean((g - d["final_elev"]) ** 2))) < 1e-4 * rm      # the file is written with %.5f
    xs = np.loadtxt(tmp_path / "run0" / "prediction_data" / "pred_xslc.txt")
    ys = np.loadtxt(tmp_path / "run0" / "prediction_data" / "pred_yslc.txt")
    assert xs.shape == (8, 123) and ys.shape == (8, 123) and np.all(xs[0] == 0)
    assert np.loadtxt(tmp_path / "run0" / "prediction_data" / "mean_pred_erdp_pts_15000.txt").shape == (10,)


def test_speculative_sampler_is_the_same_chain(crater_fast, tmp_path):
    """sampler(speculate=K) evaluates K proposals ahead assuming rejections: the chain, the files it writes and the state
    of both RNG streams afterwards must be identical to the sequential sampler's."""
    import filecmp
    import os
    res, states = [], []
    for rep, K in enumerate((1, 16)):
        mc = _mc(crater_fast, 16, samples=60)
        mc.filename = str(tmp_path / ("run%d" % rep))
        np.random.seed(5); random.seed(6)
        res.append(mc.sampler(speculate=K))
        states.append((np.random.get_state()[1].copy(), np.random.get_state()[2], random.getstate()))
        launches = mc.engine.launches
        res[-1] = res[-1] + (launches,)
    a, b = res
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert 0 < a[3] < 100.0, "the chain must contain acceptances and rejections for the test to mean anything"
    assert np.array_equal(states[0][0], states[1][0]) and states[0][1] == states[1][1] and states[0][2] == states[1][2]
    for f in ("exp_data.txt", os.path.join("prediction_data", "mean_pred_elev_15000.txt"),
              os.path.join("prediction_data", "mean_pred_erdp_pts.txt")):
        assert filecmp.cmp(tmp_path / "run0" / f, tmp_path / "run1" / f, shallow=False), f
    assert b[4] < a[4]                                        # fewer launches


def test_speculative_sampler_chains_are_the_same_chains(crater_fast):
    """sampler_chains(speculate=K): K proposals ahead per chain and launch; identical chains for any K."""
    outs = []
    for K in (1, 8):
        mc = _mc(crater_fast, 16, samples=40)
        mc._evals = 0
        outs.append((mc.sampler_chains(2, np_seed=21, py_seed=22, speculate=K), mc.engine.launches, mc._evals))
    (a, la, ea), (b, lb, eb) = outs
    for k in ("pos_rain", "pos_erod", "pos_likl", "accept_ratio"):
        assert np.array_equal(a[k], b[k]), k
    assert 0 < a["accept_ratio"].sum() < 200.0 and lb < la and ea == eb
