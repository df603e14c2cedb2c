"""Per-stage entry points of the C-ABI (bl_fill, bl_sfd, bl_stack, bl_discharge, bl_basins, bl_drainage, bl_streampower,
bl_diffuse, bl_interp) on caller arrays, both device paths, against

  * ALL of tests/golden/sfd_reference_vectors.npz - what the reference's own sfd.c returns (4 surfaces x 7 arrays),
  * the hand-derived micro-fixtures of tests/micro.py (tests/golden/README.md),
  * the oracle's routines on mid-run crater_fast arrays.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


@pytest.mark.parametrize("generic", [False, True])
def test_stages_match_all_reference_sfd_vectors(crater_fast, generic):
    from bayeslands_b200.engine import Engine
    from make_golden_sfd import surfaces
    d, inp, m = crater_fast
    g = np.load(os.path.join(ROOT, "tests", "golden", "sfd_reference_vectors.npz"))
    assert int(g["N"]) == m.N
    eng = Engine(m, inp, 1, force_generic=generic)
    for t, (z, fill) in enumerate(surfaces(m)):
        base1, rcv1 = eng.directions_base(z)
        assert np.array_equal(rcv1, g["rcv1_%d" % t]) and np.array_equal(base1, g["base1_%d" % t]), (generic, t)
        base, rcv, maxh, maxdep = eng.directions(fill, z)
        assert np.array_equal(rcv, g["rcv_%d" % t]) and np.array_equal(base, g["base_%d" % t]), (generic, t)
        assert np.array_equal(maxh, g["maxh_%d" % t]) and np.array_equal(maxdep, g["maxdep_%d" % t]), (generic, t)
        assert np.array_equal(eng.diffusion(z), g["diff_%d" % t]), (generic, t, "diffusion")


@pytest.mark.parametrize("generic", [False, True])
def test_stages_match_oracle_routines_mid_run(crater_fast, generic):
    """Every routine of a time step, fed with the oracle's own intermediate arrays of step 12 of a crater_fast run."""
    from bayeslands_b200.engine import Engine
    from oracle import oracle
    d, inp, m = crater_fast
    eng = Engine(m, inp, 1, force_generic=generic)
    rain, erod, seed = 2.2, 6.1e-5, 5
    om = oracle.OracleModel(m, inp, seed=seed)
    om.reset(rain, erod, seed=seed)
    for _ in range(11):
        om.streamflow(); om.sediment_flux(15000.0)
    step = om.steps
    z = om.z.copy()
    om.streamflow()
    sea = om.sea(om.tNow)
    assert np.array_equal(eng.pitfilling(z, sea), om.fillH)
    _, rcv, maxh, _ = eng.directions(om.fillH, z)
    assert np.array_equal(rcv, om.rcv) and np.array_equal(maxh, om.maxh)
    assert np.array_equal(eng.directions_base(z)[1], om.rcv1)
    stack, base = eng.stack_build(om.rcv, shuffle=1, seed=seed, step=step)
    assert np.array_equal(stack, om.stack) and np.array_equal(base, om.base)
    stack1, base1 = eng.stack_build(om.rcv1, shuffle=0)
    assert np.array_equal(stack1, om.stack1) and np.array_equal(base1, om.base1)
    assert np.array_equal(eng.discharge(om.rcv, rain), om.Q)
    pid, vol = eng.basinparameters(om.rcv1, z, om.fillH)
    assert np.array_equal(pid, om.pitID) and np.array_equal(vol, om.pitVol)
    d1, d2 = eng.basindrainage(om.pitID, om.pitVol, om.rcv, om.fillH, sea)
    assert np.array_equal(d1, om.pitDrain) and np.array_equal(d2, om.allDrain)
    pre = {k: getattr(om, k).copy() for k in ("rcv", "pitID", "pitVol", "pitDrain", "maxh", "Q", "fillH")}
    om.sediment_flux(15000.0)
    depo, ero = eng.streampower(pre["rcv"], pre["pitID"], pre["pitVol"], pre["pitDrain"], pre["maxh"], pre["Q"], pre["fillH"], z,
                                erod, sea, om.last_dt, shuffle=1, seed=seed, step=step)
    assert np.array_equal(depo, om.depo) and np.array_equal(ero, om.ero)
    grid = eng.interpolate(om.z)
    assert np.array_equal(grid, om.interpolate(om.z))


@pytest.mark.parametrize("generic", [False, True])
def test_micro_fixtures_on_the_device(generic):
    """The hand-derived fixtures F1 (capped lake, sweep order), F2 (two lakes: stack, basins, drainage chain, discharge,
    two-pit cascade, land-pit fill) and F3 (drainage de-dup hit) through the per-stage entry points and through bl_step."""
    import micro
    from bayeslands_b200.engine import Engine
    # F1
    for name, order in (("ascending", None), ("descending", [5, 4, 3, 2])):
        m = micro.chain_mesh(micro.F1_Z, order)
        eng = Engine(m, micro.inputs(fillmax=15.0), 1, force_generic=generic)
        assert np.array_equal(eng.pitfilling(m.elevation, -1000.0)[m.extra["ids"][1:5]], micro.f1_expected(name)), (name, "stage")
        eng.reset(np.array([1.0]), np.array([1e-5]), np.array([1], dtype=np.int64))
        eng.step(5000.0)
        assert np.array_equal(eng.field("fillH", 0)[m.extra["ids"][1:5]], micro.f1_expected(name)), (name, "step")
    # F2, full step
    m = micro.chain_mesh(micro.F2_Z)
    eng = Engine(m, micro.inputs(fillmax=200.0), 1, shuffle_mode=0, force_generic=generic)
    eng.reset(np.array([micro.F2_RAIN]), np.array([micro.F2_EROD]), np.array([1], dtype=np.int64))
    eng.step(5000.0)
    dt = micro.f2_cfl_dt()
    assert eng.scalars()["last_dt"][0] == dt and eng.scalars()["status"][0] == 0
    exp = micro.f2_derive(dt)
    for k in ("fillH", "rcv1", "rcv", "stack", "stack1", "pitID", "pitVol", "pitDrain", "allDrain", "Q", "ero", "depo", "sedflux"):
        assert np.array_equal(eng.field(k, 0), np.asarray(exp[k])), ("F2 step", k)
    # F2, routine by routine on the caller's arrays
    z = m.elevation
    e0 = micro.f2_derive(300.0)
    assert np.array_equal(eng.pitfilling(z, micro.F2_SEA), e0["fillH"])
    assert np.array_equal(eng.directions(e0["fillH"], z)[1], e0["rcv"]) and np.array_equal(eng.directions_base(z)[1], e0["rcv1"])
    assert np.array_equal(eng.stack_build(e0["rcv"])[0], e0["stack"]) and np.array_equal(eng.stack_build(e0["rcv1"])[0], e0["stack1"])
    pid, vol = eng.basinparameters(e0["rcv1"], z, e0["fillH"])
    assert np.array_equal(pid, e0["pitID"]) and np.array_equal(vol, e0["pitVol"])
    d1, d2 = eng.basindrainage(e0["pitID"], e0["pitVol"], e0["rcv"], e0["fillH"], micro.F2_SEA)
    assert np.array_equal(d1, e0["pitDrain"]) and np.array_equal(d2, e0["allDrain"])
    assert np.array_equal(eng.discharge(e0["rcv"], micro.F2_RAIN), e0["Q"])
    maxh = eng.directions(e0["fillH"], z)[2]
    depo, ero = eng.streampower(e0["rcv"], e0["pitID"], e0["pitVol"], e0["pitDrain"], maxh, e0["Q"], e0["fillH"], z, micro.F2_EROD,
                                micro.F2_SEA, 300.0)
    assert e0["load6"] > e0["pitVol"][6]                    # the two-pit cascade: lake 2 -> lake 1 -> boundary
    assert np.array_equal(depo, e0["depo"]) and np.array_equal(ero, e0["ero"])
    # F3: de-dup hit (a cycle in the pit graph: the device takes its literal sequential fallback)
    f = micro.F3
    m8 = micro.chain_mesh([0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0])
    eng8 = Engine(m8, micro.inputs(), 1, force_generic=generic)
    d1, d2 = eng8.basindrainage(f["pitID"], f["pitVol"], f["rcv"], f["fillH"], -1000.0)
    assert d1.tolist() == f["pitDrain"] and d2.tolist() == f["pitDrain"]


@pytest.mark.parametrize("generic", [False, True])
def test_micro_marine_fixtures_on_the_device(generic):
    """F4a / F4b (marine_distribution: both exits of the walk, restart from the entry node) and F5 (two sub-steps of the
    river-borne marine diffusion with every branch of diffmarine) and F7 (a coastal lake overflowing into the sea: marine branch of
    the pit cascade) through bl_step, against the hand derivation."""
    import micro
    from bayeslands_b200.engine import Engine
    cases = (("F4a", micro.F4A_Z, micro.inputs_marine(5, 0.9), micro.f4a_derive, ("Q", "ero", "depo", "sedflux", "z")),
             ("F4b", micro.F4B_Z, micro.inputs_marine(1, 0.1), micro.f4b_derive, ("fillH", "Q", "ero", "depo", "sedflux", "z")),
             ("F5", micro.F4A_Z, micro.inputs_marine(5, 0.9, micro.F5_CDR), micro.f5_derive, ("ero", "depo", "sedflux", "z", "cumdiff")),
             ("F7", micro.F7_Z, micro.inputs_marine(5, 0.9), micro.f7_derive,
              ("fillH", "Q", "pitVol", "pitDrain", "ero", "depo", "sedflux", "z")))
    for name, Z, inp, derive, keys in cases:
        m = micro.chain_mesh(Z)
        eng = Engine(m, inp, 1, shuffle_mode=0, force_generic=generic)
        eng.reset(np.array([micro.F4_RAIN]), np.array([micro.F4_EROD]), np.array([1], dtype=np.int64))
        eng.step(5000.0)
        sc = eng.scalars()
        assert sc["last_dt"][0] == micro.F4_DT and sc["status"][0] == 0, (name, sc)
        exp = derive()
        for k in keys:
            assert np.array_equal(eng.field(k, 0), np.asarray(exp[k])), (name, generic, k)


@pytest.mark.parametrize("generic", [False, True])
def test_micro_overfilled_closed_basin_on_the_device(generic):
    """F6: getid1 finds the overfilled closed lake, the step shrinks from 353 a to 54 a and stream power runs again."""
    import micro
    from bayeslands_b200.engine import Engine
    m = micro.chain_mesh(micro.F6_Z)
    eng = Engine(m, micro.inputs(fillmax=micro.F6_TH), 1, shuffle_mode=0, force_generic=generic)
    eng.reset(np.array([micro.F6_RAIN]), np.array([micro.F6_EROD]), np.array([1], dtype=np.int64))
    eng.step(5000.0)
    exp = micro.f6_derive()
    sc = eng.scalars()
    assert sc["last_dt"][0] == exp["dt"] == 54.0 and sc["status"][0] == 0, sc
    for k in ("fillH", "Q", "pitVol", "pitDrain", "ero", "depo", "sedflux", "z"):
        assert np.array_equal(eng.field(k, 0), np.asarray(exp[k])), (generic, k)


@pytest.mark.parametrize("generic", [False, True])
def test_micro_flow_cfl_slope_exponent_on_the_device(generic):
    """F8: the CFL step with n = 2 (hand value 1767 a); the step itself against the oracle (pow(slope, 2) is libm on the host
    and the CUDA math library on the device: 1e-12 relative instead of bit equality for the erosion)."""
    import micro
    from bayeslands_b200.engine import Engine
    from oracle import oracle
    inp = micro.inputs(fillmax=200.0)
    inp.SPLn = micro.F8_N
    m = micro.chain_mesh(micro.F2_Z)
    eng = Engine(m, inp, 1, shuffle_mode=0, force_generic=generic)
    eng.reset(np.array([micro.F2_RAIN]), np.array([micro.F2_EROD]), np.array([1], dtype=np.int64))
    eng.step(5000.0)
    sc = eng.scalars()
    assert sc["last_dt"][0] == micro.f8_cfl_dt()[0] == 1767.0 and sc["status"][0] == 0, sc
    om = oracle.OracleModel(m, inp, seed=1, shuffle_mode=0)
    om.reset(micro.F2_RAIN, micro.F2_EROD, seed=1)
    om.streamflow(); om.sediment_flux(5000.0)
    for k in ("ero", "depo", "z"):
        np.testing.assert_allclose(eng.field(k, 0), np.asarray(getattr(om, k)), rtol=1e-12, atol=0.0, err_msg=k)
