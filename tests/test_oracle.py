"""CPU tests of the oracle (C restatement) - pinned against the reference where the reference can be had.

* sfd.c is the one reference source that compiles here: oracle/_ref/libsfd_ref.so (built from
  /root/reference by oracle/Makefile) pins or_directions_base / or_directions / or_diffusion bit for bit.
* The Fortran routines have no reference tests: checked through hand cases and invariants.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def _grid_mesh(nx, ny, z, dx=1.0):
    """Small regular mesh with the D8 stencil through the product mesh builder."""
    from bayeslands_b200 import mesh
    return mesh.build_regular(nx, ny, dx, z, btype="fixed")


class _Inp:
    SPLm, SPLn, CDa, CDm, CDr, minDT, maxDT = 0.5, 1.0, 0.02, 0.05, 0.0, 1.0, 5000.0
    diffprop, diffnb, depo, btype, fillmax, SPLero = 0.9, 5, 1, "fixed", 15.0, 5e-5
    Hillslope, tStart, tEnd, tDisplay, seapos, seafile = True, 0.0, 15000.0, 5000.0, -1000.0, None
    rainVal = np.array([1.0]); rainTime = np.array([[0.0, 15000.0]]); tectTime = np.array([[1e9, 2e9]])


def test_sfd_3x3_hand_case():
    # 3x3 grid probe of sfd.c (SURVEY.md section 4 idea): lowest-elevation neighbour wins
    L = oracle.lib()
    z = np.array([9., 8., 7., 9.5, 6.5, 6., 10., 9., 5.])
    # 4-neighbour lists on a 3x3 grid, order: E, N, W, S (any fixed order)
    ngb = np.full((9, 20), -2, dtype=np.int32)
    for j in range(3):
        for i in range(3):
            k, q = j * 3 + i, 0
            for dj, di in ((0, 1), (1, 0), (0, -1), (-1, 0)):
                jj, ii = j + dj, i + di
                if 0 <= jj < 3 and 0 <= ii < 3:
                    ngb[k, q] = jj * 3 + ii; q += 1
    m = oracle._Mesh(); m.N = 9; m.bounds = 0; m.ngb = ngb.ctypes.data_as(c_ip)
    base = np.zeros(9, dtype=np.int32); rcv = np.zeros(9, dtype=np.int32)
    L.or_directions_base(C.byref(m), z.ctypes.data_as(c_dp), base.ctypes.data_as(c_ip), rcv.ctypes.data_as(c_ip))
    assert rcv.tolist() == [1, 4, 5, 4, 5, 8, 7, 8, 8]  # lowest neighbour, not steepest
    assert (base >= 0).sum() == 1 and base[8] == 8
    ref = oracle.ref_sfd()
    if ref is not None:
        b2 = np.zeros(9, dtype=np.int32); r2 = np.zeros(9, dtype=np.int32)
        gids = np.arange(9, dtype=np.int32)
        ed = np.zeros((9, 20))
        ref.directions_base(z.ctypes.data_as(c_dp), ngb.ctypes.data_as(c_ip), ed.ctypes.data_as(c_dp),
                            ed.ctypes.data_as(c_dp), gids.ctypes.data_as(c_ip), b2.ctypes.data_as(c_ip),
                            r2.ctypes.data_as(c_ip), 9, 9)
        assert np.array_equal(r2, rcv) and np.array_equal(b2, base)


@pytest.mark.skipif(oracle.ref_sfd() is None, reason="oracle/_ref not built (reference absent)")
def test_sfd_routines_match_reference_sfd_c(crater_fast):
    """or_directions_base / or_directions / or_diffusion == the reference's own sfd.c, bit for bit, on the
    crater mesh with perturbed (tie-rich and random) surfaces."""
    d, inp, m = crater_fast
    L, ref = oracle.lib(), oracle.ref_sfd()
    om = oracle.OracleModel(m, inp)
    N = m.N
    rng = np.random.default_rng(0)
    gids = np.arange(N, dtype=np.int32)
    ngb, ed, di = om._keep["ngb"], om._keep["vor"], om._keep["edge"]
    for trial in range(4):
        z = m.elevation + (rng.normal(0, 2.0, N) if trial else 0.0)
        if trial == 2:
            z = np.round(z)                     # many exact ties: first-in-list must win
        z = np.ascontiguousarray(z)
        fill = np.ascontiguousarray(np.maximum(z, z.mean()))
        for use_fill in (False, True):
            b = np.zeros(N, dtype=np.int32); r = np.zeros(N, dtype=np.int32)
            b2 = np.zeros(N, dtype=np.int32); r2 = np.zeros(N, dtype=np.int32)
            if not use_fill:
                L.or_directions_base(C.byref(om.cm), z.ctypes.data_as(c_dp), b.ctypes.data_as(c_ip), r.ctypes.data_as(c_ip))
                ref.directions_base(z.ctypes.data_as(c_dp), ngb.ctypes.data_as(c_ip), ed.ctypes.data_as(c_dp),
                                    di.ctypes.data_as(c_dp), gids.ctypes.data_as(c_ip), b2.ctypes.data_as(c_ip),
                                    r2.ctypes.data_as(c_ip), N, N)
            else:
                mh = np.zeros(N); md = np.zeros(N); mh2 = np.zeros(N); md2 = np.zeros(N)
                L.or_directions(C.byref(om.cm), fill.ctypes.data_as(c_dp), z.ctypes.data_as(c_dp), b.ctypes.data_as(c_ip),
                                r.ctypes.data_as(c_ip), mh.ctypes.data_as(c_dp), md.ctypes.data_as(c_dp))
                ref.directions(fill.ctypes.data_as(c_dp), z.ctypes.data_as(c_dp), ngb.ctypes.data_as(c_ip),
                               ed.ctypes.data_as(c_dp), di.ctypes.data_as(c_dp), gids.ctypes.data_as(c_ip),
                               b2.ctypes.data_as(c_ip), r2.ctypes.data_as(c_ip), mh2.ctypes.data_as(c_dp),
                               md2.ctypes.data_as(c_dp), N, N)
                assert np.array_equal(mh, mh2) and np.array_equal(md, md2)
            assert np.array_equal(r, r2) and np.array_equal(b, b2)
        df = np.zeros(N); df2 = np.zeros(N)
        L.or_diffusion(C.byref(om.cm), z.ctypes.data_as(c_dp), df.ctypes.data_as(c_dp))
        bord2 = om._keep["borders2"]
        ref.diffusion(z.ctypes.data_as(c_dp), bord2.ctypes.data_as(c_ip), ngb.ctypes.data_as(c_ip), ed.ctypes.data_as(c_dp),
                      di.ctypes.data_as(c_dp), gids.ctypes.data_as(c_ip), df2.ctypes.data_as(c_dp), N, N)
        assert np.array_equal(df, df2)


def test_pairwise_sum_is_numpy_sum():
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 8, 9, 127, 128, 129, 1000, 4097, 15129):
        a = np.ascontiguousarray(rng.normal(0, 1e3, n) ** 3)
        assert L.or_pairwise_sum(a.ctypes.data_as(c_dp), n) == (np.sum(a) if n else 0.0)


def test_interp_matches_numpy_average(crater_fast):
    d, inp, m = crater_fast
    om = oracle.OracleModel(m, inp)
    rng = np.random.default_rng(1)
    v = rng.normal(50, 20, m.N)
    got = om.interpolate(v).ravel()
    idx, w = m.grid_idx, m.grid_w
    with np.errstate(invalid="ignore"):
        want = np.average(v[idx], weights=w, axis=1)
    on = m.grid_exact == 1
    want[on] = v[idx[on, 0]]
    assert np.array_equal(got, want)


def test_step_invariants(crater_fast):
    """Invariants SURVEY.md section 4 asks for: stack is a permutation with every node after its receiver,
    sum of Q at the bases = sum(area*rain), fillH >= z, pit volumes consistent."""
    d, inp, m = crater_fast
    om = oracle.OracleModel(m, inp, seed=5)
    om.reset(1.5, 5e-5, seed=5)
    area = m.FVmesh.control_volumes
    for step in range(4):
        om.streamflow()
        for stack, rcv in ((om.stack, om.rcv), (om.stack1, om.rcv1)):
            assert np.array_equal(np.sort(stack), np.arange(m.N))
            pos = np.empty(m.N, dtype=np.int64); pos[stack] = np.arange(m.N)
            nb = rcv != np.arange(m.N)
            assert (pos[rcv[nb]] < pos[nb]).all()
        assert (om.fillH >= om.z).all()
        bases = om.rcv == np.arange(m.N)
        rain = np.where(np.arange(m.N) >= m.boundsPt, 1.5, 0.0)
        assert np.isclose(om.Q[bases].sum(), (area * rain).sum(), rtol=1e-12)
        wet = om.fillH > om.z
        pits = np.flatnonzero(om.pitVol >= 0)
        for p in pits[:20]:
            mem = (om.pitID == p) & wet
            assert np.isclose(om.pitVol[p], -1 + ((om.fillH - om.z) * area)[mem].sum(), rtol=1e-9, atol=1e-9)
        om.sediment_flux(15000.0)
    assert om.tNow > 0 and np.isfinite(om.z).all()


def test_fill_fixed_point_residual(crater_fast):
    """fill_mode=1 (run to convergence) is the greatest fixed point of W = max(z, min(min_nb W + eps, cap))."""
    d, inp, m = crater_fast
    om = oracle.OracleModel(m, inp, fill_mode=1)
    om.streamflow()
    W, z = om.fillH.copy(), om.z.copy()
    ngb = m.FVmesh.neighbours
    Wn = np.where(ngb >= 0, W[np.maximum(ngb, 0)], np.inf).min(1)
    F = np.maximum(z, np.minimum(Wn + 0.01, np.maximum(z, -1000.0) + inp.fillmax))
    inner = np.arange(m.N) >= m.boundsPt
    assert np.array_equal(W[inner], F[inner]) and np.array_equal(W[~inner], z[~inner])


def test_shuffle_is_permutation_and_seeded():
    L = oracle.lib()
    base = np.arange(100, 400, 3, dtype=np.int32)
    keys = np.zeros(len(base), dtype=np.uint64)
    outs = []
    for seed, step in ((1, 0), (1, 1), (2, 0), (1, 0)):
        b = base.copy()
        L.or_shuffle_base(b.ctypes.data_as(c_ip), len(b), C.c_uint64(seed), C.c_uint64(step), keys.ctypes.data_as(C.c_void_p))
        assert np.array_equal(np.sort(b), base)
        outs.append(b)
    assert np.array_equal(outs[0], outs[3]) and not np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[0], outs[2])


def test_sea_level_matches_scipy_interp1d():
    from scipy import interpolate
    import os
    from conftest import ROOT
    sl = np.load(os.path.join(ROOT, "data", "etopo_fast.npz"))["sealevel"]
    f = interpolate.interp1d(sl[:, 0], sl[:, 1], kind="linear")
    p = oracle._Params(); t = np.ascontiguousarray(sl[:, 0]); v = np.ascontiguousarray(sl[:, 1])
    p.nsea = len(t); p.seat = t.ctypes.data_as(c_dp); p.seav = v.ctypes.data_as(c_dp)
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for x in list(rng.uniform(0, 5e5, 200)) + [0.0, 100.0, 5e5, 250.0, 499999.0]:
        assert L.or_sea(C.byref(p), float(x)) == float(f(x))
    assert L.or_sea(C.byref(p), -5.0) == sl[0, 1] and L.or_sea(C.byref(p), 9e9) == sl[-1, 1]


# ---------------------------------------------------------------------------------------------------
# hand-derived micro-fixtures (tests/micro.py, derivations in tests/golden/README.md): they pin the restatement of the
# Fortran routines (no gfortran here) to values followed by hand from the reference's source
# ---------------------------------------------------------------------------------------------------
def test_micro_f1_capped_lake_sweep_order_and_early_stop():
    """PDalgo.f90:58-79: capped writes do not raise `change`, so the sweep can stop early and its result depends on the
    node order: the same 4-node profile fills differently when the ids ascend left-to-right or right-to-left."""
    import micro
    from oracle import oracle
    for name, order, sweeps in (("ascending", None, 1), ("descending", [5, 4, 3, 2], 2)):
        m = micro.chain_mesh(micro.F1_Z, order)
        om = oracle.OracleModel(m, micro.inputs(fillmax=15.0), seed=1)
        om.reset(1.0, 1e-5, seed=1)
        om.streamflow()
        assert np.array_equal(om.fillH[m.extra["ids"][1:5]], micro.f1_expected(name)), name
        assert om.fill_sweeps == sweeps


def test_micro_f2_two_lakes_full_step():
    """Stack order, basin ids / volumes, the drainage chain 6 -> 3 -> 0, discharge, stream power with an overflowing lake
    and the proportional land-pit fill of one full time step, array by array against the hand derivation."""
    import micro
    from oracle import oracle
    m = micro.chain_mesh(micro.F2_Z)
    om = oracle.OracleModel(m, micro.inputs(fillmax=200.0), seed=1, shuffle_mode=0)
    om.reset(micro.F2_RAIN, micro.F2_EROD, seed=1)
    om.streamflow(); om.sediment_flux(5000.0)
    dt = micro.f2_cfl_dt()
    assert om.last_dt == dt == 176.0
    exp = micro.f2_derive(dt)
    assert exp["load3"] > exp["pitVol"][3] > 0          # lake 1 overflows to the boundary in this step
    for k in ("fillH", "rcv1", "rcv", "stack", "stack1", "pitID", "pitVol", "pitDrain", "allDrain", "Q", "ero", "depo", "sedflux"):
        assert np.array_equal(np.asarray(getattr(om, k)), np.asarray(exp[k])), k


def test_micro_f2_two_pit_cascade_routine():
    """FLOWalgo.f90:967-1034 at dt = 300 a: lake 2 receives more than its volume, the excess enters lake 1, whose own load
    then overflows to the boundary node."""
    import micro
    from oracle import oracle
    m = micro.chain_mesh(micro.F2_Z)
    om = oracle.OracleModel(m, micro.inputs(fillmax=200.0), seed=1, shuffle_mode=0)
    om.reset(micro.F2_RAIN, micro.F2_EROD, seed=1)
    om.streamflow()
    exp = micro.f2_derive(300.0)
    assert exp["load6"] > exp["pitVol"][6] and exp["depo"][6] == exp["pitVol"][6] and exp["depo"][3] == exp["pitVol"][3] and exp["depo"][0] > 0
    depo, ero = om.streampower_routine(om.stack, om.rcv, om.pitID, om.pitVol, om.pitDrain, om.maxh, om.Q, om.fillH, om.z,
                                       micro.F2_EROD, micro.F2_SEA, 300.0)
    assert np.array_equal(depo, exp["depo"]) and np.array_equal(ero, exp["ero"])


def test_micro_f3_drainage_dedup_hit():
    """FLOWalgo.f90:190-234: a walk that comes back to a pit already in chainDrain walks through it (and overwrites the
    first pit's target with the terminal it finally reaches)."""
    import micro
    from oracle import oracle
    m = micro.chain_mesh([0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0])       # any 8-node mesh: the routine reads only the arrays
    om = oracle.OracleModel(m, micro.inputs(), seed=1)
    f = micro.F3
    d1, d2 = om.basindrainage_routine(f["pitID"], f["pitVol"], f["rcv"], f["fillH"], -1000.0)
    assert d1.tolist() == f["pitDrain"] and d2.tolist() == f["pitDrain"]


def _marine_step(Z, inp):
    import micro
    from oracle import oracle
    m = micro.chain_mesh(Z)
    om = oracle.OracleModel(m, inp, seed=1, shuffle_mode=0)
    om.reset(micro.F4_RAIN, micro.F4_EROD, seed=1)
    om.streamflow()
    fillH = om.fillH.copy()
    om.sediment_flux(5000.0)
    assert om.last_dt == micro.F4_DT
    return om, fillH


def test_micro_f4_marine_distribution_both_exits():
    """PDalgo.f90:136-269: F4a (ramp) - the share stays / the rest moves to the lowest neighbour; F4b (sill) - the walk
    ends in a marine pit, raises it above its rim and restarts from the entry node."""
    import micro
    om, _ = _marine_step(micro.F4A_Z, micro.inputs_marine(5, 0.9))
    exp = micro.f4a_derive()
    assert exp["diffsed"][3] > 9.0 and exp["diffsed"][4] > 9.9 and exp["z"][4] < 0.0      # the ramp keeps both under water
    for k in ("Q", "ero", "depo", "sedflux", "z"):
        assert np.array_equal(np.asarray(getattr(om, k)), np.asarray(exp[k])), ("F4a", k)
    assert np.array_equal(om.cumdiff, exp["sedflux"])
    om, fillH = _marine_step(micro.F4B_Z, micro.inputs_marine(1, 0.1))
    exp = micro.f4b_derive()
    assert exp["z"][3] == micro.F4B_Z[1] + 0.1                                            # the pit bottom ends 0.1 m above the sill
    assert np.array_equal(fillH, exp["fillH"])
    for k in ("Q", "ero", "depo", "sedflux", "z"):
        assert np.array_equal(np.asarray(getattr(om, k)), np.asarray(exp[k])), ("F4b", k)


def test_micro_f5_marine_diffusion_substeps():
    """buildFlux.py:198-247 + FLOWalgo.f90:332-388: two sub-steps (200 a, 4 a) that take every branch of diffmarine: loss to
    a lower neighbour, gain from a higher marine neighbour, loss to a ghost node."""
    import micro
    om, _ = _marine_step(micro.F4A_Z, micro.inputs_marine(5, 0.9, micro.F5_CDR))
    exp = micro.f5_derive()
    kinds = {(k, s) for k, _, s in exp["trace"]}
    assert kinds == {("gain", 200.0), ("loss", 200.0), ("gain", 4.0), ("loss", 4.0), ("ghost-loss", 4.0)}
    for k in ("ero", "depo", "sedflux", "z", "cumdiff"):
        assert np.array_equal(np.asarray(getattr(om, k)), np.asarray(exp[k])), ("F5", k)


def test_micro_f6_overfilled_closed_basin_reruns_the_step():
    """FLOWalgo.f90:1047-1086 + flowNetwork.py:662-700: a capped lake that drains to itself would receive more than it
    holds: the step shrinks from 353 a to 54 a and stream power runs a second time."""
    import micro
    from oracle import oracle
    m = micro.chain_mesh(micro.F6_Z)
    om = oracle.OracleModel(m, micro.inputs(fillmax=micro.F6_TH), seed=1, shuffle_mode=0)
    om.reset(micro.F6_RAIN, micro.F6_EROD, seed=1)
    om.streamflow(); om.sediment_flux(5000.0)
    exp = micro.f6_derive()
    assert (exp["cfl"], exp["dt"]) == (353.0, 54.0) and om.last_dt == 54.0 and om.reruns == 1
    for k in ("fillH", "Q", "pitVol", "pitDrain", "ero", "depo", "sedflux", "z"):
        assert np.array_equal(np.asarray(getattr(om, k)), np.asarray(exp[k])), k


def test_micro_f7_coastal_lake_overflows_into_the_sea():
    """FLOWalgo.f90:979-991 (marine branch of the pit cascade), :207-210 (a land pit drains to the first node below sea
    level), PDalgo.f90:167-170 (marine_distribution leaves the domain through a ghost node)."""
    import micro
    om, fillH = _marine_step(micro.F7_Z, micro.inputs_marine(5, 0.9))
    exp = micro.f7_derive()
    assert exp["depo"][4] == exp["pitVol"][4] and exp["depo"][2] > 3.0e5
    assert np.array_equal(fillH, exp["fillH"])
    for k in ("Q", "pitVol", "pitDrain", "ero", "depo", "sedflux", "z"):
        assert np.array_equal(np.asarray(getattr(om, k)), np.asarray(exp[k])), k


def test_micro_f8_flow_cfl_with_slope_exponent():
    """FLOWalgo.f90:299-330 with n = 2: the slope factor (dz / dist)^(n - 1) of the CFL condition, on the filled surface."""
    import micro
    from oracle import oracle
    inp = micro.inputs(fillmax=200.0)
    inp.SPLn = micro.F8_N
    om = oracle.OracleModel(micro.chain_mesh(micro.F2_Z), inp, seed=1, shuffle_mode=0)
    om.reset(micro.F2_RAIN, micro.F2_EROD, seed=1)
    om.streamflow(); om.sediment_flux(5000.0)
    dt, per_node = micro.f8_cfl_dt()
    assert dt == 1767.0 and om.reruns == 0 and om.last_dt == dt
    assert per_node[3] > 1.0e5 and per_node[6] > 1.0e5          # lake nodes: 1 cm of head on the filled surface


def test_micro_f9_eps_puddle_on_a_boundary_tree_drops_its_load():
    """FLOWalgo.f90:150-163 + :897-967: a node flooded by less than eps on the tree of a ghost node carries the ghost's pit
    id (area 0): the 411 851 m3 that reach it are neither routed on nor deposited."""
    import micro
    from oracle import oracle
    m = micro.chain_mesh(micro.F9_Z)
    om = oracle.OracleModel(m, micro.inputs(fillmax=200.0), seed=1, shuffle_mode=0)
    om.reset(micro.F9_RAIN, micro.F9_EROD, seed=1)
    om.streamflow()
    exp = micro.f9_derive()
    for k in ("fillH", "pitID", "pitVol", "Q"):
        assert np.array_equal(np.asarray(getattr(om, k)), np.asarray(exp[k])), k
    om.sediment_flux(5000.0)
    assert om.last_dt == exp["dt"] == 223.0 and om.reruns == 0 and exp["lost"] > 4.1e5
    for k in ("ero", "depo", "sedflux", "z"):
        assert np.array_equal(np.asarray(getattr(om, k)), np.asarray(exp[k])), k
