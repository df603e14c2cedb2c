"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same inputs.

Bar: bit-exact for integer/index outputs (receivers, stack order, pit ids, drainage) AND for every fp64
array, because the kernels keep the oracle's operation order and are compiled without FMA contraction.
"""
import numpy as np
import pytest

from conftest import CRATER_FAST_COORDS, ETOPO_FAST_COORDS

pytestmark = pytest.mark.gpu

INT_FIELDS = ["rcv1", "rcv", "stack1", "stack", "pitID", "pitDrain", "allDrain"]
FP_FIELDS = ["fillH", "maxh", "pitVol", "Q", "ero", "depo", "sedflux", "z", "cumdiff", "cumhill"]

PROPOSALS = [(1.5, 5e-5), (0.3, 3.1e-5), (2.9, 6.9e-5), (0.0, 5e-5), (1.98, 4.36e-5), (1.1, 9e-5)]


def _engine(crater_fast, B, **kw):
    from bayeslands_b200.engine import Engine
    d, inp, m = crater_fast
    return Engine(m, inp, B, real_elev=d["final_elev"], erdp_coords=CRATER_FAST_COORDS, **kw)


@pytest.mark.parametrize("shuffle", [0, 1])
def test_stepwise_all_stages_bit_exact(crater_fast, shuffle):
    from oracle import oracle
    d, inp, m = crater_fast
    B = len(PROPOSALS)
    eng = _engine(crater_fast, B, shuffle_mode=shuffle)
    rain = np.array([p[0] for p in PROPOSALS]); erod = np.array([p[1] for p in PROPOSALS])
    seeds = np.arange(B, dtype=np.int64) + 17
    eng.reset(rain, erod, seeds)
    oms = []
    for b in range(B):
        om = oracle.OracleModel(m, inp, pow_mode=1, shuffle_mode=shuffle, fill_mode=0, seed=int(seeds[b]))
        om.reset(rain[b], erod[b], seed=int(seeds[b]))
        oms.append(om)
    for step in range(12):
        eng.step(15000.0)
        sc = eng.scalars()
        for b, om in enumerate(oms):
            om.streamflow(); om.sediment_flux(15000.0)
            for f in INT_FIELDS:
                g = eng.field(f, b); o = getattr(om, f)
                assert np.array_equal(g, o), "step %d proposal %d %s: %d mismatches" % (step, b, f, (g != o).sum())
            for f in FP_FIELDS:
                g = eng.field(f, b); o = getattr(om, f)
                assert np.array_equal(g, o), "step %d proposal %d %s: max abs diff %g" % (step, b, f, np.abs(g - o).max())
            assert sc["tNow"][b] == om.tNow and sc["last_dt"][b] == om.last_dt
            assert sc["status"][b] == 0


def test_full_run_blackbox_bit_exact(crater_fast):
    from oracle import oracle
    from bayeslands_b200 import hostparams
    d, inp, m = crater_fast
    B = len(PROPOSALS)
    eng = _engine(crater_fast, B)
    rain = np.array([p[0] for p in PROPOSALS]); erod = np.array([p[1] for p in PROPOSALS])
    seeds = np.arange(B, dtype=np.int64) * 3 + 1
    eng.reset(rain, erod, seeds)
    times = hostparams.sim_interval(15000)
    pts, sse, elev, erdp = eng.run_schedule(times, grids=True)
    pts, sse, elev, erdp = pts.cpu().numpy(), sse.cpu().numpy(), elev.cpu().numpy(), erdp.cpu().numpy()
    sc = eng.scalars()
    assert (sc["status"] == 0).all()
    for b in range(B):
        om = oracle.OracleModel(m, inp, seed=int(seeds[b]))
        ev, ed, pv = om.blackbox(rain[b], erod[b], 15000, CRATER_FAST_COORDS, seed=int(seeds[b]))
        assert sc["steps"][b] == om.steps and sc["tNow"][b] == 15000.0
        assert np.array_equal(eng.field("z", b), om.z)
        assert np.array_equal(eng.field("cumdiff", b), om.cumdiff)
        for ci, T in enumerate(times):
            assert np.array_equal(elev[b, ci].reshape(ev[T].shape), ev[T])
            assert np.array_equal(erdp[b, ci].reshape(ed[T].shape), ed[T])
            assert np.array_equal(pts[b, ci], pv[T])
        ref = np.sum(np.square(ev[times[-1]] - d["final_elev"]))
        assert abs(sse[b] - ref) <= 1e-11 * ref


ETOPO_PROPOSALS = [(1.5, 5e-6), (2.9, 6.9e-6), (0.5, 3e-6), (0.0, 5e-6)]


@pytest.mark.parametrize("generic", [0, 1])
def test_etopo_stepwise_bit_exact(etopo_fast, generic):
    """etopo path: sea-level curve, 'slope' boundary, marine deposition (marine_distribution), river-borne marine
    diffusion sub-steps, overflow-to-sea fallback - every stage array against the oracle, both device paths."""
    from bayeslands_b200.engine import Engine
    from oracle import oracle
    d, inp, m = etopo_fast
    B = len(ETOPO_PROPOSALS)
    eng = Engine(m, inp, B, real_elev=d["final_elev"], erdp_coords=ETOPO_FAST_COORDS, sealevel=d["sealevel"],
                 force_generic=bool(generic))
    rain = np.array([p[0] for p in ETOPO_PROPOSALS]); erod = np.array([p[1] for p in ETOPO_PROPOSALS])
    seeds = np.arange(B, dtype=np.int64) + 5
    eng.reset(rain, erod, seeds)
    oms = []
    for b in range(B):
        om = oracle.OracleModel(m, inp, seed=int(seeds[b]), sealevel=d["sealevel"])
        om.reset(rain[b], erod[b], seed=int(seeds[b]))
        oms.append(om)
    for step in range(25):
        eng.step(500000.0)
        sc = eng.scalars()
        for b, om in enumerate(oms):
            om.streamflow(); om.sediment_flux(500000.0)
            for f in INT_FIELDS:
                g = eng.field(f, b); o = getattr(om, f)
                assert np.array_equal(g, o), "step %d proposal %d %s: %d mismatches" % (step, b, f, (g != o).sum())
            for f in FP_FIELDS:
                g = eng.field(f, b); o = getattr(om, f)
                assert np.array_equal(g, o), "step %d proposal %d %s: max abs diff %g" % (step, b, f, np.abs(g - o).max())
            assert sc["tNow"][b] == om.tNow and sc["last_dt"][b] == om.last_dt and sc["status"][b] == 0


def test_etopo_full_run_bit_exact(etopo_fast):
    from bayeslands_b200.engine import Engine
    from bayeslands_b200 import hostparams
    from oracle import oracle
    d, inp, m = etopo_fast
    B = len(ETOPO_PROPOSALS)
    eng = Engine(m, inp, B, real_elev=d["final_elev"], erdp_coords=ETOPO_FAST_COORDS, sealevel=d["sealevel"])
    rain = np.array([p[0] for p in ETOPO_PROPOSALS]); erod = np.array([p[1] for p in ETOPO_PROPOSALS])
    seeds = np.arange(B, dtype=np.int64) + 50
    eng.reset(rain, erod, seeds)
    times = hostparams.sim_interval(500000)
    pts, sse, elev, erdp = eng.run_schedule(times, grids=True)
    pts, elev = pts.cpu().numpy(), elev.cpu().numpy()
    sc = eng.scalars()
    assert (sc["status"] == 0).all()
    for b in range(B):
        om = oracle.OracleModel(m, inp, seed=int(seeds[b]), sealevel=d["sealevel"])
        ev, ed, pv = om.blackbox(rain[b], erod[b], 500000, ETOPO_FAST_COORDS, seed=int(seeds[b]))
        assert sc["steps"][b] == om.steps
        assert np.array_equal(eng.field("z", b), om.z) and np.array_equal(eng.field("cumdiff", b), om.cumdiff)
        for ci, T in enumerate(times):
            assert np.array_equal(elev[b, ci].reshape(ev[T].shape), ev[T]) and np.array_equal(pts[b, ci], pv[T])


def test_crater_generic_path_bit_exact(crater_fast):
    """The generic (L2-resident) device path stays bit-exact too: it serves meshes too large for shared memory."""
    from oracle import oracle
    d, inp, m = crater_fast
    eng = _engine(crater_fast, 2, force_generic=True)
    rain = np.array([1.5, 2.7]); erod = np.array([5e-5, 3.3e-5]); seeds = np.array([3, 4], dtype=np.int64)
    eng.reset(rain, erod, seeds)
    oms = [oracle.OracleModel(m, inp, seed=int(s)) for s in seeds]
    for b, om in enumerate(oms):
        om.reset(rain[b], erod[b], seed=int(seeds[b]))
    for step in range(6):
        eng.step(15000.0)
        for b, om in enumerate(oms):
            om.streamflow(); om.sediment_flux(15000.0)
            for f in INT_FIELDS + FP_FIELDS:
                assert np.array_equal(eng.field(f, b), getattr(om, f)), (step, b, f)


@pytest.mark.parametrize("tpb", [256, 512, 768, 1024])
def test_crater_every_cta_size_bit_exact(crater_fast, tpb):
    """The device code is instantiated once per CTA size (csrc/bl_v256.cu / bl_v512.cu / bl_v1024.cu); `opt_tpb` forces
    one.  Block-level primitives (scans, ordered compaction, warp ballots) must give the same bits at every size."""
    from oracle import oracle
    d, inp, m = crater_fast
    eng = _engine(crater_fast, 2, tpb=tpb)
    rain = np.array([0.9, 2.2]); erod = np.array([6e-5, 4e-5]); seeds = np.array([8, 9], dtype=np.int64)
    eng.reset(rain, erod, seeds)
    oms = [oracle.OracleModel(m, inp, seed=int(s)) for s in seeds]
    for b, om in enumerate(oms):
        om.reset(rain[b], erod[b], seed=int(seeds[b]))
    for step in range(5):
        eng.step(15000.0)
        sc = eng.scalars()
        for b, om in enumerate(oms):
            om.streamflow(); om.sediment_flux(15000.0)
            for f in INT_FIELDS + FP_FIELDS:
                assert np.array_equal(eng.field(f, b), getattr(om, f)), (tpb, step, b, f)
            assert sc["tNow"][b] == om.tNow and sc["status"][b] == 0


@pytest.mark.parametrize("generic", [0, 1])
def test_synthetic_uplift_regular_grid_bit_exact(generic):
    """BASELINE config 5 in miniature: regular grid (D8 stencil as the neighbour list), noisy flat start, uplift
    map applied every step (buildFlux.py:312-313), purely erosive (dep = 0), row-major numbering (deep Gauss-Seidel
    dependency levels -> exercises the level schedule of the fill)."""
    import os
    from bayeslands_b200 import config, mesh, forcing
    from bayeslands_b200.engine import Engine
    from oracle import oracle
    from conftest import ROOT
    nx, ny, dx = (24, 18, 100.0) if not generic else (40, 28, 100.0)   # 58 / 94 Gauss-Seidel levels
    rng = np.random.default_rng(1234)
    z0 = 10.0 + rng.uniform(0, 1, (ny, nx))
    m = mesh.build_regular(nx, ny, dx, z0, btype="fixed")
    xml = str(np.load(os.path.join(ROOT, "data", "crater_fast.npz"))["xml"])
    xml = xml.replace("<end>15000.</end>", "<end>200000.</end>").replace("<rend>15000.</rend>", "<rend>200000.</rend>")
    xml = xml.replace("<display>5000.</display>", "<display>50000.</display>").replace("<fillmax>15.</fillmax>", "<dep>0</dep>")
    xml = xml.replace("<caerial>2.e-2</caerial>", "<caerial>8.e-1</caerial>").replace("<cmarine>5.e-2</cmarine>", "<cmarine>1.</cmarine>")
    inp = config.parse_xml(xml)
    assert inp.depo == 0 and inp.fillmax == 200.0
    X, Y = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    bump = 1e-3 * 200000.0 * np.sin(np.pi * X / (nx - 1)) * np.sin(np.pi * Y / (ny - 1))      # cumulative uplift (m)
    disp = forcing.load_tecto_map(m, bump.T.ravel(order="F"), 0.0, 200000.0)
    assert disp.shape == (m.N,) and disp.max() <= 1.0001e-3 and disp.min() >= 0.0
    B = 3
    rain = np.array([0.6, 1.5, 2.4]); erod = np.array([3.2e-6, 5e-6, 6.8e-6]); seeds = np.array([1, 2, 3], dtype=np.int64)
    eng = Engine(m, inp, B, disp=disp, force_generic=bool(generic))
    eng.reset(rain, erod, seeds)
    oms = [oracle.OracleModel(m, inp, seed=int(s), disp=disp) for s in seeds]
    for b, om in enumerate(oms):
        om.reset(rain[b], erod[b], seed=int(seeds[b]))
    for step in range(8):
        eng.step(200000.0)
        sc = eng.scalars()
        for b, om in enumerate(oms):
            om.streamflow(); om.sediment_flux(200000.0)
            for f in INT_FIELDS + FP_FIELDS:
                assert np.array_equal(eng.field(f, b), getattr(om, f)), (step, b, f)
            assert sc["tNow"][b] == om.tNow and sc["status"][b] == 0


def test_first_step_receivers_match_reference_sfd_vectors(crater_fast):
    """The real-surface receivers of the first time step are `sfd.directions_base` of the initial elevation: the CUDA
    path (through the C-ABI) must return exactly what the reference's own sfd.c returned (tests/golden)."""
    import os
    from bayeslands_b200.engine import Engine
    from conftest import ROOT
    d, inp, m = crater_fast
    g = np.load(os.path.join(ROOT, "tests", "golden", "sfd_reference_vectors.npz"))
    for generic in (False, True):
        eng = Engine(m, inp, 2, force_generic=generic)
        eng.reset(np.array([1.5, 0.5]), np.array([5e-5, 4e-5]), np.array([1, 2], dtype=np.int64))
        eng.step(15000.0)
        for b in range(2):
            assert np.array_equal(eng.field("rcv1", b), g["rcv1_0"]), (generic, b)
