"""Lattice engine (`bl_grid_*`, BASELINE config 5) against the CPU oracle, bit for bit, through the C-ABI."""
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

FP = ("z", "fillH", "Q", "cumdiff", "cumhill")


def synthetic(nx, ny, dx=100.0, btype="fixed", T=200000.0, seed=1234, amp=1.0):
    """SURVEY.md 8d C5 in miniature: z0 = 10 + U(0,1), sin-bump uplift, dep = 0, mountain.xml creep values."""
    from bayeslands_b200 import config, mesh, forcing, lattice
    rng = np.random.default_rng(seed)
    z0 = 10.0 + amp * rng.uniform(0, 1, (ny, nx))
    m = mesh.build_regular(nx, ny, dx, z0, btype=btype, order="colour")
    xml = str(np.load(os.path.join(ROOT, "data", "crater_fast.npz"))["xml"])
    xml = xml.replace("<end>15000.</end>", "<end>%r</end>" % T).replace("<rend>15000.</rend>", "<rend>%r</rend>" % T)
    xml = xml.replace("<display>5000.</display>", "<display>%r</display>" % (T / 4)).replace("<fillmax>15.</fillmax>", "<dep>0</dep>")
    xml = xml.replace("<caerial>2.e-2</caerial>", "<caerial>8.e-1</caerial>").replace("<cmarine>5.e-2</cmarine>", "<cmarine>1.</cmarine>")
    if btype != "fixed":
        xml = xml.replace("<boundary>fixed</boundary>", "<boundary>%s</boundary>" % btype)
    inp = config.parse_xml(xml)
    assert inp.depo == 0 and inp.btype == btype
    X, Y = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    bump = 1e-3 * T * np.sin(np.pi * X / (nx - 1)) * np.sin(np.pi * Y / (ny - 1))
    disp = forcing.load_tecto_map(m, bump.T.ravel(order="F"), 0.0, T)
    return m, inp, disp, lattice.from_mesh(m, disp)


@pytest.mark.parametrize("btype,nx,ny", [("fixed", 64, 40), ("slope", 40, 36), ("fixed", 130, 96), ("fixed", 128, 6)])
def test_lattice_stepwise_bit_exact(btype, nx, ny):
    from bayeslands_b200.grid_engine import GridEngine
    from oracle import oracle
    m, inp, disp, lat = synthetic(nx, ny, btype=btype)
    rain = np.array([0.6, 1.5, 2.4]); erod = np.array([3.2e-6, 5e-6, 6.8e-6])
    eng = GridEngine(lat, inp, 3, debug=True)
    eng.reset(rain, erod)
    oms = [oracle.OracleModel(m, inp, seed=1, disp=disp) for _ in range(3)]
    for b, om in enumerate(oms):
        om.reset(rain[b], erod[b], seed=1)
    for step in range(6):
        eng.step(200000.0)
        sc = eng.scalars()
        for b, om in enumerate(oms):
            om.streamflow(); om.sediment_flux(200000.0)
            assert np.array_equal(eng.field("rcv", b), om.rcv), (step, b, "rcv")
            for f in FP:
                assert np.array_equal(eng.field(f, b), getattr(om, f)), (step, b, f)
            assert sc["tNow"][b] == om.tNow and sc["steps"][b] == step + 1


def test_lattice_full_run_and_likelihood():
    """run_schedule over the blackBox checkpoints: ragged adaptive stepping, grids and SSE against the oracle."""
    from bayeslands_b200.grid_engine import GridEngine
    from oracle import oracle
    T = 60000.0
    m, inp, disp, lat = synthetic(36, 28, T=T)
    rain = np.array([0.3, 1.1, 1.9, 2.7]); erod = np.array([3.0e-6, 4.5e-6, 5.0e-5, 2.0e-4])  # the last two are flow-CFL bound
    truth = oracle.OracleModel(m, inp, seed=1, disp=disp)
    truth.reset(1.5, 5e-6, seed=1)
    truth.run_to_time(T)
    ext = lat.ext
    real = truth.z[ext]                                   # lattice-order "observed" topography
    eng = GridEngine(lat, inp, 4, real_elev=real)
    eng.reset(rain, erod)
    times = [T / 4, T / 2, 3 * T / 4, T]
    _, sse, elev, erdp = eng.run_schedule(times, grids=True)
    sse = sse.cpu().numpy(); elev = elev.cpu().numpy(); erdp = erdp.cpu().numpy()
    sc = eng.scalars()
    for b in range(4):
        om = oracle.OracleModel(m, inp, seed=1, disp=disp)
        om.reset(rain[b], erod[b], seed=1)
        for ci, t in enumerate(times):
            om.run_to_time(t)
            assert np.array_equal(elev[b, ci], om.z[ext].ravel()), (b, ci)
            assert np.array_equal(erdp[b, ci], om.cumdiff[ext].ravel()), (b, ci)
        assert sc["tNow"][b] == T and sc["steps"][b] == om.steps
        want = float(np.sum((om.z[ext] - real) ** 2))
        assert abs(sse[b] - want) <= 1e-11 * max(1.0, abs(want))
    assert len(set(sc["steps"].tolist())) > 1            # the proposals really took different numbers of steps


def test_lattice_engine_rejects_out_of_scope():
    from bayeslands_b200 import _lib, config, lattice
    from bayeslands_b200.grid_engine import GridEngine
    xml = str(np.load(os.path.join(ROOT, "data", "crater_fast.npz"))["xml"])
    inp = config.parse_xml(xml)                            # depo = 1
    lat = lattice.regular(16, 12, 100.0, np.zeros((12, 16)))
    with pytest.raises(_lib.BlError):
        GridEngine(lat, inp, 2)
    with pytest.raises(ValueError):
        lattice.regular(15, 12, 100.0, np.zeros((12, 15)))


def test_lattice_full_size_properties():
    """BASELINE config 5 at its full 2048 x 2048 size (the oracle needs minutes per step there): size-independent
    properties of two time steps - depression-free filled surface, receivers = first lowest neighbour, discharge
    conservation, elevation bookkeeping, determinism."""
    import torch
    from bayeslands_b200 import synthetic
    from bayeslands_b200.grid_engine import GridEngine
    nx = ny = 2048
    T = 50000.0
    lat, inp = synthetic.build(nx, ny, T=T)
    rain, erod = synthetic.proposals(2)
    eng = GridEngine(lat, inp, 2, debug=True)
    out = []
    for rep in range(2):
        eng.reset(rain, erod)
        eng.step(T)
        zs = [eng.field("z", b, mesh_order=False) for b in range(2)]
        eng.step(T)
        out.append([eng.field(f, 1, mesh_order=False) for f in ("z", "fillH", "Q", "cumdiff", "cumhill", "rcv")])
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)                                     # bit-reproducible run to run
    z, W, Q, cd, ch, rcv = out[0]
    sc = eng.scalars()
    assert sc["steps"].tolist() == [2, 2] and np.all(sc["fill_sweeps"] > 100)
    # the surface the second step filled is the first step's result
    z1 = zs[1]
    assert np.all(W >= z1)
    Wp = np.pad(W, 1, constant_values=np.inf)
    nbmin = np.full_like(W, np.inf)
    first = np.zeros(W.shape, dtype=np.int64)
    LY, LX = W.shape
    J, I = np.meshgrid(np.arange(LY), np.arange(LX), indexing="ij")
    low = W.copy(); lowidx = (J * LX + I)
    for dj, di in [(0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1)]:
        nb = Wp[1 + dj:1 + dj + LY, 1 + di:1 + di + LX]
        nbmin = np.minimum(nbmin, nb)
        better = nb < low
        low = np.where(better, nb, low)
        lowidx = np.where(better, (J + dj) * LX + (I + di), lowidx)
    assert np.array_equal(rcv.ravel(), lowidx.ravel())                  # sfd.c:55-111, strict <, first wins
    inner = np.zeros(W.shape, dtype=bool); inner[1:-1, 1:-1] = True
    wet = inner & (W > z1)
    assert np.all(W[wet] <= nbmin[wet] + 0.01 * (1 + 1e-12))            # PDalgo: W <= min_nb W + eps where W > z
    assert np.all(nbmin[inner] < W[inner])                              # no interior pit left on the filled surface
    # discharge: what leaves through the self-receivers is what rained (FLOWalgo.f90:53-81)
    base = rcv.ravel() == np.arange(LY * LX)
    total = lat.area * rain[1] * nx * ny
    assert abs(Q.ravel()[base].sum() - total) <= 1e-9 * total
    assert Q[inner].min() >= lat.area * rain[1]
    # bookkeeping: z = z0 + cumdiff + uplift (buildFlux.py:193-195, :257-272, :312-313), erosion only lowers
    t = float(sc["tNow"][1])
    assert np.allclose(z, lat.z0 + cd + lat.disp * t, rtol=0, atol=1e-9)
    assert np.all((cd - ch)[inner] <= 1e-12)
    del eng
    torch.cuda.empty_cache()


def test_mcmc_front_end_on_the_lattice_engine():
    """bayeslands_mcmc(engine=GridEngine): likelihood_batch and the (speculative) sampler over a regular-grid landscape,
    likelihoods equal to the oracle's (bl_mcmc.py:482-513 on the lattice-shaped grids)."""
    import random
    from bayeslands_b200.grid_engine import GridEngine
    from bayeslands_b200.mcmc import bayeslands_mcmc
    from oracle import oracle
    T = 20000.0
    m, inp, disp, lat = synthetic(24, 20, T=T)
    ext = lat.ext
    truth = oracle.OracleModel(m, inp, seed=1, disp=disp)
    truth.reset(1.5, 5e-6, seed=1)
    truth.run_to_time(T)
    real = truth.z[ext] + 0.3 * np.random.default_rng(3).standard_normal(ext.shape)
    coords = [[5, 7], [10, 12], [15, 3]]
    eng = GridEngine(lat, inp, 6, real_elev=real, erdp_coords=coords)
    real_pts = np.zeros((4, 3))
    mc = bayeslands_mcmc(True, T, 30, real, None, real_pts, coords, None, inp, [3e-6, 7e-6], [0.5, 2.5], [0.4, 0.6], [0.9, 1.1],
                         "lat", False, engine=eng)
    rain = np.array([0.7, 1.2, 1.5, 1.9, 2.2, 2.4]); erod = np.array([3.5e-6, 4e-6, 5e-6, 5.5e-6, 6e-6, 6.5e-6])
    lik, _ = mc.likelihood_batch(rain, erod)
    for b in range(6):
        om = oracle.OracleModel(m, inp, seed=1, disp=disp)
        om.reset(rain[b], erod[b], seed=1)
        om.run_to_time(T)
        want, _ = oracle.likelihood(om.z[ext], real, np.zeros((4, 3)), real_pts, False)
        assert abs(lik[b] - want) <= 1e-9 * abs(want), (b, lik[b], want)
    out = []
    for K in (1, 6):
        np.random.seed(1); random.seed(2)
        out.append(mc.sampler(speculate=K))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][2], out[1][2]) and out[0][3] == out[1][3]
