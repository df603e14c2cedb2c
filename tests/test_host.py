"""Host-side logic: XML subset, stop-time schedule, mesh contract, C-ABI symbols (no GPU compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_xml_defaults_and_values(crater_fast):
    d, inp, m = crater_fast
    assert inp.btype == "fixed" and inp.tEnd == 15000.0 and inp.tDisplay == 5000.0 and inp.maxDT == 5000.0
    assert inp.fillmax == 15.0 and inp.SPLm == 0.5 and inp.SPLn == 1.0 and inp.SPLero == 9e-5
    assert inp.CDa == 0.02 and inp.CDm == 0.05 and inp.CDr == 0.0 and inp.depo == 1 and inp.seapos == -1000.0
    from bayeslands_b200 import config
    e = config.parse_xml(str(np.load(os.path.join(ROOT, "data", "etopo_fast.npz"))["xml"]))
    assert e.btype == "slope" and e.Afactor == 4 and e.diffnb == 10 and e.diffprop == 0.05 and e.CDr == 10.0
    assert e.fillmax == 200.0 and e.seafile.endswith("sealevel.csv")
    with pytest.raises(ValueError):
        config.parse_xml("<badlands><time><start>0</start><end>1</end><display>1</display></time></badlands>")
    with pytest.raises(ValueError):
        config.parse_xml(str(d["xml"]).replace("<display>5000.</display>", "<display>4000.</display>"))


def test_schedule_reproduces_run_to_time_stops():
    """model.py:246-259, :447-486, :529-533: next_display is bumped inside the loop AND after every call."""
    from bayeslands_b200 import config, hostparams
    x = str(np.load(os.path.join(ROOT, "data", "crater.npz"))["xml"])
    inp = config.parse_xml(x)
    s = hostparams.Schedule(inp)
    t, out = 0.0, []
    for T in hostparams.sim_interval(50000):
        tEnd, segs = s.stops_for(t, float(T))
        out.append([g[0] for g in segs]); t = tEnd
    assert out == [[], [5000.0, 10000.0, 12500.0], [20000.0, 25000.0], [30000.0, 35000.0, 37500.0], [45000.0, 50000.0]]
    assert hostparams.sim_interval(15000).tolist() == [0, 3750, 7500, 11250, 15000]
    # clamp to the XML end time (model.py:240-243)
    s2 = hostparams.Schedule(inp)
    tEnd, segs = s2.stops_for(0.0, 1e9)
    assert tEnd == 50000.0 and segs[-1][0] == 50000.0


def test_mesh_contract(crater_fast):
    d, inp, m = crater_fast
    fv = m.FVmesh
    N = m.N
    assert m.boundsPt == 488 and fv.neighbours.shape == (N, 20) and fv.neighbours.dtype == np.int32
    assert (fv.neighbours[fv.neighbours < 0] == -2).all()
    # symmetric adjacency, ghost cells have zero area, total area = (2400 + 20)^2
    for i in range(0, N, 53):
        for j in fv.neighbours[i][fv.neighbours[i] >= 0]:
            assert i in fv.neighbours[j]
    assert (fv.control_volumes[:m.boundsPt] == 0).all() and (fv.control_volumes[m.boundsPt:] > 0).all()
    assert abs(fv.control_volumes.sum() - 2420.0 ** 2) < 1.0
    # float32 round trip (FVmethod.py:152) and output grid (bl_mcmc.py:186-194)
    assert np.array_equal(fv.control_volumes, fv.control_volumes.astype(np.float32).astype(np.float64))
    assert (m.grid_ny, m.grid_nx) == (123, 123) == d["final_elev"].shape
    # colour-ordered numbering: the Gauss-Seidel dependency depth of fillPD is the number of colours
    lev = np.zeros(N, dtype=int)
    for k in range(m.boundsPt, N):
        nb = fv.neighbours[k]; nb = nb[(nb >= m.boundsPt) & (nb < k)]
        lev[k] = 1 + (lev[nb].max() if len(nb) else 0)
    assert lev.max() <= 6
    # ghost elevations: 'fixed' copies the nearest interior neighbour (elevationTIN.py:52-75)
    assert np.array_equal(m.elevation[:m.boundsPt], m.elevation[m.parentIDs])


def test_cabi_exports_every_declared_symbol():
    from bayeslands_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "bayeslands_b200.h")).read()
    declared = set(re.findall(r"\b(bl_[a-z_]+)\s*\(", hdr)) - {"bl_get_status"}
    assert declared, "no declarations parsed"
    L = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert set(_lib.EXPORTS) <= declared
    assert _lib.load().bl_version() >= 100


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bayeslands_b200 import _lib
    from bayeslands_b200.engine import Engine
    with pytest.raises(_lib.BlError):
        Engine(None, None, 1)


def test_lattice_arrays_match_the_mesh_builder():
    """`lattice.regular` (no neighbour tables; what the 2048 x 2048 config uses) == `lattice.from_mesh` of
    `mesh.build_regular(order="colour")` + `forcing.load_tecto_map`; colour order is a pure renumbering."""
    from bayeslands_b200 import mesh, forcing, lattice
    nx, ny, dx, T = 14, 10, 100.0, 1000.0
    rng = np.random.default_rng(5)
    z0 = 10.0 + rng.uniform(0, 1, (ny, nx))
    X, Y = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    bump = 1e-3 * T * np.sin(np.pi * X / (nx - 1)) * np.sin(np.pi * Y / (ny - 1))
    for btype in ("fixed", "wall"):
        m = mesh.build_regular(nx, ny, dx, z0, btype=btype, order="colour")
        disp = forcing.load_tecto_map(m, bump.T.ravel(order="F"), 0.0, T)
        a = lattice.from_mesh(m, disp)
        b = lattice.regular(nx, ny, dx, z0, btype=btype, disp_rate=a.disp[1:-1, 1:-1])
        for f in ("dist", "vor", "edge", "z0", "flags", "parent", "disp"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), (btype, f)
        assert a.area == b.area and a.hillslope_min_edge2 == b.hillslope_min_edge2
    # row-major and colour-major meshes are the same landscape under a permutation of the node ids
    r = mesh.build_regular(nx, ny, dx, z0, order="row")
    c = mesh.build_regular(nx, ny, dx, z0, order="colour")
    perm = np.empty(r.N, dtype=np.int64)
    perm[r.extra["regular"]["ext"].ravel()] = c.extra["regular"]["ext"].ravel()
    assert np.array_equal(r.elevation, c.elevation[perm]) and np.array_equal(r.borders2, c.borders2[perm])
    with pytest.raises(ValueError):
        lattice.from_mesh(r)
    # the interior uplift rate interpolated by load_tecto_map is the map itself on a lattice = DEM grid
    assert np.allclose(a.disp[1:-1, 1:-1], bump / T, rtol=0, atol=1e-18)


def test_fv_arithmetic_on_analytic_cells():
    """FVclass.f90:605-856 through mesh.build_fv on point sets whose Voronoi cells are known in closed form: a regular
    triangular lattice (hexagonal cells: area sqrt(3)/2 a^2, six neighbours at distance a across Voronoi edges of
    a / sqrt(3)) and a square lattice (square cells; the diagonal Delaunay edges Qhull has to invent between co-circular
    points carry Voronoi edges of length zero).  Values pass through the float32 round trip of FVmethod.py:152."""
    from bayeslands_b200 import mesh as M
    a = 100.0
    f32 = lambda v: float(np.float32(v))
    # triangular lattice, 9 x 9
    pts = np.array([[a * (i + 0.5 * (j % 2)), a * np.sqrt(3.0) / 2.0 * j] for j in range(9) for i in range(9)])
    fv = M.build_fv(pts, res_cap=1.0e9)
    inner = [j * 9 + i for j in range(2, 7) for i in range(2, 7)]
    for k in inner:
        nb = fv.neighbours[k]
        deg = int((nb >= 0).sum())
        assert deg == 6 and sorted(np.round(np.hypot(*(pts[nb[:6]] - pts[k]).T), 6)) == [a] * 6
        assert np.allclose(fv.edge_length[k, :6], a, rtol=1e-6) and np.allclose(fv.vor_edges[k, :6], a / np.sqrt(3.0), rtol=1e-6)
        assert abs(fv.control_volumes[k] - np.sqrt(3.0) / 2.0 * a * a) <= 1e-6 * a * a
        assert fv.control_volumes[k] == f32(fv.control_volumes[k])
    hull = [k for k in range(81) if k // 9 in (0, 8)]            # bottom and top row: on the convex hull
    assert (fv.control_volumes[hull] == 0.0).all()               # unbounded cells export zero area (FVclass.f90:760)
    # square lattice, 7 x 7
    pts = np.array([[a * i, a * j] for j in range(7) for i in range(7)], dtype=np.float64)
    fv = M.build_fv(pts, res_cap=1.0e9)
    for k in [j * 7 + i for j in range(2, 5) for i in range(2, 5)]:
        nb = fv.neighbours[k]; nb = nb[nb >= 0]
        dist = np.hypot(*(pts[nb] - pts[k]).T)
        axis = np.isclose(dist, a)
        assert axis.sum() == 4 and np.isclose(dist[~axis], a * np.sqrt(2.0)).all()
        assert np.allclose(fv.vor_edges[k, :len(nb)][axis], a, rtol=1e-6) and np.allclose(fv.vor_edges[k, :len(nb)][~axis], 0.0, atol=1e-6 * a)
        assert abs(fv.control_volumes[k] - a * a) <= 1e-6 * a * a


def test_bench_reference_arm_prints_its_line():
    """`bench.py --impl reference` (the CPU-oracle arm the driver runs next to the GPU arm) on the host cores: one JSON line
    with the metric, the arm's own cpu_baseline and a zero-copy e2e block.  Also guards the fork-safety of the config
    arrays (they are read eagerly: a lazily read NpzFile shared between the pool workers corrupted their zlib streams)."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    import bench
    d = bench.load_config("crater_fast")[0]
    assert isinstance(d, dict) and isinstance(d["final_elev"], np.ndarray)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "crater_fast",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mcmc_forward_model_evals_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
