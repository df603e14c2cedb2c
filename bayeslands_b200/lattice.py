"""Lattice description of a regular-grid landscape for the lattice engine (`bl_grid_*`, BASELINE config 5).

Two ways to get one:
  * `from_mesh(mesh, disp)`  - from `mesh.build_regular(..., order="colour")` arrays, checking that the mesh really
    is the regular D8 lattice the engine assumes (used by the parity tests, which need the mesh for the oracle);
  * `regular(nx, ny, dx, z, btype, disp_map, ...)` - the same arrays written down directly, without the (N, 20)
    neighbour tables (4.2 M nodes would need > 2 GB of them); `tests/test_host.py` checks both agree.

Everything is in lattice order: the (ny+2, nx+2) extended lattice including the ghost ring, row-major, x fastest.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import mesh as meshmod

# D8 slots in the neighbour order of mesh.build_regular: (dj, di)
SLOTS = [(0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1)]


@dataclass
class Lattice:
    nx: int
    ny: int
    dx: float
    dist: np.ndarray          # (8,) |xy(d) - xy(r)|
    vor: np.ndarray           # (8,)
    edge: np.ndarray          # (8,)
    area: float
    z0: np.ndarray            # (ny+2, nx+2)
    flags: np.ndarray         # (ny+2, nx+2) uint8: bit0 borders, bit1 borders2
    parent: np.ndarray        # (nb,) int32 lattice cell of each ghost's parent
    disp: Optional[np.ndarray]  # (ny+2, nx+2) or None
    hillslope_min_edge2: float
    ext: Optional[np.ndarray] = None    # (ny+2, nx+2) mesh ids (only when built from a mesh)

    @property
    def shape(self):
        return (self.ny + 2, self.nx + 2)

    @property
    def nb(self):
        return 2 * (self.nx + 2) + 2 * self.ny


def _ghost_cells(nx, ny):
    """Lattice cell (J, I) of ghost id g, raster2TIN order (south row, x = -dx column, north row, x = nx*dx column)."""
    J = np.concatenate((np.zeros(nx + 2, int), np.arange(1, ny + 1), np.full(nx + 2, ny + 1), np.arange(1, ny + 1)))
    I = np.concatenate((np.arange(nx + 2), np.zeros(ny, int), np.arange(nx + 2), np.full(ny, nx + 1)))
    return J, I


def _slot_geometry(dx):
    dist = np.array([np.sqrt((dx * di) ** 2 + (dx * dj) ** 2) for dj, di in SLOTS])
    edge = np.array([dx if (di == 0 or dj == 0) else dx * np.sqrt(2.0) for dj, di in SLOTS])
    vor = np.array([dx if (di == 0 or dj == 0) else 0.0 for dj, di in SLOTS])
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    return dist, f32(vor), f32(edge)


def regular(nx: int, ny: int, dx: float, z: np.ndarray, btype: str = "fixed",
            disp_rate: Optional[np.ndarray] = None) -> Lattice:
    """Lattice arrays of `mesh.build_regular(nx, ny, dx, z, btype, order="colour")` (+ `forcing.load_tecto_map`'s
    rate when `disp_rate` (ny, nx), m/a on the interior nodes, is given) without building the mesh."""
    if nx % 2 or ny % 2:
        raise ValueError("the lattice engine needs even nx, ny")
    z = np.asarray(z, dtype=np.float64).reshape(ny, nx)
    dist, vor, edge = _slot_geometry(float(dx))
    Jg, Ig = _ghost_cells(nx, ny)
    # parent = nearest non-ghost neighbour: the axis neighbour, or the diagonal one at the 4 corners
    # (surface/elevationTIN.py:22-208 picks argmin(edge_length) among the non-ghost neighbours)
    Jp, Ip = np.clip(Jg, 1, ny), np.clip(Ig, 1, nx)
    z0 = np.zeros((ny + 2, nx + 2))
    z0[1:-1, 1:-1] = z
    thetype = 1 if btype in ("slope", "outlet", "wall1") else 0
    if thetype == 0:
        z0[Jg, Ig] = z0[Jp, Ip]
    else:
        # second node along the same direction: nearest non-ghost neighbour of the parent by edge length = its first
        # axis neighbour in slot order (elevationTIN.py:64-80); handled by the mesh builder for small lattices only
        raise NotImplementedError("slope-type boundaries: build the lattice with from_mesh()")
    if btype == "wall":
        z0[Jg, Ig] = 1.0e7
    flags = np.zeros((ny + 2, nx + 2), dtype=np.uint8)
    flags[1:-1, 1:-1] |= 1                                        # borders: area > 0 inside the DEM box
    # borders2: strictly inside the DEM perimeter by more than 1 m (buildFlux.py:149-151)
    xs, ys = dx * np.arange(nx), dx * np.arange(ny)
    in2x = (xs > xs.min() + 1.0) & (xs < xs.max() - 1.0)
    in2y = (ys > ys.min() + 1.0) & (ys < ys.max() - 1.0)
    flags[1:-1, 1:-1] |= (2 * (in2y[:, None] & in2x[None, :])).astype(np.uint8)
    disp = None
    if disp_rate is not None:
        disp = np.zeros((ny + 2, nx + 2))
        disp[1:-1, 1:-1] = np.asarray(disp_rate, dtype=np.float64).reshape(ny, nx)
        disp[Jg, Ig] = disp[Jp, Ip]                                # forceSim.disp_border (forceSim.py:400-449)
    area = float(np.float32(dx * dx))
    return Lattice(nx=nx, ny=ny, dx=float(dx), dist=dist, vor=vor, edge=edge, area=area, z0=z0, flags=flags,
                   parent=(Jp * (nx + 2) + Ip).astype(np.int32), disp=disp,
                   hillslope_min_edge2=float(np.amin(edge[edge > 0.0] ** 2)))


def from_mesh(mesh, disp: Optional[np.ndarray] = None) -> Lattice:
    """Lattice view of a `mesh.build_regular(..., order="colour")` mesh; raises if the mesh is anything else."""
    info = mesh.extra.get("regular")
    if info is None or info["order"] != "colour":
        raise ValueError("the lattice engine needs a mesh from build_regular(..., order='colour')")
    nx, ny, dx, ext = info["nx"], info["ny"], info["dx"], info["ext"]
    fv = mesh.FVmesh
    if not np.array_equal(ext, meshmod.regular_ext_ids(nx, ny, "colour")):
        raise ValueError("unexpected node numbering")
    # neighbour table == D8 lattice in slot order
    LY, LX = ny + 2, nx + 2
    J, I = np.meshgrid(np.arange(LY), np.arange(LX), indexing="ij")
    want = np.full((mesh.N, meshmod.KP), -2, dtype=np.int64)
    cnt = np.zeros(mesh.N, dtype=np.int64)
    for dj, di in SLOTS:
        jj, ii = J + dj, I + di
        ok = (jj >= 0) & (jj < LY) & (ii >= 0) & (ii < LX)
        src, dst = ext[J[ok], I[ok]], ext[jj[ok], ii[ok]]
        want[src, cnt[src]] = dst
        cnt[src] += 1
    if not np.array_equal(want, fv.neighbours):
        raise ValueError("neighbour table is not the D8 lattice")
    dist, vor, edge = _slot_geometry(dx)
    xy = fv.node_coords
    inner = ext[2:-2, 2:-2].ravel()                                # nodes whose 8 neighbours are all interior
    for s, (dj, di) in enumerate(SLOTS):
        nbr = fv.neighbours[inner, s]
        d = np.sqrt((xy[inner, 0] - xy[nbr, 0]) ** 2 + (xy[inner, 1] - xy[nbr, 1]) ** 2)
        if not (np.all(d == dist[s]) and np.all(fv.edge_length[inner, s] == edge[s]) and np.all(fv.vor_edges[inner, s] == vor[s])):
            raise ValueError("slot geometry is not uniform")
    # every receiver distance the engine can meet (also next to / on the ghost ring) must be the slot's distance
    Jn, In = np.meshgrid(np.arange(LY), np.arange(LX), indexing="ij")
    for s, (dj, di) in enumerate(SLOTS):
        jj, ii = Jn + dj, In + di
        ok = (jj >= 0) & (jj < LY) & (ii >= 0) & (ii < LX)
        a, b = ext[Jn[ok], In[ok]], ext[jj[ok], ii[ok]]
        d = np.sqrt((xy[a, 0] - xy[b, 0]) ** 2 + (xy[a, 1] - xy[b, 1]) ** 2)
        if not np.all(d == dist[s]):
            raise ValueError("node spacing is not exactly uniform in floating point")
    area = fv.control_volumes
    if not np.all(area[ext[1:-1, 1:-1]] == area[ext[1, 1]]) or np.any(area[: mesh.boundsPt] != 0.0):
        raise ValueError("control volumes are not uniform")
    flags = (mesh.borders[ext] != 0).astype(np.uint8) | (2 * (mesh.borders2[ext] != 0)).astype(np.uint8)
    inv = np.empty(mesh.N, dtype=np.int64)
    inv[ext.ravel()] = np.arange(LY * LX)
    return Lattice(nx=nx, ny=ny, dx=dx, dist=dist, vor=vor, edge=edge, area=float(area[ext[1, 1]]),
                   z0=np.ascontiguousarray(mesh.elevation[ext]), flags=np.ascontiguousarray(flags),
                   parent=inv[mesh.parentIDs].astype(np.int32),
                   disp=None if disp is None else np.ascontiguousarray(np.asarray(disp, dtype=np.float64)[ext]),
                   hillslope_min_edge2=mesh.hillslope_min_edge2, ext=ext)
