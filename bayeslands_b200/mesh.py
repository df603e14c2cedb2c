"""Mesh / finite-volume arrays contract - host side, built once.

The reference rebuilds this per proposal (`bl_mcmc.py:126-129` -> `pyBadlands/model.py:46-154` ->
`simulation/buildMesh.py:20-176`); here it is built once per config and shared by every proposal.

What is followed from the reference:
  * DEM CSV layout, ghost ring ("bounds", ids 0..boundsPt-1) then DEM perimeter ("edges") node order
    (`surface/raster2TIN.py:110-207`);
  * neighbour order per node = partners of edges where the node is endpoint 1, then endpoint 2
    (`libUtils/FVclass.f90:662-724`), -2 padding, KP = 20 slots (`simulation/buildMesh.py:87-90`);
  * Voronoi dual: cell area as the fan sum of 0.5*|cross| (`FVclass.f90:785-792`), Voronoi edge length,
    Delaunay edge length, border cells (unbounded Voronoi cell) zeroed (`:705`, `:754`, `:760`, `:800`),
    export symmetrisation (`:580-600`), float32 round trip (`surface/FVmethod.py:152,236,273`) and the
    edge-length cap (`FVmethod.py:348-349`);
  * TIN elevations by bilinear `interpn` of the DEM (`surface/elevationTIN.py:210-243`),
    ghost elevations / parentIDs (`elevationTIN.py:22-208`, `model.py:127-132`);
  * masks borders/borders2/insideIDs (`simulation/buildFlux.py:132-152`);
  * the TIN->grid inverse-distance table of `bl_mcmc.interpolateArray` (`bl_mcmc.py:167-215`).

What is NOT reproducible: the interior node positions.  The reference gets them from Shewchuk's Triangle
(`triangle.triangulate(..., 'Dqa<area>')`, `raster2TIN.py:224-225`), which is absent here and unpinned.
We place interior nodes on a hexagonal lattice with the same target cell area and triangulate with
Qhull (`scipy.spatial.Delaunay`).  Parity of the hot path is defined GIVEN these arrays.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
from scipy.interpolate import interpn
from scipy.spatial import Delaunay, cKDTree

KP = 20  # MAX_NEIGHBOURS, libUtils/sfd.c:14


@dataclass
class RecGrid:
    """Subset of `raster2TIN` attributes the hot path reads (raster2TIN.py:86-109)."""
    regX: np.ndarray
    regY: np.ndarray
    regZ: np.ndarray       # (rnx, rny), order 'F' reshape of the DEM z column (raster2TIN.py:169)
    resEdges: float
    areaDel: float
    nx: int
    ny: int
    rnx: int
    rny: int
    boundsPt: int
    edgesPt: int


@dataclass
class FVmesh:
    """`FVmethod` fields consumed downstream (surface/FVmethod.py:41-54; buildMesh.py:87-100)."""
    node_coords: np.ndarray      # (N, 2)
    neighbours: np.ndarray       # (N, 20) int32, -2 padded
    edge_length: np.ndarray      # (N, 20) Delaunay edge lengths
    vor_edges: np.ndarray        # (N, 20) Voronoi edge lengths
    control_volumes: np.ndarray  # (N,)  Voronoi areas (0 on border cells)


@dataclass
class MeshArrays:
    recGrid: RecGrid
    FVmesh: FVmesh
    elevation: np.ndarray        # initial TIN elevation, ghosts assigned
    parentIDs: np.ndarray        # (boundsPt,) int32
    borders: np.ndarray          # (N,) int32  1 on insideIDs
    borders2: np.ndarray         # (N,) int32  1 strictly inside the DEM perimeter
    insideIDs: np.ndarray
    btype: str
    # TIN -> regular grid table (bl_mcmc.interpolateArray)
    grid_nx: int = 0
    grid_ny: int = 0
    grid_idx: Optional[np.ndarray] = None    # (G, 3) int32
    grid_w: Optional[np.ndarray] = None      # (G, 3) 1/distance (inf on exact hits)
    grid_exact: Optional[np.ndarray] = None  # (G,) int32 1 where distance[:,0] == 0
    hillslope_min_edge2: float = 0.0         # min(edge_length[>0]**2) for diffLinear.dt_stability
    extra: dict = field(default_factory=dict)

    @property
    def N(self):
        return self.FVmesh.node_coords.shape[0]

    @property
    def boundsPt(self):
        return self.recGrid.boundsPt


# ---------------------------------------------------------------------------------------------
# node placement
# ---------------------------------------------------------------------------------------------
def _raster_edges(dem_xyz, Afactor):
    """raster2TIN._raster_edges (raster2TIN.py:110-169), resRecFactor == 1 as the reference calls it."""
    rectX, rectY, rectZ = dem_xyz[:, 0], dem_xyz[:, 1], dem_xyz[:, 2]
    resDEM = rectX[1] - rectX[0]
    minX, maxX, minY, maxY = rectX.min(), rectX.max(), rectY.min(), rectY.max()
    resEdges = max(resDEM * 1, resDEM)
    areaDel = max(resEdges ** 2, (resEdges ** 2) * Afactor)
    nx = int(round((maxX - minX) / resEdges + 1))
    ny = int(round((maxY - minY) / resEdges + 1))
    e_x = np.linspace(minX, maxX, nx)
    south = np.column_stack((e_x, np.full(nx, minY)))
    north = np.column_stack((e_x, np.full(nx, maxY)))
    e_y = np.linspace(minY + resEdges, maxY - resEdges, ny - 2)
    east = np.column_stack((np.full(ny - 2, minX), e_y))
    west = np.column_stack((np.full(ny - 2, maxX), e_y))
    edges = np.vstack((south, east, north, west))
    rnx = int(round((maxX - minX) / resDEM + 1))
    rny = int(round((maxY - minY) / resDEM + 1))
    regX = np.linspace(minX, maxX, rnx)
    regY = np.linspace(minY, maxY, rny)
    regZ = np.reshape(rectZ, (rnx, rny), order="F")
    # ghost ring, raster2TIN._TIN_ghosts_bounds (raster2TIN.py:171-207)
    b_x = np.linspace(minX - resEdges, maxX + resEdges, nx + 2)
    bsouth = np.column_stack((b_x, np.full(nx + 2, minY - resEdges)))
    bnorth = np.column_stack((b_x, np.full(nx + 2, maxY + resEdges)))
    b_y = np.linspace(minY, maxY, ny)
    beast = np.column_stack((np.full(ny, minX - resEdges), b_y))
    bwest = np.column_stack((np.full(ny, maxX + resEdges), b_y))
    bounds = np.vstack((bsouth, beast, bnorth, bwest))
    rg = RecGrid(regX=regX, regY=regY, regZ=regZ, resEdges=float(resEdges), areaDel=float(areaDel),
                 nx=nx, ny=ny, rnx=rnx, rny=rny, boundsPt=len(bounds), edgesPt=len(edges))
    return rg, bounds, edges


def _hex_interior(rg: RecGrid, spacing_factor=1.18):
    """Interior nodes on a hexagonal lattice standing in for Triangle's Steiner points.

    Triangle 'q a<A>' yields triangles whose mean area is ~0.6 A; an equilateral lattice of edge
    a = 1.18 sqrt(A) has that cell area (sqrt(3)/4 a^2 = 0.60 A).
    """
    a = spacing_factor * np.sqrt(rg.areaDel)
    minX, maxX, minY, maxY = rg.regX[0], rg.regX[-1], rg.regY[0], rg.regY[-1]
    margin = 0.6 * min(a, rg.resEdges * 1.5)
    h = a * np.sqrt(3.0) / 2.0
    cx, cy = 0.5 * (minX + maxX), 0.5 * (minY + maxY)
    kx = int(np.ceil((maxX - minX) / a)) + 2
    ky = int(np.ceil((maxY - minY) / h)) + 2
    pts = []
    for r in range(-ky, ky + 1):
        y = cy + r * h
        if not (minY + margin <= y <= maxY - margin):
            continue
        off = 0.0 if (r % 2 == 0) else 0.5 * a
        xs = cx + off + a * np.arange(-kx, kx + 1)
        xs = xs[(xs >= minX + margin) & (xs <= maxX - margin)]
        pts.append(np.column_stack((xs, np.full(len(xs), y))))
    return np.vstack(pts)


def _grid_interior(rg: RecGrid):
    """Interior nodes = strictly interior DEM points (regular lattice)."""
    X, Y = np.meshgrid(rg.regX[1:-1], rg.regY[1:-1], indexing="xy")
    return np.column_stack((X.ravel(), Y.ravel()))


# ---------------------------------------------------------------------------------------------
# Delaunay / Voronoi duality
# ---------------------------------------------------------------------------------------------
def build_fv(points: np.ndarray, res_cap: float, simplices: Optional[np.ndarray] = None,
             tri_neighbors: Optional[np.ndarray] = None) -> FVmesh:
    """Finite-volume arrays of a point set (FVclass.f90:605-856, :554-603; FVmethod.py:283-356)."""
    N = len(points)
    if simplices is None:
        tri = Delaunay(points)
        simplices, tri_neighbors = tri.simplices, tri.neighbors
    S = simplices.astype(np.int64)
    P = points
    # circumcentres
    A, B, C = P[S[:, 0]], P[S[:, 1]], P[S[:, 2]]
    d = 2.0 * (A[:, 0] * (B[:, 1] - C[:, 1]) + B[:, 0] * (C[:, 1] - A[:, 1]) + C[:, 0] * (A[:, 1] - B[:, 1]))
    a2, b2, c2 = (A ** 2).sum(1), (B ** 2).sum(1), (C ** 2).sum(1)
    ux = (a2 * (B[:, 1] - C[:, 1]) + b2 * (C[:, 1] - A[:, 1]) + c2 * (A[:, 1] - B[:, 1])) / d
    uy = (a2 * (C[:, 0] - B[:, 0]) + b2 * (A[:, 0] - C[:, 0]) + c2 * (B[:, 0] - A[:, 0])) / d
    cc = np.column_stack((ux, uy))
    # every (triangle, opposite-vertex k) gives the edge (s[k+1], s[k+2]) and the triangle across it
    T = len(S)
    e0 = np.concatenate([S[:, 1], S[:, 2], S[:, 0]])
    e1 = np.concatenate([S[:, 2], S[:, 0], S[:, 1]])
    t_self = np.tile(np.arange(T), 3)
    t_opp = np.concatenate([tri_neighbors[:, 0], tri_neighbors[:, 1], tri_neighbors[:, 2]])
    lo, hi = np.minimum(e0, e1), np.maximum(e0, e1)
    # one record per undirected edge: keep the copy seen from the lower-numbered triangle (or hull)
    keep = (t_opp < 0) | (t_self < t_opp)
    lo, hi, t_self, t_opp = lo[keep], hi[keep], t_self[keep], t_opp[keep]
    order = np.lexsort((hi, lo))           # edge list sorted by (endpoint1, endpoint2)
    lo, hi, t_self, t_opp = lo[order], hi[order], t_self[order], t_opp[order]
    hull = t_opp < 0
    vlen = np.zeros(len(lo))
    vlen[~hull] = np.hypot(*(cc[t_self[~hull]] - cc[t_opp[~hull]]).T)
    dlen = np.hypot(*(P[lo] - P[hi]).T)
    border = np.zeros(N, dtype=bool)       # unbounded Voronoi cell  <=> touches a hull edge
    border[lo[hull]] = True
    border[hi[hull]] = True
    # cell areas: fan of triangles (node, cc_t, cc_t') over the incident Delaunay edges
    area = np.zeros(N)
    ok = ~hull
    for end in (lo, hi):
        v1 = cc[t_self[ok]] - P[end[ok]]
        v2 = cc[t_opp[ok]] - P[end[ok]]
        np.add.at(area, end[ok], 0.5 * np.abs(v1[:, 0] * v2[:, 1] - v2[:, 0] * v1[:, 1]))
    area[border] = 0.0
    # neighbour lists: partners where node is endpoint 1 (ascending), then where it is endpoint 2
    ngb = np.full((N, KP), -2, dtype=np.int32)
    elen = np.zeros((N, KP))
    vor = np.zeros((N, KP))
    cnt = np.zeros(N, dtype=np.int64)
    for src, dst in ((lo, hi), (hi, lo)):
        # edges are sorted by (lo, hi); for the (hi, lo) pass re-sort by (hi, lo)
        o = np.lexsort((dst, src))
        s, t, vl, dl = src[o], dst[o], vlen[o], dlen[o]
        first = np.r_[0, np.flatnonzero(np.diff(s)) + 1]
        run = np.arange(len(s)) - np.repeat(first, np.diff(np.r_[first, len(s)]))
        slot = cnt[s] + run
        if slot.max() >= KP:
            raise ValueError("a node has more than %d neighbours" % KP)
        ngb[s, slot] = t
        # a border cell exports 0 for its own Voronoi edges and distances (FVclass.f90:760,800)
        own = ~border[s]
        elen[s[own], slot[own]] = dl[own]
        vor[s[own], slot[own]] = vl[own]
        np.add.at(cnt, s, 1)
    # export symmetrisation (FVclass.f90:580-600): fill zero distances from the other side and copy
    # the neighbour's Voronoi length onto border cells
    deg = (ngb >= 0).sum(1)
    for i in np.flatnonzero(border):
        for o in range(deg[i]):
            j = ngb[i, o]
            o2 = np.flatnonzero(ngb[j, :deg[j]] == i)[0]
            if elen[i, o] == 0.0:
                elen[i, o] = elen[j, o2]
            vor[i, o] = vor[j, o2]
    # float32 round trip of the MPI gather (FVmethod.py:152,236,273) and the edge cap (:348-349)
    area = area.astype(np.float32).astype(np.float64)
    elen = elen.astype(np.float32).astype(np.float64)
    vor = vor.astype(np.float32).astype(np.float64)
    elen[elen > 2.0 * np.sqrt(2.0) * res_cap] = np.sqrt(2.0) * res_cap
    return FVmesh(node_coords=np.ascontiguousarray(points, dtype=np.float64), neighbours=ngb,
                  edge_length=elen, vor_edges=vor, control_volumes=area)


# ---------------------------------------------------------------------------------------------
# boundary elevations (elevationTIN.py:22-208)
# ---------------------------------------------------------------------------------------------
def _boundary_elevation(elevation, neighbours, edge_length, boundPts, btype):
    thetype = 1 if btype in ("slope", "outlet", "wall1") else 0
    elevation = elevation.copy()
    elevation[:boundPts] = 1.0e7
    missed = []
    parentID = np.zeros(boundPts, dtype=np.int64)
    for i in range(boundPts):
        ngbhs = neighbours[i, :]
        ids = np.where(ngbhs >= boundPts)[0]
        if len(ids) == 0:
            missed.append(i)
            continue
        lsel = edge_length[i, ids]
        picked = int(np.argmin(lsel))
        id1, ln1 = int(ngbhs[ids[picked]]), lsel[picked]
        parentID[i] = id1
        if thetype == 0:
            elevation[i] = elevation[id1]
        else:
            ngbhs2 = neighbours[id1, :]
            ids2 = np.where(ngbhs2 >= boundPts)[0]
            lsel2 = edge_length[id1, ids2]
            if len(lsel2) > 0:
                p2 = int(np.argmin(lsel2))
                id2, ln2 = int(ngbhs2[ids2[p2]]), lsel2[p2]
                elevation[i] = (elevation[id1] - elevation[id2]) * (ln2 + ln1) / ln2 + elevation[id2]
            else:
                missed.append(i)
    for i in missed:
        ngbhs = neighbours[i, :]
        ids = np.where((elevation[ngbhs] < 9.0e6) & (ngbhs >= 0))[0]
        if len(ids) == 0:
            raise ValueError("Error while getting boundary elevation for point '%d'." % i)
        picked = int(np.argmin(edge_length[i, ids]))
        elevation[i] = elevation[ngbhs[ids[picked]]]
        parentID[i] = ngbhs[ids[picked]]
    if thetype == 1:
        elevation[:boundPts] -= 0.5
    if btype == "wall":
        elevation[:boundPts] = 1.0e7
    return elevation, parentID


def _masks(fv: FVmesh, rg: RecGrid):
    """buildFlux.py:132-152 (open-box tests; margins of 1 m make the edge convention irrelevant)."""
    xy, area = fv.node_coords, fv.control_volumes
    x0, x1, y0, y1 = rg.regX.min(), rg.regX.max(), rg.regY.min(), rg.regY.max()
    in1 = (xy[:, 0] > x0 - 1.0) & (xy[:, 0] < x1 + 1.0) & (xy[:, 1] > y0 - 1.0) & (xy[:, 1] < y1 + 1.0)
    inside = np.flatnonzero((area > 0.0) & in1)
    borders = np.zeros(len(area), dtype=np.int32)
    borders[inside] = 1
    in2 = (xy[:, 0] > x0 + 1.0) & (xy[:, 0] < x1 - 1.0) & (xy[:, 1] > y0 + 1.0) & (xy[:, 1] < y1 - 1.0)
    borders2 = in2.astype(np.int32)
    return borders, borders2, inside


def interp_table(xy: np.ndarray):
    """Mesh-constant part of `bl_mcmc.interpolateArray` (bl_mcmc.py:183-199)."""
    x, y = xy[:, 0], xy[:, 1]
    dx = x[1] - x[0]
    nx = int((x.max() - x.min()) / dx + 1)
    ny = int((y.max() - y.min()) / dx + 1)
    xi = np.linspace(x.min(), x.max(), nx)
    yi = np.linspace(y.min(), y.max(), ny)
    xi, yi = np.meshgrid(xi, yi)
    xyi = np.dstack([xi.flatten(), yi.flatten()])[0]
    tree = cKDTree(np.column_stack((x, y)))
    distances, indices = tree.query(xyi, k=3)
    with np.errstate(divide="ignore"):
        w = 1.0 / distances
    exact = (distances[:, 0] == 0).astype(np.int32)
    return nx, ny, indices.astype(np.int32), w, exact


def _finish(points, rg, btype, Afactor, simplices=None, tri_neighbors=None, elev_interior=None):
    fv = build_fv(points, rg.resEdges * Afactor, simplices, tri_neighbors)
    N = len(points)
    elev = np.full(N, -1.0e6)
    if elev_interior is None:
        elev[rg.boundsPt:] = interpn((rg.regX, rg.regY), rg.regZ, points[rg.boundsPt:, :2], method="linear")
    else:
        elev[rg.boundsPt:] = elev_interior
    elev, parent = _boundary_elevation(elev, fv.neighbours, fv.edge_length, rg.boundsPt, btype)
    # model.py:127-132: a parent that is itself a ghost is re-mapped to the nearest non-ghost node
    re_id = np.where(parent < len(parent))[0]
    if len(re_id) > 0:
        tree = cKDTree(points[len(parent):, :2])
        _, ind = tree.query(points[re_id, :2], k=1)
        parent[re_id] = ind + len(parent)
    borders, borders2, inside = _masks(fv, rg)
    nx, ny, idx, w, exact = interp_table(points)
    el = fv.edge_length
    m = MeshArrays(recGrid=rg, FVmesh=fv, elevation=elev, parentIDs=parent.astype(np.int32),
                   borders=borders, borders2=borders2, insideIDs=inside, btype=btype,
                   grid_nx=nx, grid_ny=ny, grid_idx=idx, grid_w=w, grid_exact=exact,
                   hillslope_min_edge2=float(np.amin(el[el > 0.0] ** 2)))
    return m


def build_from_dem(dem_xyz: np.ndarray, btype: str = "slope", Afactor: int = 1,
                   interior: str = "hex", colour_order: bool = True) -> MeshArrays:
    """DEM (x y z rows) -> TIN + FV arrays.  `interior` = 'hex' (TIN-like) or 'grid' (DEM points)."""
    rg, bounds, edges = _raster_edges(np.asarray(dem_xyz, dtype=np.float64), Afactor)
    inner = _hex_interior(rg) if interior == "hex" else _grid_interior(rg)
    pts = np.vstack((bounds, edges, inner))
    if colour_order:
        pts = pts[_colour_order(pts, rg.boundsPt)]
    return _finish(pts, rg, btype, Afactor)


def _dsatur(N: int, nb: int, indptr, indices) -> np.ndarray:
    """DSATUR colouring of the non-ghost nodes (ids >= nb): always colour next the node that sees the most distinct
    colours among its neighbours (ties: higher degree, then lower id), with the smallest colour not yet seen."""
    import heapq
    colour = np.full(N, -1, dtype=np.int64)
    adj = [indices[indptr[v]:indptr[v + 1]] for v in range(N)]
    adj = [a[a >= nb] for a in adj]
    sat = [set() for _ in range(N)]
    todo = np.zeros(N, dtype=bool)
    todo[nb:] = True
    heap = [(0, -len(adj[v]), v) for v in range(nb, N)]
    heapq.heapify(heap)
    left = N - nb
    while left:
        s, _, v = heapq.heappop(heap)
        if not todo[v] or -s != len(sat[v]):
            continue                                    # stale heap entry
        c = 0
        while c in sat[v]:
            c += 1
        colour[v] = c
        todo[v] = False
        left -= 1
        for w in adj[v]:
            if todo[w] and c not in sat[w]:
                sat[w].add(c)
                heapq.heappush(heap, (-len(sat[w]), -len(adj[w]), w))
    return colour


def _colour_order(points: np.ndarray, nb: int) -> np.ndarray:
    """Node numbering that keeps fillPD's Gauss-Seidel sweep shallow.

    `PDalgo.fillPD` sweeps nodes in ascending id and reads the already-updated values of lower-numbered
    neighbours (libUtils/PDalgo.f90:49-81).  Its result therefore depends on the numbering, which in the
    reference is whatever order Triangle emitted.  We number the non-ghost nodes colour by colour of a
    DSATUR graph colouring: no two nodes of a colour are adjacent, so one sweep is `ncolours` parallel
    passes on the GPU and still equals the sequential sweep bit for bit (csrc/bl_fast.cuh: fill_bitmap).
    DSATUR finds the three colours of the triangular interior lattice (plus a few dozen boundary-ring nodes in a
    fourth); first-fit greedy needed five, which cost the sweep both more passes and more sweeps to converge
    (crater: 185 -> 124 passes per fill).
    Ghost nodes must stay first (ids < boundsPt, raster2TIN.py:220-226).
    """
    tri = Delaunay(points)
    indptr, indices = tri.vertex_neighbor_vertices
    N = len(points)
    colour = _dsatur(N, nb, indptr, indices)
    inner = np.arange(nb, N)
    order = np.concatenate((np.arange(nb), inner[np.lexsort((inner, colour[inner]))]))
    # Gauss-Seidel level of each node under that numbering (1 + max level of its lower-numbered non-ghost
    # neighbours); numbering by (level, id) keeps the levels unchanged and makes every level a contiguous id
    # range, which the device sweep exploits.
    rank = np.empty(N, dtype=np.int64)
    rank[order] = np.arange(N)
    level = np.zeros(N, dtype=np.int64)
    for v in order[nb:]:
        nbr = indices[indptr[v]:indptr[v + 1]]
        lower = nbr[(rank[nbr] >= nb) & (rank[nbr] < rank[v])]
        level[v] = 1 + (level[lower].max() if len(lower) else 0)
    inner2 = order[nb:]
    return np.concatenate((np.arange(nb), inner2[np.lexsort((rank[inner2], level[inner2]))]))


def regular_interior_ids(nx: int, ny: int, order: str = "row") -> np.ndarray:
    """(ny, nx) ids (0-based, before adding boundsPt) of the interior nodes of a regular lattice.

    "row": row-major, x fastest.  "colour": by 2x2 parity class c = (j&1)*2 + (i&1), row-major inside a class -
    no two D8 neighbours share a class, so fillPD's ascending-id Gauss-Seidel sweep (PDalgo.f90:49-81) has 4
    dependency levels whatever the lattice size (what the lattice engine emulates exactly)."""
    if order == "row":
        return np.arange(nx * ny, dtype=np.int64).reshape(ny, nx)
    if order != "colour":
        raise ValueError("order must be 'row' or 'colour'")
    if nx % 2 or ny % 2:
        raise ValueError("colour order needs even nx, ny")
    j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    c = (j & 1) * 2 + (i & 1)
    return c * ((nx // 2) * (ny // 2)) + (j >> 1) * (nx // 2) + (i >> 1)


def regular_ext_ids(nx: int, ny: int, order: str = "row") -> np.ndarray:
    """(ny+2, nx+2) mesh id of every cell of the extended lattice (ghost ring in raster2TIN order first)."""
    nb = 2 * (nx + 2) + 2 * ny
    ext = np.full((ny + 2, nx + 2), -1, dtype=np.int64)
    ext[0, :] = np.arange(nx + 2)
    ext[1:-1, 0] = nx + 2 + np.arange(ny)
    ext[-1, :] = nx + 2 + ny + np.arange(nx + 2)
    ext[1:-1, -1] = 2 * (nx + 2) + ny + np.arange(ny)
    ext[1:-1, 1:-1] = nb + regular_interior_ids(nx, ny, order)
    return ext


def build_regular(nx: int, ny: int, dx: float, z: np.ndarray, btype: str = "fixed", order: str = "row") -> MeshArrays:
    """Synthetic regular-grid landscape (BASELINE config 5): D8 stencil as the neighbour list.

    Documented deviation: the reference has no regular-grid mode.  Node order: ghost
    ring first (same order as raster2TIN), then the nx*ny grid in `order` (see regular_interior_ids).  FV
    arrays are analytic: area dx^2, Voronoi edge dx for the 4 axis neighbours and 0 for the diagonals.
    """
    z = np.asarray(z, dtype=np.float64).reshape(ny, nx)
    xs = dx * np.arange(nx)
    ys = dx * np.arange(ny)
    b_x = dx * np.arange(-1, nx + 1)
    bsouth = np.column_stack((b_x, np.full(nx + 2, -dx)))
    bnorth = np.column_stack((b_x, np.full(nx + 2, ny * dx)))
    beast = np.column_stack((np.full(ny, -dx), ys))
    bwest = np.column_stack((np.full(ny, nx * dx), ys))
    bounds = np.vstack((bsouth, beast, bnorth, bwest))
    nb = len(bounds)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    # id lookup on the (ny+2, nx+2) extended lattice
    ext = regular_ext_ids(nx, ny, order)
    N = nb + nx * ny
    pts = np.zeros((N, 2))
    pts[:nb] = bounds
    pts[ext[1:-1, 1:-1].ravel(), 0] = X.ravel()
    pts[ext[1:-1, 1:-1].ravel(), 1] = Y.ravel()
    ngb = np.full((N, KP), -2, dtype=np.int32)
    elen = np.zeros((N, KP)); vor = np.zeros((N, KP))
    offs = [(0, 1, dx, dx), (1, 1, dx * np.sqrt(2.0), 0.0), (1, 0, dx, dx), (1, -1, dx * np.sqrt(2.0), 0.0),
            (0, -1, dx, dx), (-1, -1, dx * np.sqrt(2.0), 0.0), (-1, 0, dx, dx), (-1, 1, dx * np.sqrt(2.0), 0.0)]
    J, I = np.meshgrid(np.arange(ny + 2), np.arange(nx + 2), indexing="ij")
    ids = ext
    cnt = np.zeros(N, dtype=np.int64)
    ghost = np.zeros(N, dtype=bool); ghost[:nb] = True
    for dj, di, dl, vl in offs:
        jj, ii = J + dj, I + di
        ok = (jj >= 0) & (jj < ny + 2) & (ii >= 0) & (ii < nx + 2)
        src = ids[J[ok], I[ok]]; dst = ids[jj[ok], ii[ok]]
        s = cnt[src]
        ngb[src, s] = dst
        gg = ghost[src] & ghost[dst]
        elen[src, s] = np.where(gg, 0.0, dl)
        vor[src, s] = np.where(gg, 0.0, vl)
        cnt[src] += 1
    area = np.full(N, dx * dx); area[:nb] = 0.0
    area = area.astype(np.float32).astype(np.float64)
    elen = elen.astype(np.float32).astype(np.float64)
    vor = vor.astype(np.float32).astype(np.float64)
    fv = FVmesh(node_coords=pts, neighbours=ngb, edge_length=elen, vor_edges=vor, control_volumes=area)
    rg = RecGrid(regX=xs, regY=ys, regZ=z.T.copy(), resEdges=float(dx), areaDel=float(dx * dx), nx=nx, ny=ny,
                 rnx=nx, rny=ny, boundsPt=nb, edgesPt=2 * nx + 2 * (ny - 2))
    elev = np.full(N, -1.0e6)
    elev[ext[1:-1, 1:-1].ravel()] = z.ravel()
    elev, parent = _boundary_elevation(elev, ngb, elen, nb, btype)
    borders, borders2, inside = _masks(fv, rg)
    # grid table: every (ny+2, nx+2) output cell is exactly a node
    idx = np.zeros(((ny + 2) * (nx + 2), 3), dtype=np.int32)
    idx[:, 0] = ext.ravel()
    w = np.zeros((len(idx), 3)); w[:, 0] = np.inf
    exact = np.ones(len(idx), dtype=np.int32)
    return MeshArrays(recGrid=rg, FVmesh=fv, elevation=elev, parentIDs=parent.astype(np.int32), borders=borders,
                      borders2=borders2, insideIDs=inside, btype=btype, grid_nx=nx + 2, grid_ny=ny + 2,
                      grid_idx=idx, grid_w=w, grid_exact=exact,
                      hillslope_min_edge2=float(np.amin(elen[elen > 0.0] ** 2)),
                      extra=dict(regular=dict(nx=nx, ny=ny, dx=float(dx), order=order, ext=ext)))
