"""Host-side forcing set-up for the hot path (`pyBadlands/forcing/forceSim.py` subset, built once per config).

* `load_tecto_map`  - `forceSim.load_Tecto_map` (`forceSim.py:451-494`): cumulative vertical displacement map on
  the regular grid -> rate (m/a) on the TIN nodes by bilinear `interpn`, divided by the event duration.
* `disp_border`     - `forceSim.disp_border` (`forceSim.py:400-449`): ghost nodes take the rate of their nearest
  non-ghost neighbour.
The device engine applies `elevation += disp * dt` every step (`simulation/buildFlux.py:312-313`).
"""
from __future__ import annotations

import numpy as np
from scipy import interpolate


def disp_border(disp, neighbours, edge_length, boundPts):
    disp = np.array(disp, dtype=np.float64)
    disp[:boundPts] = 1.0e7
    missed = []
    for i in range(boundPts):
        ngbhs = neighbours[i, :]
        ids = np.where(ngbhs >= boundPts)[0]
        if len(ids) == 1:
            disp[i] = disp[ngbhs[ids[0]]]
        elif len(ids) > 1:
            picked = int(np.argmin(edge_length[i, ids]))
            disp[i] = disp[ngbhs[ids[picked]]]
        else:
            missed.append(i)
    for i in missed:
        ngbhs = neighbours[i, :]
        ids = np.where((disp[ngbhs] < 9.0e6) & (ngbhs >= 0))[0]
        if len(ids) == 0:
            raise ValueError("Error while getting boundary displacement for point '%d'." % i)
        picked = int(np.argmin(edge_length[i, ids]))
        disp[i] = disp[ngbhs[ids[picked]]]
    return disp


def load_tecto_map(mesh, disp_map, t_start, t_end):
    """Displacement rate [N] (m/a) for one tectonic event covering [t_start, t_end] (model.py:285-292)."""
    rg, fv = mesh.recGrid, mesh.FVmesh
    dt = float(t_end) - float(t_start)
    if dt <= 0:
        raise ValueError("Problem computing the displacements rate for the event.")
    rect = np.reshape(np.asarray(disp_map, dtype=np.float64), (len(rg.regX), len(rg.regY)), order="F")
    nb = mesh.boundsPt
    ldisp = np.full(mesh.N, -1.0e6)
    ldisp[nb:] = interpolate.interpn((rg.regX, rg.regY), rect, fv.node_coords[nb:, :2], method="linear") / dt
    return disp_border(ldisp, fv.neighbours, fv.edge_length, nb)
