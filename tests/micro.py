"""Hand-derivable micro-fixtures for the routines the reference keeps in Fortran (no gfortran here: SURVEY.md F1).

A *chain mesh* is a 1-D landscape: ghost node 0 at the left end, ghost node 1 at the right end (boundsPt = 2), interior
nodes in between, each adjacent to its left and right neighbour only (neighbour order: left, right).  Cells are
100 m x 100 m (area 1e4 m^2, Voronoi and Delaunay edge lengths 100 m), so every quantity of a time step can be
followed by hand.  `order[p]` gives the node id at position p = 1 .. L-1 (the sweep order of fillPD is the id order).

The expected values and their derivations are in tests/golden/README.md; both the CPU oracle and the CUDA path (per-stage
entry points and bl_step) are checked against them (tests/test_oracle.py, tests/test_gpu_stages.py).
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DX = 100.0


def chain_mesh(z_by_pos, order=None, btype="fixed"):
    """z_by_pos: elevations at positions 0 .. L (positions 0 and L are the ghost nodes).  Returns MeshArrays."""
    from bayeslands_b200 import mesh as M
    z_by_pos = np.asarray(z_by_pos, dtype=np.float64)
    L = len(z_by_pos) - 1
    N = L + 1
    if order is None:
        order = list(range(2, N))                  # ids ascend from left to right
    ids = np.empty(L + 1, dtype=np.int64)          # id at position p
    ids[0], ids[L] = 0, 1
    ids[1:L] = order
    assert sorted(ids.tolist()) == list(range(N))
    pos = np.empty(N, dtype=np.int64)
    pos[ids] = np.arange(L + 1)
    ngb = np.full((N, M.KP), -2, dtype=np.int32)
    edge = np.zeros((N, M.KP)); vor = np.zeros((N, M.KP))
    for k in range(N):
        p = pos[k]
        q = 0
        for pp in (p - 1, p + 1):
            if 0 <= pp <= L:
                ngb[k, q] = ids[pp]; edge[k, q] = DX; vor[k, q] = DX
                q += 1
    area = np.where(np.arange(N) < 2, 0.0, DX * DX)
    xy = np.column_stack((pos * DX, np.zeros(N)))
    fv = M.FVmesh(node_coords=xy, neighbours=ngb, edge_length=edge, vor_edges=vor, control_volumes=area)
    regX = np.arange(1, L) * DX
    rg = M.RecGrid(regX=regX, regY=np.array([0.0]), regZ=np.zeros((L - 1, 1)), resEdges=DX, areaDel=DX * DX, nx=L - 1, ny=1,
                   rnx=L - 1, rny=1, boundsPt=2, edgesPt=0)
    borders = (area > 0).astype(np.int32)
    borders2 = borders.copy()
    borders2[ids[1]] = 0; borders2[ids[L - 1]] = 0           # the DEM perimeter is excluded from hillslope diffusion
    parent = np.array([ids[1], ids[L - 1]], dtype=np.int32)
    elev = z_by_pos[pos]
    m = M.MeshArrays(recGrid=rg, FVmesh=fv, elevation=elev.copy(), parentIDs=parent, borders=borders, borders2=borders2,
                     insideIDs=np.where(borders > 0)[0], btype=btype, grid_nx=1, grid_ny=1,
                     grid_idx=np.array([[2, 2, 2]], dtype=np.int32), grid_w=np.array([[1.0, 1.0, 1.0]]),
                     grid_exact=np.array([1], dtype=np.int32), hillslope_min_edge2=DX * DX)
    m.extra["pos"] = pos
    m.extra["ids"] = ids
    return m


def inputs(fillmax=15.0, dep=1, caerial=0.0, cmarine=0.0, tend=5000.0):
    """Input of Examples/crater_fast with the given knobs (no hillslope creep by default, so a step is fluvial only)."""
    from bayeslands_b200 import config
    xml = str(np.load(os.path.join(ROOT, "data", "crater_fast.npz"))["xml"])
    xml = xml.replace("<fillmax>15.</fillmax>", "<fillmax>%r</fillmax>" % float(fillmax))
    xml = xml.replace("<caerial>2.e-2</caerial>", "<caerial>%r</caerial>" % float(caerial))
    xml = xml.replace("<cmarine>5.e-2</cmarine>", "<cmarine>%r</cmarine>" % float(cmarine))
    xml = xml.replace("<end>15000.</end>", "<end>%r</end>" % float(tend)).replace("<rend>15000.</rend>", "<rend>%r</rend>" % float(tend))
    inp = config.parse_xml(xml)
    assert inp.fillmax == float(fillmax) and inp.CDa == float(caerial)
    if not dep:
        inp.depo = 0
    return inp


# ---------------------------------------------------------------------------------------------------
# F1  capped lake: the sweep order changes the answer and the loop stops one sweep early (PDalgo.f90:58-79)
# positions:      0(g0)  1     2     3     4    5(g1)
F1_Z = [0.0, 20.0, 0.0, -10.0, 30.0, 50.0]
EPS, TH = 0.01, 15.0


def f1_expected(order):
    """fillH by position 1..4, derived by hand in tests/golden/README.md (F1)."""
    if order == "ascending":       # ids ascend left to right: one sweep, every write is a dry-out or a capped write
        return [20.0, 0.0 + TH, -10.0 + TH, 30.0]
    if order == "descending":      # ids ascend right to left: sweep 1 leaves 20, 5 + eps, 5, 45; sweep 2 dries position 4
        return [20.0, 5.0 + EPS, -10.0 + TH, 30.0]
    raise ValueError(order)


# ---------------------------------------------------------------------------------------------------
# F2  a staircase with two lakes: stack order, basins, drainage chain, discharge, stream power with a two-pit cascade
#     (the upper lake overflows into the lower one, which overflows to the boundary) and the proportional land-pit fill.
# positions:   0(g0)  1    2    3     4     5     6     7     8    9(g1)      ids: g0 = 0, g1 = 1, position p -> id p + 1
F2_Z = [0.0, 10.0, 4.0, 12.0, 20.0, 14.0, 16.0, 25.0, 40.0, 60.0]
F2_RAIN, F2_EROD, F2_SEA = 1.0, 2.0e-3, -1000.0
AREA = DX * DX


def f2_derive(dt):
    """Every array of one time step of F2, written out node by node from the reference's formulas
    (tests/golden/README.md, F2).  ids: g0 = 0, g1 = 1, position p -> p + 1."""
    import math
    z = {0: 0.0, 1: 60.0}
    for p in range(1, 9):
        z[p + 1] = F2_Z[p]
    # fillPD, ascending ids, sweep 1 (PDalgo.f90:49-81); sweep 2 changes nothing
    W = dict(z)
    W[2] = 10.0                         # z >= W(g0) + eps: dries out
    W[3] = W[2] + EPS                   # lake 1: spills over position 1
    W[4] = 12.0; W[5] = 20.0
    W[6] = W[5] + EPS                   # lake 2: spills over position 4
    W[7] = W[6] + EPS
    W[8] = 25.0; W[9] = 40.0
    rcv1 = {0: 0, 1: 9, 2: 0, 3: 3, 4: 3, 5: 4, 6: 6, 7: 6, 8: 7, 9: 8}     # lowest neighbour on z (sfd.c:113-153)
    rcv = {0: 0, 1: 9, 2: 0, 3: 2, 4: 3, 5: 4, 6: 5, 7: 6, 8: 7, 9: 8}      # lowest neighbour on W (sfd.c:55-111)
    stack = [0, 2, 3, 4, 5, 6, 7, 8, 9, 1]       # pre-order DFS from base 0, donors ascending (FLOWstack.f90:24-40)
    stack1 = [0, 2, 3, 4, 5, 6, 7, 8, 9, 1]      # bases 0, 3, 6 in ascending order
    pitID = {k: -1 for k in range(10)}
    pitID[0], pitID[3], pitID[6], pitID[7] = 0, 3, 6, 6                      # FLOWalgo.f90:150-163
    pitVol = {k: -1.0 for k in range(10)}
    pitVol[3] = -1.0 + (W[3] - z[3]) * AREA
    pitVol[6] = (-1.0 + (W[6] - z[6]) * AREA) + (W[7] - z[7]) * AREA         # stack1 order: 6 then 7
    pitDrain = {k: -1 for k in range(10)}
    pitDrain[6], pitDrain[3], pitDrain[0] = 3, 0, 0                          # chain 6 -> 3 -> 0 (FLOWalgo.f90:190-234)
    # discharge, reverse stack (FLOWalgo.f90:70-75): q0 = area * rain on interior nodes
    Q = {1: 0.0}
    acc = 0.0
    for k in (9, 8, 7, 6, 5, 4, 3, 2):
        Q[k] = AREA * F2_RAIN + (Q[1] if k == 9 else acc)
        acc = Q[k]
    Q[0] = 0.0 + Q[2]
    # stream power (FLOWalgo.f90:635-714, :875-892): eroding nodes 9, 8, 5, 4, 2
    def erode(d):
        r = rcv[d]
        dh = 0.95 * (z[d] - z[r])
        slp = dh / DX
        spl = -F2_EROD * 1.0 * 1.0 * math.sqrt(Q[d]) * slp
        tot = 0.0 + spl
        if -tot * dt > dh:
            spl = spl * (dh / (-tot * dt))
        return spl * dt * AREA
    ero = {k: 0.0 for k in range(10)}
    depo = {k: 0.0 for k in range(10)}
    for k in (9, 8, 5, 4, 2):
        ero[k] = 0.0 + erode(k)
    qs9 = -ero[9] + 0.0
    load6 = -ero[8] + qs9                 # enters lake 2 at node 7 (pit 6)
    qs5 = -ero[5] + 0.0
    load3 = -ero[4] + qs5                 # enters lake 1 at node 3 (pit 3)
    vol = dict(pitVol)
    # cascade (FLOWalgo.f90:967-1034), events in reverse stack order: lake 2 first
    def cascade(pit, load):
        while load > 0.0:
            have = 0.0 + depo[pit]
            if have + load <= vol[pit] or pitDrain[pit] == pit:
                depo[pit] = depo[pit] + load
                return
            frac = load / load
            depo[pit] = depo[pit] + (vol[pit] - have) * frac
            load = (load - (vol[pit] - have)) * frac
            vol[pit] = 0.0 + depo[pit]
            pit = pitDrain[pit]
    cascade(6, load6)
    cascade(3, load3)
    # land-pit fill (flowNetwork.py:750-763): thickness (W - z) scaled so that the pit's members hold depo[pit]
    th = {k: 0.0 for k in range(10)}
    t3 = (W[3] - z[3]) * 1.0
    th[3] = t3 * ((0.0 + depo[3]) / (t3 * AREA))
    t6, t7 = (W[6] - z[6]) * 1.0, (W[7] - z[7]) * 1.0
    f6 = (0.0 + depo[6]) / (t6 * AREA + t7 * AREA)
    th[6], th[7] = t6 * f6, t7 * f6
    sedflux = {k: 0.0 for k in range(10)}
    for k in range(2, 10):
        sedflux[k] = ero[k] / AREA + th[k]
    tolist = lambda dct: [dct[k] for k in range(10)]
    return dict(fillH=tolist(W), rcv1=tolist(rcv1), rcv=tolist(rcv), stack=stack, stack1=stack1, pitID=tolist(pitID),
                pitVol=tolist(pitVol), pitDrain=tolist(pitDrain), allDrain=tolist(pitDrain), Q=tolist(Q), ero=tolist(ero),
                depo=tolist(depo), sedflux=tolist(sedflux), load6=load6, load3=load3)


def f2_cfl_dt():
    """dt of the first step: floor of min dist / (K sqrt(Q)) over the flowing nodes (FLOWalgo.f90:299-330, buildFlux.py:169-177)."""
    import math
    return float(math.floor(DX / (F2_EROD * math.sqrt(8.0 * AREA * F2_RAIN))))


# ---------------------------------------------------------------------------------------------------
# F3  basindrainage with a chainDrain de-duplication hit (FLOWalgo.f90:190-234): two pits whose walks reach each other.
# Hand-made arrays (not a surface): pit 1 (fillH 9) walks 1 -> 3, node 3 lies in pit 2: drain[1] = 2, jump to 2; 2 -> 4, node 4
# lies in pit 1, not yet in the chain: drain[2] = 1, jump to 1; 1 -> 3 again, pit 2 IS in the chain now: walk through, 3 -> 5,
# 5 is a self-receiver: drain[1] = 5 (overwrites 2).  Pit 2 is then found already drained.
F3 = dict(N=8,
          rcv=[0, 3, 4, 5, 6, 5, 6, 7],
          pitID=[-1, 1, 2, 2, 1, -1, -1, -1],
          pitVol=[-1.0, 5.0, 3.0, -1.0, -1.0, -1.0, -1.0, -1.0],
          fillH=[0.0, 9.0, 8.0, 7.5, 7.0, 1.0, 2.0, 3.0],
          pitDrain=[-1, 5, 1, -1, -1, -1, -1, -1])


# ---------------------------------------------------------------------------------------------------
# F4  river sediment reaching the sea: marine_distribution (PDalgo.f90:136-269), both ways out of its walk
# F5  the same step followed by the river-borne marine diffusion sub-steps (buildFlux.py:198-247, FLOWalgo.f90:332-388)
# positions:    0(g0)   1     2     3    4    5     6    7(g1)       ids: g0 = 0, g1 = 1, position p -> id p + 1
F4A_Z = [-40.0, -30.0, -20.0, -10.0, 5.0, 20.0, 40.0, 60.0]        # a ramp into the sea (sea level 0): diffnb 5, diffprop 0.9
F4B_Z = [-40.0, -5.0, -20.0, -10.0, 5.0, 20.0, 40.0, 60.0]         # a sill at position 1: marine pit at positions 2, 3; diffnb 1, diffprop 0.1
F4_RAIN, F4_EROD, F4_SEA, F4_DT = 1.0, 2.0e-3, 0.0, 204.0
F5_CDR = 0.5                                                        # CFLms = 0.01 * 100^2 / 0.5 = 200 a: sub-steps of 200 a and 4 a


def inputs_marine(diffnb, diffprop, criver=0.0):
    inp = inputs(fillmax=200.0)
    inp.seapos, inp.diffnb, inp.diffprop, inp.CDr = F4_SEA, int(diffnb), float(diffprop), float(criver)
    return inp


def _f4_fluvial(Z):
    """Erosion of the three land nodes (ids 7, 6, 5) and the load that reaches the first marine node (id 4) at dt = 204 a
    (FLOWalgo.f90:635-714, :875-964): id 5 drains into the sea, its dh is 0.99 (z - sea) (:652)."""
    import math
    z = {0: Z[0], 1: Z[7]}
    for p in range(1, 7):
        z[p + 1] = Z[p]
    Q = {7: AREA * F4_RAIN + 0.0}
    Q[6] = AREA * F4_RAIN + Q[7]
    Q[5] = AREA * F4_RAIN + Q[6]
    Q[4] = AREA * F4_RAIN + Q[5]
    Q[3] = AREA * F4_RAIN + Q[4]
    Q[2] = AREA * F4_RAIN + Q[3]
    Q[0] = 0.0 + Q[2]
    Q[1] = 0.0
    assert float(math.floor(DX / (F4_EROD * math.sqrt(Q[2])))) == F4_DT          # flowcfl (FLOWalgo.f90:299-330), id 2 has the largest Q
    ero = {k: 0.0 for k in range(8)}
    for d, dh in ((7, 0.95 * (z[7] - z[6])), (6, 0.95 * (z[6] - z[5])), (5, 0.99 * (z[5] - F4_SEA))):
        spl = -F4_EROD * 1.0 * 1.0 * math.sqrt(Q[d]) * (dh / DX)
        assert -(0.0 + spl) * F4_DT <= dh                                          # no capping (:706-712)
        ero[d] = 0.0 + spl * F4_DT * AREA
    load = -ero[5] + (0.0 + (-ero[6] + (0.0 + (-ero[7] + 0.0))))                   # sedfl handed down ids 7 -> 6 -> 5 -> 4
    return z, Q, ero, load


def _finish(z, ero, elev):
    """deposition = diffsed = elev - elevation (PDalgo.f90:263), sedflux = erosion + deposition (flowNetwork.py:787-791),
    elevation += sedflux (buildFlux.py:193)."""
    n = len(z)
    diffsed = [elev[k] - z[k] for k in range(n)]
    sedflux = [0.0 if k < 2 else ero[k] / AREA + (0.0 + diffsed[k]) for k in range(n)]
    znew = [z[k] + sedflux[k] for k in range(n)]
    return diffsed, sedflux, znew


def f4a_derive():
    """Ramp: 5 rounds of load / 5 start at id 4 (water depth 10 m).  Rounds 1, 2: the share fits under diffprop x the
    room below sea level and stays.  Rounds 3-5: id 4 keeps diffprop x its room, the rest moves to its lowest neighbour
    id 3 (:249-253), which has room for it (:201-204)."""
    z, Q, ero, load = _f4_fluvial(F4A_Z)
    e = dict(z)
    share = load / 5.0                                   # seavol / float(diffnbmax) (:160)
    for rnd in range(5):
        sd = share
        maxz = min(max(e[3], e[5]), F4_SEA)              # highest neighbour, clipped to sea level (:171-176)
        vol = max(0.0, 0.9 * (maxz - e[4]) * AREA)
        assert sd / AREA >= 1.0e-2
        if sd < vol:                                     # rounds 1, 2
            assert rnd < 2
            e[4] = e[4] + sd / AREA
            continue
        assert rnd >= 2
        sd = sd - vol
        e[4] = e[4] + vol / AREA
        assert sd > 0.0 and e[3] < e[4] and e[3] < e[5]  # lowest neighbour is id 3, lower than id 4: pass the rest on
        sd3 = 0.0 + sd
        maxz = min(max(e[2], e[4]), F4_SEA)
        vol3 = max(0.0, 0.9 * (maxz - e[3]) * AREA)
        assert sd3 / AREA >= 1.0e-2 and sd3 < vol3
        e[3] = e[3] + sd3 / AREA
    elev = [e[k] for k in range(8)]
    ero_l = [ero[k] for k in range(8)]
    depo = [0.0] * 8
    depo[4] = 0.0 + load
    diffsed, sedflux, znew = _finish([z[k] for k in range(8)], ero_l, elev)
    return dict(Q=[Q[k] for k in range(8)], ero=ero_l, depo=depo, diffsed=diffsed, sedflux=sedflux, z=znew, load=load)


def f4b_derive():
    """Sill: one round with the whole load (diffnb 1), diffprop 0.1.  id 4 keeps 0.1 x 10 m, passes on to id 3 (bottom of
    the marine pit), which keeps 0.1 x 15 m, has no lower neighbour (:214) and more left than lifts it 0.1 m above its
    highest neighbour: it is raised by dh (:216-218), the rest goes BACK to the start node id 4 (:219-228), which keeps
    0.1 x 9 m, has no lower neighbour either and takes the remainder (:231-234)."""
    z, Q, ero, load = _f4_fluvial(F4B_Z)
    e = dict(z)
    sd = load / 1.0
    vol = max(0.0, 0.1 * (min(max(e[3], e[5]), F4_SEA) - e[4]) * AREA)
    assert sd / AREA >= 1.0e-2 and sd >= vol
    sd = sd - vol
    e[4] = e[4] + vol / AREA
    assert sd > 0.0 and e[3] < e[4]
    sd3 = 0.0 + sd                                       # id 3
    vol3 = max(0.0, 0.1 * (min(max(e[2], e[4]), F4_SEA) - e[3]) * AREA)
    assert sd3 / AREA >= 1.0e-2 and sd3 >= vol3
    sd3 = sd3 - vol3
    e[3] = e[3] + vol3 / AREA
    assert sd3 > 0.0 and e[2] >= e[3] and e[4] >= e[3]  # no lower neighbour
    dh = max(e[2], e[4]) - e[3] + 0.1                   # the unclipped highest neighbour (:207-213)
    assert sd3 > dh * AREA
    e[3] = e[3] + dh
    sd3 = sd3 - dh * AREA
    sd = 0.0 + sd3                                       # back at id 4 (seadep(4) was zeroed when it passed its rest on)
    vol = max(0.0, 0.1 * (min(max(e[3], e[5]), F4_SEA) - e[4]) * AREA)
    assert sd / AREA >= 1.0e-2 and sd >= vol
    sd = sd - vol
    e[4] = e[4] + vol / AREA
    assert sd > 0.0 and e[3] >= e[4] and e[5] >= e[4]
    dh = max(e[3], e[5]) - e[4] + 0.1
    assert sd <= dh * AREA
    e[4] = e[4] + sd / AREA
    elev = [e[k] for k in range(8)]
    ero_l = [ero[k] for k in range(8)]
    depo = [0.0] * 8
    depo[4] = 0.0 + load
    W = [z[k] for k in range(8)]
    W[3] = z[2] + EPS                                    # fillPD: the marine pit fills to its sill (PDalgo.f90:58-79)
    W[4] = W[3] + EPS
    diffsed, sedflux, znew = _finish([z[k] for k in range(8)], ero_l, elev)
    return dict(Q=[Q[k] for k in range(8)], ero=ero_l, depo=depo, diffsed=diffsed, sedflux=sedflux, z=znew, load=load, fillH=W)


def f5_derive():
    """F4a followed by the marine diffusion of the river deposits with criver = 0.5: two sub-steps (200 a, 4 a).  coeff is
    frozen before the loop (buildFlux.py:205): criver / area on ids 2, 3, 4 (below sea level after the fluvial update).
    A node loses to a lower neighbour when it carries more than 0.1 m of river deposits (FLOWalgo.f90:367), gains from a
    higher marine neighbour that does (:369); towards a ghost node only the loss exists (:372-377)."""
    f = f4a_derive()
    z = list(f["z"])
    H = list(f["diffsed"])                               # np.sum(deposition, axis=1): the marine deposits of this step
    cum = list(f["sedflux"])
    assert z[2] < F4_SEA and z[3] < F4_SEA and z[4] < F4_SEA and z[5] >= F4_SEA
    c = 1.0 / AREA * F5_CDR
    left, right = {2: 0, 3: 2, 4: 3}, {2: 3, 3: 4, 4: 5}
    ghost = {0}
    trace = []
    for step in (200.0, 4.0):                            # min(CFLms, diffstep); mindt never binds (checked below)
        d = {}
        for g in (2, 3, 4):
            acc = 0.0
            for nb in (left[g], right[g]):
                flx = DX * (z[nb] - z[g]) / DX
                if nb in ghost:
                    if H[g] > 0.1 and z[g] > z[nb]:
                        acc = acc + c * flx
                        trace.append(("ghost-loss", g, step))
                elif H[g] > 0.1 and z[g] > z[nb]:
                    acc = acc + c * flx
                    trace.append(("loss", g, step))
                elif H[nb] > 0.1 and z[g] < z[nb] and z[nb] < F4_SEA:
                    acc = acc + c * flx
                    trace.append(("gain", g, step))
            assert not (acc < 0.0 and acc * step < -H[g])
            d[g] = acc
        for g in (2, 3, 4):
            H[g] += d[g] * step; z[g] += d[g] * step; cum[g] += d[g] * step
    return dict(z=z, cumdiff=cum, trace=trace, ero=f["ero"], depo=f["depo"], sedflux=f["sedflux"])


# ---------------------------------------------------------------------------------------------------
# F6  overfilled closed basin: getid1 and the re-run of the stream power law with the shortened step
#     (FLOWalgo.f90:1047-1086, flowNetwork.py:662-700)
# positions:   0(g0)  1     2     3     4    5(g1)      ids: g0 = 0, g1 = 1, position p -> id p + 1;  fill_TH = 5 m
F6_Z = [0.0, 30.0, 10.0, 30.0, 50.0, 70.0]
F6_RAIN, F6_EROD, F6_TH = 1.0, 2.0e-3, 5.0


def f6_derive():
    """The lake at id 3 is capped 5 m above its bottom, 15 m below its rim: it is its own receiver on the filled surface and
    drains to itself (a closed basin).  With the CFL step (353 a) it would receive 6.5 x its volume: getid1 finds it
    (deposit > pitVol > 0, allDrain = pitID), the step shrinks to floor(353 x pitVol / deposit) = 54 a and stream power
    runs again."""
    import math
    z = {0: 0.0, 1: 70.0, 2: 30.0, 3: 10.0, 4: 30.0, 5: 50.0}
    W = dict(z)
    W[3] = z[3] + F6_TH                                   # 30 + eps would stand 20.01 m deep: capped (PDalgo.f90:66-69)
    Q = {5: AREA * F6_RAIN + 0.0}
    Q[4] = AREA * F6_RAIN + Q[5]
    Q[3] = AREA * F6_RAIN + Q[4]
    Q[2] = AREA * F6_RAIN + 0.0
    Q[0] = 0.0 + Q[2]
    Q[1] = 0.0
    pitVol3 = -1.0 + (W[3] - z[3]) * AREA
    cfl = float(math.floor(DX / (F6_EROD * math.sqrt(Q[4]))))      # id 4 (Q = 2 A) is the fastest flowing node
    rcv = {2: 0, 4: 3, 5: 4}

    def spl(d):
        dh = 0.95 * (z[d] - z[rcv[d]])
        return -F6_EROD * 1.0 * 1.0 * math.sqrt(Q[d]) * (dh / DX), dh

    def run(dt):
        ero = {k: 0.0 for k in range(6)}
        for d in (5, 4, 2):
            s, dh = spl(d)
            assert -(0.0 + s) * dt <= dh
            ero[d] = 0.0 + s * dt * AREA
        depo3 = 0.0 + (0.0 + (-ero[4] + (0.0 + (-ero[5] + 0.0))))   # everything that enters the closed lake stays (:1003-1008)
        return ero, depo3
    ero, depo3 = run(cfl)
    assert depo3 > pitVol3 > 0.0
    newdt = cfl * (pitVol3 / (0.0 + (depo3 + 0.0)))                 # flowNetwork.py:683-685
    newdt = float(round(newdt - 0.5)) if newdt > 1.0 else newdt    # :687-688 (python-2 round)
    ero, depo3 = run(newdt)
    assert depo3 <= pitVol3
    th = (W[3] - z[3]) * 1.0                                        # land-pit fill (flowNetwork.py:750-763)
    th = th * ((0.0 + depo3) / (th * AREA))
    sed = [0.0, 0.0] + [ero[k] / AREA + (th if k == 3 else 0.0) for k in range(2, 6)]
    return dict(cfl=cfl, dt=newdt, fillH=[W[k] for k in range(6)], Q=[Q[k] for k in range(6)],
                pitVol=[-1.0, -1.0, -1.0, pitVol3, -1.0, -1.0], pitDrain=[-1, -1, -1, 3, -1, -1],
                ero=[ero[k] for k in range(6)], depo=[0.0, 0.0, 0.0, depo3, 0.0, 0.0], sedflux=sed,
                z=[z[k] + sed[k] for k in range(6)])


# ---------------------------------------------------------------------------------------------------
# F7  a coastal lake overflowing into the sea: the marine branch of the pit cascade (FLOWalgo.f90:979-991), the drain of a
#     land pit at the first node below sea level (:207-210), marine_distribution leaving the domain through a ghost node
# positions:    0(g0)   1    2    3    4     5     6    7(g1)       ids: g0 = 0, g1 = 1, position p -> id p + 1; sea level 0
F7_Z = [-40.0, -20.0, 8.0, 2.0, 20.0, 40.0, 60.0, 80.0]


def f7_derive():
    """Lake at id 4 (bottom 2 m, rim id 3 at 8 m, fillH = 8 + eps, volume 60 099 m3); id 3 drains into the sea node id 2.
    basindrainage stops at the first node below sea level: pitDrain[4] = 2, which is no pit.  The lake receives 307 992 m3:
    it keeps its volume, the cascade moves on to `pitDrain`, finds fillH < sea there and, the donor (id 4) standing above
    sea level, hands the excess to the donor's RECEIVER's flux (:986-989): id 3 then carries it (plus its own erosion,
    dh = 0.99 (z - sea)) to id 2, where everything deposits (32 m under 20 m of water).  marine_distribution: rounds 1-3 stay
    at id 2, rounds 4 and 5 leave 0.9 x the room and pass the rest to the lowest neighbour - the ghost node g0, where it is
    dropped (border < 1, PDalgo.f90:167-170)."""
    import math
    z = {0: F7_Z[0], 1: F7_Z[7]}
    for p in range(1, 7):
        z[p + 1] = F7_Z[p]
    W = dict(z)
    W[4] = z[3] + EPS
    Q = {7: AREA * F4_RAIN + 0.0}
    for k in (6, 5, 4, 3, 2):
        Q[k] = AREA * F4_RAIN + Q[k + 1]
    Q[0] = 0.0 + Q[2]
    Q[1] = 0.0
    assert float(math.floor(DX / (F4_EROD * math.sqrt(Q[2])))) == F4_DT
    pitVol4 = -1.0 + (W[4] - z[4]) * AREA
    ero = {k: 0.0 for k in range(8)}
    for d, dh in ((7, 0.95 * (z[7] - z[6])), (6, 0.95 * (z[6] - z[5])), (5, 0.95 * (z[5] - z[4])), (3, 0.99 * (z[3] - F4_SEA))):
        spl = -F4_EROD * 1.0 * 1.0 * math.sqrt(Q[d]) * (dh / DX)
        assert -(0.0 + spl) * F4_DT <= dh
        ero[d] = 0.0 + spl * F4_DT * AREA
    load4 = -ero[5] + (0.0 + (-ero[6] + (0.0 + (-ero[7] + 0.0))))
    assert load4 > pitVol4
    depo = {k: 0.0 for k in range(8)}
    frac = load4 / load4
    depo[4] = 0.0 + (pitVol4 - 0.0) * frac                          # the lake keeps its volume (:1016-1020)
    rest = (load4 - (pitVol4 - 0.0)) * frac
    sedfl3 = 0.0 + rest                                             # fillH(pitDrain) < sea, z(donor) >= sea: to the receiver (:986-989)
    load2 = -ero[3] + sedfl3
    depo[2] = 0.0 + (0.0 + load2)
    # land-pit fill of the lake (flowNetwork.py:750-763)
    th = (W[4] - z[4]) * (depo[4] / (0.0 + depo[4]))
    th4 = th * ((0.0 + depo[4]) / (th * AREA))
    # marine_distribution from id 2 (water depth 20 m): neighbours g0 (-40) and id 3 (8 m -> clipped to sea level)
    e2 = z[2]
    share = load2 / 5.0
    exits = []
    for rnd in range(5):
        vol = max(0.0, 0.9 * (min(max(z[0], z[3]), F4_SEA) - e2) * AREA)
        assert share / AREA >= 1.0e-2
        if share < vol:
            e2 = e2 + share / AREA
            exits.append("stays")
        else:
            e2 = e2 + vol / AREA
            assert share - vol > 0.0 and z[0] < e2                  # the rest goes to g0 and is dropped
            exits.append("ghost")
    assert exits == ["stays", "stays", "stays", "ghost", "ghost"]
    sed = [0.0] * 8
    sed[2] = 0.0 / AREA + (0.0 + (e2 - z[2]))
    sed[3] = ero[3] / AREA + 0.0
    sed[4] = 0.0 / AREA + th4
    for k in (5, 6, 7):
        sed[k] = ero[k] / AREA + 0.0
    return dict(fillH=[W[k] for k in range(8)], Q=[Q[k] for k in range(8)], pitVol=[-1.0] * 4 + [pitVol4] + [-1.0] * 3,
                pitDrain=[-1, -1, -1, -1, 2, -1, -1, -1], ero=[ero[k] for k in range(8)], depo=[depo[k] for k in range(8)],
                sedflux=sed, z=[z[k] + sed[k] for k in range(8)])


# ---------------------------------------------------------------------------------------------------
# F8  flowcfl with a slope exponent n != 1 (FLOWalgo.f90:299-330): dt = min over the flowing nodes of
#     dist / (K Q^m (dz / dist)^(n - 1)), dz taken on the FILLED surface.  F2's staircase with n = 2.
F8_N = 2.0


def f8_cfl_dt():
    import math
    e = f2_derive(1.0)
    W, rcv = e["fillH"], e["rcv"]
    per_node = {}
    for d in range(2, 10):
        dz = W[d] - W[rcv[d]]
        assert dz > 0.0
        per_node[d] = DX / (F2_EROD * math.sqrt(e["Q"][d]) * (dz / DX))          # (dz / dist) ** (n - 1), n - 1 = 1
    assert min(per_node, key=per_node.get) == 2                                    # the outlet node: Q = 8 A, slope 0.1
    return float(math.floor(min(per_node.values()))), per_node


# ---------------------------------------------------------------------------------------------------
# F9  an eps-puddle on the tree of a boundary node (FLOWalgo.f90:150-163, :897-905, :955-967): a node less than eps above its
#     receiver is "flooded" by fillPD (W = W(rcv) + eps > z).  basinparameters gives it the pit id of the tree's root - the
#     ghost node, area 0 - so in streampower the load that reaches it is a pit deposit (waterH > 0) whose pit has no area:
#     neither the routing branch (:955) nor the cascade (:967) fires and the load leaves the model.
# positions:    0(g0)  1     2        3     4     5    6(g1)         ids: g0 = 0, g1 = 1, position p -> id p + 1
F9_Z = [0.0, 10.0, 10.0005, 12.0, 50.0, 90.0, 130.0]
F9_RAIN, F9_EROD = 1.0, 2.0e-3


def f9_derive():
    import math
    n = 7
    z = {0: F9_Z[0], 1: F9_Z[6]}
    for p in range(1, 6):
        z[p + 1] = F9_Z[p]
    W = dict(z)
    W[3] = z[2] + EPS                                             # 10.0005 < 10 + eps: flooded by 9.5 mm
    Q = {6: AREA * F9_RAIN + 0.0}
    for k in (5, 4, 3, 2):
        Q[k] = AREA * F9_RAIN + Q[k + 1]
    Q[0] = 0.0 + Q[2]
    Q[1] = 0.0
    dt = float(math.floor(DX / (F9_EROD * math.sqrt(Q[2]))))     # id 2 carries the largest discharge
    pitID = [0, -1, -1, 0, -1, -1, -1]                            # id 3 belongs to the "pit" of the boundary node g0
    pitVol = [-1.0 + (W[3] - z[3]) * AREA] + [-1.0] * 6
    ero = {k: 0.0 for k in range(n)}
    for d in (6, 5, 4, 2):                                        # id 3 is under water: no incision (:692)
        dh = 0.95 * (z[d] - z[0 if d == 2 else d - 1])
        spl = -F9_EROD * 1.0 * 1.0 * math.sqrt(Q[d]) * (dh / DX)
        assert -(0.0 + spl) * dt <= dh
        ero[d] = 0.0 + spl * dt * AREA
    lost = -ero[4] + (0.0 + (-ero[5] + (0.0 + (-ero[6] + 0.0))))  # reaches id 3 and is dropped
    sed = [0.0, 0.0] + [ero[k] / AREA + 0.0 for k in range(2, n)]
    return dict(dt=dt, fillH=[W[k] for k in range(n)], pitID=pitID, pitVol=pitVol, Q=[Q[k] for k in range(n)],
                ero=[ero[k] for k in range(n)], depo=[0.0] * n, sedflux=sed, z=[z[k] + sed[k] for k in range(n)], lost=lost)
