"""ctypes front-end of the CPU ORACLE (oracle/bl_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` legs may import
this module.  The product package `bayeslands_b200` never does ("parity unpinned" at per-kernel level
except sfd.c, see bl_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from bayeslands_b200 import hostparams

HERE = os.path.dirname(os.path.abspath(__file__))
KP = 20
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class _Mesh(C.Structure):
    _fields_ = [("N", C.c_int), ("bounds", C.c_int), ("ngb", c_ip), ("edge", c_dp), ("vor", c_dp),
                ("area", c_dp), ("xy", c_dp), ("borders", c_ip), ("borders2", c_ip), ("indomain", c_ip),
                ("parent", c_ip)]


class _Params(C.Structure):
    _fields_ = [("spl_m", C.c_double), ("spl_n", C.c_double), ("eps", C.c_double), ("fill_th", C.c_double),
                ("CDa", C.c_double), ("CDm", C.c_double), ("CDr", C.c_double), ("minDT", C.c_double),
                ("maxDT", C.c_double), ("hill_cfl", C.c_double), ("ms_cfl", C.c_double),
                ("diffprop", C.c_double), ("diffnb", C.c_int), ("depo", C.c_int), ("btype", C.c_int),
                ("pow_mode", C.c_int), ("shuffle_mode", C.c_int), ("fill_mode", C.c_int),
                ("seed", C.c_uint64), ("seapos", C.c_double), ("nsea", C.c_int), ("seat", c_dp),
                ("seav", c_dp), ("tStart", C.c_double), ("tEnd", C.c_double), ("apply_disp", C.c_int),
                ("disp", c_dp)]


class _State(C.Structure):
    _fields_ = [("N", C.c_int), ("tNow", C.c_double), ("rain", C.c_double), ("erod", C.c_double),
                ("step", C.c_long), ("fill_sweeps", C.c_long), ("rerun", C.c_long),
                ("z", c_dp), ("cumdiff", c_dp), ("cumhill", c_dp),
                ("fillH", c_dp), ("maxh", c_dp), ("maxdep", c_dp), ("Q", c_dp), ("lay", c_dp), ("pitVol", c_dp),
                ("rcv", c_ip), ("rcv1", c_ip), ("stack", c_ip), ("stack1", c_ip), ("donors", c_ip),
                ("donors1", c_ip), ("pitID", c_ip), ("pitDrain", c_ip), ("allDrain", c_ip),
                ("nbase", C.c_int), ("nbase1", C.c_int), ("base", c_ip), ("base1", c_ip),
                ("ero", c_dp), ("depo", c_dp), ("sedflux", c_dp), ("last_dt", C.c_double),
                ("last_cfl", C.c_double),
                ("w0", c_dp), ("w1", c_dp), ("w2", c_dp), ("w3", c_dp), ("w4", c_dp), ("w5", c_dp),
                ("i0", c_ip), ("i1", c_ip), ("i2", c_ip), ("i3", c_ip), ("i4", c_ip)]


def build(force=False):
    """Compile the C restatement (and oracle/_ref from the reference when it is present)."""
    so = os.path.join(HERE, "libbl_oracle.so")
    src = os.path.join(HERE, "bl_oracle.c")
    if force or not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libbl_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/pyBadlands/libUtils"):
        ref = os.path.join(HERE, "_ref", "libsfd_ref.so")
        if force or not os.path.isfile(ref):
            subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.or_state_create.restype = C.POINTER(_State)
        L.or_state_create.argtypes = [C.c_int]
        L.or_state_destroy.argtypes = [C.POINTER(_State)]
        L.or_state_reset.argtypes = [C.POINTER(_State), c_dp, C.c_double, C.c_double, C.c_double]
        L.or_sea.restype = C.c_double
        L.or_sea.argtypes = [C.POINTER(_Params), C.c_double]
        L.or_streamflow.argtypes = [C.POINTER(_Mesh), C.POINTER(_Params), C.POINTER(_State)]
        L.or_sediment_flux.argtypes = [C.POINTER(_Mesh), C.POINTER(_Params), C.POINTER(_State), C.c_double]
        L.or_sediment_flux.restype = C.c_int
        L.or_run_to.argtypes = [C.POINTER(_Mesh), C.POINTER(_Params), C.POINTER(_State), C.c_int, c_dp]
        L.or_run_to.restype = C.c_int
        L.or_pitfilling.restype = C.c_long
        L.or_pairwise_sum.restype = C.c_double
        L.or_pairwise_sum.argtypes = [c_dp, C.c_long]
        L.or_flowcfl.restype = C.c_double
        L.or_sse.restype = C.c_double
        L.or_sse.argtypes = [c_dp, c_dp, C.c_long, c_dp]
        L.or_interp.argtypes = [C.c_int, c_ip, c_dp, c_ip, c_dp, c_dp]
        _lib = L
    return _lib


def ref_sfd():
    """The reference's own sfd.c compiled into oracle/_ref (None if it was never built)."""
    p = os.path.join(HERE, "_ref", "libsfd_ref.so")
    return C.CDLL(p) if os.path.isfile(p) else None


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def indomain_mask(mesh):
    rg, xy = mesh.recGrid, mesh.FVmesh.node_coords
    x0, x1, y0, y1 = rg.regX.min() - 1.0, rg.regX.max() + 1.0, rg.regY.min() - 1.0, rg.regY.max() + 1.0
    return ((xy[:, 0] > x0) & (xy[:, 0] < x1) & (xy[:, 1] > y0) & (xy[:, 1] < y1)).astype(np.int32)


class OracleModel:
    """One proposal's forward model on the CPU (Model + flowNetwork + buildFlux restated)."""

    def __init__(self, mesh, inp, pow_mode=1, shuffle_mode=1, fill_mode=0, seed=0, sealevel=None,
                 disp=None):
        self.L = lib()
        self.mesh, self.inp = mesh, inp
        fv = mesh.FVmesh
        self.N = mesh.N
        self._keep = dict(
            ngb=np.ascontiguousarray(fv.neighbours, dtype=np.int32),
            edge=np.ascontiguousarray(fv.edge_length, dtype=np.float64),
            vor=np.ascontiguousarray(fv.vor_edges, dtype=np.float64),
            area=np.ascontiguousarray(fv.control_volumes, dtype=np.float64),
            xy=np.ascontiguousarray(fv.node_coords[:, :2], dtype=np.float64),
            borders=np.ascontiguousarray(mesh.borders, dtype=np.int32),
            borders2=np.ascontiguousarray(mesh.borders2, dtype=np.int32),
            indomain=indomain_mask(mesh),
            parent=np.ascontiguousarray(mesh.parentIDs, dtype=np.int32),
            z0=np.ascontiguousarray(mesh.elevation, dtype=np.float64),
        )
        k = self._keep
        self.cm = _Mesh(self.N, mesh.boundsPt, _ip(k["ngb"]), _dp(k["edge"]), _dp(k["vor"]), _dp(k["area"]),
                        _dp(k["xy"]), _ip(k["borders"]), _ip(k["borders2"]), _ip(k["indomain"]), _ip(k["parent"]))
        d = hostparams.derive(inp, mesh)
        self.derived = d
        p = _Params()
        p.spl_m, p.spl_n, p.eps, p.fill_th = inp.SPLm, inp.SPLn, d.eps, d.fill_th
        p.CDa, p.CDm, p.CDr = inp.CDa, inp.CDm, inp.CDr
        p.minDT, p.maxDT, p.hill_cfl, p.ms_cfl = inp.minDT, inp.maxDT, d.hill_cfl, d.ms_cfl
        p.diffprop, p.diffnb, p.depo, p.btype = inp.diffprop, inp.diffnb, inp.depo, d.btype
        p.pow_mode, p.shuffle_mode, p.fill_mode, p.seed = pow_mode, shuffle_mode, fill_mode, seed
        p.seapos = inp.seapos
        if sealevel is not None:
            k["seat"] = np.ascontiguousarray(sealevel[:, 0], dtype=np.float64)
            k["seav"] = np.ascontiguousarray(sealevel[:, 1], dtype=np.float64)
            p.nsea, p.seat, p.seav = len(k["seat"]), _dp(k["seat"]), _dp(k["seav"])
        else:
            p.nsea = 0
        p.tStart, p.tEnd = inp.tStart, inp.tEnd
        if disp is not None:
            k["disp"] = np.ascontiguousarray(disp, dtype=np.float64)
            p.apply_disp, p.disp = 1, _dp(k["disp"])
        self.cp = p
        self.st = self.L.or_state_create(self.N)
        self.sched = hostparams.Schedule(inp)
        self.reset(float(inp.rainVal[0]) if len(inp.rainVal) else 0.0, inp.SPLero)

    def __del__(self):
        try:
            self.L.or_state_destroy(self.st)
        except Exception:
            pass

    # ---- state access -------------------------------------------------------------------
    def _arr(self, name, dtype=np.float64, n=None):
        ptr = getattr(self.st.contents, name)
        return np.ctypeslib.as_array(ptr, shape=(n or self.N,))

    def __getattr__(self, name):
        if name in ("z", "cumdiff", "cumhill", "fillH", "maxh", "maxdep", "Q", "lay", "pitVol", "rcv", "rcv1",
                    "stack", "stack1", "donors", "donors1", "pitID", "pitDrain", "allDrain", "ero", "depo",
                    "sedflux"):
            return self._arr(name)
        raise AttributeError(name)

    @property
    def base(self):
        return self._arr("base", n=self.st.contents.nbase).copy()

    @property
    def base1(self):
        return self._arr("base1", n=self.st.contents.nbase1).copy()

    @property
    def tNow(self):
        return self.st.contents.tNow

    @property
    def steps(self):
        return self.st.contents.step

    @property
    def fill_sweeps(self):
        return self.st.contents.fill_sweeps

    @property
    def reruns(self):
        return self.st.contents.rerun

    @property
    def last_dt(self):
        return self.st.contents.last_dt

    # ---- control -------------------------------------------------------------------------
    def reset(self, rain, erod, z0=None, seed=None):
        z = self._keep["z0"] if z0 is None else np.ascontiguousarray(z0, dtype=np.float64)
        self._keep["zcur"] = z
        if seed is not None:
            self.cp.seed = seed
        self.L.or_state_reset(self.st, _dp(z), float(rain), float(erod), float(self.inp.tStart))
        self.sched = hostparams.Schedule(self.inp)

    def sea(self, t):
        return self.L.or_sea(C.byref(self.cp), float(t))

    def streamflow(self):
        self.L.or_streamflow(C.byref(self.cm), C.byref(self.cp), self.st)

    def sediment_flux(self, tStop):
        return self.L.or_sediment_flux(C.byref(self.cm), C.byref(self.cp), self.st, float(tStop))

    def run_to_time(self, tEnd):
        tEnd, segs = self.sched.stops_for(self.tNow, tEnd)
        stops = np.array([s[0] for s in segs], dtype=np.float64)
        if len(stops):
            return self.L.or_run_to(C.byref(self.cm), C.byref(self.cp), self.st, len(stops), _dp(stops))
        return 0

    # ---- single routines on explicit arrays (micro-fixtures, per-stage parity tests) --------
    def streampower_routine(self, stack, rcv, pitID, pitVol, pitDrain, maxh, Q, fillH, z, erod, sea, dt):
        """`flowcompute.streampower` (FLOWalgo.f90:583-1044) -> cdepo, cero."""
        N = self.N
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        stack, rcv, pitID, pitDrain = i32(stack), i32(rcv), i32(pitID), i32(pitDrain)
        pitVol, maxh, Q, fillH, z = f64(pitVol), f64(maxh), f64(Q), f64(fillH), f64(z)
        maxdep = np.zeros(N); depo = np.zeros(N); ero = np.zeros(N); sedfl = np.zeros(N); pv = np.zeros(N)
        self.L.or_streampower.restype = C.c_int
        self.L.or_streampower.argtypes = [C.POINTER(_Mesh), C.POINTER(_Params), C.c_int, c_ip, c_ip, c_ip, c_dp, c_ip, c_dp, c_dp,
                                          c_dp, c_dp, c_dp, C.c_double, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp]
        self.L.or_streampower(C.byref(self.cm), C.byref(self.cp), N, _ip(stack), _ip(rcv), _ip(pitID), _dp(pitVol), _ip(pitDrain),
                              _dp(maxh), _dp(maxdep), _dp(Q), _dp(fillH), _dp(z), float(erod), float(sea), float(dt), _dp(depo),
                              _dp(ero), _dp(sedfl), _dp(pv))
        return depo, ero

    def basindrainage_routine(self, pitID, pitVol, rcv, fillH, sea):
        """`compute_parameters_depression` drainage part (flowNetwork.py:558-568): pit order + basindrainage(all)."""
        N = len(pitID)
        pitID = np.ascontiguousarray(pitID, dtype=np.int32); rcv = np.ascontiguousarray(rcv, dtype=np.int32)
        fillH = np.ascontiguousarray(fillH, dtype=np.float64)
        pIDs = np.ascontiguousarray(np.where(np.asarray(pitVol) >= 0.0)[0], dtype=np.int32)
        order = np.ascontiguousarray(np.argsort(fillH[pIDs], kind="stable")[::-1], dtype=np.int32)
        d1 = np.zeros(N, dtype=np.int32); d2 = np.zeros(N, dtype=np.int32)
        ch = np.zeros(max(1, len(pIDs)), dtype=np.int32)
        self.L.or_basindrainage.argtypes = [C.c_int, c_ip, c_ip, c_ip, c_ip, c_dp, C.c_double, C.c_int, c_ip, c_ip]
        self.L.or_basindrainageall.argtypes = [C.c_int, c_ip, c_ip, c_ip, c_ip, C.c_int, c_ip, c_ip]
        self.L.or_basindrainage(len(pIDs), _ip(order), _ip(pitID), _ip(rcv), _ip(pIDs), _dp(fillH), float(sea), N, _ip(d1), _ip(ch))
        self.L.or_basindrainageall(len(pIDs), _ip(order), _ip(pitID), _ip(rcv), _ip(pIDs), N, _ip(d2), _ip(ch))
        return d1, d2

    # ---- bl_mcmc.blackBox / interpolateArray (bl_mcmc.py:97-215) --------------------------
    def interpolate(self, v):
        m = self.mesh
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty(m.grid_nx * m.grid_ny)
        self.L.or_interp(len(out), _ip(np.ascontiguousarray(m.grid_idx)), _dp(np.ascontiguousarray(m.grid_w)),
                         _ip(np.ascontiguousarray(m.grid_exact)), _dp(v), _dp(out))
        return out.reshape(m.grid_ny, m.grid_nx)

    def blackbox(self, rain, erod, simtime, erdp_coords, seed=None):
        self.reset(rain, erod, seed=seed)
        elev_vec, erdp_vec, pts_vec = {}, {}, {}
        for T in hostparams.sim_interval(simtime):
            self.run_to_time(float(T))
            elev = self.interpolate(self.z)
            erdp = self.interpolate(self.cumdiff)
            elev_vec[T], erdp_vec[T] = elev, erdp
            pts_vec[T] = np.array([erdp[c[0], c[1]] for c in erdp_coords])
        return elev_vec, erdp_vec, pts_vec


def likelihood(pred_elev_final, real_elev, pred_pts, real_pts, likl_sed, sed_weight=50.0):
    """bl_mcmc.likelihoodFunc arithmetic (bl_mcmc.py:488-508) in numpy, as the reference evaluates it."""
    tausq = np.sum(np.square(pred_elev_final - real_elev)) / real_elev.size
    l_elev = -0.5 * np.log(2 * np.pi * tausq) - 0.5 * np.square(pred_elev_final - real_elev) / tausq
    lik = np.sum(l_elev)
    if likl_sed:
        lp = 0.0
        for i in range(1, pred_pts.shape[0]):
            t2 = np.sum(np.square(pred_pts[i] - real_pts[i])) / real_pts.shape[1]
            lp += np.sum(-0.5 * np.log(2 * np.pi * t2) - 0.5 * np.square(pred_pts[i] - real_pts[i]) / t2)
        lik = lik + lp * sed_weight
    return lik, tausq
