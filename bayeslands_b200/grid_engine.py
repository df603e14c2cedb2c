"""Lattice engine front end: B proposals of a regular-grid landscape too large for one CTA (BASELINE config 5).

Same role as `engine.Engine` (the batched `Model.run_to_time` of `bl_mcmc.blackBox`, bl_mcmc.py:126-152) over the
`bl_grid_*` half of the C-ABI: grid-wide kernels batched over proposals instead of one CTA per landscape.  Scope: the
purely erosive configuration (`flow.depo = 0`, constant sea level) - see `csrc/bl_grid.cuh`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, hostparams
from .lattice import Lattice


def _dp(a):
    return a.ctypes.data_as(_lib.c_dp)


def _ip(a):
    return a.ctypes.data_as(_lib.c_ip)


class GridEngine:
    """One context = one lattice + one parameter set + B proposals resident on `device`."""

    def __init__(self, lat: Lattice, inp, B: int, device: int = 0, real_elev: Optional[np.ndarray] = None,
                 erdp_coords: Optional[Sequence] = None, debug: bool = False, fill_group: int = 0,
                 fill_blocks_per_sm: int = 0):
        if not torch.cuda.is_available():
            raise _lib.BlError("bayeslands_b200.GridEngine needs a CUDA device (no CPU fallback)")
        self.L = _lib.load()
        self.lat, self.inp, self.B, self.device = lat, inp, int(B), int(device)
        self.tdev = torch.device("cuda", self.device)
        self.LY, self.LX = lat.shape
        self.G = self.LY * self.LX
        self.grid_ny, self.grid_nx = self.LY, self.LX          # the output grid of interpolateArray is the lattice itself
        k = self._keep = dict(z0=np.ascontiguousarray(lat.z0, dtype=np.float64),
                              flags=np.ascontiguousarray(lat.flags, dtype=np.uint8),
                              parent=np.ascontiguousarray(lat.parent, dtype=np.int32))
        lt = _lib.BlLattice()
        lt.nx, lt.ny, lt.area = lat.nx, lat.ny, lat.area
        for s in range(8):
            lt.dist[s], lt.vor[s], lt.edge[s] = float(lat.dist[s]), float(lat.vor[s]), float(lat.edge[s])
        lt.z0, lt.flags, lt.parent = _dp(k["z0"]), k["flags"].ctypes.data_as(C.POINTER(C.c_uint8)), _ip(k["parent"])
        if lat.disp is not None:
            k["disp"] = np.ascontiguousarray(lat.disp, dtype=np.float64)
            lt.disp = _dp(k["disp"])
        if real_elev is not None:
            k["real"] = np.ascontiguousarray(real_elev, dtype=np.float64).ravel()
            if k["real"].size != self.G:
                raise ValueError("real_elev has %d cells, lattice has %d" % (k["real"].size, self.G))
            lt.real_elev = _dp(k["real"])
        d = hostparams.derive(inp, lat)
        p = _lib.BlParams()
        p.spl_m, p.spl_n, p.eps, p.fill_th = inp.SPLm, inp.SPLn, d.eps, d.fill_th
        p.CDa, p.CDm, p.CDr = inp.CDa, inp.CDm, inp.CDr
        p.minDT, p.maxDT, p.hill_cfl, p.ms_cfl = inp.minDT, inp.maxDT, d.hill_cfl, d.ms_cfl
        p.diffprop, p.diffnb, p.depo, p.btype = inp.diffprop, inp.diffnb, inp.depo, d.btype
        p.shuffle_mode, p.seapos = 0, inp.seapos
        p.opt_grid_group = int(fill_group)                  # proposals filled together (0: library default)
        p.opt_debug = int(fill_blocks_per_sm) << 4          # blocks per SM of the cooperative fill loop (0: occupancy limit)
        if inp.seafile is not None:
            raise _lib.BlError("the lattice engine takes a constant sea level only")
        self.npts = 0
        if erdp_coords is not None:
            cells = np.array([int(c[0]) * self.LX + int(c[1]) for c in erdp_coords], dtype=np.int32)
            k["pts"] = cells
            self.npts = len(cells)
            p.npts, p.pts_cell = self.npts, _ip(cells)
        self._lt, self._p = lt, p
        nbytes = self.L.bl_grid_workspace_bytes(C.byref(lt), self.B, int(debug))
        if nbytes == 0:
            raise _lib.BlError("bl_grid_workspace_bytes failed (nx, ny must be even)")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.tdev)
        ctx = C.c_void_p()
        _lib.check(self.L.bl_grid_create(C.byref(lt), C.byref(p), self.B, self.device,
                                         C.c_void_p(self.workspace.data_ptr()), nbytes, int(debug), C.byref(ctx)))
        self.ctx = ctx
        self.sched = hostparams.Schedule(inp)
        self.tNow = float(inp.tStart)
        self.d_rain = torch.zeros(self.B, dtype=torch.float64, device=self.tdev)
        self.d_erod = torch.zeros(self.B, dtype=torch.float64, device=self.tdev)
        self.out_pts = torch.zeros((self.B, _lib.BL_NCKPT_MAX, _lib.BL_NPTS_MAX), dtype=torch.float64, device=self.tdev)
        self.out_sse = torch.zeros(self.B, dtype=torch.float64, device=self.tdev)

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.L.bl_grid_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    def reset(self, rain, erod, seeds=None):
        """`seeds` is accepted for interface parity with `Engine.reset` and ignored: no stack order is built here, so the
        base shuffle (flowNetwork.py:316) has nothing to act on."""
        def to_dev(x, buf):
            if isinstance(x, torch.Tensor):
                buf.copy_(x.to(dtype=torch.float64), non_blocking=True)
            else:
                buf.copy_(torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64), non_blocking=True)
        to_dev(rain, self.d_rain)
        to_dev(erod, self.d_erod)
        _lib.check(self.L.bl_grid_reset(self.ctx, C.c_void_p(self.d_rain.data_ptr()), C.c_void_p(self.d_erod.data_ptr()),
                                        float(self.inp.tStart), self._stream()))
        self.sched = hostparams.Schedule(self.inp)
        self.tNow = float(self.inp.tStart)

    def run_schedule(self, times: Sequence[float], grids: bool = False):
        """As `Engine.run_schedule`: returns (out_pts [B, len(times), npts], out_sse [B], elev, erdp)."""
        stops, ck = [], []
        for ci, T in enumerate(times):
            tEnd, segs = self.sched.stops_for(self.tNow, float(T))
            if not segs:
                segs = [(self.tNow, -1, -1)]
            for si, s in enumerate(segs):
                stops.append(s[0])
                ck.append(ci if si == len(segs) - 1 else -1)
            self.tNow = max(self.tNow, tEnd)
        nck = len(times)
        if nck > _lib.BL_NCKPT_MAX:
            raise ValueError("at most %d checkpoints per call" % _lib.BL_NCKPT_MAX)
        elev = erdp = None
        if grids:
            elev = torch.empty((self.B, nck, self.G), dtype=torch.float64, device=self.tdev)
            erdp = torch.empty((self.B, nck, self.G), dtype=torch.float64, device=self.tdev)
        st = np.ascontiguousarray(stops, dtype=np.float64)
        ckp = np.ascontiguousarray(ck, dtype=np.int32)
        _lib.check(self.L.bl_grid_run(self.ctx, len(st), _dp(st), _ip(ckp), C.c_void_p(self.out_pts.data_ptr()),
                                      C.c_void_p(self.out_sse.data_ptr()),
                                      C.c_void_p(elev.data_ptr()) if grids else None,
                                      C.c_void_p(erdp.data_ptr()) if grids else None, nck, self._stream()))
        return self.out_pts[:, :nck, :self.npts], self.out_sse, elev, erdp

    def run_to_time(self, tEnd: float):
        tEnd, segs = self.sched.stops_for(self.tNow, float(tEnd))
        if segs:
            st = np.ascontiguousarray([s[0] for s in segs], dtype=np.float64)
            _lib.check(self.L.bl_grid_run(self.ctx, len(st), _dp(st), None, None, None, None, None, 0, self._stream()))
            self.tNow = tEnd

    def step(self, tStop: float):
        """Exactly one time step for every proposal still short of `tStop`."""
        _lib.check(self.L.bl_grid_step(self.ctx, float(tStop), self._stream()))

    def field(self, name: str, proposal: int = 0, mesh_order: bool = True) -> np.ndarray:
        """One per-proposal array: lattice order (ny+2, nx+2), or mesh-id order [N] when the lattice came from a mesh
        (`rcv` is then mesh ids too)."""
        fid = _lib.FIELDS[name]
        out = np.empty((self.LY, self.LX), dtype=np.float64 if fid < 32 else np.int32)
        _lib.check(self.L.bl_grid_get_field(self.ctx, fid, int(proposal), out.ctypes.data_as(C.c_void_p), self._stream()))
        ext = self.lat.ext
        if not mesh_order or ext is None:
            return out
        res = np.empty(self.G, dtype=out.dtype)
        if fid >= 32:
            res[ext.ravel()] = ext.ravel()[out.ravel()]
        else:
            res[ext.ravel()] = out.ravel()
        return res

    def scalars(self):
        tnow = np.empty(self.B); steps = np.empty(self.B, dtype=np.int64)
        sweeps = np.empty(self.B, dtype=np.int64); ldt = np.empty(self.B)
        _lib.check(self.L.bl_grid_get_scalars(self.ctx, tnow.ctypes.data_as(C.c_void_p), steps.ctypes.data_as(C.c_void_p),
                                              sweeps.ctypes.data_as(C.c_void_p), ldt.ctypes.data_as(C.c_void_p),
                                              self._stream()))
        return dict(tNow=tnow, steps=steps, fill_sweeps=sweeps, last_dt=ldt)

    @property
    def launches(self) -> int:
        return int(self.L.bl_grid_launch_count(self.ctx))
