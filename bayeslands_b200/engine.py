"""Batched forward-model engine: B independent (rain, erodibility) proposals on one GPU.

Thin host layer over the C-ABI: PyTorch owns every device buffer (one workspace tensor + outputs) and the
CUDA stream; the library (`csrc/bl_engine.cu`) runs the whole `Model.run_to_time` loop on the device.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, hostparams


def _dp(a):
    return a.ctypes.data_as(_lib.c_dp)


def _ip(a):
    return a.ctypes.data_as(_lib.c_ip)


def indomain_mask(mesh):
    """`flow.domain.contains_points(xycoords)` (simulation/buildFlux.py:135-138)."""
    rg, xy = mesh.recGrid, mesh.FVmesh.node_coords
    x0, x1, y0, y1 = rg.regX.min() - 1.0, rg.regX.max() + 1.0, rg.regY.min() - 1.0, rg.regY.max() + 1.0
    return ((xy[:, 0] > x0) & (xy[:, 0] < x1) & (xy[:, 1] > y0) & (xy[:, 1] < y1)).astype(np.int32)


class Engine:
    """One context = one mesh + one parameter set + B proposals resident on `device`."""

    def __init__(self, mesh, inp, B: int, device: int = 0, real_elev: Optional[np.ndarray] = None,
                 erdp_coords: Optional[Sequence] = None, shuffle_mode: int = 1, sealevel=None, disp=None,
                 tpb: int = 0, force_generic: bool = False, l2_persist: bool = True,
                 debug: int = 0, compact: Optional[bool] = None):
        """Engine options (`bl_params.opt_*`; the library reads no environment variables): `tpb` threads per CTA
        (0 = automatic), `force_generic` the L2-resident device path, `compact` the 8-bytes-per-node shared-memory layout (None = automatic: chosen when it
        lets two landscapes share an SM; True = wherever it fits; False = never)."""
        if not torch.cuda.is_available():
            raise _lib.BlError("bayeslands_b200.Engine needs a CUDA device (no CPU fallback)")
        self.L = _lib.load()
        self.mesh, self.inp, self.B, self.device = mesh, inp, int(B), int(device)
        self.tdev = torch.device("cuda", self.device)
        fv = mesh.FVmesh
        self.N = mesh.N
        self.G = mesh.grid_nx * mesh.grid_ny
        k = self._keep = dict(
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
            gidx=np.ascontiguousarray(mesh.grid_idx, dtype=np.int32),
            gw=np.ascontiguousarray(np.where(np.isfinite(mesh.grid_w), mesh.grid_w, 0.0), dtype=np.float64),
            gex=np.ascontiguousarray(mesh.grid_exact, dtype=np.int32),
        )
        m = _lib.BlMesh()
        m.N, m.bounds = self.N, mesh.boundsPt
        m.ngb, m.edge, m.vor, m.area, m.xy = _ip(k["ngb"]), _dp(k["edge"]), _dp(k["vor"]), _dp(k["area"]), _dp(k["xy"])
        m.borders, m.borders2, m.indomain, m.parent = _ip(k["borders"]), _ip(k["borders2"]), _ip(k["indomain"]), _ip(k["parent"])
        m.z0 = _dp(k["z0"])
        if disp is not None:
            k["disp"] = np.ascontiguousarray(disp, dtype=np.float64)
            m.disp = _dp(k["disp"])
        m.G, m.grid_idx, m.grid_w, m.grid_exact = self.G, _ip(k["gidx"]), _dp(k["gw"]), _ip(k["gex"])
        d = hostparams.derive(inp, mesh)
        p = _lib.BlParams()
        p.spl_m, p.spl_n, p.eps, p.fill_th = inp.SPLm, inp.SPLn, d.eps, d.fill_th
        p.CDa, p.CDm, p.CDr = inp.CDa, inp.CDm, inp.CDr
        p.minDT, p.maxDT, p.hill_cfl, p.ms_cfl = inp.minDT, inp.maxDT, d.hill_cfl, d.ms_cfl
        p.diffprop, p.diffnb, p.depo, p.btype = inp.diffprop, inp.diffnb, inp.depo, d.btype
        p.shuffle_mode, p.seapos = shuffle_mode, inp.seapos
        if sealevel is not None:
            k["seat"] = np.ascontiguousarray(sealevel[:, 0], dtype=np.float64)
            k["seav"] = np.ascontiguousarray(sealevel[:, 1], dtype=np.float64)
            p.nsea, p.seat, p.seav = len(k["seat"]), _dp(k["seat"]), _dp(k["seav"])
        if real_elev is not None:
            k["real"] = np.ascontiguousarray(real_elev, dtype=np.float64).ravel()
            if k["real"].size != self.G:
                raise ValueError("real_elev has %d cells, grid has %d" % (k["real"].size, self.G))
            p.real_elev = _dp(k["real"])
        self.npts = 0
        if erdp_coords is not None:
            cells = np.array([int(c[0]) * mesh.grid_nx + int(c[1]) for c in erdp_coords], dtype=np.int32)
            k["pts"] = cells
            self.npts = len(cells)
            p.npts, p.pts_cell = self.npts, _ip(cells)
        p.opt_tpb = int(tpb)
        p.opt_flags = ((_lib.OPT_FORCE_GENERIC if force_generic else 0) | (0 if l2_persist else _lib.OPT_NO_L2_PERSIST)
                       | (0 if compact is None else (_lib.OPT_COMPACT if compact else _lib.OPT_NO_COMPACT)))
        p.opt_debug = int(debug)
        self._m, self._p = m, p
        nbytes = self.L.bl_workspace_bytes(C.byref(m), C.byref(p), self.B)
        if nbytes == 0:
            raise _lib.BlError("bl_workspace_bytes failed")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.tdev)
        ctx = C.c_void_p()
        _lib.check(self.L.bl_ctx_create(C.byref(m), C.byref(p), self.B, self.device,
                                        C.c_void_p(self.workspace.data_ptr()), nbytes, C.byref(ctx)))
        self.ctx = ctx
        self.sched = hostparams.Schedule(inp)
        self.tNow = float(inp.tStart)
        # persistent device-side parameter / output buffers
        self.d_rain = torch.zeros(self.B, dtype=torch.float64, device=self.tdev)
        self.d_erod = torch.zeros(self.B, dtype=torch.float64, device=self.tdev)
        self.d_seed = torch.zeros(self.B, dtype=torch.int64, device=self.tdev)
        self.out_pts = torch.zeros((self.B, _lib.BL_NCKPT_MAX, _lib.BL_NPTS_MAX), dtype=torch.float64, device=self.tdev)
        self.out_sse = torch.zeros(self.B, dtype=torch.float64, device=self.tdev)

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.L.bl_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    # ---- Model() + load_xml + parameter pokes of blackBox (bl_mcmc.py:126-140) --------------------
    def reset(self, rain, erod, seeds=None):
        """rain/erod: host arrays or device tensors of length B."""
        def to_dev(x, buf, dt):
            if isinstance(x, torch.Tensor):
                buf.copy_(x.to(dtype=dt), non_blocking=True)
            else:
                buf.copy_(torch.as_tensor(np.ascontiguousarray(x), dtype=dt), non_blocking=True)
        to_dev(rain, self.d_rain, torch.float64)
        to_dev(erod, self.d_erod, torch.float64)
        if seeds is None:
            self.d_seed.zero_()
        else:
            to_dev(seeds, self.d_seed, torch.int64)
        _lib.check(self.L.bl_reset(self.ctx, C.c_void_p(self.d_rain.data_ptr()), C.c_void_p(self.d_erod.data_ptr()),
                                   C.c_void_p(self.d_seed.data_ptr()), float(self.inp.tStart), self._stream()))
        self.sched = hostparams.Schedule(self.inp)
        self.tNow = float(self.inp.tStart)

    # ---- Model.run_to_time ------------------------------------------------------------------------
    def run_schedule(self, times: Sequence[float], grids: bool = False):
        """Run every proposal through `times` (the sim_interval of blackBox), treating each as a checkpoint.

        Returns (out_pts [B, len(times), npts] device tensor view, out_sse [B], elev, erdp) where elev/erdp are
        [B, len(times), G] device tensors when `grids` else None.
        """
        stops, ck = [], []
        for ci, T in enumerate(times):
            tEnd, segs = self.sched.stops_for(self.tNow, float(T))
            if not segs:            # run_to_time(t <= tNow): nothing to integrate, still a checkpoint
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
        _lib.check(self.L.bl_run(self.ctx, len(st), _dp(st), _ip(ckp), C.c_void_p(self.out_pts.data_ptr()),
                                 C.c_void_p(self.out_sse.data_ptr()),
                                 C.c_void_p(elev.data_ptr()) if grids else None,
                                 C.c_void_p(erdp.data_ptr()) if grids else None, nck, self._stream()))
        return self.out_pts[:, :nck, :self.npts], self.out_sse, elev, erdp

    def run_to_time(self, tEnd: float):
        tEnd, segs = self.sched.stops_for(self.tNow, float(tEnd))
        if segs:
            st = np.ascontiguousarray([s[0] for s in segs], dtype=np.float64)
            _lib.check(self.L.bl_run(self.ctx, len(st), _dp(st), None, None, None, None, None, 0, self._stream()))
            self.tNow = tEnd

    def step(self, tStop: float):
        """Exactly one time step for every proposal (per-stage parity entry point)."""
        _lib.check(self.L.bl_step(self.ctx, float(tStop), self._stream()))

    # ---- per-stage entry points: caller arrays in, fresh arrays out, like the f2py routines ----------
    def _dev(self, a, dt):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(self.tdev)

    def _ptr(self, t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def pitfilling(self, elev, sealevel):
        """`PDalgo.pdstack.pitfilling(elev, 0, sealevel)` (elevationTIN.py:304) -> fillH."""
        z = self._dev(elev, torch.float64); out = torch.empty_like(z)
        _lib.check(self.L.bl_fill(self.ctx, self._ptr(z), float(sealevel), self._ptr(out), self._stream()))
        return out.cpu().numpy()

    def directions(self, fillH, elev):
        """`sfd.directions(fillH, elev, ...)` (flowNetwork.py:310) -> base, receivers, maxh, maxdep."""
        w = self._dev(fillH, torch.float64); z = self._dev(elev, torch.float64)
        base = torch.empty(self.N, dtype=torch.int32, device=self.tdev); rcv = torch.empty_like(base)
        maxh = torch.empty_like(z); maxdep = torch.empty_like(z)
        _lib.check(self.L.bl_sfd(self.ctx, self._ptr(w), self._ptr(z), self._ptr(base), self._ptr(rcv), self._ptr(maxh),
                                 self._ptr(maxdep), self._stream()))
        return base.cpu().numpy(), rcv.cpu().numpy(), maxh.cpu().numpy(), maxdep.cpu().numpy()

    def directions_base(self, elev):
        """`sfd.directions_base(elev, ...)` (flowNetwork.py:298) -> base1, receivers1."""
        z = self._dev(elev, torch.float64)
        base = torch.empty(self.N, dtype=torch.int32, device=self.tdev); rcv = torch.empty_like(base)
        _lib.check(self.L.bl_sfd(self.ctx, None, self._ptr(z), self._ptr(base), self._ptr(rcv), None, None, self._stream()))
        return base.cpu().numpy(), rcv.cpu().numpy()

    def stack_build(self, receivers, shuffle=0, seed=0, step=0):
        """`FLWnetwork.fstack.build` (flowNetwork.py:393) -> stack, ordered bases."""
        r = self._dev(receivers, torch.int32)
        stack = torch.empty_like(r); base = torch.empty_like(r); nb = torch.zeros(1, dtype=torch.int32, device=self.tdev)
        _lib.check(self.L.bl_stack(self.ctx, self._ptr(r), int(shuffle), int(seed), int(step), self._ptr(stack), self._ptr(base),
                                   self._ptr(nb), self._stream()))
        return stack.cpu().numpy(), base.cpu().numpy()[:int(nb.item())]

    def discharge(self, receivers, rain):
        """`flowcompute.discharge` with the uniform-rain source `area * rain` (flowNetwork.py:418-443) -> Q."""
        r = self._dev(receivers, torch.int32); q = torch.empty(self.N, dtype=torch.float64, device=self.tdev)
        _lib.check(self.L.bl_discharge(self.ctx, self._ptr(r), float(rain), self._ptr(q), self._stream()))
        return q.cpu().numpy()

    def basinparameters(self, receivers1, elev, fillH):
        """`flowcompute.basinparameters` (flowNetwork.py:549) -> pitID, pitVolume."""
        r = self._dev(receivers1, torch.int32); z = self._dev(elev, torch.float64); w = self._dev(fillH, torch.float64)
        pid = torch.empty_like(r); vol = torch.empty_like(z)
        _lib.check(self.L.bl_basins(self.ctx, self._ptr(r), self._ptr(z), self._ptr(w), self._ptr(pid), self._ptr(vol), self._stream()))
        return pid.cpu().numpy(), vol.cpu().numpy()

    def basindrainage(self, pitID, pitVolume, receivers, fillH, sealevel):
        """`flowcompute.basindrainage` + `basindrainageall` (flowNetwork.py:558-568) -> pitDrain, allDrain."""
        pid = self._dev(pitID, torch.int32); vol = self._dev(pitVolume, torch.float64)
        r = self._dev(receivers, torch.int32); w = self._dev(fillH, torch.float64)
        d1 = torch.empty_like(r); d2 = torch.empty_like(r)
        _lib.check(self.L.bl_drainage(self.ctx, self._ptr(pid), self._ptr(vol), self._ptr(r), self._ptr(w), float(sealevel),
                                      self._ptr(d1), self._ptr(d2), self._stream()))
        return d1.cpu().numpy(), d2.cpu().numpy()

    def streampower(self, receivers, pitID, pitVolume, pitDrain, maxh, Q, fillH, elev, erod, sealevel, dt, shuffle=0,
                    seed=0, step=0):
        """`flowcompute.streampower` (flowNetwork.py:656) -> cdepo, cero (m^3)."""
        r = self._dev(receivers, torch.int32); pid = self._dev(pitID, torch.int32); vol = self._dev(pitVolume, torch.float64)
        pdr = self._dev(pitDrain, torch.int32); mh = self._dev(maxh, torch.float64); q = self._dev(Q, torch.float64)
        w = self._dev(fillH, torch.float64); z = self._dev(elev, torch.float64)
        depo = torch.empty_like(z); ero = torch.empty_like(z)
        _lib.check(self.L.bl_streampower(self.ctx, self._ptr(r), self._ptr(pid), self._ptr(vol), self._ptr(pdr), self._ptr(mh), self._ptr(q),
                                         self._ptr(w), self._ptr(z), float(erod), float(sealevel), float(dt), int(shuffle), int(seed),
                                         int(step), self._ptr(depo), self._ptr(ero), self._stream()))
        return depo.cpu().numpy(), ero.cpu().numpy()

    def diffusion(self, elev):
        """`sfd.diffusion(elev, borders2, ...)` (flowNetwork.py:148) -> diffusion flux."""
        z = self._dev(elev, torch.float64); out = torch.empty_like(z)
        _lib.check(self.L.bl_diffuse(self.ctx, self._ptr(z), self._ptr(out), self._stream()))
        return out.cpu().numpy()

    def interpolate(self, v):
        """`bl_mcmc.interpolateArray` (bl_mcmc.py:167-215) on the mesh-constant table -> (ny, nx) grid."""
        x = self._dev(v, torch.float64); out = torch.empty(self.G, dtype=torch.float64, device=self.tdev)
        _lib.check(self.L.bl_interp(self.ctx, self._ptr(x), self._ptr(out), self._stream()))
        return out.cpu().numpy().reshape(self.mesh.grid_ny, self.mesh.grid_nx)

    # ---- inspection ---------------------------------------------------------------------------------
    def field(self, name: str, proposal: int = 0) -> np.ndarray:
        fid = _lib.FIELDS[name]
        out = np.empty(self.N, dtype=np.float64 if fid < 32 else np.int32)
        _lib.check(self.L.bl_get_field(self.ctx, fid, int(proposal), out.ctypes.data_as(C.c_void_p), self._stream()))
        return out

    def scalars(self):
        tnow = np.empty(self.B); steps = np.empty(self.B, dtype=np.int64)
        status = np.empty(self.B, dtype=np.uint32); ldt = np.empty(self.B)
        _lib.check(self.L.bl_get_scalars(self.ctx, tnow.ctypes.data_as(C.c_void_p), steps.ctypes.data_as(C.c_void_p),
                                         status.ctypes.data_as(C.c_void_p), ldt.ctypes.data_as(C.c_void_p),
                                         self._stream()))
        return dict(tNow=tnow, steps=steps, status=status, last_dt=ldt)

    def status(self) -> np.ndarray:
        """Per-proposal status bits (`BL_ST_*`) of the evaluations run since the last reset."""
        st = np.empty(self.B, dtype=np.uint32)
        _lib.check(self.L.bl_get_scalars(self.ctx, None, None, st.ctypes.data_as(C.c_void_p), None, self._stream()))
        return st

    def phase_cycles(self) -> np.ndarray:
        """[B, BL_NPHASE] SM cycles per phase (see _lib.PHASES), accumulated since reset."""
        out = np.zeros((self.B, _lib.BL_NPHASE), dtype=np.int64)
        _lib.check(self.L.bl_get_phase_cycles(self.ctx, out.ctypes.data_as(C.c_void_p), self._stream()))
        return out

    @property
    def launches(self) -> int:
        return int(self.L.bl_launch_count(self.ctx))

    def info(self) -> dict:
        """Launch configuration the context selected (`bl_ctx_info`)."""
        a = (C.c_int32 * 4)()
        _lib.check(self.L.bl_ctx_info(self.ctx, a))
        return dict(tpb=int(a[0]), path=("generic", "shared", "compact")[int(a[1])], dyn_smem=int(a[2]),
                    ctas_per_sm=int(a[3]))
