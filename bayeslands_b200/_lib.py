"""ctypes binding of the C-ABI (`include/bayeslands_b200.h`, `libbayeslands_b200.so`).

The CUDA library is the only compute path: if it is missing, import fails loudly - there is no CPU
fallback (the CPU oracle under oracle/ is test infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BAYESLANDS_B200_LIB", os.path.join(HERE, "libbayeslands_b200.so"))   # override: profiling builds

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)

BL_KP = 20
BL_NCKPT_MAX = 8
BL_NPTS_MAX = 16
BL_NPHASE = 48
PHASES = ["fill", "sfd", "stack1", "basins", "stack", "drainage", "discharge", "cfl", "streampower",
          "newdt", "erodep", "apply"]

# bl_field
F_Z, F_CUMDIFF, F_CUMHILL, F_FILLH, F_MAXH, F_Q, F_PITVOL, F_ERO, F_DEPO, F_SEDFLUX = range(10)
F_RCV, F_RCV1, F_STACK, F_STACK1, F_PITID, F_PITDRAIN, F_ALLDRAIN = range(32, 39)
FIELDS = dict(z=F_Z, cumdiff=F_CUMDIFF, cumhill=F_CUMHILL, fillH=F_FILLH, maxh=F_MAXH, Q=F_Q, pitVol=F_PITVOL,
              ero=F_ERO, depo=F_DEPO, sedflux=F_SEDFLUX, rcv=F_RCV, rcv1=F_RCV1, stack=F_STACK, stack1=F_STACK1,
              pitID=F_PITID, pitDrain=F_PITDRAIN, allDrain=F_ALLDRAIN)

ST_CASCADE_GUARD, ST_DRAIN_CYCLE, ST_NAN = 1, 4, 8
ST_FATAL = ST_CASCADE_GUARD | ST_NAN      # evaluations whose result must not enter Metropolis
OPT_FORCE_GENERIC, OPT_NO_L2_PERSIST, OPT_COMPACT, OPT_NO_COMPACT = 1, 2, 8, 16


class BlMesh(C.Structure):
    _fields_ = [("N", C.c_int32), ("bounds", C.c_int32), ("ngb", c_ip), ("edge", c_dp), ("vor", c_dp),
                ("area", c_dp), ("xy", c_dp), ("borders", c_ip), ("borders2", c_ip), ("indomain", c_ip),
                ("parent", c_ip), ("z0", c_dp), ("disp", c_dp), ("G", C.c_int32), ("grid_idx", c_ip),
                ("grid_w", c_dp), ("grid_exact", c_ip)]


class BlParams(C.Structure):
    _fields_ = [("spl_m", C.c_double), ("spl_n", C.c_double), ("eps", C.c_double), ("fill_th", C.c_double),
                ("CDa", C.c_double), ("CDm", C.c_double), ("CDr", C.c_double), ("minDT", C.c_double),
                ("maxDT", C.c_double), ("hill_cfl", C.c_double), ("ms_cfl", C.c_double),
                ("diffprop", C.c_double), ("diffnb", C.c_int32), ("depo", C.c_int32), ("btype", C.c_int32),
                ("shuffle_mode", C.c_int32), ("seapos", C.c_double), ("nsea", C.c_int32), ("seat", c_dp),
                ("seav", c_dp), ("real_elev", c_dp), ("npts", C.c_int32), ("pts_cell", c_ip),
                ("opt_tpb", C.c_int32), ("opt_flags", C.c_int32), ("opt_grid_group", C.c_int32),
                ("opt_debug", C.c_int32)]


class BlLattice(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("dist", C.c_double * 8), ("vor", C.c_double * 8),
                ("edge", C.c_double * 8), ("area", C.c_double), ("z0", c_dp), ("flags", C.POINTER(C.c_uint8)),
                ("disp", c_dp), ("parent", c_ip), ("real_elev", c_dp)]


EXPORTS = ("bl_last_error", "bl_version", "bl_workspace_bytes", "bl_ctx_create", "bl_ctx_destroy", "bl_reset",
           "bl_run", "bl_step", "bl_get_field", "bl_get_scalars", "bl_get_phase_cycles", "bl_launch_count", "bl_ctx_info",
           "bl_fill", "bl_sfd", "bl_stack", "bl_discharge", "bl_basins", "bl_drainage", "bl_streampower", "bl_diffuse",
           "bl_interp",
           "bl_grid_workspace_bytes", "bl_grid_create", "bl_grid_destroy", "bl_grid_reset", "bl_grid_run",
           "bl_grid_step", "bl_grid_get_field", "bl_grid_get_scalars", "bl_grid_launch_count")

_lib = None


class BlError(RuntimeError):
    pass


def load():
    """Load libbayeslands_b200.so (built in-tree by `__graft_entry__.build()`); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise BlError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise BlError("libbayeslands_b200.so does not export %s" % name)
    L.bl_last_error.restype = C.c_char_p
    L.bl_version.restype = C.c_int
    L.bl_workspace_bytes.restype = C.c_size_t
    L.bl_workspace_bytes.argtypes = [C.POINTER(BlMesh), C.POINTER(BlParams), C.c_int]
    L.bl_ctx_create.argtypes = [C.POINTER(BlMesh), C.POINTER(BlParams), C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                C.POINTER(C.c_void_p)]
    L.bl_ctx_destroy.argtypes = [C.c_void_p]
    L.bl_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    L.bl_run.argtypes = [C.c_void_p, C.c_int, c_dp, c_ip, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                         C.c_void_p]
    L.bl_step.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    L.bl_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.bl_get_scalars.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bl_get_phase_cycles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.bl_launch_count.restype = C.c_int64
    L.bl_launch_count.argtypes = [C.c_void_p]
    L.bl_ctx_info.restype = C.c_int
    L.bl_ctx_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    vp, dbl, i32, u64, i64 = C.c_void_p, C.c_double, C.c_int, C.c_uint64, C.c_int64
    L.bl_fill.argtypes = [vp, vp, dbl, vp, vp]
    L.bl_sfd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.bl_stack.argtypes = [vp, vp, i32, u64, i64, vp, vp, vp, vp]
    L.bl_discharge.argtypes = [vp, vp, dbl, vp, vp]
    L.bl_basins.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.bl_drainage.argtypes = [vp, vp, vp, vp, vp, dbl, vp, vp, vp]
    L.bl_streampower.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, dbl, dbl, dbl, i32, u64, i64, vp, vp, vp]
    L.bl_diffuse.argtypes = [vp, vp, vp, vp]
    L.bl_interp.argtypes = [vp, vp, vp, vp]
    L.bl_grid_workspace_bytes.restype = C.c_size_t
    L.bl_grid_workspace_bytes.argtypes = [C.POINTER(BlLattice), C.c_int, C.c_int]
    L.bl_grid_create.argtypes = [C.POINTER(BlLattice), C.POINTER(BlParams), C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                 C.c_int, C.POINTER(C.c_void_p)]
    L.bl_grid_destroy.argtypes = [C.c_void_p]
    L.bl_grid_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    L.bl_grid_run.argtypes = [C.c_void_p, C.c_int, c_dp, c_ip, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                              C.c_void_p]
    L.bl_grid_step.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    L.bl_grid_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.bl_grid_get_scalars.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bl_grid_launch_count.restype = C.c_int64
    L.bl_grid_launch_count.argtypes = [C.c_void_p]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise BlError("libbayeslands_b200: status %d: %s" % (rc, load().bl_last_error().decode()))
