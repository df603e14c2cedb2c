#!/usr/bin/env python
"""Golden vectors from the reference's OWN code: `libUtils/sfd.c` (the one reference source that compiles here, built by
`oracle/Makefile ref` from where it lies under /root/reference into oracle/_ref/libsfd_ref.so) run on the crater_fast mesh.

    python tools/make_golden_sfd.py        ->  tests/golden/sfd_reference_vectors.npz

Inputs are regenerated from (mesh, seed) by the tests; only the reference's outputs are stored:
  directions_base(z)            -> base1, rcv1                      (sfd.c:113-153)
  directions(fillH, z)          -> base, rcv, maxh, maxdep           (sfd.c:55-111)
  diffusion(z, borders2, ...)   -> diff                             (sfd.c:155-185)
for four surfaces: the initial elevation, two random perturbations and one rounded (tie-rich) surface, and - for the GPU
parity test - directions_base on the initial elevation, which is the real-surface receiver array of the first time step.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayeslands_b200 import config, mesh          # noqa: E402
from oracle import oracle                          # noqa: E402

c_dp, c_ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)


def surfaces(m):
    """The four input surfaces (also used by tests/test_golden.py)."""
    rng = np.random.default_rng(20261020)
    out = []
    for trial in range(4):
        z = m.elevation + (rng.normal(0, 2.0, m.N) if trial else 0.0)
        if trial == 2:
            z = np.round(z)
        z = np.ascontiguousarray(z)
        out.append((z, np.ascontiguousarray(np.maximum(z, z.mean()))))
    return out


def main():
    ref = oracle.ref_sfd()
    if ref is None:
        raise SystemExit("oracle/_ref/libsfd_ref.so is missing: run `make -C oracle ref` where /root/reference exists")
    d = np.load(os.path.join(ROOT, "data", "crater_fast.npz"))
    inp = config.parse_xml(str(d["xml"]))
    m = mesh.build_from_dem(d["dem_xyz"], inp.btype, inp.Afactor)
    N = m.N
    fv = m.FVmesh
    ngb = np.ascontiguousarray(fv.neighbours, dtype=np.int32)
    vor = np.ascontiguousarray(fv.vor_edges, dtype=np.float64)
    edge = np.ascontiguousarray(fv.edge_length, dtype=np.float64)
    bord2 = np.ascontiguousarray(m.borders2, dtype=np.int32)
    gids = np.arange(N, dtype=np.int32)
    out = {}
    for t, (z, fill) in enumerate(surfaces(m)):
        b1 = np.zeros(N, dtype=np.int32); r1 = np.zeros(N, dtype=np.int32)
        ref.directions_base(z.ctypes.data_as(c_dp), ngb.ctypes.data_as(c_ip), vor.ctypes.data_as(c_dp), edge.ctypes.data_as(c_dp),
                            gids.ctypes.data_as(c_ip), b1.ctypes.data_as(c_ip), r1.ctypes.data_as(c_ip), N, N)
        b = np.zeros(N, dtype=np.int32); r = np.zeros(N, dtype=np.int32); mh = np.zeros(N); md = np.zeros(N)
        ref.directions(fill.ctypes.data_as(c_dp), z.ctypes.data_as(c_dp), ngb.ctypes.data_as(c_ip), vor.ctypes.data_as(c_dp),
                       edge.ctypes.data_as(c_dp), gids.ctypes.data_as(c_ip), b.ctypes.data_as(c_ip), r.ctypes.data_as(c_ip),
                       mh.ctypes.data_as(c_dp), md.ctypes.data_as(c_dp), N, N)
        df = np.zeros(N)
        ref.diffusion(z.ctypes.data_as(c_dp), bord2.ctypes.data_as(c_ip), ngb.ctypes.data_as(c_ip), vor.ctypes.data_as(c_dp),
                      edge.ctypes.data_as(c_dp), gids.ctypes.data_as(c_ip), df.ctypes.data_as(c_dp), N, N)
        out.update({"base1_%d" % t: b1, "rcv1_%d" % t: r1, "base_%d" % t: b, "rcv_%d" % t: r, "maxh_%d" % t: mh,
                    "maxdep_%d" % t: md, "diff_%d" % t: df})
    out["N"] = np.int64(N)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sfd_reference_vectors.npz"), **out)
    print("wrote tests/golden/sfd_reference_vectors.npz (N = %d)" % N)


if __name__ == "__main__":
    main()
