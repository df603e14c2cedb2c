#!/usr/bin/env python
"""Import the reference's shipped example DATA (not sources) as compact fixtures.

Run once in the build container (where /root/reference exists):

    python tools/import_reference_data.py

Writes `data/<config>.npz` with the inputs the hot path needs (DEM, sea-level curve, observed
final_elev / final_erdp / final_erdp_pts, likelihood-surface sweeps `exp_data.txt`) so that nothing in
tests/, bench.py or smoke() has to read /root/reference at run time (it does not exist on the GPU box).

Provenance of each array (all under /root/reference/Examples/<config>/):
  dem_xyz            data/res_crater.csv | data/nodes.csv      (x y z rows, x fastest)
  sealevel           data/sealevel.csv                          (t z rows)        [etopo only]
  final_elev         data/final_elev.txt                        (ny+2, nx+2) grid + Gaussian noise
  final_erdp         data/final_erdp.txt
  final_erdp_pts     data/final_erdp_pts.txt                    (5, 10)
  surf_false/true    *liklsurf_s_{false,true}/exp_data.txt      (2500, >=3): rain erod likl ...
"""
import os
import sys
import numpy as np

REF = "/root/reference/Examples"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")

CONFIGS = {
    "crater_fast": dict(dem="data/res_crater.csv", surf=("c_f_liklsurf_s_false", "c_f_liklsurf_s_true")),
    "crater": dict(dem="data/res_crater.csv", surf=("c_liklsurf_s_false", "c_liklsurf_s_true")),
    "etopo_fast": dict(dem="data/nodes.csv", sea="data/sealevel.csv",
                       surf=("e_f_liklsurf_s_false", "e_f_liklsurf_s_true")),
    "etopo": dict(dem="data/nodes.csv", sea="data/sealevel.csv",
                  surf=("e_liklsurf_s_false", "e_liklsurf_s_true")),
}


def _load_surf(path):
    if not os.path.isfile(path):
        return np.zeros((0, 3))
    rows = []
    with open(path) as f:
        for line in f:
            p = line.replace("[", " ").replace("]", " ").split()
            if len(p) >= 3:
                rows.append([float(v) for v in p[:5]])
    n = min(len(r) for r in rows)
    return np.array([r[:n] for r in rows])


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, c in CONFIGS.items():
        d = os.path.join(REF, name)
        arrs = {}
        arrs["dem_xyz"] = np.loadtxt(os.path.join(d, c["dem"]))
        if "sea" in c:
            arrs["sealevel"] = np.loadtxt(os.path.join(d, c["sea"]))
        for k in ("final_elev", "final_erdp", "final_erdp_pts"):
            arrs[k] = np.loadtxt(os.path.join(d, "data", k + ".txt"))
        arrs["surf_false"] = _load_surf(os.path.join(d, c["surf"][0], "exp_data.txt"))
        arrs["surf_true"] = _load_surf(os.path.join(d, c["surf"][1], "exp_data.txt"))
        with open(os.path.join(d, [f for f in os.listdir(d) if f.endswith(".xml")][0])) as f:
            arrs["xml"] = np.array(f.read())
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        print(name, {k: getattr(v, "shape", None) for k, v in arrs.items()})
    # mountain: template for the synthetic uplift config (uplift map semantics only)
    d = os.path.join(REF, "mountain")
    np.savez_compressed(os.path.join(OUT, "mountain.npz"),
                        dem_xyz=np.loadtxt(os.path.join(d, "data/nodes.csv")),
                        uplift=np.loadtxt(os.path.join(d, "data/uplift.csv")))


if __name__ == "__main__":
    sys.exit(main())
