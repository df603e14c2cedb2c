"""XML configuration for the hot path.

Mirrors the subset of `pyBadlands/forcing/xmlParser.py` (reference file:line in comments) that the
five BASELINE configs exercise: <grid>, <time>, <sea>, <tectonic> (vertical maps only),
<precipitation> (uniform values), <sp_law>, <creep>.  Attribute names are the reference's
(`input.SPLero`, `input.tDisplay`, ...) so that host code written against `model.input` keeps working.
Everything the hot path never reads (strata, waves, carbonates, flexure, 3-D displacements, rivers,
orographic rain, erodibility layers) is rejected loudly rather than silently ignored.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_UNSUPPORTED = ("strata", "rivers", "waveglobal", "wavesed", "erocoeffs", "flexure", "carb",
                "carbonate", "pelagic", "sedfluxfunction", "rockdef", "compaction")


@dataclass
class Input:
    """Flat attribute bag with the reference's defaults (xmlParser.py:41-215)."""
    demfile: Optional[str] = None
    btype: str = "slope"            # xmlParser.py:42, :240-246
    Afactor: int = 1                # :248-254 (resfactor)
    nopit: int = 0
    udw: int = 0
    tStart: float = 0.0
    tEnd: float = 0.0
    tDisplay: float = 0.0
    minDT: float = 1.0              # :328-332
    maxDT: float = 1.0e6            # :333-338 (defaults to tDisplay)
    seapos: float = 0.0             # :365-371
    seafile: Optional[str] = None   # :373-380
    disp3d: bool = False
    tectTime: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))
    tectFile: List[Optional[str]] = field(default_factory=list)
    rainTime: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))
    rainVal: np.ndarray = field(default_factory=lambda: np.zeros(0))
    rainMap: List[Optional[str]] = field(default_factory=list)
    spl: bool = False
    depo: int = 0                   # :877-884 (dep, default 1 inside <sp_law>)
    slp_cr: float = 0.0
    perc_dep: float = 0.0
    fillmax: float = 200.0          # :902-906
    SPLm: float = 0.5
    SPLn: float = 1.0
    SPLero: float = 0.0
    diffnb: int = 5                 # :926-930
    diffprop: float = 0.9           # :931-938
    CDa: float = 0.0
    CDm: float = 0.0
    CDr: float = 0.0
    Sc: float = 0.0
    Hillslope: bool = False
    outDir: str = "output"
    # fixed by scope (never set by the five configs)
    incisiontype: int = 0
    bedslptype: int = 0
    mp: float = 0.0
    erolays: Optional[int] = None
    laytime: float = 0.0
    rockNb: int = 0


def _f(node, tag, default=None, conv=float):
    e = node.find(tag)
    if e is None:
        return default
    return conv(e.text)


def parse_xml(source: str, base_dir: Optional[str] = None, check_files: bool = False) -> Input:
    """Parse an XML file path or XML text into an `Input`.

    `base_dir` is prepended to relative data paths (the reference resolves them against the cwd).
    """
    if os.path.isfile(source):
        root = ET.parse(source).getroot()
    else:
        root = ET.fromstring(source)
    for tag in _UNSUPPORTED:
        if root.find(tag) is not None:
            raise ValueError("XML structure <%s> is outside the B200 hot path (SURVEY.md section 2)" % tag)

    inp = Input()

    def _path(p):
        if p is None:
            return None
        p = p.strip()
        if base_dir is not None and not os.path.isabs(p):
            p = os.path.join(base_dir, p)
        if check_files and not os.path.isfile(p):
            raise ValueError("file is missing or the given path is incorrect: %s" % p)
        return p

    # ---- grid (xmlParser.py:228-275)
    grid = root.find("grid")
    if grid is None:
        raise ValueError("Error in the XmL file: grid structure definition is required!")
    inp.demfile = _path(_f(grid, "demfile", None, str))
    bt = _f(grid, "boundary", "slope", str)
    bt = bt.strip()
    if bt not in ("slope", "flat", "wall", "fixed", "outlet", "wall1"):
        raise ValueError("Error in the definition of the grid structure: Boundary type is either: flat, slope or wall")
    if bt in ("outlet", "wall1"):
        raise ValueError("boundary type %r is outside the B200 hot path" % bt)
    inp.btype = bt
    inp.Afactor = max(1, _f(grid, "resfactor", 1, int))
    inp.nopit = 1 if _f(grid, "nopit", 0, int) >= 1 else 0
    inp.udw = 1 if _f(grid, "udw", 0, int) >= 1 else 0
    if inp.nopit or inp.udw:
        raise ValueError("<nopit>/<udw> are outside the B200 hot path")

    # ---- time (xmlParser.py:277-340)
    t = root.find("time")
    if t is None:
        raise ValueError("Error in the XmL file: time structure definition is required!")
    if t.find("restart") is not None:
        raise ValueError("restart is outside the B200 hot path")
    for tag, attr in (("start", "tStart"), ("end", "tEnd"), ("display", "tDisplay")):
        v = _f(t, tag)
        if v is None:
            raise ValueError("Error in the definition of the simulation time: %s time declaration is required" % tag)
        setattr(inp, attr, v)
    if inp.tStart > inp.tEnd:
        raise ValueError("Error in the definition of the simulation time: start time is greater than end time!")
    if (inp.tEnd - inp.tStart) % inp.tDisplay != 0:
        raise ValueError("Error in the definition of the simulation time: display time needs to be a multiple of simulation time.")
    inp.minDT = _f(t, "mindt", 1.0)
    inp.maxDT = _f(t, "maxdt", inp.tDisplay)

    # ---- sea (xmlParser.py:363-383)
    sea = root.find("sea")
    if sea is not None:
        inp.seapos = _f(sea, "position", 0.0)
        inp.seafile = _path(_f(sea, "curve", None, str))

    # ---- tectonic: vertical displacement maps only (xmlParser.py:385-495)
    tec = root.find("tectonic")
    if tec is not None:
        if _f(tec, "disp3d", 0, int) != 0:
            raise ValueError("3-D displacements are outside the B200 hot path")
        nb = _f(tec, "events", None, int)
        if not nb:
            raise ValueError("The number of tectonic events needs to be defined.")
        tm, fl = [], []
        for d in tec.iter("disp"):
            a, b = _f(d, "dstart"), _f(d, "dend")
            if a is None or b is None or a >= b:
                raise ValueError("Displacement event start and end time values are not properly defined.")
            f_ = _path(_f(d, "dfile", None, str))
            if f_ is None:
                raise ValueError("Displacement event is missing file argument.")
            tm.append((a, b)); fl.append(f_)
        if len(tm) != nb:
            raise ValueError("Number of events %d does not match with the number of declared displacement parameters %d." % (nb, len(tm)))
        # continuous series with None-gaps (:453-487)
        T, F = [], []
        if tm[0][0] > inp.tStart:
            T.append((inp.tStart, tm[0][0])); F.append(None)
        T.append(tm[0]); F.append(fl[0])
        for p in range(1, nb):
            if tm[p][0] > tm[p - 1][1]:
                T.append((tm[p - 1][1], tm[p][0])); F.append(None)
            T.append(tm[p]); F.append(fl[p])
        if tm[-1][1] < inp.tEnd:
            T.append((tm[-1][1], inp.tEnd)); F.append(None)
        inp.tectTime = np.array(T, dtype=float); inp.tectFile = F
    else:
        inp.tectTime = np.array([[inp.tEnd + 1.0e5, inp.tEnd + 2.0e5]])  # :488-494
        inp.tectFile = [None]

    # ---- precipitation: uniform values (xmlParser.py:568-760)
    pr = root.find("precipitation")
    if pr is not None:
        nb = _f(pr, "climates", None, int)
        if nb is None:
            raise ValueError("The number of climatic events needs to be defined.")
        tm, vals = [], []
        for c in pr.iter("rain"):
            for bad in ("map", "ortime", "rzmax"):
                if c.find(bad) is not None:
                    raise ValueError("rain maps / orographic rain are outside the B200 hot path")
            a, b = _f(c, "rstart"), _f(c, "rend")
            if a is None or b is None or a >= b:
                raise ValueError("Climate event start and end time values are not properly defined.")
            tm.append((a, b)); vals.append(_f(c, "rval", 0.0))
        if len(tm) != nb:
            raise ValueError("Number of climates %d does not match with the number of declared rain parameters %d." % (nb, len(tm)))
        T, V = [], []
        if tm[0][0] > inp.tStart:
            T.append((inp.tStart, tm[0][0])); V.append(0.0)
        T.append(tm[0]); V.append(vals[0])
        for p in range(1, nb):
            if tm[p][0] > tm[p - 1][1]:
                T.append((tm[p - 1][1], tm[p][0])); V.append(0.0)
            T.append(tm[p]); V.append(vals[p])
        if tm[-1][1] < inp.tEnd:
            T.append((tm[-1][1], inp.tEnd)); V.append(0.0)
        inp.rainTime = np.array(T, dtype=float); inp.rainVal = np.array(V, dtype=float)
        inp.rainMap = [None] * len(V)
    else:
        inp.rainTime = np.array([[inp.tStart, inp.tEnd]]); inp.rainVal = np.zeros(1); inp.rainMap = [None]

    # ---- sp_law (xmlParser.py:872-945)
    sp = root.find("sp_law")
    if sp is not None:
        inp.spl = True
        inp.depo = _f(sp, "dep", 1, int)
        inp.slp_cr = _f(sp, "slp_cr", 0.0)
        if inp.slp_cr != 0.0:
            raise ValueError("<slp_cr> (forced alluvial deposition) is outside the B200 hot path")
        inp.perc_dep = _f(sp, "perc_dep", 0.0)
        inp.fillmax = _f(sp, "fillmax", 200.0)
        inp.SPLm = _f(sp, "m", 0.5)
        inp.SPLn = _f(sp, "n", 1.0)
        inp.SPLero = _f(sp, "erodibility", 0.0)
        inp.diffnb = _f(sp, "diffnb", 5, int)
        inp.diffprop = _f(sp, "diffprop", 0.9)
        if not (0.0 < inp.diffprop < 1.0):
            raise ValueError("Proportion of marine sediment deposited on downstream nodes needs to be range between ]0,1[")
    else:
        inp.depo = 0; inp.SPLm = 1.0; inp.SPLn = 1.0; inp.SPLero = 0.0; inp.diffnb = 5

    # ---- creep (xmlParser.py:1001-1040)
    cr = root.find("creep")
    if cr is not None:
        inp.CDa = _f(cr, "caerial", 0.0)
        inp.CDm = _f(cr, "cmarine", 0.0)
        inp.Sc = max(0.0, _f(cr, "cslp", 0.0))
        if inp.Sc != 0.0:
            raise ValueError("non-linear creep (<cslp>) is outside the B200 hot path")
        inp.CDr = _f(cr, "criver", 0.0)
        inp.Hillslope = True

    out = root.find("outfolder")
    if out is not None and out.text:
        inp.outDir = out.text.strip()
    return inp
