"""Host-side scalars and time bookkeeping shared by every proposal (python, once per config).

* `derive(inp, mesh)`  - constants the reference computes once per model: hillslope CFLs
  (`hillslope/diffLinear.py:40-75`, `:115-148`), boundary code, PD parameters
  (`surface/elevationTIN.py:245`, `libUtils/PDalgo.f90:271-298`).
* `Schedule`           - the stop-time logic of `Model.run_to_time` (`pyBadlands/model.py:246-259`,
  `:275-292`, `:447-486`, `:529-533`).  Because every sub-step is clipped to `tStop - tNow`
  (`simulation/buildFlux.py:175`) and all times are integer-valued, each proposal reaches every stop
  exactly, so the sequence of stops is the same for all proposals and can be computed on the host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

BTYPE_CODE = {"fixed": 0, "slope": 1, "flat": 2, "wall": 3}


@dataclass
class Derived:
    hill_cfl: float
    ms_cfl: float
    btype: int
    eps: float
    fill_th: float


def derive(inp, mesh) -> Derived:
    e2 = mesh.hillslope_min_edge2
    maxCD = max(inp.CDa, inp.CDm)
    if inp.Hillslope:
        hill = 0.05 * e2 / maxCD if maxCD > 0.0 else 1.0e6              # diffLinear.py:66-70
        ms = 0.01 * e2 if inp.CDr == 0.0 else 0.01 * e2 / inp.CDr       # diffLinear.py:140-144
    else:
        hill = inp.tEnd - inp.tStart                                      # buildFlux.py:166-167
        ms = 0.01 * e2
    return Derived(hill_cfl=float(hill), ms_cfl=float(ms), btype=BTYPE_CODE[inp.btype], eps=0.01,
                   fill_th=float(inp.fillmax))


class Schedule:
    """`force.next_*` state of one Model across successive `run_to_time` calls."""

    def __init__(self, inp):
        self.inp = inp
        self.started = False
        self.next_rain = self.next_disp = self.next_display = 0.0
        self.rain_event = -1
        self.disp_event = -1

    def _rain_event(self, t):
        T = self.inp.rainTime
        return int(np.sum((T[:, 1] - t) <= 0))                            # forceSim.py:318-319

    def stops_for(self, tNow: float, tEnd: float):
        """Stop times a `run_to_time(tEnd)` call steps through, with the rain / displacement event active
        on each segment.  Returns (tEnd_clamped, [(tStop, rain_event, disp_event or -1), ...])."""
        inp = self.inp
        tEnd = min(tEnd, inp.tEnd)                                        # model.py:240-243
        if not self.started:                                              # model.py:246-259
            self.next_rain = inp.rainTime[0, 0]
            self.next_disp = inp.tectTime[0, 0]
            self.next_display = inp.tStart
            self.started = True
        far = inp.tEnd + 1000.0                                           # next_layer / flexure / wave / carb
        segs = []
        t = tNow
        while t < tEnd:
            if self.next_rain <= t and self.next_rain < inp.tEnd:         # model.py:275-280
                self.rain_event = self._rain_event(t)
                self.next_rain = inp.rainTime[self.rain_event, 1]         # forceSim.py:337
            if self.next_disp <= t and self.next_disp < inp.tEnd:         # model.py:285-292
                T = inp.tectTime
                self.disp_event = int(np.sum((T[:, 1] - t) <= 0))
                self.next_disp = T[self.disp_event, 1]
            if t >= self.next_display:                                    # model.py:447-469
                self.next_display += inp.tDisplay
            tStop = min(self.next_display, far, tEnd, self.next_disp, self.next_rain)   # model.py:484-486
            segs.append((float(tStop), self.rain_event, self.disp_event))
            t = tStop
        self.next_display += inp.tDisplay                                 # model.py:529-533 (udw == 0)
        return tEnd, segs


def sim_interval(simtime) -> np.ndarray:
    """`np.arange(0, simtime+1, simtime/4)` with python-2 integer division (bl_mcmc.py:94)."""
    if float(simtime).is_integer():
        step = int(simtime) // 4
        return np.arange(0, int(simtime) + 1, step)
    return np.arange(0, simtime + 1, simtime / 4)
