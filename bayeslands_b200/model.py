"""`pyBadlands.model.Model` look-alike (reference: `pyBadlands/model.py:18-562`) for the hot path.

Keeps the surface `bl_mcmc.blackBox` uses (`bl_mcmc.py:126-152`):

    model = Model()
    model.load_xml(run_nb, filename, verbose=False, muted=False)   -> (initial_erod, initial_rain)
    model.input.SPLero = erod ; model.flow.erodibility.fill(erod) ; model.force.rainVal[:] = rain
    model.run_to_time(tEnd, profile=False, verbose=False, muted=False)
    model.FVmesh.node_coords ; model.elevation ; model.cumdiff ; model.cumhill ; model.tNow

Differences: one Model drives B >= 1 proposals (`load_xml(..., batch=B)`); `flow.erodibility` and
`force.rainVal` keep their reference shapes for B = 1 and gain a leading batch axis otherwise; the device
state is created lazily at the first `run_to_time` so that the parameter pokes between `load_xml` and
`run_to_time` are honoured exactly like the reference (rain is loaded inside the time loop, model.py:275-280).
Everything `muted=False` would write (HDF5/XDMF checkpoints) is out of scope.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import config, mesh as meshmod
from .engine import Engine


class _Flow(object):
    """The attributes of `flowNetwork` the drivers poke (flow/flowNetwork.py:40-115)."""

    def __init__(self, model):
        self._m = model
        self.erodibility = None
        self.xycoords = None
        self.parentIDs = None
        self.borders = self.borders2 = self.insideIDs = None
        self.mindt = 1.0
        self.spl = True
        self.depo = 1

    def _field(self, name, b=0):
        return self._m._field(name, b)

    receivers = property(lambda s: s._field("rcv"))
    receivers1 = property(lambda s: s._field("rcv1"))
    stack = property(lambda s: s._field("stack"))
    stack1 = property(lambda s: s._field("stack1"))
    discharge = property(lambda s: s._field("Q"))
    maxh = property(lambda s: s._field("maxh"))
    pitID = property(lambda s: s._field("pitID"))
    pitVolume = property(lambda s: s._field("pitVol"))
    pitDrain = property(lambda s: s._field("pitDrain"))
    allDrain = property(lambda s: s._field("allDrain"))


class _Force(object):
    """`forceSim` subset (forcing/forceSim.py:226-347)."""

    def __init__(self, inp, sealevel):
        self.rainVal = np.array(inp.rainVal, dtype=np.float64)
        self.T_rain = inp.rainTime
        self.T_disp = inp.tectTime
        self.sea0 = inp.seapos
        self.seacurve = sealevel
        self.sealevel = inp.seapos

    def getSea(self, time):
        if self.seacurve is None:
            self.sealevel = self.sea0
        else:
            t = min(max(time, self.seacurve[:, 0].min()), self.seacurve[:, 0].max())
            self.sealevel = float(np.interp(t, self.seacurve[:, 0], self.seacurve[:, 1]))


class Model(object):
    """State object for the pyBadlands model (B200 engine underneath)."""

    def __init__(self):
        self.tNow = 0.0
        self.outputStep = 0
        self.disp = None
        self.applyDisp = False
        self.simStarted = False
        self.initial_erod = []
        self.initial_rain = []
        self.input = None
        self._engine: Optional[Engine] = None
        self._batch = 1

    # ---------------------------------------------------------------------------------------------
    def load_xml(self, run_nb, filename, verbose=False, muted=False, batch=1, device=0, base_dir=None,
                 dem_xyz=None, sealevel=None, mesh=None, seed=None):
        """Load an XML configuration file (model.py:46-92).  `filename` may also be the XML text."""
        self.input = config.parse_xml(filename, base_dir=base_dir)
        self.tNow = self.input.tStart
        self._batch = int(batch)
        self._device = int(device)
        self._seed = seed
        if mesh is None:
            if dem_xyz is None:
                if not self.input.demfile or not os.path.isfile(self.input.demfile):
                    raise ValueError("DEM file is missing or the given path is incorrect.")
                dem_xyz = np.loadtxt(self.input.demfile)
            mesh = meshmod.build_from_dem(dem_xyz, self.input.btype, self.input.Afactor)
        if sealevel is None and self.input.seafile is not None:
            sealevel = np.loadtxt(self.input.seafile)
        self.mesh = mesh
        self.recGrid = mesh.recGrid
        self.FVmesh = mesh.FVmesh
        self.totPts = mesh.N
        self.inIDs = np.arange(mesh.boundsPt, mesh.N)
        self.inGIDs = self.lGIDs = np.arange(mesh.N)
        self.force = _Force(self.input, sealevel)
        self.flow = _Flow(self)
        shape = (mesh.N,) if self._batch == 1 else (self._batch, mesh.N)
        self.flow.erodibility = np.full(shape, self.input.SPLero)          # model.py:117-118
        self.flow.xycoords = mesh.FVmesh.node_coords[:, :2]
        self.flow.parentIDs = mesh.parentIDs
        self.flow.borders, self.flow.borders2, self.flow.insideIDs = mesh.borders, mesh.borders2, mesh.insideIDs
        self.flow.mindt = self.input.minDT
        self.flow.depo = self.input.depo
        if self._batch > 1:
            self.force.rainVal = np.tile(self.force.rainVal, (self._batch, 1))
        self._sealevel = sealevel
        self._engine = None
        self.simStarted = False
        return self.initial_erod, self.initial_rain

    # ---------------------------------------------------------------------------------------------
    def _start(self):
        inp = self.input
        rv = np.asarray(self.force.rainVal, dtype=np.float64)
        rain = rv.reshape(self._batch, -1)
        if not np.all(rain == rain[:, :1]):
            raise ValueError("per-event rain values differ: only one uniform rain value per proposal is supported")
        ero = np.asarray(self.flow.erodibility, dtype=np.float64).reshape(self._batch, -1)
        if not np.all(ero == ero[:, :1]):
            raise ValueError("flow.erodibility must be uniform per proposal (erodibility maps are out of scope)")
        self._engine = Engine(self.mesh, inp, self._batch, device=self._device, sealevel=self._sealevel)
        seeds = None if self._seed is None else (np.arange(self._batch, dtype=np.int64) + int(self._seed))
        self._engine.reset(rain[:, 0].copy(), ero[:, 0].copy(), seeds)
        self.simStarted = True

    def run_to_time(self, tEnd, profile=False, verbose=False, muted=False):
        """Run the simulation to `tEnd` years (model.py:218-558)."""
        assert self.input is not None, "DEM file has not been loaded. Configure one in your XML file or call the build_mesh function."
        if tEnd > self.input.tEnd:
            print("Specified end time is greater than the one used in the XML input file and has been adjusted!")
            tEnd = self.input.tEnd
        if not self.simStarted:
            self._start()
        self._engine.run_to_time(float(tEnd))
        self.tNow = self._engine.tNow
        self.force.getSea(self.tNow)
        self.outputStep += 1

    def ncpus(self):
        return 1

    # ---------------------------------------------------------------------------------------------
    def _field(self, name, b=0):
        if self._engine is None:
            if name == "z":
                return np.array(self.mesh.elevation)
            if name in ("cumdiff", "cumhill"):
                return np.zeros(self.mesh.N)
            raise AttributeError("%s is available after the first run_to_time" % name)
        if self._batch == 1:
            return self._engine.field(name, 0)
        return np.stack([self._engine.field(name, i) for i in range(self._batch)])

    elevation = property(lambda s: s._field("z"))
    cumdiff = property(lambda s: s._field("cumdiff"))
    cumhill = property(lambda s: s._field("cumhill"))
    fillH = property(lambda s: s._field("fillH"))

    @property
    def steps(self):
        return None if self._engine is None else self._engine.scalars()["steps"]
