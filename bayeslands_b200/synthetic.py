"""BASELINE config 5: synthetic regular-grid uplift + rain landscape.

nx x ny nodes, dx = 100 m, z0 = 10 + U(0,1) (numpy default_rng(1234)), uplift = 1 mm/a x sin-bump that vanishes on
the edge (what bl_topogenr.checkUplift enforces, bl_topogenr.py:395-433), uniform rain ~ U(0.5, 2.5) and erodibility
~ U(3e-6, 7e-6) per proposal (default_rng(5678)), purely erosive (dep = 0), CDa = 0.8, CDm = 1.0 (mountain.xml), the
D8 stencil as the neighbour list (the reference has no regular-grid mode: documented deviation).
"""
from __future__ import annotations

import numpy as np

from . import config, lattice

XML = """<?xml version="1.0" encoding="UTF-8"?>
<badlands xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance">
  <grid><demfile>synthetic</demfile><boundary>fixed</boundary><resfactor>1</resfactor></grid>
  <time><start>0.</start><end>{T!r}</end><display>{disp!r}</display></time>
  <precipitation><climates>1</climates>
    <rain><rstart>0.</rstart><rend>{T!r}</rend><rval>1.</rval></rain></precipitation>
  <sp_law><dep>0</dep><m>0.5</m><n>1.</n><erodibility>5.e-6</erodibility></sp_law>
  <creep><caerial>8.e-1</caerial><cmarine>1.</cmarine></creep>
</badlands>
"""


def inputs(T: float = 50000.0):
    """Parsed XML-equivalent input of the synthetic config (tEnd = T years, 4 display intervals)."""
    return config.parse_xml(XML.format(T=float(T), disp=float(T) / 4.0), check_files=False)


def surface(nx: int, ny: int):
    return 10.0 + np.random.default_rng(1234).uniform(0.0, 1.0, (ny, nx))


def uplift_rate(nx: int, ny: int):
    """(ny, nx) m/a: 1 mm/a at the centre, exactly 0 on the edge."""
    sx = np.sin(np.pi * np.arange(nx) / (nx - 1)); sx[0] = sx[-1] = 0.0
    sy = np.sin(np.pi * np.arange(ny) / (ny - 1)); sy[0] = sy[-1] = 0.0
    return 1.0e-3 * sy[:, None] * sx[None, :]


def proposals(B: int, offset: int = 0):
    rng = np.random.default_rng(5678 + offset)
    return rng.uniform(0.5, 2.5, B), rng.uniform(3.0e-6, 7.0e-6, B)


def build(nx: int = 2048, ny: int = 2048, dx: float = 100.0, T: float = 50000.0):
    """(Lattice, Input) of the config; no (N, 20) tables are built, so 2048 x 2048 costs a few hundred MB of host RAM."""
    lat = lattice.regular(nx, ny, dx, surface(nx, ny), "fixed", disp_rate=uplift_rate(nx, ny))
    return lat, inputs(T)
