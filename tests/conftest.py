import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def crater_fast():
    """(data, Input, MeshArrays) of Examples/crater_fast, from the committed fixture data/crater_fast.npz."""
    import numpy as np
    from bayeslands_b200 import config, mesh
    d = np.load(os.path.join(ROOT, "data", "crater_fast.npz"))
    inp = config.parse_xml(str(d["xml"]))
    m = mesh.build_from_dem(d["dem_xyz"], inp.btype, inp.Afactor)
    return d, inp, m


CRATER_FAST_COORDS = [[60, 60], [72, 66], [85, 73], [90, 75], [44, 86], [100, 80], [88, 69], [79, 91], [96, 77], [42, 49]]


@pytest.fixture(scope="session")
def etopo_fast():
    """(data, Input, MeshArrays) of Examples/etopo_fast (sea-level curve, slope boundary, marine deposition)."""
    import numpy as np
    from bayeslands_b200 import config, mesh
    d = np.load(os.path.join(ROOT, "data", "etopo_fast.npz"))
    inp = config.parse_xml(str(d["xml"]))
    m = mesh.build_from_dem(d["dem_xyz"], inp.btype, inp.Afactor)
    return d, inp, m


ETOPO_FAST_COORDS = [[42, 10], [39, 8], [75, 51], [59, 13], [40, 5], [6, 20], [14, 66], [4, 40], [68, 40], [72, 44]]
