"""Edge cases of the C-ABI / engine on the GPU: degenerate batches, ragged run_to_time sequences, error paths."""
import ctypes as C

import numpy as np
import pytest

from conftest import CRATER_FAST_COORDS

pytestmark = pytest.mark.gpu


def test_single_proposal_and_ragged_run_to_time_sequence(crater_fast):
    """B = 1; run_to_time in uneven increments (and a no-op call) must land on the same state as one call,
    because every stop is reached exactly (buildFlux.py:175) - checked against the oracle doing the same calls."""
    from bayeslands_b200.engine import Engine
    from oracle import oracle
    d, inp, m = crater_fast
    eng = Engine(m, inp, 1)
    eng.reset(np.array([1.2]), np.array([6e-5]), np.array([9], dtype=np.int64))
    om = oracle.OracleModel(m, inp, seed=9); om.reset(1.2, 6e-5, seed=9)
    for T in (0.0, 100.0, 100.0, 2600.0, 5000.0, 5001.0, 9999.0):
        eng.run_to_time(T); om.run_to_time(T)
        sc = eng.scalars()
        assert sc["tNow"][0] == om.tNow == T and sc["steps"][0] == om.steps
        assert np.array_equal(eng.field("z", 0), om.z)


def test_checkpoint_at_time_zero_and_zero_rain(crater_fast):
    from bayeslands_b200.engine import Engine
    from oracle import oracle
    d, inp, m = crater_fast
    eng = Engine(m, inp, 2, real_elev=d["final_elev"], erdp_coords=CRATER_FAST_COORDS)
    eng.reset(np.array([0.0, 0.0]), np.array([3e-5, 7e-5]))
    pts, sse, elev, erdp = eng.run_schedule([0], grids=True)        # no time step at all: interpolation only
    om = oracle.OracleModel(m, inp)
    assert np.array_equal(elev[0, 0].cpu().numpy().reshape(123, 123), om.interpolate(m.elevation))
    assert float(erdp.abs().max()) == 0.0 and float(pts.abs().max()) == 0.0
    pts, sse, elev, erdp = eng.run_schedule([15000], grids=True)
    assert np.array_equal(elev[0].cpu().numpy(), elev[1].cpu().numpy())       # rain = 0: erodibility is irrelevant
    assert sse[0].item() == sse[1].item()


def test_engine_reuse_after_reset_is_deterministic(crater_fast):
    from bayeslands_b200.engine import Engine
    d, inp, m = crater_fast
    eng = Engine(m, inp, 3, real_elev=d["final_elev"], erdp_coords=CRATER_FAST_COORDS)
    rain = np.array([1.5, 2.2, 0.7]); erod = np.array([5e-5, 4e-5, 6e-5]); seeds = np.array([4, 5, 6], dtype=np.int64)
    outs = []
    for rep in range(2):
        eng.reset(rain, erod, seeds)
        pts, sse, _, _ = eng.run_schedule([0, 3750, 7500])
        outs.append((pts.cpu().numpy().copy(), sse.cpu().numpy().copy(), eng.field("z", 1)))
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
    # a different shuffle seed changes the stack order but, away from pit overflows, not the physics much
    eng.reset(rain, erod, seeds + 100)
    eng.run_schedule([0, 3750, 7500])
    assert np.abs(eng.field("z", 1) - outs[0][2]).max() < 1.0


def test_error_paths(crater_fast):
    import torch
    from bayeslands_b200 import _lib
    from bayeslands_b200.engine import Engine
    d, inp, m = crater_fast
    L = _lib.load()
    eng = Engine(m, inp, 1)
    # too many checkpoints per call
    with pytest.raises(ValueError):
        eng.run_schedule(list(range(0, 9000, 1000)))
    # bad arguments to the raw ABI
    assert L.bl_run(eng.ctx, 0, None, None, None, None, None, None, 0, None) < 0
    assert b"bl_run" in L.bl_last_error()
    assert L.bl_get_field(eng.ctx, 999, 0, C.c_void_p(1), None) < 0
    # workspace too small / asymmetric neighbour list are rejected at create time
    ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
    ctx = C.c_void_p()
    assert L.bl_ctx_create(C.byref(eng._m), C.byref(eng._p), 1, 0, C.c_void_p(ws.data_ptr()), 1024, C.byref(ctx)) == -3
    bad = eng._keep["ngb"].copy()
    j = bad[m.boundsPt + 5, 0]
    bad[j][bad[j] == m.boundsPt + 5] = bad[j][0] if bad[j][0] != m.boundsPt + 5 else bad[j][1]
    mm = _lib.BlMesh.from_buffer_copy(eng._m)
    mm.ngb = bad.ctypes.data_as(_lib.c_ip)
    big = torch.empty(int(L.bl_workspace_bytes(C.byref(eng._m), C.byref(eng._p), 1)), dtype=torch.uint8, device="cuda")
    assert L.bl_ctx_create(C.byref(mm), C.byref(eng._p), 1, 0, C.c_void_p(big.data_ptr()), big.numel(), C.byref(ctx)) == -1
    assert b"symmetric" in L.bl_last_error()
    with pytest.raises(ValueError):
        Engine(m, inp, 1, real_elev=np.zeros((5, 5)))
