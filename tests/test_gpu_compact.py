"""Compact shared-memory layout (csrc/bl_fastc.cuh: 8 bytes per node, two landscapes per SM) against the CPU oracle.

Same bar as test_gpu_parity.py: every stage array bit-exact, step by step and after full runs, at every CTA size the layout
can be forced on, including the generic fallback it takes when the receiver tree is deeper than its tables.
"""
import numpy as np
import pytest

from conftest import CRATER_FAST_COORDS
from test_gpu_parity import FP_FIELDS, INT_FIELDS, PROPOSALS

pytestmark = pytest.mark.gpu


def _engine(crater_fast, B, **kw):
    from bayeslands_b200.engine import Engine
    d, inp, m = crater_fast
    return Engine(m, inp, B, real_elev=d["final_elev"], erdp_coords=CRATER_FAST_COORDS, **kw)


def test_crater_layout_follows_the_batch_size(crater_fast):
    """More landscapes than SMs: two compact 512-thread CTAs per SM.  Up to one wave: one 1024-thread CTA per SM in the
    17-byte layout (a lone landscape finishes a time step sooner that way)."""
    import torch
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    info = _engine(crater_fast, 2 * nsm).info()
    assert info["path"] == "compact" and info["tpb"] == 512 and info["ctas_per_sm"] >= 2, info
    info = _engine(crater_fast, 2).info()
    assert info["path"] == "shared" and info["tpb"] == 1024, info
    assert _engine(crater_fast, 2 * nsm, compact=False).info()["path"] == "shared"


@pytest.mark.parametrize("shuffle", [0, 1])
def test_compact_stepwise_all_stages_bit_exact(crater_fast, shuffle):
    from oracle import oracle
    d, inp, m = crater_fast
    B = len(PROPOSALS)
    eng = _engine(crater_fast, B, shuffle_mode=shuffle, tpb=512, compact=True)
    assert eng.info()["path"] == "compact"
    rain = np.array([p[0] for p in PROPOSALS]); erod = np.array([p[1] for p in PROPOSALS])
    seeds = np.arange(B, dtype=np.int64) + 17
    eng.reset(rain, erod, seeds)
    oms = []
    for b in range(B):
        om = oracle.OracleModel(m, inp, pow_mode=1, shuffle_mode=shuffle, fill_mode=0, seed=int(seeds[b]))
        om.reset(rain[b], erod[b], seed=int(seeds[b]))
        oms.append(om)
    for step in range(12):
        eng.step(15000.0)
        sc = eng.scalars()
        for b, om in enumerate(oms):
            om.streamflow(); om.sediment_flux(15000.0)
            for f in INT_FIELDS:
                g = eng.field(f, b); o = getattr(om, f)
                assert np.array_equal(g, o), "step %d proposal %d %s: %d mismatches" % (step, b, f, (g != o).sum())
            for f in FP_FIELDS:
                g = eng.field(f, b); o = getattr(om, f)
                assert np.array_equal(g, o), "step %d proposal %d %s: max abs diff %g" % (step, b, f, np.abs(g - o).max())
            assert sc["tNow"][b] == om.tNow and sc["last_dt"][b] == om.last_dt
            assert sc["status"][b] == 0


@pytest.mark.parametrize("tpb,debug", [(256, 0), (384, 0), (768, 0), (1024, 0), (512, 8), (512, 16), (512, 32), (512, 256)])
def test_compact_every_cta_size_and_deep_tree_fallback(crater_fast, tpb, debug):
    """`compact=True` forces the layout at any CTA size; debug bit 8 lowers the depth limit of its breadth-first tables to
    12 levels, so that every step takes the generic discharge / stream-power fallback."""
    from oracle import oracle
    d, inp, m = crater_fast
    eng = _engine(crater_fast, 2, tpb=tpb, compact=True, debug=debug)
    assert eng.info()["path"] == "compact"
    rain = np.array([0.9, 2.2]); erod = np.array([6e-5, 4e-5]); seeds = np.array([8, 9], dtype=np.int64)
    eng.reset(rain, erod, seeds)
    oms = [oracle.OracleModel(m, inp, seed=int(s)) for s in seeds]
    for b, om in enumerate(oms):
        om.reset(rain[b], erod[b], seed=int(seeds[b]))
    for step in range(5):
        eng.step(15000.0)
        sc = eng.scalars()
        for b, om in enumerate(oms):
            om.streamflow(); om.sediment_flux(15000.0)
            for f in INT_FIELDS + FP_FIELDS:
                assert np.array_equal(eng.field(f, b), getattr(om, f)), (tpb, step, b, f)
            assert sc["tNow"][b] == om.tNow and sc["status"][b] == 0


def test_compact_full_run_equals_shared_layout(crater_fast):
    """Full blackBox runs of 24 proposals: the compact layout and the 17-bytes-per-node layout give the same bits."""
    from bayeslands_b200 import hostparams
    B = 24
    rng = np.random.default_rng(5)
    rain = rng.uniform(0.0, 3.0, B); erod = rng.uniform(3e-5, 7e-5, B); seeds = np.arange(B, dtype=np.int64) + 100
    times = hostparams.sim_interval(15000)
    out = []
    for compact in (True, False):
        eng = _engine(crater_fast, B, compact=compact)
        eng.reset(rain, erod, seeds)
        pts, sse, elev, erdp = eng.run_schedule(times, grids=True)
        sc = eng.scalars()
        assert (sc["status"] == 0).all()
        out.append((pts.cpu().numpy(), sse.cpu().numpy(), elev.cpu().numpy(), erdp.cpu().numpy(), sc["steps"].copy(),
                    np.stack([eng.field("z", b) for b in range(B)]), np.stack([eng.field("cumdiff", b) for b in range(B)])))
    a, b = out
    assert np.array_equal(a[4], b[4])
    for i in (0, 2, 3, 5, 6):
        assert np.array_equal(a[i], b[i]), i
    assert np.allclose(a[1], b[1], rtol=1e-11, atol=0)
