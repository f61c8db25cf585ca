"""brush-model dynamic volumes (R_DrawBrushModelVolumes M:63326-63499) on the device, vertex
and index buffers bit-identical to the oracle (pinned to the reference function in
tests/test_oracle_vs_reference.py::test_brush_model_volumes)."""
import numpy as np
import pytest

from berserkerquake2_b200 import _capi as c, host
from scenes import BrushModel, bits

pytestmark = pytest.mark.gpu


def test_brush_volumes_match_oracle(gpu_ctx, oracle):
    rng = np.random.RandomState(4)
    models = [BrushModel(8, 3, 0), BrushModel(5, 2, 1, open_edges=3), BrushModel(12, 4, 2), BrushModel(3, 1, 3, open_edges=1),
              BrushModel(64, 40, 4)]
    assert models[4].n_polys > 2500
    ids = [gpu_ctx.upload_brush(m.numverts, m.verts, m.neighbours, m.fence) for m in models]
    pairs, lits, want = [], [], []
    for k in range(40):
        mi = k % len(models); m = models[mi]
        origin = rng.randn(3) * 50
        angles = np.zeros(3) if k % 3 == 0 else rng.uniform(-180, 180, 3)
        light_world = origin + rng.randn(3) * 120
        lm = host.point_to_model(light_world, origin, angles)
        lit = m.lit(lm)
        if k == 7: lit[:] = 1
        if k == 8: lit[:] = 0
        pairs.append((ids[mi], tuple(lm), np.float32(32) * np.float32(300.0)))
        lits.append(lit)
        want.append(oracle.brush_volume(m, lit, lm, 300.0))
    got = gpu_ctx.brush_volumes(np.array(pairs, c.BRUSH_PAIR_DTYPE), np.concatenate(lits))
    for k, ((vc, ic), (wv, wi)) in enumerate(zip(got, want)):
        assert np.array_equal(bits(vc), bits(wv)), k
        assert np.array_equal(ic, wi), k
    assert sum(len(w[1]) for w in want) > 50000
    # a call without the large brush, and brushes of a few hundred polygons
    small = [k for k in range(40) if k % len(models) != 4]
    got_s = gpu_ctx.brush_volumes(np.array([pairs[k] for k in small], c.BRUSH_PAIR_DTYPE), np.concatenate([lits[k] for k in small]))
    for (vc, ic), k in zip(got_s, small):
        assert np.array_equal(bits(vc), bits(want[k][0])) and np.array_equal(ic, want[k][1]), k
    for slices, bands in ((16, 12), (20, 12), (28, 8), (36, 7)):
        m = BrushModel(slices, bands, 11)
        bid = gpu_ctx.upload_brush(m.numverts, m.verts, m.neighbours, m.fence)
        lm = np.array([60.0, -45.0, 30.0], np.float32)
        lit = m.lit(lm)
        (vc, ic), = gpu_ctx.brush_volumes(np.array([(bid, tuple(lm), np.float32(9600.0))], c.BRUSH_PAIR_DTYPE), lit)
        wv, wi = oracle.brush_volume(m, lit, lm, 300.0)
        assert np.array_equal(bits(vc), bits(wv)) and np.array_equal(ic, wi), (slices, bands, m.n_polys, int(m.first[-1]))
