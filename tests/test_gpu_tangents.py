"""the tangent-space anorm bytes computed on the device, bit-identical to the oracle's
restatement of the loader loops (M:51661-51804, M:50993-51014), whose every inner function is pinned to the
reference object code (tests/test_oracle_vs_reference.py), and to the golden byte arrays that went through the
reference's own light-lerp functions."""
import numpy as np
import pytest

from berserkerquake2_b200 import synth
from scenes import SMOOTH_COSINE
from test_golden import G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,frames", [((10, 7), 4), ((16, 17), 6), ((5, 3), 3), ((62, 34), 2)])
def test_md2_smooth_and_flat(gpu_ctx, oracle, shape, frames):
    blob = synth.md2(shape[0], shape[1], frames, 77)
    st = synth.md2_st(blob)
    t, b = oracle.md2_tangents_smooth(blob, st, SMOOTH_COSINE)
    _, gt, gb = gpu_ctx.gpu_md2_tangents(blob, st, SMOOTH_COSINE, smooth=True)
    assert np.array_equal(gt, t) and np.array_equal(gb, b)
    n, t, b = oracle.md2_tangents_flat(blob, st)
    gn, gt, gb = gpu_ctx.gpu_md2_tangents(blob, st, SMOOTH_COSINE, smooth=False)
    assert np.array_equal(gn, n) and np.array_equal(gt, t) and np.array_equal(gb, b)


def test_md2_coincident_vertices_and_degenerate_triangles(gpu_ctx, oracle):
    """Vertices sharing byte coordinates (merged when their anorms are within r_smoothangle, else not), repeated
    corners, zero-area triangles (division by zero in VecsForTris -> NaN -> index 0)."""
    rng = np.random.RandomState(3)
    for trial in range(6):
        blob = synth.md2(12, 9, 3, 100 + trial)
        V = int(blob[:68].view("<i4")[6])
        tris = synth.md2_tris(blob)
        for f in range(3):
            fv = synth.md2_frame_verts(blob, f)
            for _ in range(6 + trial):                      # copy positions; sometimes the normal index too
                a, b = rng.randint(V, size=2)
                fv[b, :3] = fv[a, :3]
                if rng.rand() < 0.5:
                    fv[b, 3] = fv[a, 3]
            a, b, c_ = rng.randint(V, size=3)               # a three-vertex coincident group
            fv[b, :] = fv[a, :]; fv[c_, :] = fv[a, :]
        tris[5, 0:3] = tris[5, [0, 0, 2]]; tris[9, 0:3] = tris[9, 0]
        st = synth.md2_st(blob)
        t, b = oracle.md2_tangents_smooth(blob, st, SMOOTH_COSINE)
        _, gt, gb = gpu_ctx.gpu_md2_tangents(blob, st, SMOOTH_COSINE, smooth=True)
        assert np.array_equal(gt, t) and np.array_equal(gb, b), trial
        n, t, b = oracle.md2_tangents_flat(blob, st)
        gn, gt, gb = gpu_ctx.gpu_md2_tangents(blob, st, SMOOTH_COSINE, smooth=False)
        assert np.array_equal(gn, n) and np.array_equal(gt, t) and np.array_equal(gb, b), trial


@pytest.mark.parametrize("shape", [(12, 9), (32, 33), (3, 2)])
def test_md3(gpu_ctx, oracle, shape):
    verts, idx, st = synth.md3_mesh(shape[0], shape[1], 5, 9)
    want = verts.copy()
    oracle.md3_tangents(want, st, idx)
    got = verts.copy()
    gpu_ctx.gpu_md3_tangents(got, st, idx)
    assert np.array_equal(got["tangent"], want["tangent"]) and np.array_equal(got["binormal"], want["binormal"])
    assert np.array_equal(got["normal"], verts["normal"]) and np.array_equal(got["point"], verts["point"])
    assert (want["tangent"] != 0).any()


def test_golden_byte_arrays(gpu_ctx):
    """The byte arrays stored in the golden file were fed to the reference's GL_DrawAliasFrameLerpLightARB[_vp]."""
    for kind in ("smooth", "flat"):
        blob = G[f"light/md2_{kind}/blob"]
        st = synth.md2_st(blob)
        n, t, b = gpu_ctx.gpu_md2_tangents(blob, st, SMOOTH_COSINE, smooth=(kind == "smooth"))
        assert np.array_equal(t, G[f"light/md2_{kind}/tangents"]) and np.array_equal(b, G[f"light/md2_{kind}/binormals"])
        if kind == "flat":
            assert np.array_equal(n, G["light/md2_flat/normals"])
    v = np.ascontiguousarray(G["md3/mesh0/verts"]).view(synth.MD3_VERTEX).reshape(G["md3/mesh0/verts"].shape[:2]).copy()
    want_t, want_b = v["tangent"].copy(), v["binormal"].copy()
    v["tangent"] = 0; v["binormal"] = 0
    _, _, st = synth.md3_mesh(12, 9, 1, 0)
    gpu_ctx.gpu_md3_tangents(v, st, G["md3/mesh0/indexes"])
    assert np.array_equal(v["tangent"], want_t) and np.array_equal(v["binormal"], want_b)
