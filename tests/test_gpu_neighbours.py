"""the neighbour tables built on the device, bit-identical to the reference's builders
(findneighbour M:61782-61824 under M:51643-51653; R_BuildTriangleNeighbors M:68646-68696) -- here against the
oracle, which tests/test_oracle_vs_reference.py pins to the reference object code on the same meshes, and
against the golden tables the reference produced."""
import numpy as np
import pytest

from berserkerquake2_b200 import synth
from test_golden import G

pytestmark = pytest.mark.gpu


def vandalise(tris, rng, n):
    T, V = len(tris), int(tris[:, :3].max()) + 1
    for _ in range(n):
        k, a = rng.randint(5), rng.randint(T)
        if k == 0: tris[a, :3] = tris[rng.randint(T), :3]                 # duplicate triangle: 3-triangle edges
        elif k == 1: tris[a, rng.randint(3)] = rng.randint(V)             # open edges
        elif k == 2: tris[a, :3] = tris[a, [0, 1, 0]]                     # two edges with the same vertex pair
        elif k == 3: tris[a, :3] = tris[rng.randint(T), [2, 1, 0]]        # flipped copy
        else: tris[a, :3] = tris[a, 0]                                    # fully degenerate (x, x, x)


@pytest.mark.parametrize("shape", [(16, 17), (8, 5), (5, 3), (3, 2), (24, 13), (62, 34)])
def test_gpu_neighbours_match_oracle(gpu_ctx, oracle, shape):
    rng = np.random.RandomState(shape[0] * 31 + shape[1])
    for trial in range(8):
        blob = synth.md2(shape[0], shape[1], 1, trial)
        tris = synth.md2_tris(blob)
        vandalise(tris, rng, trial * 3)
        assert np.array_equal(gpu_ctx.gpu_md2_neighbours(blob), oracle.md2_neighbours(blob)), trial
        ix = tris[:, :3].astype(np.uint16)
        assert np.array_equal(gpu_ctx.gpu_md3_neighbours(ix), oracle.md3_neighbours(ix)), trial


def test_gpu_neighbours_md3_max_size(gpu_ctx, oracle):
    verts, idx, st = synth.md3_mesh(60, 69, 1, 0)                  # V = 4082, T = 8160: just under the MD3 mesh limits
    assert len(idx) == 8160 and verts.shape[1] == 4082
    rng = np.random.RandomState(5)
    idx = idx.copy()
    for _ in range(20):
        idx[rng.randint(8160)] = idx[rng.randint(8160)]
    assert np.array_equal(gpu_ctx.gpu_md3_neighbours(idx), oracle.md3_neighbours(idx))


def test_gpu_neighbours_golden(gpu_ctx):
    blob = G["md2_broken/blob"]
    assert np.array_equal(gpu_ctx.gpu_md2_neighbours(blob), G["md2_broken/neighbours"])
    ix = synth.md2_tris(blob)[:, :3].astype(np.uint16)
    assert np.array_equal(gpu_ctx.gpu_md3_neighbours(ix), G["md2_broken/md3style_neighbours"])
    for k in range(3):
        assert np.array_equal(gpu_ctx.gpu_md3_neighbours(G[f"md3/mesh{k}/indexes"]), G[f"md3/mesh{k}/neighbours"])
    assert np.array_equal(gpu_ctx.gpu_md2_neighbours(synth.md2(16, 17, 198, 0)), G["md2_c1/neighbours"])


def test_upload_with_device_built_neighbours(gpu_ctx, oracle):
    """bq2_upload_md2 / bq2_upload_md3 with neighbours == NULL + the build flag: the whole path with no host
    precompute gives the same index buffers."""
    import berserkerquake2_b200 as bq2
    from scenes import Md2Model, Md3Model, world_scene, bits
    m = Md2Model(oracle, 12, 9, frames=4, seed=9)
    mid = gpu_ctx.upload_md2(m.blob, "gpu")
    md3 = Md3Model(oracle, 10, 7, frames=4, num_meshes=2, seed=9)
    mid3 = gpu_ctx.upload_md3(md3.frame_translate, [dict(m_, neighbours="gpu") for m_ in md3.meshes])
    we, wl, pe, pl = world_scene(4, 2, 4, seed=3)
    we["model"] = [mid, mid3, mid, mid3]
    ents, pairs = gpu_ctx.build_records(we, wl, pe, pl)
    res = gpu_ctx.frame(ents, pairs)
    slot = 0
    for p in range(len(pe)):
        e = int(pe[p]); W = we[e]
        if int(W["model"]) == mid:
            w = oracle.md2_pair(m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0, idx=m.idx)
            assert np.array_equal(res.indices(slot), w["indices"]); slot += 1
        else:
            for mesh in md3.meshes:
                w = oracle.md3_mesh_pair(mesh["verts"], mesh["indexes"], mesh["neighbours"], md3.frame_translate,
                                         int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                         W["oldorigin"], W["angles"], wl["origin"][pl[p]], 300.0)
                assert np.array_equal(res.indices(slot), w["indices"]); slot += 1
