"""The drop-in boundary, end to end: the reference's own R_DrawAliasShadowVolume (main.c:63713), UNMODIFIED,
linked with integration/bq2_dropin.c in front of it, so that the R_CalcAliasFrameLerp /
R_CalcAliasFrameLerpShadow / R_CalcAliasFrameShadowVolume it calls are the CUDA-backed ones
(oracle/_ref/libbq2ref_dropin.so, built by oracle/ref/build_ref.sh).  Its result globals are compared with the
golden vectors the stock reference produced, and with the stock reference library when that is present."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from scenes import bits
from test_golden import G, md2_case, md3_case, same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dropin():
    r = oracle_lib.Reference.load(oracle_lib.REF_DROPIN_SO)
    if r is None:
        pytest.skip("oracle/_ref/libbq2ref_dropin.so not built (needs /root/reference at build time)")
    r.lib.bq2_dropin_last_error.restype = C.c_char_p
    return r


@pytest.mark.parametrize("name", ["md2_small", "md2_c1"])
def test_reference_draw_function_runs_on_the_gpu(dropin, name):
    blob, nb, we, wl, pe, pl = md2_case(name)
    V, T, F = oracle_lib.md2_dims(blob)
    used = np.unique(oracle_lib.md2_idx(blob))
    before = dropin.lib.bq2_dropin_kernel_launches()
    for p in range(len(pe)):
        W = we[pe[p]]
        r = dropin.md2_pair(blob, nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                            W["oldorigin"], W["angles"], wl["origin"][pl[p]], float(wl["radius"][pl[p]]))
        assert r is not None, dropin.lib.bq2_dropin_last_error()
        assert same(r["lerped"], G[f"{name}/pair{p}/lerped"])                       # tempVertexArray
        assert same(r["extruded"][used], G[f"{name}/pair{p}/extruded"][used])       # extrudedVerts
        assert np.array_equal(r["viss"], G[f"{name}/pair{p}/viss"])                 # viss[]
        assert same(r["tris_verts"], G[f"{name}/pair{p}/tris_verts"])               # TrisVerts / TrisCounter
        assert same(r["light_local"], G[f"{name}/pair{p}/light_local"])
    assert dropin.lib.bq2_dropin_kernel_launches() - before == len(pe)              # one fused launch per call pair


def test_lerp_entry_point_alone(dropin, ref):
    """R_CalcAliasFrameLerp (ambient pass, M:50303) incl. the shell variant, against the stock reference."""
    blob, nb, we, wl, pe, pl = md2_case("md2_small")
    for flags in (0, 0x400, 0x800 | 0x4):
        a = dropin.md2_lerp(blob, 2, 1, 0.375, (1, 2, 3), (2, 2, 2.5), (0, 25, 0), flags=flags)
        b = ref.md2_lerp(blob, 2, 1, 0.375, (1, 2, 3), (2, 2, 2.5), (0, 25, 0), flags=flags)
        assert same(a, b), hex(flags)


def test_flag_filter_still_the_references(dropin):
    blob, nb, we, wl, pe, pl = md2_case("md2_small")
    z = np.zeros(3, np.float32)
    assert dropin.md2_pair(blob, nb, 1, 0, 0.5, z, z, z, (100, 50, 60), 300.0, flags=0x20) is None     # RF_TRANSLUCENT
    assert dropin.md2_pair(blob, nb, 1, 0, 0.5, z, z, z, (100, 50, 60), 300.0, flags=0) is not None


def test_md3_per_mesh_body(dropin):
    """bq2_dropin_md3_mesh = the compute block of R_DrawMD3AliasShadowVolume (M:63885-63991)."""
    ft, meshes, we, wl = md3_case()
    for p in range(6):
        W = we[p // 2]
        done, res = dropin.dropin_md3(ft, meshes, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"],
                                      W["oldorigin"], W["angles"], wl["origin"][p], 300.0)
        assert done == 2 and res[1] is None, dropin.lib.bq2_dropin_last_error()
        for k in (0, 2):
            for key in ("lerped", "extruded", "tris_verts"):
                assert same(res[k][key], G[f"md3/pair{p}/mesh{k}/{key}"]), (p, k, key)
            assert np.array_equal(res[k]["viss"], G[f"md3/pair{p}/mesh{k}/viss"])
