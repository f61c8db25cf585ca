"""the caller loop R_DrawEntsShadowVolumes (main.c:64041-64133) with the early-outs of the volume
builders (main.c:63722-63727, 63811-63816).  The engine's loop is run WHOLE by the reference object code over a
cl_lightvisedicts list that holds every combination of model kind, entity flags and model flags, for every combination
of r_mirror, the five cvars and the three passes; the product's filter (bq2_host_entity_casts, the same masks the world
path and bq2_build_records_opts use) must name the same entities in the same order."""
import ctypes as C
import itertools

import numpy as np
import pytest

import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import _capi as c, synth, host
from scenes import Md2Model, world_scene

RF = dict(VIEWER=0x2, WEAPON=0x4, FULLBRIGHT=0x8, TRANSLUCENT=0x20, BEAM=0x80, SHELL_RED=0x400, SHELL_GREEN=0x800,
          SHELL_BLUE=0x1000, NOSELF=0x2000, NOCAST=0x80000, DISTORT=0x10000000)
ENT_FLAGS = [0, RF["VIEWER"], RF["WEAPON"], RF["FULLBRIGHT"], RF["TRANSLUCENT"], RF["BEAM"], RF["SHELL_RED"], RF["SHELL_GREEN"],
             RF["SHELL_BLUE"], RF["NOCAST"], RF["DISTORT"], RF["VIEWER"] | RF["NOCAST"], 0x40 | 0x100]     # last: bits the path ignores
MODEL_FLAGS = [0, RF["NOSELF"], RF["NOCAST"], RF["DISTORT"], RF["NOSELF"] | RF["DISTORT"], 0x4000000]      # last: RF_SMOOTHVECS (ignored)


def entity_list():
    kinds, ef, mf = [], [], []
    for k, e, m in itertools.product((0, 1, 2), ENT_FLAGS, MODEL_FLAGS):
        kinds.append(k); ef.append(e); mf.append(m)
    return np.array(kinds, np.int32), np.array(ef, np.int32), np.array(mf, np.int32)


def cvar_combinations():
    for mirror, ms, mm, ps, sh, de, pas in itertools.product((0, 1), (0.0, 1.0, 2.0), (0.0, 1.0), (0.0, 1.0), (0.0, 1.0, 2.0),
                                                             (0.0, 1.0), (c.PASS_ALL, c.PASS_SELF, c.PASS_NOSELF)):
        yield mirror, ms, mm, ps, sh, de, pas


def reference_order(ref, blob, nb, kinds, ef, mf, mirror, ms, mm, ps, sh, de, pas):
    fn = ref.lib.bq2ref_ents_shadow_volumes
    fn.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_float] * 5 + [C.c_int, C.c_int, C.c_void_p]
    order = np.zeros(len(kinds), np.int32)
    n = fn(len(kinds), c.ptr(kinds), c.ptr(ef), c.ptr(mf), c.ptr(blob), c.ptr(nb), mirror, ms, mm, ps, sh, de,
           1 if pas == c.PASS_SELF else 0, 1 if pas == c.PASS_ALL else 0, c.ptr(order))
    assert n >= 0
    return order[:n].tolist()


def test_filter_matches_the_engines_loop_for_every_cvar_combination(ref, oracle):
    m = Md2Model(oracle, 6, 5, frames=2, seed=1)
    kinds, ef, mf = entity_list()
    assert len(kinds) <= 256
    combos = 0
    for mirror, ms, mm, ps, sh, de, pas in cvar_combinations():
        want = reference_order(ref, m.blob, np.ascontiguousarray(m.nb, np.int32), kinds, ef, mf, mirror, ms, mm, ps, sh, de, pas)
        o = bq2.Context.render_opts(r_mirror=mirror, r_mirror_shadows=ms, r_mirror_models=mm, r_playershadow=ps, r_shadows=sh,
                                    r_drawentities=de, pass_=pas)
        got = [i for i in range(len(kinds))
               if c.lib.bq2_host_entity_casts(int(ef[i]) & 0xffffffff, int(mf[i]) & 0xffffffff, 1 if kinds[i] == 2 else 0, C.byref(o))]
        assert got == want, (mirror, ms, mm, ps, sh, de, pas, sorted(set(got) ^ set(want))[:8])
        combos += 1
    assert combos == 432
    # NULL options = the engine's defaults
    o = bq2.Context.render_opts()
    assert (o.r_shadows, o.r_playershadow, o.r_mirror_shadows, o.r_mirror_models, o.r_drawentities, o.r_mirror, o.pass_) == (2, 0, 2, 1, 1, 0, 0)
    for i in range(len(kinds)):
        k = 1 if kinds[i] == 2 else 0
        assert c.lib.bq2_host_entity_casts(int(ef[i]), int(mf[i]), k, None) == c.lib.bq2_host_entity_casts(int(ef[i]), int(mf[i]), k, C.byref(o))
    # the older two-condition helper is the builder's own early-out (M:63722-63727) and nothing else
    assert host.casts_shadow(RF["VIEWER"], False, 1.0, 2.0) and not host.casts_shadow(RF["VIEWER"], False, 1.0, 1.0)
    assert host.casts_shadow(RF["VIEWER"], True, 0.0, 0.0) and not host.casts_shadow(RF["BEAM"], True, 1.0, 2.0)


@pytest.mark.gpu
def test_batched_routes_produce_the_engines_pair_set(ref, oracle, gpu_ctx):
    """Both batched routes under render options + model flags: a pair gets a volume exactly when the engine's loop draws
    one for its entity (alias kinds; brush models go through bq2_brush_volumes), and the volumes are the usual ones."""
    kinds, ef, mf = entity_list()
    sel = np.nonzero(kinds == 0)[0]                                # one MD2 entity per (entity flags, model flags)
    m = Md2Model(oracle, 8, 7, frames=3, seed=2)
    mids = {}
    for f in MODEL_FLAGS:
        mids[f] = m.upload(gpu_ctx)
        gpu_ctx.set_model_rf_flags(mids[f], f)
    n = len(sel)
    we, wl, pe, pl = world_scene(n, 2, 3, seed=5)
    we["rf_flags"] = ef[sel].astype(np.uint32); we["model"] = [mids[int(f)] for f in mf[sel]]
    shell = (we["rf_flags"] & (RF["SHELL_RED"] | RF["SHELL_GREEN"] | RF["SHELL_BLUE"])) != 0
    picks = [(0, 2.0, 1.0, 0.0, 2.0, 1.0, c.PASS_ALL), (0, 2.0, 1.0, 1.0, 2.0, 1.0, c.PASS_SELF), (0, 2.0, 1.0, 1.0, 2.0, 1.0, c.PASS_NOSELF),
             (1, 2.0, 1.0, 0.0, 0.0, 1.0, c.PASS_ALL), (1, 1.0, 1.0, 0.0, 2.0, 1.0, c.PASS_ALL), (1, 2.0, 0.0, 0.0, 2.0, 1.0, c.PASS_SELF),
             (0, 2.0, 1.0, 0.0, 1.0, 1.0, c.PASS_ALL), (0, 2.0, 1.0, 0.0, 2.0, 0.0, c.PASS_ALL)]
    seen_some = False
    for mirror, ms, mm, ps, sh, de, pas in picks:
        want = set(reference_order(ref, m.blob, np.ascontiguousarray(m.nb, np.int32), kinds[sel].copy(), ef[sel].copy(), mf[sel].copy(),
                                   mirror, ms, mm, ps, sh, de, pas))
        o = bq2.Context.render_opts(r_mirror=mirror, r_mirror_shadows=ms, r_mirror_models=mm, r_playershadow=ps, r_shadows=sh,
                                    r_drawentities=de, pass_=pas)
        res = gpu_ctx.world_frame(we, wl, pe, pl, opts=o)                                # records built on the device
        casting = set(int(pe[p]) for p in range(len(pe)) if res.index_counts[p] > 0)
        assert casting == want, (mirror, ms, mm, ps, sh, de, pas)
        assert all(res.index_counts[p] > 0 for p in range(len(pe)) if int(pe[p]) in want)     # both lights of each
        world_indices = {p: res.indices(p) for p in range(len(pe)) if int(pe[p]) in want}     # (before the next frame reuses the arrays)
        ents, pairs = gpu_ctx.build_records(we, wl, pe, pl, opts=o)                       # records built on the host
        assert set(int(e) for e in pairs["entity"]) == want and len(pairs) == 2 * len(want)
        if want:
            seen_some = True
            res_h = gpu_ctx.frame(ents, pairs)
            k = 0
            for p in range(len(pe)):
                if int(pe[p]) not in want:
                    continue
                assert np.array_equal(res_h.indices(k), world_indices[p])                # same volume on both routes
                k += 1
            e0 = min(want)
            assert not shell[e0]
            W = we[e0]; p0 = int(np.nonzero(pe == e0)[0][0])
            w = oracle.md2_pair(m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"], W["oldorigin"],
                                W["angles"], wl["origin"][pl[p0]], 300.0, idx=m.idx)
            assert np.array_equal(world_indices[p0], w["indices"])
    assert seen_some
