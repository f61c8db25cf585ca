"""The per-pair verification hash has three implementations that must be one function: the device kernel
(bq2_hash_results), the reference driver (oracle/ref/ref_driver.c, from the engine's globals) and numpy
(oracle_lib.pair_hash, from arrays).  Here, without a GPU: the driver's against numpy's over the reference's own outputs."""
import numpy as np

import oracle_lib
from scenes import Md2Model, world_scene


def test_reference_driver_hash_equals_numpy_hash(ref, oracle):
    models = [Md2Model(oracle, frames=6, seed=40 + s) for s in range(3)]
    we, wl, pe, pl = world_scene(12, 3, 6, n_models=3, seed=99)
    hashes, total = oracle_lib.reference_pair_hashes([m.blob for m in models], [m.nb for m in models], we, wl, pe, pl, procs=2)
    assert len(hashes) == len(pe) and len(set(hashes.tolist())) == len(pe)
    tot = 0
    for p in range(len(pe)):
        W = we[int(pe[p])]; m = models[int(W["model"])]
        r = ref.md2_pair(m.blob, m.nb, int(W["frame"]), int(W["oldframe"]), float(W["backlerp"]), W["origin"], W["oldorigin"],
                         W["angles"], wl["origin"][pl[p]], float(wl["radius"][pl[p]]))
        assert hashes[p] == oracle_lib.pair_hash(r["lerped"], r["extruded"], r["viss"], r["tris_verts"]), p
        tot += len(r["tris_verts"])
    assert tot == total
    # one flipped bit anywhere changes the hash
    r["extruded"].view(np.uint32)[5, 1] ^= 1
    assert hashes[-1] != oracle_lib.pair_hash(r["lerped"], r["extruded"], r["viss"], r["tris_verts"])
