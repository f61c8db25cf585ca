"""Shared builders for parity tests: synthetic models + the oracle-side precompute a loader would do."""
import numpy as np

import oracle_lib
from berserkerquake2_b200 import synth

SMOOTH_COSINE = float(np.float32(np.cos(np.deg2rad(45.0))))     # r_smoothangle 45, M:4427-4435


class Md2Model:
    def __init__(self, o, slices=16, rings=17, frames=198, seed=0, tangent_space=False, smooth=True):
        self.blob = synth.md2(slices, rings, frames, seed)
        self.V, self.T, self.F = oracle_lib.md2_dims(self.blob)
        self.idx = oracle_lib.md2_idx(self.blob)
        self.nb = o.md2_neighbours(self.blob)
        self.smooth = smooth
        self.tangents = self.binormals = self.normals = None
        if tangent_space:
            st = synth.md2_st(self.blob)
            if smooth:
                self.tangents, self.binormals = o.md2_tangents_smooth(self.blob, st, SMOOTH_COSINE)
            else:
                self.normals, self.tangents, self.binormals = o.md2_tangents_flat(self.blob, st)

    def upload(self, ctx):
        return ctx.upload_md2(self.blob, self.nb, self.tangents, self.binormals, self.normals, smooth=self.smooth)


class Md3Model:
    """num_meshes spheres side by side; tangent/binormal bytes and neighbours from the oracle precompute."""

    def __init__(self, o, slices=32, rings=33, frames=8, num_meshes=4, seed=0, casts=None):
        self.F = frames
        rng = np.random.RandomState(1000 + seed)
        self.frame_translate = (rng.rand(frames, 3).astype(np.float32) - 0.5) * 4
        self.meshes = []
        for m in range(num_meshes):
            verts, idx, st = synth.md3_mesh(slices, rings, frames, seed * 16 + m, centre=(40.0 * m, 0.0, 10.0 * m))
            o.md3_tangents(verts, st, idx)
            self.meshes.append(dict(verts=verts, indexes=idx, neighbours=o.md3_neighbours(idx), st=st,
                                    casts_shadow=True if casts is None else bool(casts[m])))

    def upload(self, ctx):
        return ctx.upload_md3(self.frame_translate, self.meshes)


def world_scene(n_ents, lights_per_ent, num_frames, n_models=1, seed=12345, step=0, radius=300.0):
    """(world entities, world lights, pair_ent, pair_light): light l of entity e is world light e*L+l."""
    import berserkerquake2_b200 as bq2
    we, lw = synth.scene(n_ents, lights_per_ent, num_frames, n_models, seed, step)
    wl = np.zeros(len(lw), bq2.WORLD_LIGHT_DTYPE)
    wl["origin"] = lw
    wl["radius"] = radius
    pe = np.repeat(np.arange(n_ents, dtype=np.int32), lights_per_ent)
    pl = np.arange(n_ents * lights_per_ent, dtype=np.int32)
    return we, wl, pe, pl


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def near_zero_facing(o_pair, idx, eps=1e-4):
    """SURVEY 8d epsilon exclusion: triangles whose facing dot lies within eps of zero (oracle-evaluated)."""
    P, tn, L = o_pair["lerped"], o_pair["trinormals"], o_pair["light_local"]
    dl = (tn * L[None, :]).sum(1)
    dp = (P[idx[:, 0]] * tn).sum(1)
    return np.abs(dl - dp) <= eps * np.maximum(1.0, np.abs(dp))
