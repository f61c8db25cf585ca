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


class BrushModel:
    """A closed convex-ish brush made of polygons with 3..S vertices (a lat-long drum: S-gon caps, quad bands,
    a few quads split into triangles), each polygon with its own vertex copies and per-edge neighbour polygon,
    as glpoly_t holds them (defs.h:3329-3338).  Vertex coordinates are quantised to 1/8 like BSP geometry."""

    def __init__(self, slices=8, bands=3, seed=0, open_edges=0):
        rng = np.random.RandomState(seed)
        S = slices
        ring = []
        for r in range(bands + 1):
            z = 40.0 * (r / bands) - 20.0
            rad = 30.0 + 6.0 * np.sin(1.3 * r + seed)
            ring.append([(rad * np.cos(2 * np.pi * s / S), rad * np.sin(2 * np.pi * s / S), z) for s in range(S)])
        polys = [list(reversed(ring[0])), list(ring[bands])]                  # bottom and top caps (outward winding)
        for r in range(bands):
            for s in range(S):
                a, b = ring[r][s], ring[r][(s + 1) % S]
                c_, d = ring[r + 1][(s + 1) % S], ring[r + 1][s]
                if rng.rand() < 0.25:
                    polys += [[a, b, c_], [a, c_, d]]
                else:
                    polys.append([a, b, c_, d])
        self.numverts = np.array([len(p) for p in polys], np.int32)
        self.first = np.concatenate([[0], np.cumsum(self.numverts)]).astype(np.int32)
        self.verts = (np.round(np.array([v for p in polys for v in p], np.float64) * 8) / 8).astype(np.float32)
        key = lambda v: tuple(np.round(np.asarray(v) * 8).astype(np.int64))
        edges = {}
        for p, poly in enumerate(polys):
            for j in range(len(poly)):
                k = tuple(sorted((key(poly[j]), key(poly[(j + 1) % len(poly)]))))
                edges.setdefault(k, []).append((p, j))
        self.neighbours = np.full(int(self.first[-1]), -1, np.int32)
        for lst in edges.values():
            if len(lst) == 2:
                (p0, j0), (p1, j1) = lst
                self.neighbours[self.first[p0] + j0] = p1
                self.neighbours[self.first[p1] + j1] = p0
        for _ in range(open_edges):
            self.neighbours[rng.randint(len(self.neighbours))] = -1
        self.n_polys = len(polys)
        self.fence = (rng.rand(self.n_polys) < 0.1).astype(np.uint8)
        self.normals = np.zeros((self.n_polys, 3)); self.centroids = np.zeros((self.n_polys, 3))
        for p in range(self.n_polys):
            v = self.verts[self.first[p]:self.first[p + 1]].astype(np.float64)
            self.normals[p] = np.cross(v[1] - v[0], v[2] - v[0]); self.centroids[p] = v.mean(0)

    def lit(self, light_model):
        """what the engine's light-marking pass stamps: polygons facing the light"""
        d = ((np.asarray(light_model, np.float64)[None, :] - self.centroids) * self.normals).sum(1)
        return (d > 0).astype(np.uint8)
