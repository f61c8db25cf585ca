"""Multi-GPU plumbing: entity-major sharding and the ONE collective of the path.

Every (entity, light) pair is independent and the lerp depends only on the entity, so rank r owns a contiguous
block of entities together with all of their lights; nothing crosses devices while a frame is computed.  After
the frame the per-shard counters {pairs, lerped verts, shadow-volume tris, index words} are all-gathered
(NCCL over NVLink on the GPU box; gloo in the CPU tests) for the throughput report."""
import numpy as np

from . import host

COUNTER_NAMES = ("pairs", "lerped_verts", "sv_tris", "index_words")


def shard_pairs(n_ents, pair_ent, pair_light, rank, world):
    """Entity range of `rank` and the sub-list of pairs that follow those entities (order preserved; the
    returned pair_ent is re-based to the shard's first entity)."""
    first, count = host.shard_entities(n_ents, rank, world)
    pair_ent = np.asarray(pair_ent, np.int32); pair_light = np.asarray(pair_light, np.int32)
    keep = (pair_ent >= first) & (pair_ent < first + count)
    return first, count, pair_ent[keep] - first, pair_light[keep], np.nonzero(keep)[0]


def gather_counters(counters, device="cpu", dist=None):
    """all_gather of one int64 row per rank -> array [world][len(counters)].  Without an initialised process
    group (single GPU) the row is returned as is."""
    import torch
    row = torch.tensor([int(x) for x in counters], dtype=torch.int64, device=device)
    if dist is None or not dist.is_initialized():
        return row.cpu().numpy()[None, :]
    out = [torch.zeros_like(row) for _ in range(dist.get_world_size())]
    dist.all_gather(out, row)
    return torch.stack(out).cpu().numpy()


def max_over_ranks(values, device="cpu", dist=None):
    import torch
    t = torch.tensor([float(x) for x in values], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().tolist()
