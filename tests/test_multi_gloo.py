"""The N>1 path on CPU: world_size-2 gloo.  Entity-major sharding + the single collective (all_gather of the
per-shard counters).  The per-rank "frame" here only COUNTS (what the kernel would report) -- it runs no
geometry, the geometry needs a GPU -- so this covers the host logic the GPU box runs under NCCL."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_ents, lights, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from berserkerquake2_b200 import shard
    rng = np.random.RandomState(0)                      # the same global pair list on every rank
    pe = np.sort(rng.randint(0, n_ents, n_ents * lights)).astype(np.int32)
    pl = rng.randint(0, 16, len(pe)).astype(np.int32)
    first, count, spe, spl, kept = shard.shard_pairs(n_ents, pe, pl, rank, world)
    assert ((spe >= 0) & (spe < max(count, 1))).all() and np.array_equal(pl[kept], spl)
    counters = [len(spe), count * 258, int((spe + first).sum()), int(kept.sum())]
    allc = shard.gather_counters(counters, "cpu", dist)
    tmax = shard.max_over_ranks([float(rank + 1), 10.0 - rank], "cpu", dist)
    if rank == 0:
        q.put((allc.tolist(), tmax, len(pe), int(pe.sum()), int(np.arange(len(pe)).sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_ents,lights", [(1024, 4), (7, 3)])
def test_world_size_2_sharding_and_counter_gather(n_ents, lights):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + n_ents) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_ents, lights, q)) for r in range(2)]
    for p in procs:
        p.start()
    allc, tmax, n_pairs, ent_sum, pos_sum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allc = np.array(allc)
    assert allc.shape == (2, 4)
    assert allc[:, 0].sum() == n_pairs                  # every pair is owned by exactly one rank
    assert allc[:, 1].sum() == n_ents * 258             # every entity is lerped exactly once
    assert allc[:, 2].sum() == ent_sum and allc[:, 3].sum() == pos_sum
    assert tmax == [2.0, 10.0]                          # max over ranks
