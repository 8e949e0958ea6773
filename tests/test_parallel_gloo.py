"""World-size-2 gloo test (CPU) of the multi-GPU host logic: block partition, AABB all-gather, and the rule that
every unordered pair is emitted by exactly one rank (the one owning the lower sorted position)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from ncollide_b200.parallel import shard_range

    for n in (0, 1, 7, 8, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def _worker(rank, world, port, n, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ncollide_b200.parallel import all_gather_rows, shard_range
    from ncollide_b200.scenes import config_scene
    from oracle.pyoracle import Oracle

    orc = Oracle()
    s = config_scene(3, n)
    fat_all = orc.compute_aabbs(s)  # what every rank would have after the exchange
    b, e = shard_range(n, world, rank)
    # each rank only "computes" its own block ...
    lo = torch.zeros((n, 4), dtype=torch.float32)
    hi = torch.zeros((n, 4), dtype=torch.float32)
    lo[b:e, :3] = torch.from_numpy(fat_all[b:e, :3])
    hi[b:e, :3] = torch.from_numpy(fat_all[b:e, 3:])
    # ... and the all-gather replicates the rest
    all_gather_rows(lo, b, e, world)
    all_gather_rows(hi, b, e, world)
    fat = np.concatenate([lo[:, :3].numpy(), hi[:, :3].numpy()], axis=1)
    assert np.array_equal(fat, fat_all)
    # sorted order (any fixed order works for the ownership rule): by x centre, ties by handle
    order = np.lexsort((np.arange(n), fat[:, 0] + fat[:, 3]))
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    pairs = orc.broad_phase(fat, s.groups, mode=1)
    qb, qe = shard_range(n, world, rank)
    owner_pos = np.minimum(pos[pairs[:, 0]], pos[pairs[:, 1]])
    mine = pairs[(owner_pos >= qb) & (owner_pos < qe)]
    np.save(os.path.join(tmp, f"pairs_{rank}.npy"), mine)
    dist.barrier()
    if rank == 0:
        got = np.concatenate([np.load(os.path.join(tmp, f"pairs_{r}.npy")) for r in range(world)])
        key = lambda p: np.unique(np.sort(p, axis=1), axis=0)
        assert len(got) == len(pairs), "a pair was emitted twice or dropped"
        assert np.array_equal(key(got), key(pairs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 1501])
def test_two_rank_gloo_sharding(tmp_path, n):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
