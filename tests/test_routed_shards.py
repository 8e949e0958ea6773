"""Routed multi-GPU sharding (ncb_world_update_routed + ncollide_b200.parallel.ShardedWorld.routed_plan).

GPU tests replay all ranks inside one process (one context per rank on cuda:0, the collectives performed by
parallel.run_plans_lockstep): the union of the ranks' pairs is the full pair set of a single-context update, no pair is
reported twice, manifolds are identical — also when every rank holds only the poses of its own block (they travel with the
records) and when the bucket capacities start too small (grow + repeat).  The CPU test runs the same collective sequence over
gloo with a numpy stand-in for the device stages, checking the protocol (owner / ghost / reporting rules)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from ncollide_b200 import _ffi
from ncollide_b200.parallel import ShardedWorld, run_plans_lockstep, shard_range
from ncollide_b200.scenes import config_scene, make_world_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("world1", "world2", "normal", "depth", "f1", "f2")


@pytest.fixture(scope="module")
def ctx():
    from ncollide_b200.world import Context

    c = Context(0)
    yield c
    c.close()


def canon(p):
    p = np.sort(np.asarray(p).reshape(-1, 2), axis=1)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def _replay(s, world, with_poses, full_ctx, p2p=False):
    import torch

    from ncollide_b200.world import Context

    full_ctx.set_scene(s)
    full = full_ctx.world_fetch(full_ctx.world_update_device(s.margin))
    ctxs, sws, counts = [], [], []
    try:
        for r in range(world):
            c = Context(0)
            c.set_scene(s)
            if with_poses:  # a rank knows only the poses of its own block; the rest is poison the narrow phase must never read
                b, e = shard_range(s.n, world, r)
                pos = np.full_like(s.pos, 1.0e6)
                rot = np.zeros_like(s.rot)
                pos[b:e], rot[b:e] = s.pos[b:e], s.rot[b:e]
                c.set_positions(pos, rot)
            ctxs.append(c)
            sws.append(ShardedWorld(c, s, world, r, torch.device("cuda", 0), mode="p2p" if p2p else "routed"))
            counts.append(_ffi.UpdateCountsC())
        if p2p:
            ShardedWorld.connect_p2p_local(sws)
        for _step in range(3 if p2p else 2):  # later steps run on the remembered capacities / the next flag epochs
            res = run_plans_lockstep([sw.routed_plan(cc, with_poses) for sw, cc in zip(sws, counts)])
            per_pair, total = {}, 0
            for r in range(world):
                out = ctxs[r].world_fetch(res[r])
                assert np.all(out.pairs[:, 0] > out.pairs[:, 1])
                total += len(out.pairs)
                for i, p in enumerate(map(tuple, out.pairs.tolist())):
                    assert p not in per_pair, "pair reported by two ranks"
                    per_pair[p] = out.contacts_of(i)[list(FIELDS)].copy()
            assert total == len(full.pairs), "a pair was reported by two ranks or by none"
            assert np.array_equal(canon(np.array(list(per_pair), dtype=np.uint32)), canon(full.pairs))
            for i, p in enumerate(map(tuple, full.pairs.tolist())):
                a, b = full.contacts_of(i)[list(FIELDS)], per_pair[p]
                assert len(a) == len(b)
                for name in FIELDS:
                    assert np.array_equal(a[name], b[name]), (p, name)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,mk,with_poses", [
    (2, lambda: config_scene(3, 7001), False),
    (2, lambda: config_scene(3, 7001), True),
    (4, lambda: config_scene(2, 9000), True),
    (8, lambda: config_scene(5, 30000), False),
    (8, lambda: config_scene(3, 40000), True),
    (3, lambda: make_world_scene(5000, 91, (1, 1, 1), side=4.0, n_hulls=16, name="tiny_dense"), True),
    (4, lambda: make_world_scene(3000, 92, (1, 1, 1), side=12.0, n_hulls=16, plane=True, name="with_plane"), True),
])
def test_routed_shards_partition_the_pair_set(ctx, world, mk, with_poses):
    _replay(mk(), world, with_poses, ctx)


@pytest.mark.gpu
@pytest.mark.parametrize("world,mk,with_poses", [
    (2, lambda: config_scene(3, 7001), True),
    (8, lambda: config_scene(3, 40000), True),
    (5, lambda: config_scene(5, 30000), False),
    (4, lambda: make_world_scene(3000, 92, (1, 1, 1), side=12.0, n_hulls=16, plane=True, name="with_plane"), True),
])
def test_routed_shards_over_peer_memory(ctx, world, mk, with_poses):
    """The same partition property with the peer-memory exchange (records stored straight into the owner's buffers, flag rounds instead
    of collectives); all ranks live in this process, so the peers are mapped by raw pointer and advance stage by stage."""
    _replay(mk(), world, with_poses, ctx, p2p=True)


@pytest.mark.gpu
def test_routed_shards_grow_their_buckets(ctx, monkeypatch):
    """Bucket capacities that start far too small: stage 4 answers NCB_ROUTE_REPEAT, every rank raises them identically."""
    monkeypatch.setenv("NCB_ROUTE_SLACK", "1")
    # handle order == spatial order: almost the whole block of a rank goes to ONE owner, the worst case for equal-split buckets
    s = config_scene(3, 20000)
    order = np.lexsort((s.pos[:, 2], s.pos[:, 1], s.pos[:, 0]))
    for f in ("pos", "rot", "shape_type", "shape_param", "groups", "query_limit", "ang_pred"):
        setattr(s, f, np.ascontiguousarray(getattr(s, f)[order]))
    _replay(s, 4, True, ctx)


# ---- CPU: the collective sequence over gloo, device stages replaced by numpy ------------------------------------------------
def _bins(lo, hi, gb):
    bmin, bmax = -gb[:3], gb[3:]
    e = max(float((bmax - bmin).max()), 1e-20)
    c = ((lo + hi) * 0.5 - bmin) * (1023.0 / e)
    u = np.clip(c, 0, 1023).astype(np.uint32)

    def expand(v):
        v = v.astype(np.uint64)
        out = np.zeros_like(v)
        for b in range(10):
            out |= ((v >> b) & 1) << (3 * b)
        return out

    code = (expand(u[:, 0]) << 2) | (expand(u[:, 1]) << 1) | expand(u[:, 2])
    return (code >> 20).astype(np.int64)


def _numpy_plan(rank, world, fat, groups, orc, out):
    """The routed protocol with numpy stages on CPU tensors (same requests as ShardedWorld.routed_plan)."""
    import torch

    n = len(fat)
    b, e = shard_range(n, world, rank)
    lo, hi = fat[b:e, :3], fat[b:e, 3:]
    ctr = (lo + hi) * 0.5
    bounds = torch.from_numpy(np.concatenate([-ctr.min(axis=0), ctr.max(axis=0)]).astype(np.float32))
    yield ("all_reduce_max", bounds)
    bins = _bins(lo, hi, bounds.numpy())
    hist = torch.from_numpy(np.bincount(bins, minlength=1024).astype(np.int32))
    yield ("all_reduce_sum", hist)
    h = hist.numpy().astype(np.int64)
    total, acc, split, r = int(h.sum()), 0, [0], 1
    for k in range(1024):
        acc += int(h[k])
        while r < world and acc * world >= total * r:
            split.append(k + 1)
            r += 1
    split += [1024] * (world - len(split)) + [1025]
    owner = np.searchsorted(np.array(split[1:world]), bins, side="right")
    cap = n + 1
    rec = np.concatenate([fat[b:e], np.arange(b, e, dtype=np.float64)[:, None], owner[:, None].astype(np.float64)], axis=1)  # 8 columns
    send = np.zeros((world, cap, 8))
    region = np.full((world, 6), -np.inf)
    for q in range(world):
        m = owner == q
        send[q, 0, 0] = m.sum()
        send[q, 1 : 1 + m.sum()] = rec[m]
        if m.any():
            region[q, :3] = -lo[m].min(axis=0)
            region[q, 3:] = hi[m].max(axis=0)
    send_t, recv_t = torch.from_numpy(send.reshape(-1)), torch.zeros(world * cap * 8, dtype=torch.float64)
    yield ("all_to_all", recv_t, send_t)
    region_t = torch.from_numpy(region.reshape(-1))
    yield ("all_reduce_max", region_t)
    reg = region_t.numpy().reshape(world, 6)
    gsend = np.zeros((world, cap, 8))
    for q in range(world):
        m = (owner != q) & np.all(lo <= reg[q, 3:], axis=1) & np.all(hi >= -reg[q, :3], axis=1)
        gsend[q, 0, 0] = m.sum()
        gsend[q, 1 : 1 + m.sum()] = rec[m]
    gsend_t, grecv_t = torch.from_numpy(gsend.reshape(-1)), torch.zeros(world * cap * 8, dtype=torch.float64)
    yield ("all_to_all", grecv_t, gsend_t)
    rows = []
    for buf in (recv_t, grecv_t):
        a = buf.numpy().reshape(world, cap, 8)
        for q in range(world):
            rows.append(a[q, 1 : 1 + int(a[q, 0, 0])])
    loc = np.concatenate(rows)
    handles, owners = loc[:, 6].astype(np.int64), loc[:, 7].astype(np.int64)
    assert len(np.unique(handles)) == len(handles), "an object arrived twice"
    pairs = orc.broad_phase(np.ascontiguousarray(loc[:, :6], dtype=np.float32), np.ascontiguousarray(groups[handles]), mode=1)
    o1, o2 = owners[pairs[:, 0]], owners[pairs[:, 1]]
    mine = ((o1 == rank) & (o2 == rank)) | (((o1 == rank) | (o2 == rank)) & (np.minimum(o1, o2) == rank))
    out.append(np.stack([handles[pairs[mine, 0]], handles[pairs[mine, 1]]], axis=1))
    return len(out[-1])


def _gloo_worker(rank, world, port, n, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ncollide_b200.parallel import perform_collective
    from oracle.pyoracle import Oracle

    orc = Oracle()
    s = config_scene(3, n)
    fat = orc.compute_aabbs(s)
    out = []
    plan = _numpy_plan(rank, world, fat, s.groups, orc, out)
    try:
        while True:
            perform_collective(next(plan))
    except StopIteration:
        pass
    np.save(os.path.join(tmp, f"routed_{rank}.npy"), out[0])
    dist.barrier()
    if rank == 0:
        got = np.concatenate([np.load(os.path.join(tmp, f"routed_{r}.npy")) for r in range(world)])
        want = orc.broad_phase(fat, s.groups, mode=1)
        assert len(got) == len(want), "a pair was reported twice or dropped"
        assert np.array_equal(canon(got), canon(want))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1200])
def test_routed_protocol_two_rank_gloo(tmp_path, n):
    import torch.multiprocessing as mp

    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)


def test_routed_protocol_lockstep_numpy():
    """run_plans_lockstep (the replay helper the GPU tests use) on the numpy stand-in: 3 ranks in one process."""
    from oracle.pyoracle import Oracle

    orc = Oracle()
    s = config_scene(3, 1500)
    fat = orc.compute_aabbs(s)
    outs = [[] for _ in range(3)]
    run_plans_lockstep([_numpy_plan(r, 3, fat, s.groups, orc, outs[r]) for r in range(3)])
    got = np.concatenate([o[0] for o in outs])
    want = orc.broad_phase(fat, s.groups, mode=1)
    assert len(got) == len(want)
    assert np.array_equal(canon(got), canon(want))
