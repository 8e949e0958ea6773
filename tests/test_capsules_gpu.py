"""GPU parity tests for worlds with capsules: the CapsuleCapsule / CapsuleShape generators on the device (k_capsule,
csrc/capsule.cuh) against the CPU oracle — capsule x {capsule, ball, cuboid, hull, plane} in both operand orders, fresh-world updates,
the batched generator entry, the stepping world and the committed golden fixture.  The oracle's capsule path has no known-answer test in
the reference (no reference test involves a capsule): it is checked against closed-form geometry in tests/test_oracle_capsule.py."""
import os

import numpy as np
import pytest

from ncollide_b200.scenes import make_world_scene
from test_gpu_parity import canon, compare_manifolds

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32


def capsule_scene(n, seed, kinds, side, plane, ang=0.0):
    s = make_world_scene(n, seed, kinds, side=side, n_hulls=24, plane=plane, angular=ang)
    rng = np.random.default_rng(seed + 1)
    cap = np.zeros(s.n, dtype=bool)
    cap[:n] = np.arange(n) % 3 == 1
    s.shape_type[cap] = 4
    s.shape_param[cap, 0] = rng.uniform(0.2, 0.5, size=int(cap.sum())).astype(F)
    s.shape_param[cap, 1] = rng.uniform(0.15, 0.3, size=int(cap.sum())).astype(F)
    s.shape_param[cap, 2:] = 0
    return s


SCENES = [(2400, (1, 1, 1), 8.0, True, 0.0, 141), (1800, (0, 1, 1), 5.5, False, 0.03, 142), (1500, (1, 1, 0), 5.0, True, 0.0, 143)]


@pytest.fixture(scope="module")
def ctx():
    from ncollide_b200.world import Context

    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n,kinds,side,plane,ang,seed", SCENES)
def test_world_update_with_capsules(ctx, oracle, n, kinds, side, plane, ang, seed):
    s = capsule_scene(n, seed, kinds, side, plane, ang)
    ctx.set_hulls(s.hulls)
    res = ctx.world_update(s)
    assert res.counts["epa_overflow"] == 0 and res.counts["ref_panics"] == 0 and res.counts["stack_overflow"] == 0
    fat = oracle.compute_aabbs(s)
    assert np.array_equal(ctx.compute_aabbs(s.margin, 2).view(np.uint32), fat.view(np.uint32)), "capsule AABBs"
    want = oracle.broad_phase(fat, s.groups, mode=1)
    assert np.array_equal(canon(res.pairs), canon(want))
    compare_manifolds(res, s, oracle, f"capsules {seed}")
    t = s.shape_type
    both = (t[res.pairs[:, 0]] == 4) & (t[res.pairs[:, 1]] == 4)
    one = (t[res.pairs[:, 0]] == 4) ^ (t[res.pairs[:, 1]] == 4)
    assert np.all(res.pair_algo[both] == 7) and np.all(res.pair_algo[one] == 8)
    assert res.counts["n_algo"]["capsule_capsule"] == int(both.sum()) and res.counts["n_algo"]["capsule_shape"] == int(one.sum())
    assert both.sum() > 50 and one.sum() > 300
    for other in set(t[t != 4].tolist()):
        sel = ((t[res.pairs[:, 0]] == 4) & (t[res.pairs[:, 1]] == other)) | ((t[res.pairs[:, 0]] == other) & (t[res.pairs[:, 1]] == 4))
        assert res.manifold_count[sel].sum() > 0, f"no contact between a capsule and shape {other}"


@pytest.mark.parametrize("n,kinds,side,plane,ang,seed", SCENES[:2])
def test_generate_contacts_with_capsules_both_orders(ctx, oracle, n, kinds, side, plane, ang, seed):
    s = capsule_scene(n, seed, kinds, side, plane, ang)
    ctx.set_scene(s)
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), s.groups, mode=0)
    both = np.concatenate([pairs, pairs[:, ::-1]])
    res = ctx.generate_contacts(both)
    compare_manifolds(res, s, oracle, f"capsule generators {seed}")


def test_capsule_golden_fixture_device(ctx):
    from golden.make_golden import scene_from_npz

    z = np.load(os.path.join(HERE, "golden", "capsule_mixed_plane_300.npz"))
    s = scene_from_npz(z)
    ctx.set_scene(s)
    assert np.array_equal(ctx.compute_aabbs(s.margin, 2), z["fat_aabbs"])
    res = ctx.generate_contacts(z["pairs"])
    assert np.array_equal(res.pair_algo, z["algo"]) and np.array_equal(res.manifold_count, np.diff(z["manifold_off"]))
    idx = np.concatenate([np.arange(a, a + c) for a, c in zip(res.manifold_start, res.manifold_count)]).astype(np.int64)
    dc = res.contacts[idx]
    assert np.array_equal(dc["f1"], z["c_f1"]) and np.array_equal(dc["f2"], z["c_f2"])
    for name in ("world1", "world2", "normal", "depth"):
        assert np.allclose(dc[name], z["c_" + name], rtol=1e-4, atol=1e-5), name


def test_stepping_world_with_capsules_matches_oracle(oracle):
    """CollisionWorld::update over 6 steps of a world with capsules (persistent manifolds, contact ids, warm-started GJK, events)."""
    from ncollide_b200.world import Context
    from sim_scenario import drive
    from test_bp_persistent import DeviceSimAdapter, compare_sim_logs

    s = capsule_scene(900, 151, (1, 1, 1), 5.5, True)
    c = Context(0)
    c.set_scene(s)
    dev = drive(DeviceSimAdapter(c, s), s, steps=6, seed=9)
    ref = drive(oracle.sim(s), s, steps=6, seed=9)
    compare_sim_logs(dev, ref)
    assert sum(int((r["algo"] >= 7).sum()) for r in ref) > 500
    c.close()


def test_sensors_and_queries_refuse_capsule_worlds(ctx):
    from ncollide_b200._ffi import NcbError

    s = capsule_scene(300, 161, (1, 1, 0), 3.0, False)
    ctx.set_scene(s)
    kinds = np.zeros(s.n, dtype=np.uint8)
    kinds[3] = 1
    with pytest.raises(NcbError, match="capsule"):
        ctx.set_query_types(kinds)
