"""Proximity-only interactions: GeometricQueryType::Proximity objects ("sensors").

CPU part: the oracle's restatement of the proximity detectors against the reference's own example
(examples3d/proximity_query3d.rs: ball vs cuboid -> Intersecting / WithinMargin / Disjoint), against a numpy restatement
of proximity_ball_ball, against exact distances, and the stepping world's ProximityEvents on a hand-checkable scenario.
GPU part (-m gpu): ncb_proximity and the world update with sensors against the oracle, bit-exact statuses."""
import numpy as np
import pytest

from ncollide_b200.scenes import WorldScene, make_world_scene, with_sensors
from ncollide_b200.shapes import BALL, CUBOID, HULL, PLANE, HullLibrary

F = np.float32
INTERSECTING, WITHIN_MARGIN, DISJOINT, NONE = 0, 1, 2, 255


def two_shapes(t1, p1, pos1, t2, p2, pos2, hulls=None, dtype=F, rot1=(0, 0, 0, 1), rot2=(0, 0, 0, 1), ql=(0.0, 0.0)):
    return WorldScene(
        pos=np.array([pos1, pos2], dtype=dtype), rot=np.array([rot1, rot2], dtype=dtype),
        shape_type=np.array([t1, t2], dtype=np.uint32), shape_param=np.array([p1, p2], dtype=dtype),
        groups=None, query_limit=np.array(ql, dtype=dtype), ang_pred=np.zeros(2, dtype=dtype), hulls=hulls or HullLibrary([]),
    )


# ---- reference known-answer test: examples3d/proximity_query3d.rs ----------------------------------------------
@pytest.mark.parametrize("which", ["oracle", "oracle64"])
def test_reference_example_proximity_query3d(which, request):
    orc = request.getfixturevalue(which)
    want = {(1, 1, 1): INTERSECTING, (2, 2, 2): WITHIN_MARGIN, (3, 3, 3): DISJOINT}
    for pos, st in want.items():
        # query::proximity(&ball_pos, &ball, &cuboid_pos, &cuboid, margin = 1.0)
        s = two_shapes(BALL, [1, 0, 0, 0], pos, CUBOID, [1, 1, 1, 0], (0, 0, 0), dtype=orc.dtype)
        assert orc.query_proximity(s, 1.0) == st
        # and with the operands swapped
        s = two_shapes(CUBOID, [1, 1, 1, 0], (0, 0, 0), BALL, [1, 0, 0, 0], pos, dtype=orc.dtype)
        assert orc.query_proximity(s, 1.0) == st


def test_proximity_ball_ball_matches_numpy_restatement(oracle):
    """proximity_ball_ball.rs:8-36 restated in numpy f32 with the same operation order."""
    rng = np.random.default_rng(5)
    n = 4000
    s = make_world_scene(n, 11, (1, 0, 0), side=6.0)
    pairs = rng.integers(0, n, size=(20000, 2)).astype(np.uint32)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    margins = rng.uniform(0, 0.5, size=len(pairs)).astype(F)
    got = oracle.proximity(s, pairs, margins)
    c1, c2 = s.pos[pairs[:, 0]], s.pos[pairs[:, 1]]
    d = (c2 - c1).astype(F)
    d2 = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(F) + d[:, 2] * d[:, 2]).astype(F)
    sr = (s.shape_param[pairs[:, 0], 0] + s.shape_param[pairs[:, 1], 0]).astype(F)
    sre = (sr + margins).astype(F)
    want = np.where(d2 <= (sre * sre).astype(F), np.where(d2 <= (sr * sr).astype(F), INTERSECTING, WITHIN_MARGIN), DISJOINT)
    assert np.array_equal(got, want.astype(np.uint8))
    assert len(set(got.tolist())) == 3


def test_proximity_consistent_with_exact_distance(oracle):
    """Away from the two thresholds the status must agree with the exact signed distance (from the contact query with a
    huge prediction): Intersecting iff penetrating, WithinMargin iff 0 < distance <= margin, Disjoint beyond."""
    s = make_world_scene(1500, 21, (1, 1, 1), side=3.5, n_hulls=32)
    rng = np.random.default_rng(3)
    pairs = rng.integers(0, s.n, size=(3000, 2)).astype(np.uint32)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    margin = 1.5
    got = oracle.proximity(s, pairs, np.full(len(pairs), margin, dtype=F))
    checked = {0: 0, 1: 0, 2: 0}
    for (a, b), st in zip(pairs, got):
        two = WorldScene(pos=s.pos[[a, b]], rot=s.rot[[a, b]], shape_type=s.shape_type[[a, b]], shape_param=s.shape_param[[a, b]],
                         groups=None, query_limit=np.zeros(2, dtype=F), ang_pred=np.zeros(2, dtype=F), hulls=s.hulls)
        c = oracle.query_contact(two, 100.0)
        assert c is not None
        dist = -float(c["depth"])
        if abs(dist) < 5e-3 or abs(dist - margin) < 5e-3:
            continue
        want = INTERSECTING if dist < 0 else (WITHIN_MARGIN if dist <= margin else DISJOINT)
        assert st == want, (a, b, dist, st)
        checked[want] += 1
    assert min(checked.values()) > 20, checked


def test_plane_proximity_and_no_detector(oracle):
    """proximity_plane_support_map.rs:9-47 on hand-checkable cases; plane x plane has no detector."""
    up = [0, 1, 0, 0]
    for y, st in ((0.5, INTERSECTING), (1.0, INTERSECTING), (1.25, WITHIN_MARGIN), (1.5, WITHIN_MARGIN), (1.75, DISJOINT)):
        for other, param in ((BALL, [1, 0, 0, 0]), (CUBOID, [1, 1, 1, 0])):
            s = two_shapes(PLANE, up, (0, 0, 0), other, param, (3, y, -2))
            assert oracle.query_proximity(s, 0.5) == st, (y, other)
            s = two_shapes(other, param, (3, y, -2), PLANE, up, (0, 0, 0))
            assert oracle.query_proximity(s, 0.5) == st, (y, other)
    s = two_shapes(PLANE, up, (0, 0, 0), PLANE, up, (0, 1, 0))
    assert oracle.proximity(s, [[0, 1]])[0] == NONE


def test_narrow_phase_with_sensors_splits_pairs(oracle):
    s = with_sensors(make_world_scene(2500, 31, (1, 1, 1), side=8.0, plane=True, n_hulls=32), 0.3, 7, margin=0.1)
    fat = oracle.compute_aabbs(s)
    pairs = oracle.broad_phase(fat, s.groups)
    c, off, algo, prox = oracle.narrow_phase_kinds(s, pairs)
    sensor = (s.query_kind[pairs[:, 0]] | s.query_kind[pairs[:, 1]]).astype(bool)
    assert sensor.any() and (~sensor).any()
    assert np.all(algo[sensor] == 6) and np.all(prox[~sensor] == NONE) and np.all(np.diff(off)[sensor] == 0)
    assert np.array_equal(prox[sensor], oracle.proximity(s, pairs[sensor]))
    assert set(prox[sensor].tolist()) == {0, 1, 2}
    # contact pairs are untouched by the presence of sensors
    c0, off0, algo0, _ = oracle.narrow_phase(s, pairs[~sensor])
    assert np.array_equal(algo0, algo[~sensor]) and np.array_equal(np.diff(off0), np.diff(off)[~sensor])
    assert np.array_equal(c0["depth"], c["depth"])


def test_stepping_world_proximity_events(oracle):
    """A sensor ball approaches a cuboid, enters it, leaves and goes far away: the ProximityEvents of
    NarrowPhase::update_proximity / handle_interaction (narrow_phase.rs:108-143,266-274)."""
    s = two_shapes(BALL, [0.5, 0, 0, 0], (5, 0, 0), CUBOID, [1, 1, 1, 0], (0, 0, 0), ql=(0.25, 0.0))
    s.query_kind = np.array([1, 0], dtype=np.uint8)
    sim = oracle.sim(s)
    xs = [5.0, 1.7, 1.2, 1.62, 5.0]
    # distances ball surface -> cube: 3.5 (no pair), 0.2 (within 0.25), -0.3 (intersecting), 0.12 (within), far (pair stops)
    want_status = [None, WITHIN_MARGIN, INTERSECTING, WITHIN_MARGIN, None]
    prev = DISJOINT
    for x, st in zip(xs, want_status):
        sim.set_positions([0], [[x, 0, 0]], [[0, 0, 0, 1]])
        r = sim.step()
        assert len(r["events"]) == 0 and len(r["contacts"]) == 0
        if st is None:
            assert len(r["pairs"]) == 0
            if prev != DISJOINT:  # interference_stopped: (h1, h2, prev, Disjoint)
                assert r["prox_events"].tolist() == [[0, 1, prev, DISJOINT]] or r["prox_events"].tolist() == [[1, 0, prev, DISJOINT]]
            else:
                assert len(r["prox_events"]) == 0
            prev = DISJOINT
            continue
        assert len(r["pairs"]) == 1 and r["algo"][0] == 6 and r["prox"][0] == st
        ev = r["prox_events"].tolist()
        assert len(ev) == 1 and ev[0][2:] == [prev, st] and sorted(ev[0][:2]) == [0, 1]
        prev = st
    # an idle step regenerates nothing and emits nothing
    r = sim.step()
    assert len(r["prox_events"]) == 0


def test_stepping_world_first_step_equals_fresh_world_with_sensors(oracle):
    s = with_sensors(make_world_scene(1200, 41, (1, 1, 1), side=6.5, plane=True, n_hulls=16), 0.25, 3, margin=0.15)
    sim = oracle.sim(s)
    r = sim.step()
    # same pair set / statuses as the fresh-world narrow phase on the sim's own pairs
    c, off, algo, prox = oracle.narrow_phase_kinds(s, r["pairs"])
    assert np.array_equal(algo, r["algo"]) and np.array_equal(prox, r["prox"]) and np.array_equal(np.diff(off), np.diff(r["off"]))
    ne = int(((prox != NONE) & (prox != DISJOINT)).sum())
    assert len(r["prox_events"]) == ne and ne > 0
    assert np.all(r["prox_events"][:, 2] == DISJOINT)


def test_collision_world_mirror_builds_sensor_scenes_without_a_gpu():
    """Host logic of the mirror: GeometricQueryType.Proximity(margin) -> query_kind / query_limit of the scene it uploads."""
    from ncollide_b200.shapes import Ball, Cuboid
    from ncollide_b200.world import CollisionWorld, GeometricQueryType

    ident = (0, 0, 0, 1)
    w = CollisionWorld(0.02, ctx=object())  # no device work happens before update()
    w.add(((0, 0, 0), ident), Cuboid((1, 1, 1)), query_type=GeometricQueryType.Contacts(0.02, 0.0))
    assert w.scene().query_kind is None  # no sensor: the sensor-free path
    w.add(((2, 0, 0), ident), Ball(0.5), query_type=GeometricQueryType.Proximity(0.3))
    s = w.scene()
    assert s.query_kind.tolist() == [0, 1] and s.query_kind.dtype == np.uint8
    assert np.allclose(s.query_limit, [0.02, 0.3])
    with pytest.raises(ValueError):  # the reference asserts margin >= 0 in every proximity query
        w.add(((0, 0, 0), ident), Ball(0.5), query_type=GeometricQueryType.Proximity(-0.1))
    with pytest.raises(ValueError):
        w.add(((0, 0, 0), ident), Ball(0.5), query_type=None)
    assert list(w.proximity_pairs()) == [] and w.proximity_events() == []  # before any update


# ---- golden fixture (tests/golden/prox_mixed_plane_400.npz, made by tests/golden/make_golden.py from the oracle) --------------
def _load_prox_golden():
    import os

    from golden.make_golden import scene_from_npz

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "prox_mixed_plane_400.npz"))
    return z, scene_from_npz(z)


def check_sim_against_prox_golden(sim, z, s):
    from sim_scenario import drive

    log = drive(sim, s, steps=5, seed=1020)
    for t, r in enumerate(log):
        assert np.array_equal(r["pairs"], z[f"s{t}_pairs"]) and np.array_equal(r["algo"], z[f"s{t}_algo"]), t
        assert np.array_equal(r["off"], z[f"s{t}_off"]) and np.array_equal(r["prox"], z[f"s{t}_prox"]), t
        for k in ("prox_events", "events"):
            a, b = np.asarray(r[k]), z[f"s{t}_{k}"]
            a = a[np.lexsort(a.T[::-1])] if len(a) else a
            b = b[np.lexsort(b.T[::-1])] if len(b) else b
            assert np.array_equal(a, b), (t, k)


def test_prox_golden_fixture_oracle(oracle):
    z, s = _load_prox_golden()
    assert s.query_kind is not None and s.query_kind.any()
    fat = oracle.compute_aabbs(s)
    assert np.array_equal(fat, z["fat_aabbs"])
    pairs = oracle.broad_phase(fat, s.groups, 0)
    assert np.array_equal(pairs, z["pairs"])
    c, off, algo, prox = oracle.narrow_phase_kinds(s, pairs)
    assert np.array_equal(off, z["manifold_off"]) and np.array_equal(algo, z["algo"]) and np.array_equal(prox, z["prox"])
    for name in ("world1", "world2", "normal", "depth", "f1", "f2"):
        assert np.array_equal(c[name], z["c_" + name]), name
    assert np.array_equal(oracle.proximity(s, z["batch_pairs"], z["batch_margins"]), z["batch_prox"])
    check_sim_against_prox_golden(oracle.sim(s), z, s)


@pytest.mark.gpu
def test_prox_golden_fixture_device():
    from ncollide_b200.world import Context
    from test_bp_persistent import DeviceSimAdapter
    from test_gpu_parity import canon

    z, s = _load_prox_golden()
    c = Context(0)
    c.set_hulls(s.hulls)
    r = c.world_update(s)
    assert np.array_equal(canon(r.pairs), canon(z["pairs"]))
    order = {tuple(p): i for i, p in enumerate(map(tuple, z["pairs"].tolist()))}
    j = np.array([order[tuple(p)] for p in r.pairs.tolist()])
    assert np.array_equal(r.pair_algo, z["algo"][j]) and np.array_equal(r.proximity, z["prox"][j])
    assert np.array_equal(r.manifold_count, np.diff(z["manifold_off"])[j])
    for i in np.nonzero(r.manifold_count)[0]:
        sl = slice(z["manifold_off"][j[i]], z["manifold_off"][j[i] + 1])
        got = r.contacts_of(i)
        assert np.array_equal(got["f1"], z["c_f1"][sl]) and np.array_equal(got["f2"], z["c_f2"][sl])
        for name in ("world1", "world2", "normal", "depth"):
            assert np.allclose(got[name], z["c_" + name][sl], rtol=1e-4, atol=1e-5), (i, name)
    c.set_scene(s)
    assert np.array_equal(c.proximity(z["batch_pairs"], z["batch_margins"]), z["batch_prox"])
    check_sim_against_prox_golden(DeviceSimAdapter(Context(0), s), z, s)


# ---- GPU ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    from ncollide_b200.world import Context

    c = Context(0)
    yield c
    c.close()


SENSOR_SCENES = [
    lambda: with_sensors(make_world_scene(4000, 51, (1, 1, 1), side=10.0, plane=True, n_hulls=64, name="mixed_plane_sensors"), 0.3, 1, margin=0.1),
    lambda: with_sensors(make_world_scene(3000, 52, (1, 0, 0), side=8.0, name="balls_sensors"), 0.5, 2, margin=0.2),
    lambda: with_sensors(make_world_scene(3000, 53, (0, 1, 1), side=8.0, n_hulls=32, angular=0.05, name="convex_all_sensors"), 1.0, 3, margin=0.05),
    lambda: with_sensors(make_world_scene(2000, 54, (1, 1, 1), side=5.0, n_hulls=16, name="dense_zero_margin"), 0.4, 4, margin=0.0),
]


@pytest.mark.gpu
@pytest.mark.parametrize("mk", SENSOR_SCENES)
def test_device_proximity_batch_matches_oracle(ctx, oracle, mk):
    s = mk()
    ctx.set_scene(s)
    fat = oracle.compute_aabbs(s)
    pairs = oracle.broad_phase(fat, s.groups)
    rng = np.random.default_rng(9)
    extra = rng.integers(0, s.n, size=(5000, 2)).astype(np.uint32)
    pairs = np.concatenate([pairs, extra[extra[:, 0] != extra[:, 1]]])
    assert np.array_equal(ctx.proximity(pairs), oracle.proximity(s, pairs)), s.name
    margins = rng.uniform(0, 2.0, size=len(pairs)).astype(F)
    got, want = ctx.proximity(pairs, margins), oracle.proximity(s, pairs, margins)
    assert np.array_equal(got, want), f"{s.name}: {int((got != want).sum())} statuses differ"
    assert set(want.tolist()) >= {0, 1, 2}


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(6))
def test_device_proximity_on_adversarial_scenes(ctx, oracle, k):
    """Coincident objects, exactly touching lattices, far-away coordinates, a dense clump, mixed scales: every broad-phase pair in
    both orders as a batch, and the world update with every second object a sensor."""
    from test_gpu_parity import _adversarial_scenes, canon, compare_manifolds

    s = _adversarial_scenes()[k]
    ctx.set_scene(s)
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), s.groups)
    pairs = np.concatenate([pairs, pairs[:, ::-1]])
    for margins in (None, np.zeros(len(pairs), dtype=F), np.full(len(pairs), 0.5, dtype=F)):
        got, want = ctx.proximity(pairs, margins), oracle.proximity(s, pairs, margins)
        assert np.array_equal(got, want), (s.name, int((got != want).sum()))
    s.query_kind = (np.arange(s.n) % 2).astype(np.uint8)
    res = ctx.world_update(s)
    assert res.counts["epa_overflow"] == 0
    assert np.array_equal(canon(res.pairs), canon(oracle.broad_phase(oracle.compute_aabbs(s), s.groups, mode=1)))
    compare_manifolds(res, s, oracle, s.name + "/sensors")


@pytest.mark.gpu
def test_device_proximity_reference_example(ctx):
    for pos, st in {(1, 1, 1): INTERSECTING, (2, 2, 2): WITHIN_MARGIN, (3, 3, 3): DISJOINT}.items():
        s = two_shapes(BALL, [1, 0, 0, 0], pos, CUBOID, [1, 1, 1, 0], (0, 0, 0))
        ctx.set_scene(s)
        assert ctx.proximity([[0, 1]], [1.0])[0] == st
        assert ctx.proximity([[1, 0]], [1.0])[0] == st


@pytest.mark.gpu
@pytest.mark.parametrize("mk", SENSOR_SCENES)
def test_world_update_with_sensors_matches_oracle(ctx, oracle, mk):
    from test_gpu_parity import canon, compare_manifolds

    s = mk()
    ctx.set_hulls(s.hulls)
    for how in ("host", "device"):
        if how == "host":
            res = ctx.world_update(s)
        else:
            ctx.set_scene(s)
            res = ctx.world_fetch(ctx.world_update_device(s.margin))
        assert res.counts["epa_overflow"] == 0
        fat = oracle.compute_aabbs(s)
        want = oracle.broad_phase(fat, s.groups, mode=1)
        assert np.array_equal(canon(res.pairs), canon(want)), "sensors must not change the pair set"
        assert np.all(res.pairs[:, 0] > res.pairs[:, 1])
        compare_manifolds(res, s, oracle, f"{s.name}/{how}")
        sensor = (s.query_kind[res.pairs[:, 0]] | s.query_kind[res.pairs[:, 1]]).astype(bool)
        planes = (s.shape_type[res.pairs[:, 0]] == PLANE) & (s.shape_type[res.pairs[:, 1]] == PLANE)
        assert np.array_equal(res.pair_algo == 6, sensor & ~planes)
        assert res.counts["n_algo"]["proximity"] == int((res.pair_algo == 6).sum())
        assert sum(res.counts["n_algo"].values()) == len(res.pairs)
        p = res.proximity[res.pair_algo == 6]
        assert [res.counts["n_proximity"][k] for k in ("intersecting", "within_margin", "disjoint")] == [int((p == k).sum()) for k in range(3)]
    # the same context goes back to a sensor-free world
    s.query_kind = None
    res = ctx.world_update(s)
    assert res.proximity is None and res.counts["n_algo"]["proximity"] == 0
    compare_manifolds(res, s, oracle, f"{s.name}/no sensors")


@pytest.mark.gpu
def test_collision_world_mirror_with_sensor(oracle):
    from ncollide_b200.shapes import Ball, Cuboid
    from ncollide_b200.world import CollisionWorld, GeometricQueryType, Proximity

    w = CollisionWorld(0.02)
    ident = (0, 0, 0, 1)
    cube = w.add(((0, 0, 0), ident), Cuboid((1, 1, 1)), query_type=GeometricQueryType.Contacts(0.02, 0.0))
    s1 = w.add(((1.2, 0, 0), ident), Ball(0.5), query_type=GeometricQueryType.Proximity(0.25))   # intersecting
    s2 = w.add(((0, 1.7, 0), ident), Ball(0.5), query_type=GeometricQueryType.Proximity(0.25))   # within margin
    # margin of a pair = sum of the two query limits = 0.25 + 0.02 (narrow_phase.rs:138): 0.28 away is Disjoint, the boxes still meet
    s3 = w.add(((0, 0, -1.78), ident), Ball(0.5), query_type=GeometricQueryType.Proximity(0.25))
    b = w.add(((-1.4, 0, 0), ident), Ball(0.5), query_type=GeometricQueryType.Contacts(0.02, 0.0))  # a real contact
    w.update()
    prox = {(a, c): st for a, c, st in w.proximity_pairs(effective_only=False)}
    assert prox[(s1, cube)] == Proximity.Intersecting and prox[(s2, cube)] == Proximity.WithinMargin and prox[(s3, cube)] == Proximity.Disjoint
    assert [(a, c) for a, c, _ in w.proximity_pairs(effective_only=True)] == [(s1, cube)]
    assert sorted(w.proximity_events()) == sorted([(s1, cube, Proximity.Disjoint, Proximity.Intersecting), (s2, cube, Proximity.Disjoint, Proximity.WithinMargin)])
    contacts = list(w.contact_pairs())
    assert len(contacts) == 1 and contacts[0][:2] == (b, cube)


def _sim_compare(dev, orc):
    """Per step: same pairs / orientation / algorithm / proximity status; ProximityEvents equal as sorted rows; contact side
    through the stepping-world comparison of tests/test_bp_persistent.py."""
    from test_bp_persistent import compare_sim_logs

    compare_sim_logs(dev, orc)
    n_ev = 0
    for t, (d, o) in enumerate(zip(dev, orc)):
        assert np.array_equal(d["prox"], o["prox"]), f"step {t}: proximity statuses differ on {int((d['prox'] != o['prox']).sum())} pairs"
        de = d["prox_events"][np.lexsort(d["prox_events"].T[::-1])] if len(d["prox_events"]) else d["prox_events"]
        oe = o["prox_events"][np.lexsort(o["prox_events"].T[::-1])] if len(o["prox_events"]) else o["prox_events"]
        assert np.array_equal(de, oe), f"step {t}: proximity events differ ({len(de)} vs {len(oe)})"
        n_ev += len(oe) if t > 0 else 0
    return n_ev


@pytest.mark.gpu
@pytest.mark.parametrize("n,kinds,side,plane,seed,frac", [(1500, (1, 1, 1), 6.5, False, 3, 0.3), (4000, (1, 1, 1), 9.5, True, 4, 0.5),
                                                          (2500, (0, 1, 1), 7.0, False, 5, 1.0)])
def test_stepping_world_with_sensors_matches_oracle(oracle, n, kinds, side, plane, seed, frac):
    from ncollide_b200.world import Context
    from sim_scenario import drive
    from test_bp_persistent import DeviceSimAdapter

    s = with_sensors(make_world_scene(n, 70 + seed, kinds, side=side, n_hulls=32, plane=plane, name="sim_sensors"), frac, seed, margin=0.12)
    dev = drive(DeviceSimAdapter(Context(0), s), s, steps=7, seed=seed)
    orc = drive(oracle.sim(s), s, steps=7, seed=seed)
    assert _sim_compare(dev, orc) > 10  # status changes after the first step: the warm-started detectors and the stop events


@pytest.mark.gpu
def test_stepping_world_add_remove_with_sensors_matches_oracle(oracle):
    from ncollide_b200.world import Context
    from sim_scenario import drive_add_remove
    from test_bp_persistent import DeviceSimAdapterAR

    n, seed = 1500, 11
    side = 5.5 * (n / 800.0) ** (1 / 3)
    s = with_sensors(make_world_scene(n, 80, (1, 1, 1), side=side, n_hulls=16, name="sim_addrm_sensors"), 0.3, 5, margin=0.1)
    extra = with_sensors(make_world_scene(n // 6, 81, (1, 1, 1), side=side, hull_library=s.hulls, name="extra"), 0.5, 6, margin=0.1)
    dev = drive_add_remove(DeviceSimAdapterAR(Context(0), s), s, extra, steps=7, seed=seed)
    orc = drive_add_remove(oracle.sim(s), s, extra, steps=7, seed=seed)
    assert np.array_equal(dev[3]["new_handles"], orc[3]["new_handles"])
    _sim_compare(dev, orc)
    # sensors added to a world that had none
    s2 = make_world_scene(n, 82, (1, 1, 1), side=side, n_hulls=16, name="no_sensors_then_some")
    extra2 = with_sensors(make_world_scene(n // 6, 83, (1, 1, 1), side=side, hull_library=s2.hulls, name="extra2"), 0.6, 7, margin=0.1)
    dev = drive_add_remove(DeviceSimAdapterAR(Context(0), s2), s2, extra2, steps=7, seed=seed)
    orc = drive_add_remove(oracle.sim(s2), s2, extra2, steps=7, seed=seed)
    _sim_compare(dev, orc)
    assert sum(int((r["algo"] == 6).sum()) for r in orc) > 0
