"""Persistent broad phase (BroadPhase trait over several updates).

CPU: the reference-faithful oracle (two DBVTs, slab, pending queue, purge — oracle/bp_persistent.cpp) is pinned on the
reference's own example and checked against the tree-free set semantics the device path relies on.
GPU: the device broad phase (ncb_bp_*) against the oracle, event by event."""
import numpy as np
import pytest

from bp_scenario import DeviceAdapter, OracleAdapter, SetModel, assert_same_log, run_scenario

F32 = np.float32


def ball_box(c, r=0.5):
    c = np.asarray(c, dtype=F32)
    return np.concatenate([c - F32(r), c + F32(r)])


def example_kat(make):
    # build/ncollide3d/examples/dbvt_broad_phase3d.rs:38-60
    bp = make(0.02)
    hs = bp.create([ball_box(p) for p in [(0, 0, 0), (0, 0.5, 0), (0.5, 0, 0), (0.5, 0.5, 0)]], [(0x3FFFFFFF, 0x3FFFFFFF, 0)] * 4)
    assert hs == [0, 1, 2, 3]
    assert bp.proxy(0) is None  # not attached before the first update
    started, stopped = bp.update()
    assert bp.num() == 6 and len(started) == 6 and len(stopped) == 0
    assert all(a > b for a, b in started.tolist())  # the proxy inserted later comes first
    gone = bp.remove(np.array([0, 1], dtype=np.uint32))
    assert len(gone) == 5
    started, stopped = bp.update()
    assert bp.num() == 1 and len(started) == 0 and len(stopped) == 0
    # create_proxy stores the box as given; deferred_set_bounding_volume stores it loosened by the margin
    assert np.array_equal(bp.proxy(2), ball_box((0.5, 0, 0)))
    moved = ball_box((0.5, 0.0, 3.0))
    bp.set_bvs(np.array([2], dtype=np.uint32), moved.reshape(1, 6))
    started, stopped = bp.update()
    assert stopped.tolist() == [[2, 3]] and bp.num() == 0
    want = moved.copy()
    want[:3] += -F32(0.02)
    want[3:] += F32(0.02)
    assert np.array_equal(bp.proxy(2), want)
    # a move that stays inside the stored box changes nothing
    bp.set_bvs(np.array([2], dtype=np.uint32), (moved + F32(0.01)).reshape(1, 6))
    bp.update()
    assert np.array_equal(bp.proxy(2), want)
    # freed handles are reused, last freed first
    assert bp.create([ball_box((9, 9, 9))], [(0x3FFFFFFF, 0x3FFFFFFF, 0)]) == [1]
    assert bp.create([ball_box((9, 9, 9))], [(0x3FFFFFFF, 0x3FFFFFFF, 0)]) == [0]
    started, _ = bp.update()
    assert started.tolist() == [[0, 1]] and bp.num() == 1
    with pytest.raises(RuntimeError):
        bp.set_bvs(np.array([17], dtype=np.uint32), moved.reshape(1, 6))


def test_oracle_example_kat(oracle):
    example_kat(lambda m: OracleAdapter(oracle, m))


def test_set_model_example_kat():
    example_kat(SetModel)


@pytest.mark.parametrize("seed,groups", [(1, True), (2, False), (3, True)])
def test_oracle_matches_set_semantics(oracle, seed, groups):
    a = run_scenario(OracleAdapter(oracle, 0.05), seed, n0=250, steps=11, use_groups=groups)
    b = run_scenario(SetModel(0.05), seed, n0=250, steps=11, use_groups=groups)
    assert sum(len(r["started"]) for r in a) > 300 and sum(len(r["stopped"]) for r in a) > 30
    assert_same_log(a, b, "oracle vs set model")


def test_oracle_queries_match_brute_force(oracle):
    # interferences_with_{bounding_volume,ray,point} through the two DBVTs == a scan of the stored boxes
    a = run_scenario(OracleAdapter(oracle, 0.05), 4, n0=200, steps=6, n_queries=24)
    b = run_scenario(SetModel(0.05), 4, n0=200, steps=6, n_queries=24)
    assert sum(len(r["q_ray"]) for r in a) > 50 and sum(len(r["q_point"]) for r in a) > 5
    assert_same_log(a, b, "oracle vs set model")


@pytest.mark.gpu
def test_device_example_kat():
    from ncollide_b200.world import Context

    ctx = Context(0)
    example_kat(lambda m: DeviceAdapter(ctx, m))


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n0,steps,groups", [(1, 250, 11, True), (2, 250, 11, False), (5, 4000, 8, True), (6, 20000, 6, False)])
def test_device_matches_oracle(oracle, seed, n0, steps, groups):
    from ncollide_b200.world import Context

    ctx = Context(0)
    side = 10.0 * (n0 / 300.0) ** (1 / 3)
    a = run_scenario(OracleAdapter(oracle, 0.05), seed, n0=n0, steps=steps, side=side, use_groups=groups, n_queries=64)
    b = run_scenario(DeviceAdapter(ctx, 0.05), seed, n0=n0, steps=steps, side=side, use_groups=groups, n_queries=64)
    assert_same_log(a, b, "device vs oracle")


@pytest.mark.gpu
def test_device_handler_mirror():
    """BroadPhase.update(handler) replays the events through the reference's handler trait."""
    from ncollide_b200.world import BroadPhase, BroadPhaseInterferenceHandler, Context

    class H(BroadPhaseInterferenceHandler):
        def __init__(self):
            self.started, self.stopped = [], []

        def is_interference_allowed(self, a, b):
            return not (a == "c" or b == "c")

        def interference_started(self, a, b):
            self.started.append((a, b))

        def interference_stopped(self, a, b):
            self.stopped.append((a, b))

    bp = BroadPhase(0.02, ctx=Context(0))
    ha = bp.create_proxy(ball_box((0, 0, 0)), "a")
    hb = bp.create_proxy(ball_box((0.5, 0, 0)), "b")
    hc = bp.create_proxy(ball_box((0, 0.5, 0)), "c")
    h = H()
    bp.update(h)
    assert h.started == [("b", "a")] and bp.num_interferences() == 1
    bp.deferred_set_bounding_volume(hb, ball_box((5, 0, 0)))
    bp.deferred_set_bounding_volume(hc, ball_box((5, 0.5, 0)))
    bp.update(h)
    assert h.stopped == [("a", "b")] and bp.num_interferences() == 0
    assert bp.proxy(ha)[1] == "a"
    with pytest.raises(Exception):
        bp.deferred_set_bounding_volume(99, ball_box((0, 0, 0)))
        bp.update(h)


@pytest.mark.gpu
def test_user_pair_filter_over_steps():
    """An arbitrary BroadPhasePairFilter (broad_phase_pair_filter.rs:5-16, applied through is_interference_allowed,
    glue/update.rs:29-41) on top of the device broad phase, over random moves / removals: after every update the pairs the
    handler has been told about (started minus stopped) are exactly {stored boxes intersect AND the filter allows the pair},
    and num_interferences / pairs() agree."""
    from ncollide_b200.world import BroadPhase, BroadPhaseInterferenceHandler, Context

    rng = np.random.default_rng(77)

    def allowed(a, b):  # deterministic, symmetric, vetoes about a third of the pairs
        lo, hi = min(a, b), max(a, b)
        return (lo * 2654435761 + hi * 40503) % 3 != 0

    class H(BroadPhaseInterferenceHandler):
        def __init__(self):
            self.live = set()
            self.calls = 0

        def is_interference_allowed(self, a, b):
            self.calls += 1
            return allowed(a, b)

        def interference_started(self, a, b):
            key = (min(a, b), max(a, b))
            assert allowed(a, b) and key not in self.live
            self.live.add(key)

        def interference_stopped(self, a, b):
            key = (min(a, b), max(a, b))
            assert key in self.live, "a vetoed pair must never be reported as stopped"
            self.live.discard(key)

    bp = BroadPhase(0.05, ctx=Context(0))
    n, side = 600, 9.0
    centres = (rng.random((n, 3)) * side).astype(F32)
    handles = bp.create_proxies(np.stack([ball_box(c) for c in centres]), list(range(n)))  # user data = index
    # user data must be the HANDLE here (the filter sees the data): handles are 0..n-1 in creation order on a fresh broad phase
    assert handles.tolist() == list(range(n))
    alive = set(range(n))
    h = H()
    for step in range(6):
        bp.update(h)
        ids = sorted(alive)
        boxes = np.stack([bp.proxy(i)[0] for i in ids])
        want = set()
        for ai, a in enumerate(ids):
            hit = np.all(boxes[ai, :3] <= boxes[:, 3:], axis=1) & np.all(boxes[:, :3] <= boxes[ai, 3:], axis=1)
            for bi in np.nonzero(hit)[0]:
                b = ids[bi]
                if a < b and allowed(a, b):
                    want.add((a, b))
        assert h.live == want, f"step {step}: {len(h.live ^ want)} pairs differ"
        assert bp.num_interferences() == len(want)
        got = {(min(a, b), max(a, b)) for a, b in bp.pairs().tolist()}
        assert got == want
        # move a third of the proxies, remove a few
        move = rng.choice(ids, size=len(ids) // 3, replace=False)
        centres[move] += rng.normal(0, 0.6, size=(len(move), 3)).astype(F32)
        bp.deferred_set_bounding_volumes(move, np.stack([ball_box(centres[i]) for i in move]))
        gone = rng.choice(ids, size=10, replace=False)

        class R:
            def __init__(self, live):
                self.live = live

            def __call__(self, a, b):
                self.live.discard((min(a, b), max(a, b)))

        bp.remove(gone, R(h.live))
        alive -= set(int(g) for g in gone)
    assert h.calls > 0 and len(h.live) > 100
    bp.close()


# ---- stepping world (persistent narrow phase on top of the persistent broad phase) -----------------------------------
RTOL, ATOL = 1e-4, 1e-5


class DeviceSimAdapter:
    def __init__(self, ctx, scene):
        from ncollide_b200.world import SteppingWorld

        self.w = SteppingWorld(ctx, scene)

    def set_positions(self, handles, pos, rot):
        self.w.set_positions(handles, pos, rot)

    def step(self):
        r = self.w.update()
        keep = r["algo"] != 0  # plane x plane: a broad-phase pair without an interaction edge
        cnt = np.diff(r["off"].astype(np.int64))
        assert np.array_equal(cnt, r["count"].astype(np.int64))
        sel = np.repeat(keep, cnt)
        off = np.concatenate([[0], np.cumsum(cnt[keep])]).astype(np.uint32)
        return {"pairs": r["pairs"][keep], "algo": r["algo"][keep], "off": off, "contacts": r["contacts"][sel], "ids": r["ids"][sel],
                "events": r["events"], "counts": r["counts"], "bp_pairs": len(r["pairs"]), "prox": r["prox"][keep], "prox_events": r["prox_events"]}


def compare_sim_logs(dev, orc):
    assert len(dev) == len(orc)
    for t, (a, b) in enumerate(zip(dev, orc)):
        assert a["bp_pairs"] == b["bp_pairs"], f"step {t}: broad-phase pair count"
        assert np.array_equal(a["pairs"], b["pairs"]), f"step {t}: pairs / orientation"
        assert np.array_equal(a["algo"], b["algo"]), f"step {t}: algo"
        assert np.array_equal(a["off"], b["off"]), f"step {t}: manifold sizes ({int(np.sum(a['off'] != b['off']))} differ)"
        assert np.array_equal(a["ids"], b["ids"]), f"step {t}: contact ids"
        for f in ("f1", "f2"):
            assert np.array_equal(a["contacts"][f], b["contacts"][f]), f"step {t}: {f}"
        for f in ("world1", "world2", "normal", "depth"):
            assert np.allclose(a["contacts"][f], b["contacts"][f], rtol=RTOL, atol=ATOL), f"step {t}: {f}"
        ea = a["events"][np.lexsort((a["events"][:, 1], a["events"][:, 0], a["events"][:, 2]))] if len(a["events"]) else a["events"]
        eb = b["events"][np.lexsort((b["events"][:, 1], b["events"][:, 0], b["events"][:, 2]))] if len(b["events"]) else b["events"]
        assert np.array_equal(ea, eb), f"step {t}: contact events"
        assert a["counts"]["epa_overflow"] == 0 and a["counts"]["ref_panics"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,kinds,side,plane,seed", [(1500, (1, 1, 1), 7.0, False, 3), (4000, (1, 1, 1), 10.0, True, 4), (3000, (0, 1, 1), 8.0, False, 5),
                                                      (20000, (1, 1, 1), 18.0, False, 6)])
def test_stepping_world_matches_oracle(oracle, n, kinds, side, plane, seed):
    from ncollide_b200.scenes import make_world_scene
    from ncollide_b200.world import Context
    from sim_scenario import drive

    s = make_world_scene(n, 20 + seed, kinds, side=side, n_hulls=32, plane=plane, angular=0.02 if seed == 5 else 0.0, name="sim")
    dev = drive(DeviceSimAdapter(Context(0), s), s, steps=7, seed=seed)
    orc = drive(oracle.sim(s), s, steps=7, seed=seed)
    assert sum(len(r["events"]) for r in orc[1:]) > 10
    compare_sim_logs(dev, orc)


@pytest.mark.gpu
def test_stepping_world_first_step_equals_fresh_update(oracle):
    from ncollide_b200.scenes import make_world_scene
    from ncollide_b200.world import Context, SteppingWorld

    s = make_world_scene(5000, 31, (1, 1, 1), side=11.0, n_hulls=32, plane=True, name="sim_first")
    ctx = Context(0)
    ctx.set_scene(s)
    fresh = ctx.world_update(s)
    r = SteppingWorld(ctx, s).update()
    order = np.lexsort((np.maximum(fresh.pairs[:, 0], fresh.pairs[:, 1]), np.minimum(fresh.pairs[:, 0], fresh.pairs[:, 1])))
    assert np.array_equal(r["pairs"], fresh.pairs[order])
    assert np.array_equal(r["algo"], fresh.pair_algo[order])
    assert np.array_equal(r["count"], fresh.manifold_count[order])
    got = r["contacts"]
    want = np.concatenate([fresh.contacts_of(p) for p in order]) if len(order) else fresh.contacts
    for f in ("world1", "world2", "normal", "depth", "f1", "f2"):
        assert np.array_equal(got[f], want[f]), f


class DeviceSimAdapterAR(DeviceSimAdapter):
    def remove(self, handles):
        self.w.remove(handles)

    def set_collision_groups(self, handles, groups):
        self.w.set_collision_groups(handles, groups)

    def add(self, scene):
        return self.w.add(scene)


@pytest.mark.gpu
@pytest.mark.parametrize("n,kinds,seed", [(900, (1, 1, 1), 12), (4000, (1, 1, 1), 13)])
def test_stepping_world_group_changes_match_oracle(oracle, n, kinds, seed):
    """CollisionObject::set_collision_groups on live objects (COLLISION_GROUPS_CHANGED -> broad-phase redispatch + narrow-phase
    update, glue/update.rs:76-87): pairs, orientation, manifolds, contact ids and events equal the oracle's step by step."""
    from ncollide_b200.scenes import make_world_scene
    from ncollide_b200.world import Context
    from sim_scenario import drive_group_changes

    s = make_world_scene(n, 40 + seed, kinds, side=7.0 * (n / 900.0) ** (1 / 3), n_hulls=32, name="sim_groups")
    dev = drive_group_changes(DeviceSimAdapterAR(Context(0), s), s, steps=7, seed=seed)
    orc = drive_group_changes(oracle.sim(s), s, steps=7, seed=seed)
    assert len(orc[1]["pairs"]) < len(orc[0]["pairs"]) and len(orc[2]["pairs"]) > len(orc[1]["pairs"])
    assert sum(len(r["events"]) for r in orc[1:]) > 10
    compare_sim_logs(dev, orc)


def test_oracle_group_changes_follow_the_filter(oracle):
    """ORACLE check (CPU): after every update of a world whose objects change groups, the edge set is exactly
    {stored broad-phase boxes intersect AND CollisionGroups::can_interact_with_groups} (collision_groups.rs:353-359)."""
    from ncollide_b200.scenes import make_world_scene
    from sim_scenario import drive_group_changes

    s = make_world_scene(700, 51, (1, 1, 1), side=6.5, n_hulls=16, name="sim_groups_cpu")
    log = drive_group_changes(oracle.sim(s), s, steps=7, seed=14)

    def allowed(g, a, b):
        if a == b:
            return bool(g[a][2] & g[a][0])  # never queried here
        return bool(g[a][0] & g[b][1]) and bool(g[b][0] & g[a][1]) and not (g[a][0] & g[b][2]) and not (g[b][0] & g[a][2])

    seen_drop = False
    for t, r in enumerate(log):
        g = r["groups"]
        for a, b in r["pairs"].tolist():
            assert allowed(g, a, b), f"step {t}: pair ({a}, {b}) is not allowed by its groups"
        if t > 0 and len(r["pairs"]) < len(log[0]["pairs"]):
            seen_drop = True
    assert seen_drop
    # stopped pairs that were touching produce ContactEvent::Stopped in the step their groups change
    assert any(len(r["events"]) and (np.asarray(r["events"])[:, 2] == 0).any() for r in log[1:])


@pytest.mark.gpu
@pytest.mark.parametrize("n,kinds,seed", [(800, (1, 1, 1), 9), (3000, (1, 2, 2), 10)])
def test_stepping_world_add_remove_matches_oracle(oracle, n, kinds, seed):
    from ncollide_b200.scenes import make_world_scene
    from ncollide_b200.world import Context
    from sim_scenario import drive_add_remove

    side = 5.5 * (n / 800.0) ** (1 / 3)
    s = make_world_scene(n, 23 + seed, kinds, side=side, n_hulls=16, name="sim_addrm")
    # seed 10: the added objects use a different angular prediction than the (uniform) world: the device table is expanded
    extra = make_world_scene(n // 6, 24 + seed, (1, 1, 1) if seed == 9 else (0, 1, 1), side=side, hull_library=s.hulls,
                             angular=0.03 if seed == 10 else 0.0, name="extra")
    dev = drive_add_remove(DeviceSimAdapterAR(Context(0), s), s, extra, steps=7, seed=seed)
    orc = drive_add_remove(oracle.sim(s), s, extra, steps=7, seed=seed)
    assert np.array_equal(dev[3]["new_handles"], orc[3]["new_handles"])
    compare_sim_logs(dev, orc)


@pytest.mark.gpu
def test_sim_golden_fixture_device():
    """The committed stepping-world fixture (5 updates + ray / point queries) reproduced by the device."""
    from ncollide_b200.world import Context
    from test_sim_oracle import _load_sim_golden, check_against_sim_golden

    z, s = _load_sim_golden()

    class A(DeviceSimAdapter):
        def ray_cast(self, *a, **k):
            return self.w.ray_cast(*a, **k)

        def query(self, *a, **k):
            return self.w.query(*a, **k)

    check_against_sim_golden(A(Context(0), s), z, s, exact=False)
