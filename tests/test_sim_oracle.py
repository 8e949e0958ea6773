"""Stepping world (CollisionWorld::update over several steps), oracle side.  CPU only."""
import numpy as np

from ncollide_b200.scenes import make_world_scene
from ncollide_b200.shapes import CUBOID
from sim_scenario import drive, drive_add_remove, step_poses
from test_oracle_kat import scene_of

F32 = np.float32


def test_two_colliding_cuboids_terminates(oracle):
    # build/ncollide3d/tests/pipeline/contact_pairs.rs:8-84: push two coincident unit cubes apart along the deepest
    # contact normal until no penetration is left; the test passes when the loop ends
    s = scene_of([(CUBOID, [1, 1, 1], (0, 0, 0)), (CUBOID, [1, 1, 1], (0, 0, 0))], margin=0.0)
    sim = oracle.sim(s)
    pos, rot = s.pos.copy(), s.rot.copy()
    for it in range(50):
        r = sim.step()
        moved = {}
        for p, (h1, h2) in enumerate(r["pairs"].tolist()):
            cs = r["contacts"][r["off"][p] : r["off"][p + 1]]
            if len(cs) == 0 or h1 in moved or h2 in moved:
                continue
            c = cs[np.argmax(cs["depth"])]
            if c["depth"] == 0:
                continue
            n = c["depth"] * c["normal"]
            if (n > 0).any():
                moved[h1] = -n
            else:
                moved[h2] = n
        if not moved:
            break
        for h, v in moved.items():
            pos[h] += v
            sim.set_positions([h], pos[h : h + 1], rot[h : h + 1])
    assert it < 10
    assert it >= 1  # the first update does report the interpenetration


def test_first_step_is_the_fresh_world_update(oracle):
    s = make_world_scene(3000, 21, (1, 1, 1), side=9.0, n_hulls=32, plane=True, name="sim_first")
    r = oracle.sim(s).step()
    fat = oracle.compute_aabbs(s)
    pairs = oracle.broad_phase(fat, s.groups, mode=0)
    assert r["bp_pairs"] == len(pairs)
    order = np.lexsort((np.maximum(pairs[:, 0], pairs[:, 1]), np.minimum(pairs[:, 0], pairs[:, 1])))
    pairs = pairs[order]
    c, off, algo, _ = oracle.narrow_phase(s, pairs)
    keep = algo != 0
    assert np.array_equal(r["pairs"], pairs[keep])  # same orientation (larger handle first)
    assert np.array_equal(r["algo"], algo[keep])
    assert np.array_equal(np.diff(r["off"]), np.diff(off)[keep])
    assert r["contacts"].tobytes() == c.tobytes()
    started = r["events"][r["events"][:, 2] == 1]
    assert len(started) == int(np.sum(np.diff(r["off"]) > 0)) and np.all(r["events"][:, 2] == 1)


def test_stepping_invariants(oracle):
    s = make_world_scene(1500, 22, (1, 1, 1), side=7.0, n_hulls=24, name="sim_steps")
    log = drive(oracle.sim(s), s, steps=6, seed=3)
    # an object set that did not move keeps its manifolds bit for bit (narrow_phase.rs:181 skips the pair)
    for a, b in zip(log[:-1], log[1:]):
        quiet = ~(b["moved"][b["pairs"][:, 0]] | b["moved"][b["pairs"][:, 1]])
        ka = {tuple(p): i for i, p in enumerate(a["pairs"].tolist())}
        same = 0
        for i in np.nonzero(quiet)[0]:
            j = ka.get(tuple(b["pairs"][i].tolist()))
            if j is None:
                continue
            ca = a["contacts"][a["off"][j] : a["off"][j + 1]]
            cb = b["contacts"][b["off"][i] : b["off"][i + 1]]
            assert ca.tobytes() == cb.tobytes()
            assert np.array_equal(a["ids"][a["off"][j] : a["off"][j + 1]], b["ids"][b["off"][i] : b["off"][i + 1]])
            same += 1
        assert same > 10
    # events: Started / Stopped exactly when a pair's manifold becomes non-empty / empty (or the pair disappears)
    for a, b in zip(log[:-1], log[1:]):
        na = {tuple(p): int(a["off"][i + 1] - a["off"][i]) for i, p in enumerate(a["pairs"].tolist())}
        nb = {tuple(p): int(b["off"][i + 1] - b["off"][i]) for i, p in enumerate(b["pairs"].tolist())}
        want = set()
        for p in set(na) | set(nb):
            x, y = na.get(p, 0), nb.get(p, 0)
            if x == 0 and y > 0:
                want.add((p[0], p[1], 1))
            if x > 0 and y == 0:
                want.add((p[0], p[1], 0))
        assert set(map(tuple, b["events"].tolist())) == want
    # persisting contact ids exist (resting contacts are matched through the manifold cache)
    kept = 0
    for a, b in zip(log[:-1], log[1:]):
        ia = {(tuple(a["pairs"][p]), int(i)) for p in range(len(a["pairs"])) for i in a["ids"][a["off"][p] : a["off"][p + 1]]}
        ib = {(tuple(b["pairs"][p]), int(i)) for p in range(len(b["pairs"])) for i in b["ids"][b["off"][p] : b["off"][p + 1]]}
        kept += len(ia & ib)
    assert kept > 100


def test_add_and_remove_between_updates(oracle):
    # CollisionWorld::remove / add (world.rs:64-96,129-144): handles are recycled last-freed-first, removed objects take their
    # pairs with them without contact events, new objects interact from the next update on
    s = make_world_scene(800, 23, (1, 1, 1), side=5.5, n_hulls=16, name="sim_addrm")
    extra = make_world_scene(120, 24, (1, 1, 1), side=5.5, hull_library=s.hulls, name="extra")
    log = drive_add_remove(oracle.sim(s), s, extra, steps=7, seed=9)
    gone = set(range(s.n)) - set(np.unique(log[2]["pairs"]).tolist()) - set()
    assert len(gone) >= s.n // 10
    nh = log[3]["new_handles"]
    assert len(set(nh.tolist())) == extra.n and (nh < s.n).sum() == s.n // 10 and nh.max() == s.n + extra.n - s.n // 10 - 1
    assert np.isin(nh, log[3]["pairs"]).sum() > 30  # the new objects collide
    # no Stopped event names a removed object
    removed_first = set(range(s.n)) - set(np.unique(np.concatenate([log[1]["pairs"].ravel(), nh])).tolist())
    for ev in log[2]["events"].tolist():
        assert ev[0] not in removed_first and ev[1] not in removed_first


def _load_sim_golden():
    import os

    from golden.make_golden import scene_from_npz

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "sim_mixed_plane_400.npz"))
    return z, scene_from_npz(z)


def check_against_sim_golden(sim, z, s, exact):
    """Replays tests/golden/sim_mixed_plane_400.npz (5 updates + world queries, made by make_golden.py from the oracle)."""
    log = drive(sim, s, steps=5, seed=1010)
    for t, r in enumerate(log):
        for k in ("pairs", "algo", "off", "ids"):
            assert np.array_equal(r[k], z[f"s{t}_{k}"]), (t, k)
        ev = lambda e: e[np.lexsort((e[:, 1], e[:, 0], e[:, 2]))] if len(e) else e
        assert np.array_equal(ev(r["events"]), ev(z[f"s{t}_events"])), (t, "events")
        for n in ("f1", "f2"):
            assert np.array_equal(r["contacts"][n], z[f"s{t}_c_{n}"]), (t, n)
        for n in ("world1", "world2", "normal", "depth"):
            if exact:
                assert np.array_equal(r["contacts"][n], z[f"s{t}_c_{n}"]), (t, n)
            else:
                assert np.allclose(r["contacts"][n], z[f"s{t}_c_{n}"], rtol=1e-4, atol=1e-5), (t, n)
    idx, toi, normal, feat = sim.ray_cast(z["q_ro"], z["q_rd"], 30.0)
    assert np.array_equal(idx, z["q_idx"]) and np.array_equal(feat, z["q_feat"])
    assert np.allclose(toi, z["q_toi"], rtol=1e-4, atol=1e-5) and np.allclose(normal, z["q_normal"], rtol=1e-4, atol=1e-5)
    idx1, toi1, _, _ = sim.ray_cast(z["q_ro"], z["q_rd"], 30.0, first_only=True)
    assert np.array_equal(idx1, z["q_first_idx"]) and np.allclose(toi1, z["q_first_toi"], rtol=1e-4, atol=1e-5)
    assert np.array_equal(sim.query(2, z["q_pts"]), z["q_point_rows"])


def test_sim_golden_fixture_oracle(oracle):
    z, s = _load_sim_golden()
    check_against_sim_golden(oracle.sim(s), z, s, exact=True)
