"""Capsules in the ORACLE (groundwork for SURVEY.md §8f N3; no device code yet): Capsule AABB / support map, the Segment
ConvexPolyhedron, CapsuleCapsule / CapsuleShape manifold generators with the capsule's ContactPreprocessor, capsule ray cast, point
containment and proximity.  The reference has no known-answer test that involves a capsule, so this restatement is
"parity unpinned"; it is checked here against closed-form geometry: segment-point, segment-plane and segment-segment distances."""
import numpy as np
import pytest

from ncollide_b200.scenes import WorldScene, random_unit_quaternions
from ncollide_b200.shapes import BALL, CUBOID, PLANE, HullLibrary

CAPSULE = 4
F = np.float32
IDENT = (0, 0, 0, 1)


def scene_of(objs, ql=0.0, dtype=F):
    """objs: list of (type, param, pos, rot)"""
    n = len(objs)
    param = np.zeros((n, 4), dtype=dtype)
    for i, o in enumerate(objs):
        param[i, : len(o[1])] = o[1]
    return WorldScene(
        pos=np.array([o[2] for o in objs], dtype=dtype), rot=np.array([o[3] if len(o) > 3 else IDENT for o in objs], dtype=dtype),
        shape_type=np.array([o[0] for o in objs], dtype=np.uint32), shape_param=param, groups=None,
        query_limit=np.full(n, ql, dtype=dtype), ang_pred=np.zeros(n, dtype=dtype), hulls=HullLibrary([]), margin=0.0,
    )


def rot_matrix(q):
    i, j, k, w = [float(x) for x in q]
    return np.array([[1 - 2 * (j * j + k * k), 2 * (i * j - k * w), 2 * (i * k + j * w)],
                     [2 * (i * j + k * w), 1 - 2 * (i * i + k * k), 2 * (j * k - i * w)],
                     [2 * (i * k - j * w), 2 * (j * k + i * w), 1 - 2 * (i * i + j * j)]])


def segment_ends(pos, q, hh):
    u = rot_matrix(q)[:, 1]
    p = np.asarray(pos, dtype=np.float64)
    return p - hh * u, p + hh * u


def seg_point_dist(a, b, p):
    ab = b - a
    t = np.clip(np.dot(p - a, ab) / np.dot(ab, ab), 0, 1)
    return np.linalg.norm(p - (a + t * ab))


def seg_seg_dist(a0, a1, b0, b1, n=400):
    # dense sampling + local refinement is enough for a 1e-3 check
    s = np.linspace(0, 1, n)
    pa = a0[None] + s[:, None] * (a1 - a0)[None]
    return min(seg_point_dist(b0, b1, p) for p in pa)


def test_capsule_aabb(oracle):
    rng = np.random.default_rng(1)
    q = random_unit_quaternions(rng, 50)
    objs = [(CAPSULE, [0.7, 0.3], (1, 2, 3), IDENT)] + [(CAPSULE, [rng.uniform(0.1, 1), rng.uniform(0.05, 0.5)], rng.uniform(-5, 5, 3), q[i]) for i in range(50)]
    s = scene_of(objs)
    bb = oracle.compute_aabbs(s, mode=0)
    assert np.allclose(bb[0], [1 - 0.3, 2 - 1.0, 3 - 0.3, 1 + 0.3, 2 + 1.0, 3 + 0.3], atol=1e-6)
    for i in range(1, len(objs)):
        hh, r = s.shape_param[i, 0], s.shape_param[i, 1]
        u = np.abs(rot_matrix(s.rot[i])[:, 1])
        ext = hh * u + r
        assert np.allclose(bb[i, :3], s.pos[i] - ext, atol=2e-5) and np.allclose(bb[i, 3:], s.pos[i] + ext, atol=2e-5)


def test_capsule_ball_contact_matches_segment_point_distance(oracle):
    rng = np.random.default_rng(2)
    q = random_unit_quaternions(rng, 300)
    n_contact = 0
    for k in range(300):
        hh, rc, rb = rng.uniform(0.2, 1.0), rng.uniform(0.1, 0.4), rng.uniform(0.1, 0.5)
        cpos, bpos = rng.uniform(-0.4, 0.4, 3), rng.uniform(-0.9, 0.9, 3)
        a, b = segment_ends(cpos, q[k], hh)
        d = seg_point_dist(a, b, bpos)
        if d < 1e-3:
            continue
        for order in (0, 1):
            objs = [(CAPSULE, [hh, rc], cpos, q[k]), (BALL, [rb], bpos, IDENT)]
            s = scene_of(objs[::-1] if order else objs, ql=0.05)
            c, off, algo, _ = oracle.narrow_phase(s, [[0, 1]])
            want_depth = rc + rb - d
            if want_depth < -0.1 - 1e-4:
                assert len(c) == 0
                continue
            if want_depth <= -0.1 + 1e-4:
                continue
            assert len(c) == 1 and algo[0] == 8  # CapsuleShape
            n_contact += 1
            assert abs(c["depth"][0] - want_depth) < 2e-5
            ci, bi = (1, 0) if order else (0, 1)
            wc, wb = (c["world2"][0], c["world1"][0]) if order else (c["world1"][0], c["world2"][0])
            assert abs(seg_point_dist(a, b, wc.astype(np.float64)) - rc) < 2e-5        # on the capsule surface
            assert abs(np.linalg.norm(wb - bpos) - rb) < 2e-5                           # on the ball surface
            assert abs(np.linalg.norm(c["normal"][0]) - 1) < 1e-5
            assert abs(-np.dot(c["normal"][0], c["world2"][0] - c["world1"][0]) - c["depth"][0]) < 2e-5
            # the capsule side carries a renamed segment feature (Face 0 / 1 = the end caps, Face 2 = the cylinder)
            fc = c["f2"][0] if order else c["f1"][0]
            assert fc >> 30 == 2 and (fc & 0x3FFFFFFF) in (0, 1, 2)
    assert n_contact > 150, n_contact


def test_capsule_plane_contacts(oracle):
    up = [0, 1, 0, 0]
    # lying capsule (axis along x after a rotation about z by 90 deg), 0.25 above the plane: both end caps touch the prediction zone
    qz = (0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4))
    for order in (0, 1):
        objs = [(PLANE, up, (0, 0, 0), IDENT), (CAPSULE, [0.8, 0.3], (0.2, 0.35, -0.1), qz)]
        s = scene_of(objs[::-1] if order else objs, ql=0.05)
        c, off, algo, _ = oracle.narrow_phase(s, [[0, 1]])
        assert len(c) == 2 and algo[0] == 8
        assert np.allclose(c["depth"], 0.3 - 0.35, atol=1e-5)
        wp, wc = ("world2", "world1") if order else ("world1", "world2")
        assert np.allclose(c[wp][:, 1], 0.0, atol=1e-5) and np.allclose(c[wc][:, 1], 0.05, atol=1e-5)
        assert np.allclose(np.sort(c[wc][:, 0]), [0.2 - 0.8, 0.2 + 0.8], atol=1e-5)
        fc = c["f1"] if order else c["f2"]
        assert sorted((fc & 0x3FFFFFFF).tolist()) == [0, 1] and np.all(fc >> 30 == 2)  # Vertex(i) -> Face(i)
    # upright capsule well above the prediction distance: nothing
    s = scene_of([(PLANE, up, (0, 0, 0), IDENT), (CAPSULE, [0.8, 0.3], (0, 1.3, 0), IDENT)], ql=0.05)
    assert len(oracle.narrow_phase(s, [[0, 1]])[0]) == 0
    # upright, touching: only the lower end cap
    s = scene_of([(PLANE, up, (0, 0, 0), IDENT), (CAPSULE, [0.8, 0.3], (0, 1.0, 0), IDENT)], ql=0.05)
    c = oracle.narrow_phase(s, [[0, 1]])[0]
    assert len(c) == 1 and abs(c["depth"][0] - 0.1) < 1e-5


def test_capsule_capsule_deepest_contact_matches_segment_segment_distance(oracle):
    rng = np.random.default_rng(3)
    q = random_unit_quaternions(rng, 400)
    n_checked = 0
    for k in range(200):
        h1, h2, r1, r2 = rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0), rng.uniform(0.1, 0.3), rng.uniform(0.1, 0.3)
        p1, p2 = rng.uniform(-0.7, 0.7, 3), rng.uniform(-0.7, 0.7, 3)
        a0, a1 = segment_ends(p1, q[2 * k], h1)
        b0, b1 = segment_ends(p2, q[2 * k + 1], h2)
        d = seg_seg_dist(a0, a1, b0, b1)
        if d < 0.02:
            continue  # crossing segments: the sub-detector's EPA case, not a distance check
        s = scene_of([(CAPSULE, [h1, r1], p1, q[2 * k]), (CAPSULE, [h2, r2], p2, q[2 * k + 1])], ql=0.02)
        c, off, algo, _ = oracle.narrow_phase(s, [[0, 1]])
        want = r1 + r2 - d
        if want < -0.04 - 2e-3:
            assert len(c) == 0
            continue
        if want <= -0.04 + 2e-3:
            continue
        assert algo[0] == 7 and len(c) >= 1
        assert abs(c["depth"].max() - want) < 2e-3, (k, c["depth"], want)
        for i in range(len(c)):
            assert abs(seg_point_dist(a0, a1, c["world1"][i].astype(np.float64)) - r1) < 1e-4
            assert abs(seg_point_dist(b0, b1, c["world2"][i].astype(np.float64)) - r2) < 1e-4
            assert abs(-np.dot(c["normal"][i], c["world2"][i] - c["world1"][i]) - c["depth"][i]) < 1e-4
        n_checked += 1
    assert n_checked > 60
    # parallel capsules side by side (the reference clips two parallel edge features: whatever the count, every contact is exact)
    s = scene_of([(CAPSULE, [0.8, 0.25], (0, 0, 0), IDENT), (CAPSULE, [0.8, 0.25], (0.45, 0.1, 0), IDENT)], ql=0.02)
    c = oracle.narrow_phase(s, [[0, 1]])[0]
    assert len(c) >= 1 and np.allclose(c["depth"], 0.05, atol=1e-5) and np.allclose(c["normal"], [[1, 0, 0]] * len(c), atol=1e-6)
    assert np.allclose(c["world1"][:, 0], 0.25, atol=1e-6) and np.allclose(c["world2"][:, 0], 0.2, atol=1e-6)


def test_capsule_on_cuboid_face(oracle):
    qz = (0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4))  # axis along x
    for order in (0, 1):
        objs = [(CUBOID, [1, 0.5, 1], (0, 0, 0), IDENT), (CAPSULE, [0.6, 0.2], (0.1, 0.68, 0.2), qz)]
        s = scene_of(objs[::-1] if order else objs, ql=0.05)
        c, off, algo, _ = oracle.narrow_phase(s, [[0, 1]])
        assert algo[0] == 8 and len(c) == 2
        assert np.allclose(c["depth"], 0.2 - 0.18, atol=1e-5)
        wb, wc = ("world2", "world1") if order else ("world1", "world2")
        assert np.allclose(c[wb][:, 1], 0.5, atol=1e-5) and np.allclose(c[wc][:, 1], 0.48, atol=1e-5)
        assert np.allclose(np.sort(c[wc][:, 0]), [0.1 - 0.6, 0.1 + 0.6], atol=1e-4)


def test_capsule_ray_cast_point_query_and_proximity(oracle):
    s = scene_of([(CAPSULE, [0.8, 0.3], (0, 0, 0), IDENT), (BALL, [0.5], (2.0, 0.4, 0), IDENT)])
    hit = oracle.shape_ray_cast(s, 0, (-5, 0.2, 0), (1, 0, 0), 100.0)
    assert hit is not None and abs(hit[0] - 4.7) < 1e-3 and np.allclose(hit[1], [-1, 0, 0], atol=1e-3)
    hit = oracle.shape_ray_cast(s, 0, (0, 5, 0), (0, -1, 0), 100.0)
    assert hit is not None and abs(hit[0] - (5 - 1.1)) < 1e-3
    assert oracle.shape_ray_cast(s, 0, (-5, 1.2, 0), (1, 0, 0), 100.0) is None
    pts = np.array([[0.29, 0.5, 0], [0.31, 0.5, 0], [0, 1.09, 0], [0, 1.11, 0], [0.2, 0.95, 0.1]], dtype=F)
    assert oracle.shape_contains_point_batch(s, [0] * 5, pts).tolist() == [1, 0, 1, 0, 1]
    # capsule x ball proximity: surface distance 2.0 - 0.3 - 0.5 = 1.2
    for margin, want in ((1.3, 1), (1.1, 2)):
        assert oracle.proximity(s, [[0, 1]], [margin])[0] == want and oracle.proximity(s, [[1, 0]], [margin])[0] == want
    s2 = scene_of([(CAPSULE, [0.8, 0.3], (0, 0, 0), IDENT), (BALL, [0.5], (0.6, 0.4, 0), IDENT)])
    assert oracle.proximity(s2, [[0, 1]], [0.0])[0] == 0
    # plane x capsule: upright capsule whose lower cap is 0.2 above the plane
    s3 = scene_of([(PLANE, [0, 1, 0, 0], (0, 0, 0), IDENT), (CAPSULE, [0.8, 0.3], (0, 1.3, 0), IDENT)])
    assert [int(oracle.proximity(s3, [[0, 1]], [m])[0]) for m in (0.25, 0.15)] == [1, 2]


def test_world_with_capsules_is_consistent(oracle):
    """A mixed world (balls, cuboids, capsules, a plane) through the oracle's fresh-world path: every contact is a consistent
    (world1, world2, normal, depth) tuple and capsule pairs are dispatched to the two capsule generators."""
    rng = np.random.default_rng(5)
    n = 600
    q = random_unit_quaternions(rng, n)
    objs = []
    for i in range(n):
        t = (BALL, CUBOID, CAPSULE)[i % 3]
        param = {BALL: [rng.uniform(0.25, 0.5)], CUBOID: list(rng.uniform(0.25, 0.5, 3)), CAPSULE: [rng.uniform(0.2, 0.5), rng.uniform(0.15, 0.3)]}[t]
        objs.append((t, param, rng.uniform(0, 5.0, 3), q[i]))
    objs.append((PLANE, [0, 1, 0, 0], (0, 0, 0), IDENT))
    s = scene_of(objs, ql=0.02)
    s.margin = 0.02
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), None, mode=0)
    assert np.array_equal(np.unique(np.sort(pairs, axis=1), axis=0), np.unique(np.sort(oracle.broad_phase(oracle.compute_aabbs(s), None, mode=2), axis=1), axis=0))
    c, off, algo, _ = oracle.narrow_phase(s, pairs)
    t = s.shape_type
    cap = (t[pairs[:, 0]] == CAPSULE) | (t[pairs[:, 1]] == CAPSULE)
    both = (t[pairs[:, 0]] == CAPSULE) & (t[pairs[:, 1]] == CAPSULE)
    assert np.all(algo[both] == 7) and np.all(algo[cap & ~both] == 8) and np.all(algo[~cap] <= 5)
    assert cap.sum() > 200 and len(c) > 300
    assert np.allclose(np.linalg.norm(c["normal"], axis=1), 1, atol=1e-5)
    assert np.allclose(-np.einsum("ij,ij->i", c["normal"], c["world2"] - c["world1"]), c["depth"], atol=2e-4)
    assert np.all(c["depth"] >= -0.0401 - 1e-5)


def test_stepping_world_with_capsules(oracle):
    """The oracle's stepping world accepts capsules: step 1 equals the fresh-world narrow phase on the same pairs, later steps keep
    contact ids for resting capsule contacts."""
    import sys, os

    sys.path.insert(0, os.path.dirname(__file__))
    from sim_scenario import drive

    rng = np.random.default_rng(8)
    n = 300
    q = random_unit_quaternions(rng, n)
    objs = []
    for i in range(n):
        t = (BALL, CUBOID, CAPSULE)[i % 3]
        param = {BALL: [rng.uniform(0.25, 0.5)], CUBOID: list(rng.uniform(0.25, 0.5, 3)), CAPSULE: [rng.uniform(0.2, 0.5), rng.uniform(0.15, 0.3)]}[t]
        objs.append((t, param, rng.uniform(0, 4.0, 3), q[i]))
    s = scene_of(objs, ql=0.02)
    s.margin = 0.02
    s.groups = None
    log = drive(oracle.sim(s), s, steps=5, seed=4)
    r0 = log[0]
    c, off, algo, _ = oracle.narrow_phase(s, r0["pairs"])
    assert np.array_equal(algo, r0["algo"]) and np.array_equal(off, r0["off"]) and (algo >= 7).sum() > 50
    for name in ("world1", "world2", "normal", "depth", "f1", "f2"):
        assert np.array_equal(c[name], r0["contacts"][name])
    kept = 0
    for a, b in zip(log[:-1], log[1:]):
        ka = {(tuple(p), i) for p, lo, hi in zip(a["pairs"].tolist(), a["off"][:-1], a["off"][1:]) for i in a["ids"][lo:hi].tolist()}
        kb = {(tuple(p), i) for p, lo, hi in zip(b["pairs"].tolist(), b["off"][:-1], b["off"][1:]) for i in b["ids"][lo:hi].tolist()}
        kept += len(ka & kb)
    assert kept > 100


def test_capsule_golden_fixture(oracle):
    """tests/golden/capsule_mixed_plane_300.npz (made by tests/golden/make_golden.py): guards the capsule restatement against drift and
    is the fixture the device path will have to reproduce."""
    import os

    from golden.make_golden import scene_from_npz

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "capsule_mixed_plane_300.npz"))
    s = scene_from_npz(z)
    assert (s.shape_type == CAPSULE).sum() == 100
    fat = oracle.compute_aabbs(s)
    assert np.array_equal(fat, z["fat_aabbs"])
    pairs = oracle.broad_phase(fat, s.groups, 0)
    assert np.array_equal(pairs, z["pairs"])
    c, off, algo, _ = oracle.narrow_phase(s, pairs)
    assert np.array_equal(off, z["manifold_off"]) and np.array_equal(algo, z["algo"]) and (algo >= 7).sum() > 100
    for name in ("world1", "world2", "normal", "depth", "f1", "f2"):
        assert np.array_equal(c[name], z["c_" + name]), name
