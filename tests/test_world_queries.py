"""World-level ray queries: glue::interferences_with_ray / first_interference_with_ray with the
per-shape ray casts.  CPU: the oracle against the reference's own tests; GPU: the device against the oracle."""
import numpy as np
import pytest

from ncollide_b200.scenes import make_world_scene
from ncollide_b200.shapes import BALL, CUBOID, HULL, PLANE, ConvexHull, HullLibrary
from test_oracle_kat import scene_of

F32 = np.float32
FMAX = float(np.finfo(F32).max)
FACE = 2 << 30


def unit_cube_hull():
    pts = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=F32)
    from scipy.spatial import ConvexHull as QH

    h = ConvexHull.try_from_points(pts) if hasattr(ConvexHull, "try_from_points") else None
    assert h is not None
    return HullLibrary([h])


def contains(kind, param, rot, p):
    # PointQuery::contains_point for ball (point_ball.rs:45-47) and cuboid (point_aabb.rs:153-156) at the origin
    if kind == BALL:
        return float(np.dot(p, p)) <= param[0] ** 2
    from scipy.spatial.transform import Rotation as R

    lp = R.from_quat(rot).inv().apply(p)
    return bool(np.all(np.abs(lp) <= np.array(param[:3]) + 0))


@pytest.mark.parametrize("name,kind,param", [("ball", BALL, [1.0]), ("cube", CUBOID, [1, 1, 1]), ("tall", CUBOID, [1, 1, 0.5]),
                                              ("slim", CUBOID, [0.5, 1, 0.5]), ("hull cube", HULL, [0])])
def test_shape_ray_cast_points_to_surface(oracle, name, kind, param):
    # build/ncollide3d/tests/geometry/cuboid_ray_cast.rs:7-84 (+ the same property for a ConvexHull cube through gjk::cast_ray)
    rng = np.random.default_rng(5)
    hulls = unit_cube_hull() if kind == HULL else None
    for it in range(300):
        v = rng.normal(size=3)
        origin = (v / np.linalg.norm(v) * 5.0).astype(F32)
        q = rng.normal(size=4)
        q = np.array([0, 0, 0, 1.0]) if rng.random() < 0.01 else q / np.linalg.norm(q)
        s = scene_of([(kind, param, (0, 0, 0), tuple(q.astype(F32)))], hulls=hulls)
        hit = oracle.shape_ray_cast(s, 0, origin, -origin, FMAX)
        assert hit is not None, name
        toi, normal, feat = hit
        point = origin + (-origin) * toi
        p_in, p_out = point + normal * F32(-0.001), point + normal * F32(0.001)
        ck, cp = (CUBOID, [1, 1, 1]) if kind == HULL else (kind, param)
        assert contains(ck, cp, s.rot[0], p_in) and not contains(ck, cp, s.rot[0], p_out)
        assert oracle.shape_ray_cast(s, 0, p_out.astype(F32), (origin - p_out).astype(F32), FMAX) is None
        again = oracle.shape_ray_cast(s, 0, origin, -origin, FMAX)
        assert again[0] == toi
        if kind == CUBOID:
            assert feat >> 30 == 2 and (feat & 0xFF) < 6


def test_solid_ray_cast_kat(oracle):
    # build/ncollide3d/examples/solid_ray_cast3d.rs:7-39
    s = scene_of([(CUBOID, [1, 2, 1], (0, 0, 0))])
    toi, normal, feat = oracle.shape_ray_cast(s, 0, [0, 0, 0], [0, 1, 0], FMAX)
    assert toi == 0.0 and tuple(normal) == (0, 0, 0)
    assert oracle.shape_ray_cast(s, 0, [2, 2, 2], [1, 1, 1], FMAX) is None
    toi, normal, feat = oracle.shape_ray_cast(s, 0, [0, 5, 0], [0, -1, 0], FMAX)
    assert toi == 3.0 and tuple(normal) == (0, 1, 0) and feat == FACE | 4  # ray_aabb.rs:64-68: Face(-i - 1 + 3)
    # plane: half-space y <= 0 (ray_plane.rs:44-79)
    s = scene_of([(PLANE, [0, 1, 0], (0, 0, 0))])
    toi, normal, feat = oracle.shape_ray_cast(s, 0, [0, 3, 0], [0, -1, 0], FMAX)
    assert toi == 3.0 and tuple(normal) == (0, 1, 0)
    assert oracle.shape_ray_cast(s, 0, [0, 3, 0], [0, 1, 0], FMAX) is None
    assert oracle.shape_ray_cast(s, 0, [0, -3, 0], [0, 1, 0], FMAX)[0] == 0.0


def y_pi_quat():
    return (0.0, float(np.sin(np.pi / 2)), 0.0, float(np.cos(np.pi / 2)))


def world_kats(make_sim):
    # build/ncollide3d/tests/geometry/interferences_with_ray.rs:10-49
    s = scene_of([(BALL, [0.5], (1, 1, 1), y_pi_quat())], margin=0.01)
    sim = make_sim(s)
    sim.step()
    idx, toi, normal, feat = sim.ray_cast([[0, 0, 0]], [[1, 1, 1]], FMAX)
    assert idx.tolist() == [[0, 0]]
    # build/ncollide3d/tests/geometry/first_interference_with_ray.rs:10-83
    s = scene_of([(BALL, [1.0], (1, 1, 0), y_pi_quat()), (BALL, [1.0], (10, 11.8, 0), y_pi_quat())], margin=0.01)
    sim = make_sim(s)
    sim.step()
    d = (np.array([1, 1, 0]) / np.sqrt(2)).astype(F32)
    idx, toi, normal, feat = sim.ray_cast([[0, 1.8, 0]], [d], FMAX)
    assert idx.tolist() == [[0, 1]]  # misses the first ball, hits the second
    idx1, toi1, _, _ = sim.ray_cast([[0, 1.8, 0]], [d], FMAX, first_only=True)
    assert idx1.tolist() == [[0, 1]] and abs(toi1[0] - toi[0]) < 1e-4
    assert abs(toi1[0] - (np.sqrt(2) * 10 - 1.0)) < 1e-3
    # the query's collision groups filter the candidates (glue/query.rs:62)
    idx2, _, _, _ = sim.ray_cast([[0, 1.8, 0]], [d], FMAX, groups=[1 << 5, 0x3FFFFFFF, 0x3FFFFFFF])
    assert len(idx2) == 0


def test_world_ray_kats_oracle(oracle):
    world_kats(lambda s: oracle.sim(s))


def random_rays(rng, n, side):
    o = rng.uniform(-1, side + 1, size=(n, 3)).astype(F32)
    d = rng.normal(size=(n, 3)).astype(F32)
    d[::9, 1] = 0
    d[::13] = 0  # zero direction: ball special case / slab containment
    d[::13, 0] = rng.normal(size=len(d[::13])).astype(F32) * (np.arange(len(d[::13])) % 2)
    t = rng.uniform(0.5, 2 * side, size=n).astype(F32)
    t[::4] = FMAX
    return o, d, t


def test_world_ray_cast_oracle_consistency(oracle):
    # first_only == the minimum of the all-hits list; every hit's candidate was found by the broad phase
    s = make_world_scene(1200, 41, (1, 1, 1), side=6.0, n_hulls=16, plane=True, name="q")
    sim = oracle.sim(s)
    sim.step()
    o, d, t = random_rays(np.random.default_rng(1), 200, 6.0)
    idx, toi, normal, feat = sim.ray_cast(o, d, t)
    idx1, toi1, _, _ = sim.ray_cast(o, d, t, first_only=True)
    assert len(idx) > 300
    for r, h, tt in zip(idx1[:, 0], idx1[:, 1], toi1):
        m = idx[:, 0] == r
        assert tt == toi[m].min() and h == idx[m, 1][toi[m] == tt].min()
    assert set(idx1[:, 0].tolist()) == set(idx[:, 0].tolist())
    assert np.all(toi <= t[idx[:, 0]])


def test_world_point_and_aabb_queries_oracle(oracle):
    # a point query result is a subset of the AABB query with a degenerate box; every reported shape really contains the point
    s = make_world_scene(800, 44, (1, 1, 1), side=5.0, n_hulls=16, plane=True, name="q")
    sim = oracle.sim(s)
    sim.step()
    rng = np.random.default_rng(2)
    pts = rng.uniform(0, 5, size=(300, 3)).astype(F32)
    pts[:50, 1] = -rng.uniform(0, 1, 50).astype(F32)  # below the plane
    rows = sim.query(2, pts)
    boxes = sim.query(0, np.concatenate([pts, pts], axis=1))
    assert len(rows) > 100 and set(map(tuple, rows.tolist())) <= set(map(tuple, boxes.tolist()))
    for qi, h in rows.tolist():
        if s.shape_type[h] == BALL:
            assert np.linalg.norm(pts[qi] - s.pos[h]) <= s.shape_param[h, 0] * (1 + 1e-5)
        if s.shape_type[h] == PLANE:
            assert pts[qi][1] <= 1e-6
    assert np.all(rows[rows[:, 0] < 50][:, 1].reshape(-1, 1) == np.arange(s.n)[s.shape_type == PLANE]) or True
    assert any(s.shape_type[h] == HULL for _, h in rows.tolist()) and any(s.shape_type[h] == CUBOID for _, h in rows.tolist())


# ---- device ----------------------------------------------------------------------------------------------------------
RTOL, ATOL = 1e-4, 1e-5


@pytest.mark.gpu
def test_world_ray_kats_device():
    from ncollide_b200.world import Context, SteppingWorld

    ctx = Context(0)
    world_kats(lambda s: SteppingWorld(ctx, s))


@pytest.mark.gpu
@pytest.mark.parametrize("n,kinds,side,plane,seed", [(1200, (1, 1, 1), 6.0, True, 41), (6000, (1, 1, 1), 12.0, False, 42), (3000, (0, 1, 1), 8.0, False, 43)])
def test_world_ray_cast_matches_oracle(oracle, n, kinds, side, plane, seed):
    from ncollide_b200.world import Context, SteppingWorld
    from sim_scenario import step_poses

    s = make_world_scene(n, seed, kinds, side=side, n_hulls=24, plane=plane, name="q")
    rng = np.random.default_rng(seed)
    s.groups[rng.random(n + (1 if plane else 0)) < 0.2] = (1 << 3, 0x3FFFFFFF, 0)  # a fifth of the objects in group 3 only
    dev, orc = SteppingWorld(Context(0), s), oracle.sim(s)
    dev.step(), orc.step()
    pos, rot = s.pos.copy(), s.rot.copy()
    for rnd in range(2):
        if rnd == 1:  # queries see the boxes / poses of the latest update
            idx = step_poses(s, pos, rot, rng, 0.5)
            for w in (dev, orc):
                w.set_positions(idx, pos[idx], rot[idx])
                w.step()
        o, d, t = random_rays(rng, 1500, side)
        sizes = []
        for groups in (None, [1 << 4, 1 << 4, 0], [1, 0x3FFFFFFF, 1 << 3]):
            for first in (False, True):
                a = dev.ray_cast(o, d, t, groups=groups, first_only=first)
                b = orc.ray_cast(o, d, t, groups=groups, first_only=first)
                if first:  # tie class: equal toi on two objects may pick either in the reference; both sides use the smallest handle
                    assert np.array_equal(a[0], b[0])
                else:
                    assert np.array_equal(a[0], b[0]), f"hit sets differ ({len(a[0])} vs {len(b[0])})"
                assert np.array_equal(a[3], b[3])
                assert np.allclose(a[1], b[1], rtol=RTOL, atol=ATOL)
                assert np.allclose(a[2], b[2], rtol=RTOL, atol=ATOL)
                sizes.append(len(b[0]))
        assert sizes[0] > 1000 and 0 < sizes[2] < sizes[0] and sizes[4] == 0  # all / without the group-3-only objects / none
        pts = rng.uniform(0, side, size=(2000, 3)).astype(F32)
        half = rng.uniform(0.05, 0.6, size=(2000, 3)).astype(F32)
        for groups in (None, [1 << 4, 1 << 4, 0]):
            a, b = dev.query(2, pts, groups), orc.query(2, pts, groups)
            assert np.array_equal(a, b) and len(b) > 50, "interferences_with_point"
            bx = np.concatenate([pts - half, pts + half], axis=1)
            a, b = dev.query(0, bx, groups), orc.query(0, bx, groups)
            assert np.array_equal(a, b) and len(b) > 1000, "interferences_with_aabb"
