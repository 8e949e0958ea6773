"""ncollide2d ``RayCast for Polyline`` (the 2-D counterpart of the TriMesh ray path).

CPU: the oracle (oracle/ray.cpp, 2-D part) on the reference's own known-answer tests — build/ncollide2d/tests/geometry/ray_cast.rs, the
ten ``Segment`` cases (exact ``Some(0.0)`` / ``Some(1.0)`` / ``None``) and the geometry of ``convexpoly_raycast_fuzz`` against the same
square as a closed polyline; the reference-faithful best-first search against the brute-force definition the device follows; an
independent f64 intersection in numpy; the device source compiled for the host against the oracle, bit for bit.
GPU: ``ncb2d_polyline_ray_cast`` against the oracle, bit for bit."""
import ctypes as C

import numpy as np
import pytest

from ncollide_b200 import dim2
from ncollide_b200.scenes import make_polyline_scene

F = np.float32
FMAX = np.finfo(np.float32).max


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ---- CPU: known-answer tests -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["oracle", "oracle64"])
def test_oracle_segment_ray_kats(which, request):
    """build/ncollide2d/tests/geometry/ray_cast.rs, in file order."""
    R = request.getfixturevalue(which).segment_ray_cast
    assert R((2, 1), (2, 0), (0, 0), (0, 1)) is None  # issue_178_parallel_raycast
    assert R((2, 1), (2, -1), (0, 0), (0, 1)) is None  # parallel_raycast
    assert R((0, 1), (0, -1), (0, 0), (0, 1))[0] == 0.0  # collinear_raycast_starting_on_segment
    assert R((0, 1), (0, -1), (0, -2), (0, 1))[0] == 1.0  # collinear_raycast_starting_bellow_segment
    assert R((0, 1), (0, -1), (0, 2), (0, 1)) is None  # collinear_raycast_starting_above_segment
    assert R((0, -10), (0, 10), (-1, 0), (1, 0)) is not None  # perpendicular_raycast_starting_behind_sement
    assert R((0, -10), (0, 10), (1, 0), (1, 0)) is None  # perpendicular_raycast_starting_in_front_of_sement
    assert R((0, -10), (0, 10), (0, 3), (1, 0))[0] == 0.0  # perpendicular_raycast_starting_on_segment
    assert R((0, -10), (0, 10), (0, 11), (1, 0)) is None  # perpendicular_raycast_starting_above_segment
    assert R((0, -10), (0, 10), (0, -11), (1, 0)) is None  # perpendicular_raycast_starting_bellow_segment
    # the features and the scaled normal of the two collinear hits (ray_support_map.rs:251-270)
    toi, n, feat = R((0, 1), (0, -1), (0, -2), (0, 1))
    assert feat == (2, 1) and tuple(n) == (-2.0, 0.0)
    assert R((0, 1), (0, -1), (0, 0), (0, 1))[2] == (1, 0)


def test_oracle_square_polyline_fuzz(oracle64):
    """The rays of ray_cast.rs::convexpoly_raycast_fuzz against the same square as a closed Polyline: every ray hits the front face
    at a distance in [1, sqrt(2))."""
    sq = oracle64.polyline([[2, 1], [2, 2], [1, 2], [1, 1]], [[0, 1], [1, 2], [2, 3], [3, 0]])
    i = np.arange(10_000)
    o = np.stack([np.full(len(i), 3.0), 1.0 + i * 1e-4], axis=1)
    d = np.array([0.0, 2.0]) - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    toi, feat, n = sq.ray_cast(o, d)
    assert (toi >= 1.0).all() and (toi < np.sqrt(2.0)).all()
    assert set(np.unique(feat % 4).tolist()) <= {0, 1}  # the right edge, and the top edge for the last rays


def _scene(kind, n_edges, n_rays, seed, pose=None):
    """A polyline scene; with a pose, the rays are moved along with the polyline (rounded to f32 in the world frame)."""
    pts, edges, o, d = make_polyline_scene(kind, n_edges, n_rays, seed)
    if pose is not None:
        rot = np.array([[pose[2], -pose[3]], [pose[3], pose[2]]], dtype=np.float64)
        o = np.ascontiguousarray(o.astype(np.float64) @ rot.T + np.asarray(pose[:2], dtype=np.float64), dtype=F)
        d = np.ascontiguousarray(d.astype(np.float64) @ rot.T, dtype=F)
    return pts, edges, o, d


@pytest.mark.parametrize("kind", ["terrain", "soup"])
def test_oracle_best_first_equals_definition(oracle, kind):
    """ray_polyline.rs's best-first search (mode 0) returns the minimum toi over the accepted hits (mode 1), the same edge and normal
    outside exact ties."""
    pose = np.array([1.5, -2.0, np.cos(0.4), np.sin(0.4)], dtype=F) if kind == "soup" else None
    pts, edges, o, d = _scene(kind, 3000, 4000, 5, pose)
    pl = oracle.polyline(pts, edges)
    t0, f0, n0 = pl.ray_cast(o, d, pose=pose, mode=0)
    t1, f1, n1 = pl.ray_cast(o, d, pose=pose, mode=1)
    assert np.array_equal(bits(t0), bits(t1))
    same = f0 == f1
    assert (~same).sum() <= 5 and np.array_equal(bits(n0[same]), bits(n1[same]))
    assert (t1 >= 0).sum() > 1000 and (t1 < 0).sum() > (0 if kind == "terrain" else 100)


def test_oracle_against_numpy_intersections(oracle64):
    """ORACLE check (f64): toi == the smallest ray parameter over all segments, computed independently with a 2 x 2 solve."""
    pts, edges, o, d = _scene("soup", 600, 800, 6)
    pl = oracle64.polyline(pts, edges)
    toi, feat, n = pl.ray_cast(o, d)
    a, b = pts[edges[:, 0]].astype(np.float64), pts[edges[:, 1]].astype(np.float64)
    e = b - a
    checked = 0
    for r in range(len(o)):
        oo, dd = o[r].astype(np.float64), d[r].astype(np.float64)
        den = dd[0] * e[:, 1] - dd[1] * e[:, 0]
        ok = np.abs(den) > 1e-9
        w = a - oo
        s = np.where(ok, (w[:, 0] * e[:, 1] - w[:, 1] * e[:, 0]) / np.where(ok, den, 1), np.inf)
        t = np.where(ok, (w[:, 0] * dd[1] - w[:, 1] * dd[0]) / np.where(ok, den, 1), np.inf)
        hit = ok & (s >= 0) & (t >= 0) & (t <= 1)
        margin = ok & (np.minimum(np.abs(t), np.abs(t - 1)) < 1e-7)
        if margin.any() or (~ok).any() and np.abs(den[~ok]).max() > 0:
            continue
        if hit.any():
            k = int(np.argmin(np.where(hit, s, np.inf)))
            assert abs(toi[r] - s[k]) < 1e-9 * max(1.0, s[k]), (r, toi[r], s[k])
            assert feat[r] % len(edges) == k
            nn = n[r] / np.linalg.norm(n[r])
            assert nn @ dd <= 1e-12 and abs(nn @ e[k]) < 1e-9 * np.linalg.norm(e[k])  # faces the ray, perpendicular to the edge
        else:
            assert toi[r] < 0
        checked += 1
    assert checked > 500


def test_oracle_max_toi_limits_the_boxes_only(oracle):
    """RayCast for Segment (dim2) never compares its hit with max_toi: a polyline hit is cut off only when the edge's AABB is not
    entered within max_toi (ray_support_map.rs:219-293, ray_polyline.rs:113-146)."""
    pl = oracle.polyline([[0, 0], [10, 10]])  # one diagonal edge: its AABB is the square [0, 10]^2
    o, d = [[5.0, -1.0]], [[0.0, 1.0]]  # enters the box at toi 1, meets the segment at toi 6
    assert pl.ray_cast(o, d, max_toi=0.5)[0][0] < 0
    assert pl.ray_cast(o, d, max_toi=2.0)[0][0] == 6.0
    assert pl.ray_cast(o, d)[0][0] == 6.0


# ---- CPU: the DEVICE source compiled for the host against the oracle, bit for bit ------------------------------------------------
@pytest.fixture(scope="module")
def ray_shim():
    from test_device_source_on_host import _build_shim

    return _build_shim("libray_host.so", "ray_host.cpp")


def _vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


@pytest.mark.parametrize("kind,posed,limited", [("terrain", False, False), ("soup", True, False), ("soup", False, True), ("terrain", True, True)])
def test_device_source_equals_oracle_bit_for_bit(ray_shim, oracle, kind, posed, limited):
    pose = np.array([0.5, 2.0, np.cos(-0.7), np.sin(-0.7)], dtype=F) if posed else None
    pts, edges, o, d = _scene(kind, 2500, 3000, 9, pose)
    if edges is None:
        edges = np.stack([np.arange(len(pts) - 1), np.arange(1, len(pts))], axis=1).astype(np.uint32)
    limits = np.random.default_rng(3).uniform(2.0, 14.0, size=len(o)).astype(F) if limited else None
    n = len(o)
    toi, feat, nrm = np.zeros(n, dtype=F), np.zeros(n, dtype=np.uint32), np.zeros((n, 2), dtype=F)
    ray_shim.shim2_polyline_ray_cast(C.c_uint32(len(edges)), _vp(pts), _vp(edges), _vp(pose), C.c_uint64(n), _vp(o), _vp(d), C.c_float(FMAX),
                                     _vp(limits), _vp(toi), _vp(feat), _vp(nrm))
    ot, of, on = oracle.polyline(pts, edges).ray_cast(o, d, max_toi=limits, pose=pose, mode=1)
    assert np.array_equal(bits(toi), bits(ot)) and np.array_equal(feat, of) and np.array_equal(bits(nrm), bits(on))
    assert (ot >= 0).sum() > 500 and (ot < 0).sum() > (100 if limited else 0)


# ---- GPU ---------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    from ncollide_b200.world import Context

    c = Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n_edges,posed,limited", [("terrain", 200_000, False, False), ("soup", 150_000, True, False),
                                                       ("soup", 50_000, False, True), ("terrain", 3, True, True), ("soup", 1, False, False)])
def test_device_polyline_rays_match_oracle(ctx, oracle, kind, n_edges, posed, limited):
    pose = np.array([0.5, 2.0, np.cos(1.1), np.sin(1.1)], dtype=F) if posed else None
    pts, edges, o, d = _scene(kind, n_edges, 30_000, 12, pose)
    limits = np.random.default_rng(4).uniform(2.0, 14.0, size=len(o)).astype(F) if limited else None
    pl = dim2.Polyline(ctx, pts, edges)
    toi, feat, nrm = pl.toi_and_normal_with_ray(pose, o, d, max_toi=limits)
    op = oracle.polyline(pts, edges)
    ot, of, on = op.ray_cast(o[:6000], d[:6000], max_toi=None if limits is None else limits[:6000], pose=pose, mode=1)  # the definition
    assert np.array_equal(bits(toi[:6000]), bits(ot)) and np.array_equal(feat[:6000], of) and np.array_equal(bits(nrm[:6000]), bits(on))
    rt, rf, rn = op.ray_cast(o, d, max_toi=limits, pose=pose, mode=0)  # the reference's best-first search
    assert np.array_equal(bits(toi), bits(rt))
    same = feat == rf
    assert (~same).sum() <= 5 and np.array_equal(bits(nrm[same]), bits(rn[same]))
    if n_edges > 1000:
        assert (toi >= 0).sum() > 5000
    assert ctx.traversal_overflows() == 0
    toi2, feat2, none = pl.toi_and_normal_with_ray(pose, o, d, max_toi=limits, want_normals=False)
    assert none is None and np.array_equal(bits(toi2), bits(toi)) and np.array_equal(feat2, feat)
    pl.close()


@pytest.mark.gpu
def test_device_polyline_edge_cases(ctx):
    from ncollide_b200._ffi import NcbError

    empty = dim2.Polyline(ctx, np.zeros((1, 2), dtype=F))  # one point: no edge
    toi, feat, nrm = empty.toi_and_normal_with_ray(None, [[0, 0]], [[1, 0]])
    assert toi[0] == -1 and feat[0] == 0xFFFFFFFF and not nrm.any()
    pl = dim2.Polyline(ctx, [[0, 1], [0, -1]])
    toi, feat, nrm = pl.toi_and_normal_with_ray(None, [[0, 0], [0, -2], [0, 2]], [[0, 1]] * 3)  # the collinear known-answer tests
    assert toi.tolist() == [0.0, 1.0, -1.0]
    assert len(pl.toi_and_normal_with_ray(None, np.zeros((0, 2)), np.zeros((0, 2)))[0]) == 0
    with pytest.raises(NcbError):
        dim2.Polyline(ctx, [[0, 0], [1, 1]], [[0, 2]])  # index out of range
    with pytest.raises(NcbError):  # a Polyline is not a TriMesh
        ctx.check(ctx.lib.ncb_trimesh_set_uvs(pl.h, None), "ncb_trimesh_set_uvs")
    pl.close(), empty.close()


# ---- RayCast for the 2-D shapes (ball, cuboid, convex polygon, plane) ---------------------------------------------------------------
def random_shape_rays(n, seed, kinds=(0, 1, 2, 3)):
    """n (shape, pose, ray) triples: rays aimed near the shape from 1-4 units away, a share starting inside, some axis-aligned."""
    rng = np.random.default_rng(seed)
    sh = dim2.Shapes2D()
    t = rng.choice(kinds, size=n)
    for k in t:
        if k == 0:
            sh.ball(rng.uniform(0.2, 0.7))
        elif k == 1:
            sh.cuboid(rng.uniform(0.2, 0.7), rng.uniform(0.2, 0.7))
        elif k == 3:
            sh.plane(rng.normal(size=2))
        elif k == 4:
            a = rng.uniform(-0.6, 0.6, size=2)
            sh.segment(a, a + rng.uniform(0.2, 0.9) * np.array([np.cos(th := rng.uniform(0, 2 * np.pi)), np.sin(th)]))
        else:
            m = int(rng.integers(3, 11))
            ang = np.sort(rng.uniform(0, 2 * np.pi, size=m)) + np.arange(m) * 1e-3
            a, b = rng.uniform(0.25, 0.7, size=2)
            sh.polygon(np.stack([a * np.cos(ang), b * np.sin(ang)], axis=1))
    typ, par, pts, nrm = sh.arrays()
    c = rng.uniform(-5, 5, size=(n, 2))
    angle = rng.uniform(-np.pi, np.pi, size=n)
    angle[rng.random(n) < 0.2] = 0.0
    pose = dim2.isometry2(c, angle)
    o = c + rng.normal(size=(n, 2)) * rng.uniform(0.0, 3.0, size=(n, 1))
    target = c + rng.uniform(-0.8, 0.8, size=(n, 2))
    d = target - o
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-9)
    d *= rng.uniform(0.3, 3.0, size=(n, 1))  # ray.dir need not be a unit vector
    axis = rng.random(n) < 0.1
    d[axis] = np.where(rng.random((int(axis.sum()), 1)) < 0.5, [[0.0, -1.0]], [[1.0, 0.0]])
    d[rng.random(n) < 0.005] = 0.0  # the zero direction: None, or Some(0) from inside
    max_toi = np.where(rng.random(n) < 0.3, rng.uniform(0.2, 3.0, size=n), FMAX)
    rays = np.concatenate([o, d, max_toi[:, None]], axis=1).astype(F)
    return typ, par, pose, rays, pts


@pytest.mark.parametrize("which", ["oracle64", "oracle"])
def test_oracle_convexpoly_raycast_fuzz(which, request):
    """build/ncollide2d/tests/geometry/ray_cast.rs::convexpoly_raycast_fuzz, as written (f64 in the reference; f32 too here): every
    ray hits the front face of the square, at a distance in [1, sqrt(2))."""
    orc = request.getfixturevalue(which)
    pts = np.array([[2, 1], [2, 2], [1, 2], [1, 1]], dtype=np.float64)
    i = np.arange(10_000)
    o = np.stack([np.full(len(i), 3.0), 1.0 + i * 1e-4], axis=1)
    d = np.array([0.0, 2.0]) - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d, np.full((len(i), 1), np.finfo(orc.dtype).max)], axis=1)
    n = len(i)
    found, out, feat = orc.ray_cast2d([2] * n, [[0, 4, 0, 0]] * n, [[0, 0, 1, 0]] * n, rays, pts)
    assert found.all(), "Failed to collide with any face"
    slack = 0.0 if orc.dtype == np.float64 else 2e-6
    assert (out[:, 0] >= 1.0 - slack).all() and (out[:, 0] < np.sqrt(2.0)).all()


def test_oracle_shape_rays_against_numpy(oracle64):
    """ORACLE check (f64): toi against closed forms — circle / half-plane equations, and for cuboids and polygons the nearest
    crossing of the boundary edges (0 from inside)."""
    typ, par, pose, rays, pts = random_shape_rays(4000, 41)
    found, out, feat = oracle64.ray_cast2d(typ, par, pose, rays, pts)
    checked = {0: 0, 1: 0, 2: 0, 3: 0}
    for k in range(len(typ)):
        o, d, lim = rays[k, :2].astype(np.float64), rays[k, 2:4].astype(np.float64), float(rays[k, 4])
        m = pose[k].astype(np.float64)
        if not d.any():
            continue
        want = None  # None = miss
        if typ[k] == 0:
            dc, r = o - m[:2], float(par[k, 0])
            a, b, c = d @ d, dc @ d, dc @ dc - r * r
            if c <= 0:
                want = 0.0
            elif b <= 0 and b * b - a * c >= 0:
                want = (-b - np.sqrt(b * b - a * c)) / a
        elif typ[k] == 3:
            nw = np.array([m[2] * par[k, 0] - m[3] * par[k, 1], m[3] * par[k, 0] + m[2] * par[k, 1]], dtype=np.float64)
            dist = nw @ (o - m[:2])
            if dist < 0:
                want = 0.0
            elif nw @ d < 0:
                want = -dist / (nw @ d)
        else:
            if typ[k] == 1:
                hx, hy = float(par[k, 0]), float(par[k, 1])
                loc = np.array([[hx, hy], [-hx, hy], [-hx, -hy], [hx, -hy]])
            else:
                loc = pts[int(par[k, 0]) : int(par[k, 0]) + int(par[k, 1])].astype(np.float64)
            P = np.stack([m[2] * loc[:, 0] - m[3] * loc[:, 1] + m[0], m[3] * loc[:, 0] + m[2] * loc[:, 1] + m[1]], axis=1)
            a, e = P, np.roll(P, -1, axis=0) - P
            inside = bool(np.all(e[:, 0] * (o - a)[:, 1] - e[:, 1] * (o - a)[:, 0] >= 0))
            den = d[0] * e[:, 1] - d[1] * e[:, 0]
            ok = np.abs(den) > 1e-12
            w = a - o
            s = np.where(ok, (w[:, 0] * e[:, 1] - w[:, 1] * e[:, 0]) / np.where(ok, den, 1), np.inf)
            t = np.where(ok, (w[:, 0] * d[1] - w[:, 1] * d[0]) / np.where(ok, den, 1), np.inf)
            hit = ok & (s >= 0) & (t >= 0) & (t <= 1)
            if np.any(ok & (np.minimum(np.abs(t), np.abs(t - 1)) < 1e-6) & (s >= 0)):
                continue  # through a vertex: either answer
            if inside:
                want = 0.0
            elif hit.any():
                want = float(s[hit].min())
        if want is not None and abs(want - lim) < 1e-6 * max(1.0, lim):
            continue
        if want is None or want > lim:
            assert not found[k], (k, typ[k], want, out[k])
        else:
            assert found[k], (k, typ[k], want)
            assert abs(out[k, 0] - want) < 1e-6 * max(1.0, want), (k, typ[k], out[k, 0], want)
            if want > 0:
                assert abs(np.hypot(out[k, 1], out[k, 2]) - 1) < 1e-6 and out[k, 1:] @ d <= 1e-9
            checked[int(typ[k])] += 1
    assert min(checked.values()) > 200, checked


@pytest.fixture(scope="module")
def dim2_shim():
    from test_device_source_on_host import _build_shim

    return _build_shim("libdim2_host.so", "dim2_host.cpp")


@pytest.mark.parametrize("seed,kinds", [(51, (0, 1, 2, 3)), (52, (2,)), (53, (0, 1))])
def test_device_source_shape_rays_equal_oracle_bit_for_bit(dim2_shim, oracle, seed, kinds):
    typ, par, pose, rays, pts = random_shape_rays(30_000, seed, kinds)
    n = len(typ)
    found, out, feat = np.zeros(n, dtype=np.uint8), np.zeros((n, 3), dtype=F), np.zeros(n, dtype=np.uint32)
    dim2_shim.shim2_ray_cast(C.c_uint64(n), _vp(typ), _vp(par), _vp(pose), _vp(pts), _vp(rays), _vp(found), _vp(out), _vp(feat))
    ofound, oout, ofeat = oracle.ray_cast2d(typ, par, pose, rays, pts)
    assert np.array_equal(found, ofound) and np.array_equal(feat, ofeat)
    hit = found.astype(bool)
    assert np.array_equal(bits(out[hit]), bits(oout[hit]))
    assert hit.sum() > n // 4 and (~hit).sum() > n // 10


@pytest.mark.gpu
@pytest.mark.parametrize("seed,kinds", [(61, (0, 1, 2, 3)), (62, (2,)), (63, (0, 1, 3))])
def test_device_shape_rays_match_oracle(ctx, oracle, seed, kinds):
    typ, par, pose, rays, pts = random_shape_rays(120_000, seed, kinds)
    found, out, feat = dim2.ray_cast(ctx, typ, par, pose, rays, pts)
    ofound, oout, ofeat = oracle.ray_cast2d(typ, par, pose, rays, pts)
    assert np.array_equal(found, ofound.astype(bool)) and np.array_equal(feat, ofeat)
    assert np.array_equal(bits(out[found]), bits(oout[found])), f"{(bits(out[found]) != bits(oout[found])).sum()} words differ"
    assert found.sum() > 30_000 and (~found).sum() > 10_000


@pytest.mark.gpu
def test_device_convexpoly_raycast_fuzz(ctx):
    pts = np.array([[2, 1], [2, 2], [1, 2], [1, 1]], dtype=F)
    i = np.arange(10_000)
    o = np.stack([np.full(len(i), 3.0), 1.0 + i * 1e-4], axis=1)
    d = np.array([0.0, 2.0]) - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d, np.full((len(i), 1), FMAX)], axis=1).astype(F)
    n = len(i)
    found, out, feat = dim2.ray_cast(ctx, [2] * n, [[0, 4, 0, 0]] * n, [[0, 0, 1, 0]] * n, rays, pts)
    assert found.all() and (out[:, 0] >= 1.0 - 2e-6).all() and (out[:, 0] < np.sqrt(2.0)).all()


# ---- world ray queries of the 2-D world (interferences_with_ray / first_interference_with_ray) -------------------------------------
def _ray_world(n, seed, planes=2, with_groups=True):
    from test_dim2 import random_world

    w = random_world(n, seed, (0, 1, 2), planes=planes, with_groups=with_groups)
    rng = np.random.default_rng(seed + 1)
    lo, hi = w.pos.min(axis=0), w.pos.max(axis=0)
    m = 4000
    o = rng.uniform(lo - 1.0, hi + 1.0, size=(m, 2))
    d = rng.normal(size=(m, 2))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[rng.random(m) < 0.05] = [1.0, 0.0]
    max_toi = np.where(rng.random(m) < 0.5, rng.uniform(0.5, 6.0, size=m), FMAX)
    return w, np.concatenate([o, d, max_toi[:, None]], axis=1).astype(F)


def test_oracle_world_ray_queries_2d(oracle):
    """ORACLE properties: the first interference of a ray is the row with the smallest toi among all its interferences (ties: smallest
    handle); a query whose blacklist names every group sees nothing; every toi respects max_toi."""
    w, rays = _ray_world(1500, 81)
    idx, val, feat = oracle.world_ray_cast2d(w, rays)
    fidx, fval, ffeat = oracle.world_ray_cast2d(w, rays, first_only=True)
    assert len(idx) > 3000 and (val[:, 0] <= rays[idx[:, 0], 4]).all() and (val[:, 0] >= 0).all()
    assert np.array_equal(np.unique(idx[:, 0]), fidx[:, 0])
    order = np.lexsort((idx[:, 1], val[:, 0], idx[:, 0]))
    first = order[np.unique(idx[order, 0], return_index=True)[1]]
    assert np.array_equal(idx[first], fidx) and np.array_equal(bits(val[first]), bits(fval)) and np.array_equal(feat[first], ffeat)
    none, _, _ = oracle.world_ray_cast2d(w, rays, groups=(0x3FFFFFFF, 0x3FFFFFFF, 0x3FFFFFFF))
    assert len(none) == 0
    some, _, _ = oracle.world_ray_cast2d(w, rays, groups=(1 << 3, 0x3FFFFFFF, 1 << 5))
    assert 0 < len(some) < len(idx)


@pytest.mark.gpu
@pytest.mark.parametrize("groups", [None, (1 << 3, 0x3FFFFFFF, 1 << 5)])
def test_device_world_ray_queries_2d(ctx, oracle, groups):
    from ncollide_b200._ffi import NcbError

    w, rays = _ray_world(5000, 82)
    dim2.world_update(ctx, w)
    for first_only in (False, True):
        idx, val, feat = dim2.world_ray_cast(ctx, rays, groups=groups, first_only=first_only)
        oidx, oval, ofeat = oracle.world_ray_cast2d(w, rays, groups=groups, first_only=first_only)
        assert len(oidx) > 1000
        assert np.array_equal(idx, oidx) and np.array_equal(feat, ofeat)
        assert np.array_equal(bits(val), bits(oval)), f"{(bits(val) != bits(oval)).sum()} words differ"
    assert ctx.traversal_overflows() == 0
    from ncollide_b200.world import Context

    fresh = Context(0)
    with pytest.raises(NcbError):
        dim2.world_ray_cast(fresh, rays)  # no 2-D world on that context
    fresh.close()


# ---- point / AABB queries of the 2-D world ---------------------------------------------------------------------------------------
def _shape_points(n, seed):
    typ, par, pose, rays, pts = random_shape_rays(n, seed)
    rng = np.random.default_rng(seed + 5)
    q = pose[:, :2] + rng.normal(size=(n, 2)) * rng.uniform(0.0, 0.9, size=(n, 1))
    return typ, par, pose, np.ascontiguousarray(q, dtype=F), pts


def test_oracle_contains_point_against_numpy(oracle64):
    """ORACLE check (f64): contains_point against closed forms (disc, box, half-plane) and, for polygons, the sign of the cross products
    along the boundary."""
    typ, par, pose, q, pts = _shape_points(6000, 91)
    got = oracle64.contains_point2d(typ, par, pose, q, pts)
    seen = {0: [0, 0], 1: [0, 0], 2: [0, 0], 3: [0, 0]}
    for k in range(len(typ)):
        m = pose[k].astype(np.float64)
        d = q[k].astype(np.float64) - m[:2]
        loc = np.array([m[2] * d[0] + m[3] * d[1], -m[3] * d[0] + m[2] * d[1]])
        if typ[k] == 0:
            val = float(par[k, 0]) - np.hypot(*loc)
        elif typ[k] == 1:
            val = float(min(par[k, 0] - abs(loc[0]), par[k, 1] - abs(loc[1])))
        elif typ[k] == 3:
            val = -float(par[k, 0] * loc[0] + par[k, 1] * loc[1])
        else:
            P = pts[int(par[k, 0]) : int(par[k, 0]) + int(par[k, 1])].astype(np.float64)
            e = np.roll(P, -1, axis=0) - P
            cr = (e[:, 0] * (loc - P)[:, 1] - e[:, 1] * (loc - P)[:, 0]) / np.linalg.norm(e, axis=1)
            val = float(cr.min())
        if abs(val) < 1e-6:
            continue
        assert got[k] == (val > 0), (k, typ[k], val)
        seen[int(typ[k])][int(val > 0)] += 1
    assert min(min(v) for v in seen.values()) > 100, seen


def test_device_source_contains_point_equals_oracle(dim2_shim, oracle):
    typ, par, pose, q, pts = _shape_points(40_000, 92)
    out = np.zeros(len(typ), dtype=np.uint8)
    dim2_shim.shim2_contains_point(C.c_uint64(len(typ)), _vp(typ), _vp(par), _vp(pose), _vp(pts), _vp(q), _vp(out))
    want = oracle.contains_point2d(typ, par, pose, q, pts)
    assert np.array_equal(out.astype(bool), want) and 5000 < want.sum() < 35_000


def _world_queries(w, seed, m=3000):
    rng = np.random.default_rng(seed)
    lo, hi = w.pos.min(axis=0), w.pos.max(axis=0)
    pts = rng.uniform(lo, hi, size=(m, 2)).astype(F)
    c = rng.uniform(lo, hi, size=(m, 2))
    he = rng.uniform(0.05, 1.5, size=(m, 2))
    return pts, np.concatenate([c - he, c + he], axis=1).astype(F)


def test_oracle_world_point_and_aabb_queries_2d(oracle):
    w, _ = _ray_world(1500, 83)
    pts, boxes = _world_queries(w, 84)
    ip = oracle.world_query2d(w, "point", pts)
    ib = oracle.world_query2d(w, "aabb", boxes)
    assert len(ip) > 300 and len(ib) > 3000
    pose = np.concatenate([w.pos, w.rot], axis=1)
    inside = oracle.contains_point2d(w.type[ip[:, 1]], w.param[ip[:, 1]], pose[ip[:, 1]], pts[ip[:, 0]], w.points)
    assert inside.all()
    assert len(oracle.world_query2d(w, "point", pts, groups=(0x3FFFFFFF, 0x3FFFFFFF, 0x3FFFFFFF))) == 0
    # a box query sees at least what a point query at its centre sees
    centres = ((boxes[:, :2] + boxes[:, 2:]) * 0.5).astype(F)
    ic = oracle.world_query2d(w, "point", centres)
    assert set(map(tuple, ic.tolist())) <= set(map(tuple, ib.tolist()))


@pytest.mark.gpu
@pytest.mark.parametrize("groups", [None, (1 << 3, 0x3FFFFFFF, 1 << 5)])
def test_device_world_point_and_aabb_queries_2d(ctx, oracle, groups):
    w, _ = _ray_world(5000, 85)
    pts, boxes = _world_queries(w, 86, m=6000)
    dim2.world_update(ctx, w)
    for kind, q in (("point", pts), ("aabb", boxes)):
        got = dim2.world_query(ctx, kind, q, groups=groups)
        want = oracle.world_query2d(w, kind, q, groups=groups)
        assert len(want) > 200 and np.array_equal(got, want), (kind, len(got), len(want))
    assert ctx.traversal_overflows() == 0


# ---- Segment as a shape: rays and points -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["oracle", "oracle64"])
def test_oracle_segment_shape_ray_kats(which, request):
    """The ten Segment tests of ray_cast.rs once more, through the shape entry (type 4, identity pose) instead of the polyline's."""
    orc = request.getfixturevalue(which)
    big = np.finfo(orc.dtype).max
    cases = [((2, 1, 2, 0), (0, 0, 0, 1), None), ((2, 1, 2, -1), (0, 0, 0, 1), None), ((0, 1, 0, -1), (0, 0, 0, 1), 0.0),
             ((0, 1, 0, -1), (0, -2, 0, 1), 1.0), ((0, 1, 0, -1), (0, 2, 0, 1), None), ((0, -10, 0, 10), (-1, 0, 1, 0), 1.0),
             ((0, -10, 0, 10), (1, 0, 1, 0), None), ((0, -10, 0, 10), (0, 3, 1, 0), 0.0), ((0, -10, 0, 10), (0, 11, 1, 0), None),
             ((0, -10, 0, 10), (0, -11, 1, 0), None)]
    n = len(cases)
    found, out, feat = orc.ray_cast2d([4] * n, [c[0] for c in cases], [[0, 0, 1, 0]] * n, [list(c[1]) + [big] for c in cases])
    for k, (_, _, want) in enumerate(cases):
        assert bool(found[k]) == (want is not None), k
        if want is not None:
            assert out[k, 0] == want, (k, out[k])


@pytest.mark.parametrize("seed", [141, 142])
def test_device_source_segment_rays_and_points_equal_oracle(dim2_shim, oracle, seed):
    typ, par, pose, rays, pts = random_shape_rays(30_000, seed, kinds=(1, 4))
    n = len(typ)
    found, out, feat = np.zeros(n, dtype=np.uint8), np.zeros((n, 3), dtype=F), np.zeros(n, dtype=np.uint32)
    dim2_shim.shim2_ray_cast(C.c_uint64(n), _vp(typ), _vp(par), _vp(pose), _vp(pts), _vp(rays), _vp(found), _vp(out), _vp(feat))
    ofound, oout, ofeat = oracle.ray_cast2d(typ, par, pose, rays, pts)
    assert np.array_equal(found, ofound) and np.array_equal(feat, ofeat)
    hit = found.astype(bool)
    assert np.array_equal(bits(out[hit]), bits(oout[hit])) and (hit & (typ == 4)).sum() > 2000
    rng = np.random.default_rng(seed)
    q = (pose[:, :2] + rng.normal(size=(n, 2)) * 0.4).astype(F)
    on = rng.random(n) < 0.3  # points ON the segment (a point of it, rounded to f32): the only ones a segment can contain
    seg = typ == 4
    u = rng.random(n)[:, None]
    loc = par[:, :2] * (1 - u) + par[:, 2:] * u
    world = np.stack([pose[:, 2] * loc[:, 0] - pose[:, 3] * loc[:, 1] + pose[:, 0], pose[:, 3] * loc[:, 0] + pose[:, 2] * loc[:, 1] + pose[:, 1]], axis=1)
    q[on & seg] = world[on & seg].astype(F)
    inside = np.zeros(n, dtype=np.uint8)
    dim2_shim.shim2_contains_point(C.c_uint64(n), _vp(typ), _vp(par), _vp(pose), _vp(pts), _vp(q), _vp(inside))
    want = oracle.contains_point2d(typ, par, pose, q, pts)
    assert np.array_equal(inside.astype(bool), want) and want[seg].sum() > 200


@pytest.mark.gpu
def test_device_segment_rays_match_oracle(ctx, oracle):
    typ, par, pose, rays, pts = random_shape_rays(80_000, 143, kinds=(0, 1, 2, 3, 4))
    found, out, feat = dim2.ray_cast(ctx, typ, par, pose, rays, pts)
    ofound, oout, ofeat = oracle.ray_cast2d(typ, par, pose, rays, pts)
    assert np.array_equal(found, ofound.astype(bool)) and np.array_equal(feat, ofeat)
    assert np.array_equal(bits(out[found]), bits(oout[found])), f"{(bits(out[found]) != bits(oout[found])).sum()} words differ"
    assert (found & (typ == 4)).sum() > 1000


def test_reference_example_first_ray_intersect_2d(oracle):
    """examples2d/first_ray_intersect.rs: a world of two balls and two cuboids along the x axis (margin 0.02, Contacts(0, 0), default
    groups); the ray from the origin towards +x first meets the ball of radius 0.5 centred at (1, 0): toi == 0.5; towards -x: None."""
    sh = dim2.Shapes2D().ball(0.5).ball(0.75).cuboid(0.5, 0.75).cuboid(1.0, 0.5)
    w = dim2.World2D(sh, [[1, 0], [2, 0], [3, 0], [4, 2]], 0.0, margin=0.02, linear=0.0, angular=0.0)
    groups = (0x3FFFFFFF, 0x3FFFFFFF, 0)  # CollisionGroups::new()
    idx, val, feat = oracle.world_ray_cast2d(w, [[0, 0, 1, 0, FMAX], [0, 0, -1, 0, FMAX]], groups=groups, first_only=True)
    assert idx.tolist() == [[0, 0]] and val[0, 0] == 0.5
    every, vals, _ = oracle.world_ray_cast2d(w, [[0, 0, 1, 0, FMAX]], groups=groups)
    assert every[:, 1].tolist() == [0, 1, 2] and vals[:, 0].tolist() == [0.5, 1.25, 2.5]  # the three shapes on the axis, not the box at y = 2
