"""ncollide2d, first slice: batched ``query::contact`` between 2-D balls, cuboids and convex polygons.

CPU: the oracle (oracle/dim2.cpp) on the reference's own 2-D known-answer tests — build/ncollide2d/tests/geometry/epa2.rs (cuboid /
cuboid EPA: depth == 0.5 / 1.8 and normal == -x / -y EXACTLY; the issue-#181 walk never panics) and ball_cuboid_contact.rs (f32 and
f64, both operand orders) — and against an independent separating-axis / closest-feature computation in numpy f64.
GPU: the device (csrc/dim2.cu, ``ncb2d_contact``) against the oracle in f32 on seeded random pairs, and on the same KATs."""
import numpy as np
import pytest

from ncollide_b200 import dim2

F = np.float32


def random_pairs(n, seed, kinds=(0, 1, 2), spread=1.2):
    """n random pairs; returns (t1, p1, m1, t2, p2, m2, points, normals)."""
    rng = np.random.default_rng(seed)
    sh = dim2.Shapes2D()
    t1 = rng.choice(kinds, size=n)
    t2 = rng.choice(kinds, size=n)
    t2[(t1 == 3) & (t2 == 3)] = 1  # plane x plane has no algorithm (the reference panics)
    for t in np.concatenate([t1, t2]):
        if t == 0:
            sh.ball(rng.uniform(0.2, 0.6))
        elif t == 1:
            sh.cuboid(rng.uniform(0.2, 0.6), rng.uniform(0.2, 0.6))
        elif t == 3:
            sh.plane(rng.normal(size=2))
        elif t == 4:
            a = rng.uniform(-0.6, 0.6, size=2)
            sh.segment(a, a + rng.uniform(0.2, 0.9) * np.array([np.cos(th := rng.uniform(0, 2 * np.pi)), np.sin(th)]))
        else:
            k = int(rng.integers(3, 13))
            ang = np.sort(rng.uniform(0, 2 * np.pi, size=k))
            ang += np.arange(k) * 1e-3  # no coincident vertices
            a, b = rng.uniform(0.25, 0.6, size=2)
            sh.polygon(np.stack([a * np.cos(ang), b * np.sin(ang)], axis=1))
    typ, par, pts, nrm = sh.arrays()
    c1 = rng.uniform(-5, 5, size=(n, 2))
    c2 = c1 + rng.uniform(-spread, spread, size=(n, 2))
    degenerate = rng.random(n) < 0.03
    c2[degenerate] = c1[degenerate]  # coincident centres: the x-axis start direction
    a1, a2 = rng.uniform(-np.pi, np.pi, size=n), rng.uniform(-np.pi, np.pi, size=n)
    axis_aligned = rng.random(n) < 0.2
    a1[axis_aligned] = 0.0
    a2[axis_aligned] = rng.choice([0.0, np.pi / 2], size=int(axis_aligned.sum()))
    return typ[:n], par[:n], dim2.isometry2(c1, a1), typ[n:], par[n:], dim2.isometry2(c2, a2), pts, nrm


# ---- CPU: the oracle on the reference's known-answer tests ---------------------------------------------------------------------
@pytest.mark.parametrize("which", ["oracle64", "oracle"])
def test_oracle_epa2_cuboid_cuboid_kat(which, request):
    """build/ncollide2d/tests/geometry/epa2.rs::cuboid_cuboid_EPA (exact equalities, f64 in the reference; f32 too here)."""
    orc = request.getfixturevalue(which)
    t, p = [1, 1], [[2, 1, 0, 0]] * 2
    found, out, panics = orc.contact2d(t, p, [[3.5, 0, 1, 0], [0, 0.2, 1, 0]], t, p, [[0, 0, 1, 0]] * 2, prediction=10.0)
    assert found.tolist() == [1, 1] and panics == 0
    assert out[0, 6] == 0.5 and out[0, 4] == -1.0 and out[0, 5] == 0.0
    assert out[1, 6] == 1.8 and out[1, 4] == 0.0 and out[1, 5] == -1.0


@pytest.mark.parametrize("which", ["oracle64", "oracle"])
def test_oracle_ball_cuboid_contact_kat(which, request):
    """build/ncollide2d/tests/geometry/ball_cuboid_contact.rs: Some(contact) in both operand orders, f64 and f32."""
    orc = request.getfixturevalue(which)
    cub, ball = [0.5, 0.5, 0, 0], [0.5, 0, 0, 0]
    mc, mb = [0, 4, 1, 0], [0.0517938, 3.05178815, 1, 0]
    found, out, _ = orc.contact2d([1, 0], [cub, ball], [mc, mb], [0, 1], [ball, cub], [mb, mc], prediction=0.0)
    assert found.tolist() == [1, 1]
    assert np.allclose(out[0, 4:6], -out[1, 4:6]) and np.isclose(out[0, 6], out[1, 6]) and out[0, 6] > 0


def test_oracle_issue_181_walk_never_panics(oracle64):
    """epa2.rs::cuboids_large_size_ratio_issue_181: a (10, 10) cuboid walks and spins against a (300, 1.5) cuboid at angle 1.5,
    pushed out along the deepest contact after every query ("used to panic").  20 000 of the reference's 200 000 steps here."""
    a, b = [10, 10, 0, 0], [300, 1.5, 0, 0]
    mb = [5.0, 0.0, np.cos(1.5), np.sin(1.5)]
    px, py, angle, hits = 0.0, 0.0, 0.0, 0
    for _ in range(20000):
        px += 0.0001
        angle += 0.005
        found, out, panics = oracle64.contact2d([1], [a], [[px, py, np.cos(angle), np.sin(angle)]], [1], [b], [mb], prediction=0.0)
        assert panics == 0 and np.isfinite(out).all()
        if found[0]:
            hits += 1
            px -= out[0, 4] * out[0, 6]
            py -= out[0, 5] * out[0, 6]
    assert hits > 1000


def _world_polygon(t, p, m, pts):
    if t == 1:
        loc = np.array([[p[0], p[1]], [-p[0], p[1]], [-p[0], -p[1]], [p[0], -p[1]]], dtype=np.float64)
    elif t == 4:
        loc = np.array([[p[0], p[1]], [p[2], p[3]]], dtype=np.float64)
    else:
        loc = pts[int(p[0]) : int(p[0]) + int(p[1])].astype(np.float64)
    re, im = float(m[2]), float(m[3])
    return np.stack([re * loc[:, 0] - im * loc[:, 1] + m[0], im * loc[:, 0] + re * loc[:, 1] + m[1]], axis=1)


def _sat_signed_distance(A, B):
    """Independent check: signed distance between two convex polygons (> 0 separated, < 0 penetration depth) — the largest
    separation over the edge normals of both, refined by vertex-to-edge / vertex-to-vertex distances when separated."""
    best = -np.inf
    for P, Q in ((A, B), (B, A)):
        e = np.roll(P, -1, axis=0) - P
        nrm = np.stack([e[:, 1], -e[:, 0]], axis=1)
        nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        sep = ((Q[None, :, :] - P[:, None, :]) * nrm[:, None, :]).sum(axis=2).min(axis=1)
        best = max(best, sep.max())
    if best <= 0:
        return best
    d = np.inf
    for P, Q in ((A, B), (B, A)):  # separated: exact distance = min over vertex / edge pairs
        a, b2 = P, np.roll(P, -1, axis=0)
        ab = b2 - a
        for q in Q:
            t = np.clip(((q - a) * ab).sum(axis=1) / (ab * ab).sum(axis=1), 0, 1)
            d = min(d, np.linalg.norm(a + ab * t[:, None] - q, axis=1).min())
    return d


def test_oracle_against_separating_axes(oracle64):
    """ORACLE check (f64): depth of contact_support_map_support_map == the separating-axis answer for convex polygons / cuboids."""
    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(1500, 5, kinds=(1, 2))
    found, out, panics = oracle64.contact2d(t1, p1, m1, t2, p2, m2, pts, prediction=0.3, poly_normals=nrm)
    assert panics == 0 and found.max() <= 1
    checked = 0
    for k in range(len(t1)):
        sd = _sat_signed_distance(_world_polygon(t1[k], p1[k], m1[k], pts), _world_polygon(t2[k], p2[k], m2[k], pts))
        if sd > 0.3 + 1e-6:
            assert not found[k], (k, sd)
        elif sd < 0.3 - 1e-6:
            assert found[k], (k, sd)
            assert abs(out[k, 6] - (-sd)) < 2e-5 * max(1.0, abs(sd)), (k, out[k, 6], -sd)
            assert abs(np.hypot(out[k, 4], out[k, 5]) - 1) < 1e-6
            checked += 1
    assert checked > 500


def test_oracle_ball_polygon_against_point_distance(oracle64):
    """ORACLE check (f64): contact_ball_convex_polyhedron with a ConvexPolygon (GJK / EPA projection of the centre) against the
    distance from the centre to the polygon computed edge by edge in numpy."""
    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(1200, 8, kinds=(0, 2), spread=0.9)
    keep = (t1 == 0) & (t2 == 2)
    t1, p1, m1, t2, p2, m2 = t1[keep], p1[keep], m1[keep], t2[keep], p2[keep], m2[keep]
    found, out, panics = oracle64.contact2d(t1, p1, m1, t2, p2, m2, pts, prediction=0.1, poly_normals=nrm)
    assert panics == 0
    checked = inside_seen = 0
    for k in range(len(t1)):
        P = _world_polygon(2, p2[k], m2[k], pts)
        c, r = m1[k, :2].astype(np.float64), float(p1[k, 0])
        a, b = P, np.roll(P, -1, axis=0)
        ab = b - a
        t = np.clip(((c - a) * ab).sum(axis=1) / (ab * ab).sum(axis=1), 0, 1)
        dist = np.linalg.norm(a + ab * t[:, None] - c, axis=1).min()
        inside = bool(np.all(ab[:, 0] * (c - a)[:, 1] - ab[:, 1] * (c - a)[:, 0] >= 0))
        depth = r + dist if inside else r - dist
        if depth < -0.1 - 1e-6:
            assert not found[k]
        elif depth > -0.1 + 1e-6 and dist > 1e-6:
            assert found[k], (k, depth)
            assert abs(out[k, 6] - depth) < 1e-5, (k, out[k, 6], depth)
            checked += 1
            inside_seen += inside
    assert checked > 200 and inside_seen > 20


def test_polygon_try_new_mirror():
    """Shapes2D.polygon mirrors ConvexPolygon::try_new: unit normals per edge, vertices between collinear edges removed."""
    sh = dim2.Shapes2D().polygon([[0, 0], [1, 0], [2, 0], [2, 2], [0, 2]])  # (1, 0) lies on the edge (0,0)-(2,0)
    typ, par, pts, nrm = sh.arrays()
    assert par[0, 1] == 4 and pts[:4].tolist() == [[0, 0], [2, 0], [2, 2], [0, 2]]
    assert np.allclose(nrm[:4], [[0, -1], [1, 0], [0, 1], [-1, 0]])
    with pytest.raises(ValueError):
        dim2.Shapes2D().polygon([[0, 0], [0, 0], [1, 1]])


# ---- query::proximity ----------------------------------------------------------------------------------------------------------
def _proximity_kat_args():
    """examples2d/proximity_query2d.rs: ball(1) at (1,1) / (2,2) / (3,3) against the unit-half-extent cuboid at the origin, margin 1."""
    ball, cub, one = [1, 0, 0, 0], [1, 1, 0, 0], [0, 0, 1, 0]
    return [0] * 3, [ball] * 3, [[1, 1, 1, 0], [2, 2, 1, 0], [3, 3, 1, 0]], [1] * 3, [cub] * 3, [one] * 3


@pytest.mark.parametrize("which", ["oracle64", "oracle"])
def test_oracle_proximity_kat(which, request):
    """The reference's example asserts Intersecting, WithinMargin, Disjoint."""
    orc = request.getfixturevalue(which)
    t1, p1, m1, t2, p2, m2 = _proximity_kat_args()
    assert orc.proximity2d(t1, p1, m1, t2, p2, m2, None, 1.0).tolist() == [0, 1, 2]
    assert orc.proximity2d(t2, p2, m2, t1, p1, m1, None, 1.0).tolist() == [0, 1, 2]


def test_oracle_proximity_against_separating_axes(oracle64):
    """ORACLE check (f64): the status is the separating-axis signed distance sorted into (-inf, 0], (0, margin], (margin, inf)."""
    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(1500, 21, kinds=(1, 2))
    margins = np.random.default_rng(3).uniform(0.0, 0.4, size=len(t1))
    got = oracle64.proximity2d(t1, p1, m1, t2, p2, m2, pts, margins)
    seen = [0, 0, 0]
    for k in range(len(t1)):
        sd = _sat_signed_distance(_world_polygon(t1[k], p1[k], m1[k], pts), _world_polygon(t2[k], p2[k], m2[k], pts))
        if min(abs(sd), abs(sd - margins[k])) < 1e-6:
            continue
        want = 0 if sd < 0 else (1 if sd < margins[k] else 2)
        assert got[k] == want, (k, sd, margins[k], got[k])
        seen[want] += 1
    assert min(seen) > 100


def test_oracle_proximity_agrees_with_contact(oracle64):
    """Balls, planes and support maps: Intersecting <=> contact depth >= 0 at prediction 0; Disjoint <=> no contact at prediction = margin."""
    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(3000, 22, kinds=(0, 1, 2, 3))
    margin = 0.15
    got = oracle64.proximity2d(t1, p1, m1, t2, p2, m2, pts, margin)
    found, out, _ = oracle64.contact2d(t1, p1, m1, t2, p2, m2, pts, margin, poly_normals=nrm)
    clear = ~found.astype(bool) | (np.abs(out[:, 6]) > 1e-9) & (np.abs(out[:, 6] + margin) > 1e-9)
    want = np.where(~found.astype(bool), 2, np.where(out[:, 6] > 0, 0, 1))
    assert np.array_equal(got[clear], want[clear]), np.flatnonzero(got[clear] != want[clear])[:10]
    assert min(np.bincount(got, minlength=3)) > 200


# ---- GPU: device vs oracle -----------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    from ncollide_b200.world import Context

    c = Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_device_kats(ctx):
    t, p = [1, 1], [[2, 1, 0, 0]] * 2
    found, out, info = dim2.contact(ctx, t, p, [[3.5, 0, 1, 0], [0, 0.2, 1, 0]], t, p, [[0, 0, 1, 0]] * 2, prediction=10.0)
    assert found.all() and info == {"ref_panics": 0, "epa_overflow": 0}
    assert out[0, 6] == F(0.5) and tuple(out[0, 4:6]) == (-1.0, 0.0)
    assert out[1, 6] == F(1.8) and tuple(out[1, 4:6]) == (0.0, -1.0)
    cub, ball = [0.5, 0.5, 0, 0], [0.5, 0, 0, 0]
    mc, mb = [0, 4, 1, 0], [0.0517938, 3.05178815, 1, 0]
    found, out, _ = dim2.contact(ctx, [1, 0], [cub, ball], [mc, mb], [0, 1], [ball, cub], [mb, mc])
    assert found.all()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,kinds,prediction", [(1, (0, 1, 2), 0.0), (2, (1, 2), 0.02), (3, (0, 1), 0.3), (4, (2,), 0.05), (6, (1,), 0.0),
                                                   (7, (0, 2), 0.05), (8, (0, 1, 2, 3), 0.05)])
def test_device_contact_matches_oracle(ctx, oracle, seed, kinds, prediction):
    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(60000, seed, kinds)
    found, out, info = dim2.contact(ctx, t1, p1, m1, t2, p2, m2, pts, prediction, poly_normals=nrm)
    ofound, oout, opanics = oracle.contact2d(t1, p1, m1, t2, p2, m2, pts, prediction, poly_normals=nrm)
    assert info["epa_overflow"] == 0 and info["ref_panics"] == opanics
    assert np.array_equal(found, ofound.astype(bool)), f"{(found != ofound.astype(bool)).sum()} Some / None answers differ"
    hit = found
    assert hit.sum() > 10000 and (~hit).sum() > 1000
    assert np.allclose(out[hit], oout[hit], rtol=1e-4, atol=1e-5)
    same = (out[hit].view(np.uint32) == np.ascontiguousarray(oout[hit], dtype=np.float32).view(np.uint32)).mean()
    assert same > 0.999, f"only {same:.5f} of the contact words are bit-identical"


@pytest.mark.gpu
def test_device_refuses_bad_input(ctx):
    from ncollide_b200._ffi import NcbError

    sh = dim2.Shapes2D().ball(0.5).polygon([[0, 0], [1, 0], [0, 1]])
    typ, par, pts, nrm = sh.arrays()
    with pytest.raises(NcbError):  # a ball meets a polygon, but the polygon's normals were not given
        dim2.contact(ctx, typ[:1], par[:1], [[0, 0, 1, 0]], typ[1:], par[1:], [[0.2, 0, 1, 0]], pts)
    found, out, _ = dim2.contact(ctx, typ[:1], par[:1], [[0.3, 0.3, 1, 0]], typ[1:], par[1:], [[0.0, 0, 1, 0]], pts, poly_normals=nrm)
    assert found[0] and out[0, 6] > 0.5  # the centre is inside the triangle
    with pytest.raises(NcbError):
        dim2.contact(ctx, [7], [[1, 0, 0, 0]], [[0, 0, 1, 0]], [1], [[1, 1, 0, 0]], [[0.2, 0, 1, 0]])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,kinds", [(31, (0, 1, 2)), (32, (1, 2)), (33, (0, 1, 2, 3))])
def test_device_proximity_matches_oracle(ctx, oracle, seed, kinds):
    from ncollide_b200._ffi import NcbError

    t1, p1, m1, t2, p2, m2 = _proximity_kat_args()
    assert dim2.proximity(ctx, t1, p1, m1, t2, p2, m2, None, 1.0).tolist() == [0, 1, 2]
    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(60000, seed, kinds)
    margins = np.random.default_rng(seed).uniform(0.0, 0.4, size=len(t1)).astype(np.float32)
    got = dim2.proximity(ctx, t1, p1, m1, t2, p2, m2, pts, margins)
    want = oracle.proximity2d(t1, p1, m1, t2, p2, m2, pts, margins)
    assert min(np.bincount(want, minlength=3)) > 3000
    assert (got != want).sum() <= 2, f"{(got != want).sum()} statuses differ"  # libm sqrt / fma-free arithmetic: identical in practice
    with pytest.raises(NcbError):  # plane x plane: the reference has no algorithm for it
        dim2.proximity(ctx, [3], [[0, 1, 0, 0]], [[0, 0, 1, 0]], [3], [[0, 1, 0, 0]], [[0, 1, 1, 0]], None, 0.1)


# ---- 2-D world update ----------------------------------------------------------------------------------------------------------
def random_world(n, seed, kinds=(0, 1, 2), density=2.5, angular=0.0, linear=0.02, with_groups=False, planes=0):
    rng = np.random.default_rng(seed)
    sh = dim2.Shapes2D()
    for t in rng.choice(kinds, size=n - planes):
        if t == 0:
            sh.ball(rng.uniform(0.25, 0.5))
        elif t == 1:
            sh.cuboid(rng.uniform(0.25, 0.5), rng.uniform(0.25, 0.5))
        elif t == 4:
            a = rng.uniform(-0.4, 0.4, size=2)
            sh.segment(a, a + rng.uniform(0.3, 0.8) * np.array([np.cos(th := rng.uniform(0, 2 * np.pi)), np.sin(th)]))
        else:
            k = int(rng.integers(3, 11))
            ang = np.sort(rng.uniform(0, 2 * np.pi, size=k)) + np.arange(k) * 1e-2
            sh.polygon(np.stack([rng.uniform(0.3, 0.5) * np.cos(ang), rng.uniform(0.3, 0.5) * np.sin(ang)], axis=1))
    side = np.sqrt(n * 0.8 / density) * 1.0
    pos = rng.uniform(0, side, size=(n, 2))
    angle = rng.uniform(-np.pi, np.pi, size=n)
    angle[rng.random(n) < 0.25] = 0.0  # axis-aligned boxes: face-face contacts with two clipped points
    for k in range(planes):  # half-spaces through the scene: a floor, then tilted walls (the last handles: object 1 of their pairs)
        sh.plane((0.0, 1.0) if k == 0 else rng.normal(size=2))
        pos[n - planes + k] = (side / 2, 0.4) if k == 0 else rng.uniform(0.3 * side, 0.7 * side, size=2)
        angle[n - planes + k] = 0.0 if k == 0 else rng.uniform(-np.pi, np.pi)
    groups = None
    if with_groups:
        groups = np.tile(np.array([0x3FFFFFFF, 0x3FFFFFFF, 0], dtype=np.uint32), (n, 1))
        groups[rng.random(n) < 0.3] = (2, 0x3FFFFFFF & ~2, 0)  # members of group 1 that do not talk to each other
    return dim2.World2D(sh, pos, angle, margin=0.02, linear=linear, angular=angular, groups=groups)


def _canon(p):
    p = np.sort(np.asarray(p, dtype=np.int64).reshape(-1, 2), axis=1)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def test_oracle_world2d_properties(oracle64):
    """ORACLE check (f64, CPU): pairs == brute force over the fat boxes; every contact's depth / normal is consistent; a face-face
    contact of two axis-aligned boxes has two points; resting cuboids on a big cuboid touch with depth == overlap exactly."""
    w = random_world(400, 31)
    pairs, off, contacts, feats, panics, fat = oracle64.world_update2d(w)
    assert panics == 0
    lo, hi = fat[:, :2], fat[:, 3:5]
    brute = [(i, j) for i in range(w.n) for j in range(i) if np.all(lo[i] <= hi[j]) and np.all(lo[j] <= hi[i])]
    assert np.array_equal(_canon(pairs), _canon(brute))
    assert np.all(pairs[:, 0] > pairs[:, 1])
    n = contacts[:, 4:6]
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-9)
    assert np.allclose(contacts[:, 6], -np.einsum("ij,ij->i", n, contacts[:, 2:4] - contacts[:, 0:2]), atol=1e-9)
    assert (np.diff(off.astype(np.int64)) == 2).sum() > 5  # clipped face-face manifolds exist
    # known answer: a unit box resting 0.1 deep in a wide box: two contacts, depth 0.1, normal -y for (upper, lower) order
    sh = dim2.Shapes2D().cuboid(5, 0.5).cuboid(0.5, 0.5)
    kw = dim2.World2D(sh, [[0, 0], [0.3, 0.9]], [0, 0], margin=0.02, linear=0.02)
    pairs, off, contacts, feats, panics, _ = oracle64.world_update2d(kw)
    assert pairs.tolist() == [[1, 0]] and off.tolist() == [0, 2]
    assert np.allclose(contacts[:, 6], 0.1) and np.allclose(contacts[:, 4:6], [[0, -1], [0, -1]])
    assert sorted(np.round(contacts[:, 0], 6).tolist()) == [-0.2, 0.8]
    # known answers with a floor (half-space y <= 0, normal +y): a ball 0.3 above it (radius 0.5) and a unit box sunk 0.1 into it
    sh = dim2.Shapes2D().plane((0, 2)).ball(0.5).cuboid(0.5, 0.5)
    kw = dim2.World2D(sh, [[0, 0], [3, 0.3], [-3, 0.4]], [0, 0, 0], margin=0.02, linear=0.02)
    pairs, off, contacts, feats, panics, _ = oracle64.world_update2d(kw)
    got = {tuple(p): contacts[a:b] for p, a, b in zip(pairs.tolist(), off[:-1].tolist(), off[1:].tolist())}
    ball = got[(1, 0)]  # object 1 = the ball (larger handle): the generator is flipped, the normal points from the ball to the plane
    assert len(ball) == 1 and np.allclose(ball[0], [3, -0.2, 3, 0.0, 0, -1, 0.2])
    box = got[(2, 0)]
    assert len(box) == 2 and np.allclose(box[:, 6], 0.1) and np.allclose(box[:, 4:6], [[0, -1], [0, -1]])
    assert sorted(np.round(box[:, 0], 6).tolist()) == [-3.5, -2.5] and np.allclose(box[:, 1], -0.1) and np.allclose(box[:, 3], 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,kinds,angular,groups,planes", [(3000, 41, (0, 1, 2), 0.0, False, 0), (2500, 42, (1, 2), 0.05, True, 0),
                                                                (2000, 43, (0, 1), 0.0, False, 0), (4000, 44, (2,), 0.2, False, 0),
                                                                (50000, 45, (0, 1, 2), 0.0, False, 0), (3000, 46, (0, 1, 2), 0.0, False, 4)])
def test_device_world2d_matches_oracle(ctx, oracle, n, seed, kinds, angular, groups, planes):
    """ncb2d_world_update against the oracle: pair set and orientation exact, manifold sizes and feature ids exact, contacts within
    1e-4 / 1e-5 (and almost all words bit-identical)."""
    w = random_world(n, seed, kinds, angular=angular, with_groups=groups, planes=planes)
    r = dim2.world_update(ctx, w)
    pairs, off, ocontacts, ofeats, panics, _ = oracle.world_update2d(w)
    assert r["diag"] == {"ref_panics": panics, "epa_overflow": 0, "manifold_overflow": 0, "stack_overflow": 0}
    assert np.array_equal(_canon(r["pairs"]), _canon(pairs)) and np.all(r["pairs"][:, 0] > r["pairs"][:, 1])
    want = {tuple(p): (ocontacts[a:b], ofeats[a:b]) for p, a, b in zip(pairs.tolist(), off[:-1].tolist(), off[1:].tolist())}
    words = same = 0
    for p, st, k in zip(r["pairs"].tolist(), r["manifold_start"].tolist(), r["manifold_count"].tolist()):
        oc, of = want[tuple(p)]
        assert k == len(oc), (p, k, len(oc))
        dc, df = r["contacts"][st : st + k], r["features"][st : st + k]
        assert np.array_equal(df, of), (p, df, of)
        assert np.allclose(dc, oc, rtol=1e-4, atol=1e-5), (p, dc, oc)
        words += dc.size
        same += int((dc.view(np.uint32) == np.ascontiguousarray(oc, dtype=np.float32).view(np.uint32)).sum())
    assert len(r["contacts"]) == len(ocontacts) > n // 4
    assert same / words > 0.999


# ---- CPU: the DEVICE source compiled for the host (tests/host_shim/dim2_host.cpp) against the oracle, bit for bit ------------------
@pytest.fixture(scope="module")
def dim2_shim():
    from test_device_source_on_host import _build_shim

    return _build_shim("libdim2_host.so", "dim2_host.cpp")


def _vp(a):
    import ctypes as C

    return C.c_void_p(a.ctypes.data)


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("seed,kinds,prediction", [(11, (0, 1, 2), 0.0), (12, (1, 2), 0.02), (13, (0, 1), 0.3), (14, (2,), 0.05), (15, (0, 2), 0.1),
                                                   (16, (0, 1, 2, 3), 0.05)])
def test_device_source_contact_equals_oracle_bit_for_bit(dim2_shim, oracle, seed, kinds, prediction):
    """query::contact: the functions k_contact2d runs per pair, compiled for the host with -ffp-contract=off (= --fmad=false), give
    the oracle's answer in every bit — 2-D GJK, EPA2, the ball / cuboid / polygon projections."""
    import ctypes as C

    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(40000, seed, kinds)
    n = len(t1)
    found, out, flags = np.zeros(n, dtype=np.uint8), np.zeros((n, 7), dtype=np.float32), np.zeros(2, dtype=np.uint32)
    dim2_shim.shim2_contact(C.c_uint64(n), _vp(t1), _vp(p1), _vp(m1), _vp(t2), _vp(p2), _vp(m2), _vp(pts), _vp(nrm), C.c_float(prediction),
                            _vp(found), _vp(out), _vp(flags))
    ofound, oout, opanics = oracle.contact2d(t1, p1, m1, t2, p2, m2, pts, prediction, poly_normals=nrm)
    assert flags.tolist() == [opanics, 0]
    assert np.array_equal(found, ofound)
    hit = found.astype(bool)
    assert hit.sum() > 5000
    assert np.array_equal(_bits(out[hit]), _bits(oout[hit])), f"{(_bits(out[hit]) != _bits(oout[hit])).sum()} words differ"


@pytest.mark.parametrize("n,seed,kinds,angular,planes", [(2500, 51, (0, 1, 2), 0.0, 0), (2000, 52, (1, 2), 0.1, 0), (1500, 53, (2,), 0.3, 0),
                                                         (1800, 54, (0, 1, 2), 0.0, 3)])
def test_device_source_world2d_equals_oracle_bit_for_bit(dim2_shim, oracle, n, seed, kinds, angular, planes):
    """The 2-D world's AABBs and manifolds from the device source on the host: boxes, manifold sizes, feature ids and every contact
    word equal the oracle's."""
    import ctypes as C

    w = random_world(n, seed, kinds, angular=angular, planes=planes)
    pairs, off, ocontacts, ofeats, panics, fat = oracle.world_update2d(w)
    boxes = np.zeros((w.n, 6), dtype=np.float32)
    dim2_shim.shim2_aabbs(C.c_uint32(w.n), _vp(w.pos), _vp(w.rot), _vp(w.type), _vp(w.param), _vp(w.query_limit), _vp(w.points), _vp(w.normals),
                          C.c_float(w.margin), _vp(boxes))
    assert np.array_equal(_bits(boxes), _bits(fat))
    P = len(pairs)
    pr = np.ascontiguousarray(pairs, dtype=np.uint32)
    doff, dc, df = np.zeros(P + 1, dtype=np.uint32), np.zeros((4 * P + 16, 7), dtype=np.float32), np.zeros((4 * P + 16, 2), dtype=np.uint32)
    flags = np.zeros(3, dtype=np.uint32)
    dim2_shim.shim2_narrow.restype = C.c_uint64
    nc = dim2_shim.shim2_narrow(C.c_uint32(w.n), _vp(w.pos), _vp(w.rot), _vp(w.type), _vp(w.param), _vp(w.query_limit), _vp(w.ang_pred),
                                _vp(w.points), _vp(w.normals), C.c_uint64(P), _vp(pr), _vp(doff), _vp(dc), _vp(df), C.c_uint64(len(dc)), _vp(flags))
    assert flags.tolist() == [panics, 0, 0]
    assert np.array_equal(doff, off) and nc == len(ocontacts) > n // 4
    assert np.array_equal(df[:nc], ofeats)
    assert np.array_equal(_bits(dc[:nc]), _bits(ocontacts))


@pytest.mark.parametrize("seed,kinds", [(61, (0, 1, 2)), (62, (1, 2)), (63, (0, 1, 2, 3)), (64, (2,))])
def test_device_source_proximity_equals_oracle(dim2_shim, oracle, seed, kinds):
    """query::proximity from the device source on the host: every status equals the oracle's (GJK's proximity exits included)."""
    import ctypes as C

    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(40000, seed, kinds)
    margins = np.random.default_rng(seed).uniform(0.0, 0.4, size=len(t1)).astype(np.float32)
    margins[::7] = 0.0
    out = np.full(len(t1), 9, dtype=np.uint8)
    dim2_shim.shim2_proximity(C.c_uint64(len(t1)), _vp(t1), _vp(p1), _vp(m1), _vp(t2), _vp(p2), _vp(m2), _vp(pts), _vp(margins), _vp(out))
    want = oracle.proximity2d(t1, p1, m1, t2, p2, m2, pts, margins)
    assert np.array_equal(out, want), np.flatnonzero(out != want)[:10]
    assert min(np.bincount(want, minlength=3)) > 2000


# ---- sensors in the 2-D world (GeometricQueryType::Proximity) -------------------------------------------------------------------
def _sensor_world(n, seed, planes=0):
    w = random_world(n, seed, (0, 1, 2), planes=planes)
    w.set_sensors(np.random.default_rng(seed).random(w.n) < 0.25)
    return w


def test_oracle_world2d_sensors(oracle64):
    """ORACLE properties: a pair with a sensor carries no manifold and the status query::proximity gives for the two shapes with
    margin = the two query limits added; the other pairs' manifolds are those of the same world without sensors."""
    w = _sensor_world(2500, 71, planes=2)
    pairs, off, contacts, feats, panics, fat = oracle64.world_update2d(w)
    prox = oracle64.last_proximity2d
    sensor = (w.query_kind[pairs[:, 0]] | w.query_kind[pairs[:, 1]]).astype(bool)
    both_planes = (w.type[pairs[:, 0]] == 3) & (w.type[pairs[:, 1]] == 3)
    assert sensor.sum() > 500 and (~sensor).sum() > 500
    assert (prox[~sensor] == 255).all() and (prox[sensor & ~both_planes] <= 2).all()
    assert (np.diff(off)[sensor] == 0).all()
    i1, i2 = pairs[sensor & ~both_planes, 0], pairs[sensor & ~both_planes, 1]
    pose = lambda i: np.concatenate([w.pos[i], w.rot[i]], axis=1)  # noqa: E731
    want = oracle64.proximity2d(w.type[i1], w.param[i1], pose(i1), w.type[i2], w.param[i2], pose(i2), w.points, w.query_limit[i1] + w.query_limit[i2])
    assert np.array_equal(prox[sensor & ~both_planes], want)
    assert min(np.bincount(want, minlength=3)[:2]) > 50
    plain = random_world(2500, 71, (0, 1, 2), planes=2)
    p2, off2, c2, f2, _, _ = oracle64.world_update2d(plain)
    assert np.array_equal(p2, pairs)
    for k in np.flatnonzero(~sensor)[:400]:
        assert np.array_equal(contacts[off[k] : off[k + 1]], c2[off2[k] : off2[k + 1]])


def test_device_source_world2d_sensors_equal_oracle(dim2_shim, oracle):
    import ctypes as C

    w = _sensor_world(2500, 72, planes=2)
    pairs, off, ocontacts, ofeats, panics, fat = oracle.world_update2d(w)
    oprox = oracle.last_proximity2d
    P = len(pairs)
    pr = np.ascontiguousarray(pairs, dtype=np.uint32)
    doff, dc, df = np.zeros(P + 1, dtype=np.uint32), np.zeros((4 * P + 16, 7), dtype=np.float32), np.zeros((4 * P + 16, 2), dtype=np.uint32)
    flags, prox = np.zeros(3, dtype=np.uint32), np.zeros(P, dtype=np.uint8)
    dim2_shim.shim2_narrow_sensors.restype = C.c_uint64
    nc = dim2_shim.shim2_narrow_sensors(C.c_uint32(w.n), _vp(w.pos), _vp(w.rot), _vp(w.type), _vp(w.param), _vp(w.query_limit), _vp(w.ang_pred),
                                        _vp(w.points), _vp(w.normals), C.c_uint64(P), _vp(pr), _vp(doff), _vp(dc), _vp(df), C.c_uint64(len(dc)),
                                        _vp(flags), _vp(w.query_kind), _vp(prox))
    assert np.array_equal(prox, oprox) and (prox != 255).sum() > 500
    assert np.array_equal(doff, off) and nc == len(ocontacts)
    assert np.array_equal(_bits(dc[:nc]), _bits(ocontacts)) and np.array_equal(df[:nc], ofeats)


@pytest.mark.gpu
def test_device_world2d_sensors_match_oracle(ctx, oracle):
    from ncollide_b200._ffi import NcbError

    w = _sensor_world(6000, 73, planes=2)
    res = dim2.world_update(ctx, w)
    pairs, off, ocontacts, ofeats, panics, fat = oracle.world_update2d(w)
    oprox = oracle.last_proximity2d
    got = {tuple(p): k for k, p in enumerate(res["pairs"].tolist())}
    assert len(got) == len(pairs) and set(got) == set(map(tuple, pairs.tolist()))
    order = np.array([got[tuple(p)] for p in pairs.tolist()])
    assert np.array_equal(res["proximity"][order], oprox)
    assert (oprox != 255).sum() > 1000 and min(np.bincount(oprox[oprox != 255], minlength=3)[:2]) > 50
    assert np.array_equal(res["manifold_count"][order], np.diff(off))
    w.query_kind[0] = 7
    with pytest.raises(NcbError):
        dim2.world_update(ctx, w)


# ---- Segment as a shape (shape/segment.rs, dim2) --------------------------------------------------------------------------------
def test_oracle_segments_against_separating_axes(oracle64):
    """ORACLE check (f64): segment x cuboid / polygon / segment through contact_support_map_support_map against the separating-axis
    answer (the Minkowski difference's faces come from the polygon's edges and the segment's two sides)."""
    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(2500, 101, kinds=(1, 2, 4), spread=0.9)
    keep = (t1 == 4) | (t2 == 4)
    t1, p1, m1, t2, p2, m2 = t1[keep], p1[keep], m1[keep], t2[keep], p2[keep], m2[keep]
    found, out, panics = oracle64.contact2d(t1, p1, m1, t2, p2, m2, pts, prediction=0.3, poly_normals=nrm)
    assert panics == 0
    checked = deep = 0
    for k in range(len(t1)):
        A, B = _world_polygon(t1[k], p1[k], m1[k], pts), _world_polygon(t2[k], p2[k], m2[k], pts)
        if t1[k] == 4 and t2[k] == 4:
            da, db = A[1] - A[0], B[1] - B[0]
            if abs(da[0] * db[1] - da[1] * db[0]) < 1e-3 * np.linalg.norm(da) * np.linalg.norm(db):
                continue  # (nearly) parallel segments: a degenerate Minkowski difference
        sd = _sat_signed_distance(A, B)
        if sd > 0.3 + 1e-6:
            assert not found[k], (k, sd)
        elif sd < 0.3 - 1e-6:
            assert found[k], (k, sd)
            assert abs(out[k, 6] - (-sd)) < 2e-5 * max(1.0, abs(sd)), (k, t1[k], t2[k], out[k, 6], -sd)
            checked += 1
            deep += sd < -1e-3
    assert checked > 500 and deep > 100, (checked, deep)


def test_oracle_ball_segment_against_point_segment_distance(oracle64):
    """ORACLE check (f64): contact_ball_convex_polyhedron with a Segment: depth = radius - distance(centre, segment), both orders."""
    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(3000, 102, kinds=(0, 4), spread=0.9)
    keep = t1 != t2
    t1, p1, m1, t2, p2, m2 = t1[keep], p1[keep], m1[keep], t2[keep], p2[keep], m2[keep]
    found, out, panics = oracle64.contact2d(t1, p1, m1, t2, p2, m2, pts, prediction=0.1)
    checked = 0
    for k in range(len(t1)):
        ball_first = t1[k] == 0
        c = (m1 if ball_first else m2)[k, :2].astype(np.float64)
        r = float((p1 if ball_first else p2)[k, 0])
        S = _world_polygon(4, (p2 if ball_first else p1)[k], (m2 if ball_first else m1)[k], pts)
        ab = S[1] - S[0]
        u = np.clip(((c - S[0]) @ ab) / (ab @ ab), 0, 1)
        depth = r - np.linalg.norm(S[0] + ab * u - c)
        if abs(depth + 0.1) < 1e-6:
            continue
        assert bool(found[k]) == (depth > -0.1), (k, depth)
        if found[k]:
            assert abs(out[k, 6] - depth) < 1e-9, (k, out[k, 6], depth)
            n = out[k, 4:6] if ball_first else -out[k, 4:6]
            assert n @ (S[0] + ab * u - c) >= -1e-12  # the normal points from the ball towards the segment
            checked += 1
    assert checked > 500


def test_shapes2d_segment():
    sh = dim2.Shapes2D().segment((0, 1), (2, 3))
    typ, par, pts, nrm = sh.arrays()
    assert typ.tolist() == [4] and par[0].tolist() == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        dim2.Shapes2D().segment((1, 1), (1, 1))


@pytest.mark.parametrize("seed,kinds,prediction", [(111, (0, 1, 2, 4), 0.05), (112, (4,), 0.1), (113, (3, 4), 0.05), (114, (2, 4), 0.0)])
def test_device_source_segment_contacts_equal_oracle_bit_for_bit(dim2_shim, oracle, seed, kinds, prediction):
    import ctypes as C

    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(40000, seed, kinds)
    n = len(t1)
    found, out, flags = np.zeros(n, dtype=np.uint8), np.zeros((n, 7), dtype=np.float32), np.zeros(2, dtype=np.uint32)
    dim2_shim.shim2_contact(C.c_uint64(n), _vp(t1), _vp(p1), _vp(m1), _vp(t2), _vp(p2), _vp(m2), _vp(pts), _vp(nrm), C.c_float(prediction),
                            _vp(found), _vp(out), _vp(flags))
    ofound, oout, opanics = oracle.contact2d(t1, p1, m1, t2, p2, m2, pts, prediction, poly_normals=nrm)
    assert flags.tolist() == [opanics, 0] and np.array_equal(found, ofound)
    hit = found.astype(bool)
    assert hit.sum() > 2000 and np.array_equal(_bits(out[hit]), _bits(oout[hit]))
    margins = np.random.default_rng(seed).uniform(0.0, 0.4, size=n).astype(np.float32)
    st = np.full(n, 9, dtype=np.uint8)
    dim2_shim.shim2_proximity(C.c_uint64(n), _vp(t1), _vp(p1), _vp(m1), _vp(t2), _vp(p2), _vp(m2), _vp(pts), _vp(margins), _vp(st))
    assert np.array_equal(st, oracle.proximity2d(t1, p1, m1, t2, p2, m2, pts, margins))


@pytest.mark.parametrize("n,seed,kinds,angular,planes", [(2500, 121, (0, 1, 2, 4), 0.0, 0), (2000, 122, (1, 4), 0.1, 2), (1500, 123, (4,), 0.3, 0)])
def test_device_source_world2d_with_segments_equals_oracle(dim2_shim, oracle, n, seed, kinds, angular, planes):
    import ctypes as C

    w = random_world(n, seed, kinds, angular=angular, planes=planes)
    pairs, off, ocontacts, ofeats, panics, fat = oracle.world_update2d(w)
    boxes = np.zeros((w.n, 6), dtype=np.float32)
    dim2_shim.shim2_aabbs(C.c_uint32(w.n), _vp(w.pos), _vp(w.rot), _vp(w.type), _vp(w.param), _vp(w.query_limit), _vp(w.points), _vp(w.normals),
                          C.c_float(w.margin), _vp(boxes))
    assert np.array_equal(_bits(boxes), _bits(fat))
    P = len(pairs)
    pr = np.ascontiguousarray(pairs, dtype=np.uint32)
    doff, dc, df = np.zeros(P + 1, dtype=np.uint32), np.zeros((4 * P + 16, 7), dtype=np.float32), np.zeros((4 * P + 16, 2), dtype=np.uint32)
    flags = np.zeros(3, dtype=np.uint32)
    dim2_shim.shim2_narrow.restype = C.c_uint64
    nc = dim2_shim.shim2_narrow(C.c_uint32(w.n), _vp(w.pos), _vp(w.rot), _vp(w.type), _vp(w.param), _vp(w.query_limit), _vp(w.ang_pred),
                                _vp(w.points), _vp(w.normals), C.c_uint64(P), _vp(pr), _vp(doff), _vp(dc), _vp(df), C.c_uint64(len(dc)), _vp(flags))
    assert flags.tolist() == [panics, 0, 0]
    assert np.array_equal(doff, off) and nc == len(ocontacts) > n // 5
    assert np.array_equal(df[:nc], ofeats) and np.array_equal(_bits(dc[:nc]), _bits(ocontacts))


@pytest.mark.gpu
def test_device_segments_match_oracle(ctx, oracle):
    t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(60000, 131, (0, 1, 2, 3, 4))
    found, out, info = dim2.contact(ctx, t1, p1, m1, t2, p2, m2, pts, 0.05, poly_normals=nrm)
    ofound, oout, opanics = oracle.contact2d(t1, p1, m1, t2, p2, m2, pts, 0.05, poly_normals=nrm)
    assert info["epa_overflow"] == 0 and info["ref_panics"] == opanics and np.array_equal(found, ofound.astype(bool))
    assert np.allclose(out[found], oout[found], rtol=1e-4, atol=1e-5)
    assert (out[found].view(np.uint32) == np.ascontiguousarray(oout[found], dtype=np.float32).view(np.uint32)).mean() > 0.999
    w = random_world(6000, 132, (0, 1, 2, 4), angular=0.1, planes=2)
    res = dim2.world_update(ctx, w)
    pairs, off, ocontacts, ofeats, panics, fat = oracle.world_update2d(w)
    got = {tuple(p): k for k, p in enumerate(res["pairs"].tolist())}
    assert len(got) == len(pairs) and set(got) == set(map(tuple, pairs.tolist()))
    order = np.array([got[tuple(p)] for p in pairs.tolist()])
    assert np.array_equal(res["manifold_count"][order], np.diff(off)) and res["diag"]["ref_panics"] == panics
    for k in np.flatnonzero(np.diff(off))[:3000]:
        a = res["contacts"][res["manifold_start"][order[k]] : res["manifold_start"][order[k]] + res["manifold_count"][order[k]]]
        assert np.allclose(a, ocontacts[off[k] : off[k + 1]], rtol=1e-4, atol=1e-5), k
    from ncollide_b200._ffi import NcbError

    with pytest.raises(NcbError):  # a segment with identical end points
        dim2.contact(ctx, [4], [[1, 1, 1, 1]], [[0, 0, 1, 0]], [0], [[0.5, 0, 0, 0]], [[0.2, 0, 1, 0]])


# ---- more of the reference's own 2-D examples, pinned on the oracle and on the device source (host shim) ---------------------------
def test_reference_examples_contact_point_and_ray_queries_2d(oracle, oracle64, dim2_shim):
    """examples2d/contact_query2d.rs (depth > 0 / depth < 0 / None for the ball at (1,1), (2,2), (3,3), prediction 1),
    examples2d/solid_point_query2d.rs (cuboid (1, 2): the origin is 1 inside, (2, 2) is 1 outside — read off the ball x cuboid contact
    depth, radius 0.25) and examples2d/solid_ray_cast2d.rs (the solid cast from inside answers 0.0, the ray from (2, 2) misses)."""
    import ctypes as C

    ball, cub, one = [1, 0, 0, 0], [1, 1, 0, 0], [0, 0, 1, 0]
    args = ([0] * 3, [ball] * 3, [[1, 1, 1, 0], [2, 2, 1, 0], [3, 3, 1, 0]], [1] * 3, [cub] * 3, [one] * 3)
    for orc in (oracle, oracle64):
        found, out, _ = orc.contact2d(*args, prediction=1.0)
        assert found.tolist() == [1, 1, 0] and out[0, 6] > 0 and out[1, 6] < 0
        small, box = [0.25, 0, 0, 0], [1, 2, 0, 0]
        found, out, _ = orc.contact2d([0, 0], [small] * 2, [[0, 0, 1, 0], [2, 2, 1, 0]], [1, 1], [box] * 2, [one] * 2, prediction=2.0)
        assert found.all() and out[0, 6] == 1.25 and out[1, 6] == -0.75  # distance_to_point(.., false) == -1.0 / 1.0
        assert orc.contains_point2d([1, 1], [box] * 2, [one] * 2, [[0, 0], [2, 2]]).tolist() == [True, False]
        big = np.finfo(orc.dtype).max
        f, o, _ = orc.ray_cast2d([1, 1], [box] * 2, [one] * 2, [[0, 0, 0, 1, big], [2, 2, 1, 1, big]])
        assert f.tolist() == [1, 0] and o[0, 0] == 0.0
        # examples2d/distance_query2d.rs: the ball at (0, 1) intersects the cuboid (distance 0), at (0, 3) it is 1.0 away (epsilon 1e-7)
        found, out, _ = orc.contact2d([0, 0], [ball] * 2, [[0, 1, 1, 0], [0, 3, 1, 0]], [1, 1], [cub] * 2, [one] * 2, prediction=2.0)
        assert found.all() and out[0, 6] >= 0 and abs(-out[1, 6] - 1.0) <= 1e-7
    # the device source on the host gives the same answers
    t1, p1, m1, t2, p2, m2 = (np.ascontiguousarray(a, dtype=dt) for a, dt in zip(args, (np.uint32, F, F, np.uint32, F, F)))
    found, out, flags = np.zeros(3, dtype=np.uint8), np.zeros((3, 7), dtype=F), np.zeros(2, dtype=np.uint32)
    dim2_shim.shim2_contact(C.c_uint64(3), _vp(t1), _vp(p1), _vp(m1), _vp(t2), _vp(p2), _vp(m2), None, None, C.c_float(1.0), _vp(found), _vp(out),
                            _vp(flags))
    assert found.tolist() == [1, 1, 0] and out[0, 6] > 0 and out[1, 6] < 0


def test_reference_example_dbvt_broad_phase2d(oracle):
    """examples2d/dbvt_broad_phase2d.rs: four balls of radius 0.5 on the corners of a square of side 0.5, broad-phase margin 0.2:
    6 interferences; without the first two proxies: 1."""
    pos = np.array([[0, 0], [0, 0.5], [0.5, 0], [0.5, 0.5]], dtype=F)
    four = dim2.World2D(dim2.Shapes2D().ball(0.5).ball(0.5).ball(0.5).ball(0.5), pos, 0.0, margin=0.2, linear=0.0)
    assert len(oracle.world_update2d(four)[0]) == 6
    two = dim2.World2D(dim2.Shapes2D().ball(0.5).ball(0.5), pos[2:], 0.0, margin=0.2, linear=0.0)
    assert len(oracle.world_update2d(two)[0]) == 1


def test_oracle_world2d_contact_points_lie_on_their_shapes(oracle64):
    """ORACLE check (f64) of the 2-D manifolds: world1 lies on the boundary of object 1 and world2 on the boundary of object 2 (circle,
    box outline, polygon outline, segment, half-plane line); depth == -n . (w2 - w1); normals are unit vectors."""
    w = random_world(2500, 151, (0, 1, 2, 4), angular=0.05, planes=2)
    pairs, off, c, feats, panics, fat = oracle64.world_update2d(w)
    assert panics == 0 and len(c) > 1500

    def boundary_distance(i, p):
        m = np.array([w.pos[i, 0], w.pos[i, 1], w.rot[i, 0], w.rot[i, 1]], dtype=np.float64)
        d = p - m[:2]
        loc = np.array([m[2] * d[0] + m[3] * d[1], -m[3] * d[0] + m[2] * d[1]])
        t, par = int(w.type[i]), w.param[i].astype(np.float64)
        if t == 0:
            return abs(np.hypot(*loc) - par[0])
        if t == 3:
            return abs(par[0] * loc[0] + par[1] * loc[1])
        if t == 1:
            P = np.array([[par[0], par[1]], [-par[0], par[1]], [-par[0], -par[1]], [par[0], -par[1]]])
        elif t == 4:
            P = np.array([[par[0], par[1]], [par[2], par[3]]])
        else:
            P = w.points[int(par[0]) : int(par[0]) + int(par[1])].astype(np.float64)
        a, b = P, np.roll(P, -1, axis=0)
        ab = b - a
        u = np.clip(((loc - a) * ab).sum(axis=1) / (ab * ab).sum(axis=1), 0, 1)
        return np.linalg.norm(a + ab * u[:, None] - loc, axis=1).min()

    def convex(i):  # random_world's angular offsets can wrap the last vertex past the first one: such a polygon is not convex
        if w.type[i] != 2:  # (ConvexPolygon::try_new does not check either) and the projection of an inner point is not defined
            return True
        P = w.points[int(w.param[i, 0]) : int(w.param[i, 0]) + int(w.param[i, 1])].astype(np.float64)
        e = np.roll(P, -1, axis=0) - P
        f = np.roll(e, -1, axis=0)
        return bool(np.all(e[:, 0] * f[:, 1] - e[:, 1] * f[:, 0] > 0))

    worst, skipped = 0.0, 0
    for p, (i1, i2) in enumerate(pairs):
        if not (convex(i1) and convex(i2)):
            skipped += off[p + 1] - off[p]
            continue
        for k in range(off[p], off[p + 1]):
            w1, w2, n, depth = c[k, 0:2], c[k, 2:4], c[k, 4:6], c[k, 6]
            worst = max(worst, boundary_distance(i1, w1), boundary_distance(i2, w2))
            assert abs(np.hypot(*n) - 1) < 5e-7 and abs(depth + n @ (w2 - w1)) < 1e-6 * max(1.0, abs(depth))  # the f32 rotations / normals are unit to 1e-7
    assert worst < 1e-6 and skipped < len(c) // 4, (worst, skipped)


# ---- the committed 2-D fixture (tests/golden/dim2_world_600.npz, made by tests/golden/make_golden.py dim2) ---------------------------
def _dim2_fixture():
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dim2_world_600.npz"))
    w = dim2.World2D.__new__(dim2.World2D)
    for k in ("pos", "rot", "type", "param", "points", "normals", "query_limit", "ang_pred", "groups", "query_kind"):
        setattr(w, k, np.ascontiguousarray(z[k]))
    w.n, w.margin = len(w.type), float(z["margin"])
    return z, w


def test_dim2_golden_fixture_oracle(oracle):
    """Guards the 2-D restatement against drift: the oracle reproduces the committed world (boxes, pairs, manifolds, features, sensor
    statuses), its ray queries, the shape ray casts and the polyline casts, bit for bit."""
    z, w = _dim2_fixture()
    pairs, off, c, feats, panics, fat = oracle.world_update2d(w)
    assert np.array_equal(_bits(fat), _bits(z["fat"])) and np.array_equal(pairs, z["pairs"]) and np.array_equal(off, z["off"])
    assert np.array_equal(_bits(c), _bits(z["contacts"])) and np.array_equal(feats, z["feats"]) and panics == int(z["panics"])
    assert np.array_equal(oracle.last_proximity2d, z["prox"])
    for first, tag in ((False, "q_"), (True, "q_first_")):
        idx, val, ft = oracle.world_ray_cast2d(w, z["q_rays"], first_only=first)
        assert np.array_equal(idx, z[tag + "idx"]) and np.array_equal(_bits(val), _bits(z[tag + "val"])) and np.array_equal(ft, z[tag + "feat"])
    f, out, sf = oracle.ray_cast2d(z["s_type"], z["s_param"], z["s_pose"], z["s_rays"], z["s_points"])
    assert np.array_equal(f, z["s_found"]) and np.array_equal(_bits(out), _bits(z["s_out"])) and np.array_equal(sf, z["s_feat"])
    toi, pf, pn = oracle.polyline(z["p_points"], None).ray_cast(z["p_origins"], z["p_dirs"], mode=0)
    assert np.array_equal(_bits(toi), _bits(z["p_toi"])) and np.array_equal(pf, z["p_feat"]) and np.array_equal(_bits(pn), _bits(z["p_normal"]))


def test_dim2_golden_fixture_device_source(dim2_shim):
    """The device source on the host replays the fixture WITHOUT the oracle: boxes, manifolds, features, sensor statuses, shape rays."""
    import ctypes as C

    z, w = _dim2_fixture()
    boxes = np.zeros((w.n, 6), dtype=np.float32)
    dim2_shim.shim2_aabbs(C.c_uint32(w.n), _vp(w.pos), _vp(w.rot), _vp(w.type), _vp(w.param), _vp(w.query_limit), _vp(w.points), _vp(w.normals),
                          C.c_float(w.margin), _vp(boxes))
    assert np.array_equal(_bits(boxes), _bits(z["fat"]))
    pr = np.ascontiguousarray(z["pairs"], dtype=np.uint32)
    P = len(pr)
    doff, dc, df = np.zeros(P + 1, dtype=np.uint32), np.zeros((4 * P + 16, 7), dtype=np.float32), np.zeros((4 * P + 16, 2), dtype=np.uint32)
    flags, prox = np.zeros(3, dtype=np.uint32), np.full(P, 255, dtype=np.uint8)
    dim2_shim.shim2_narrow_sensors.restype = C.c_uint64
    nc = dim2_shim.shim2_narrow_sensors(C.c_uint32(w.n), _vp(w.pos), _vp(w.rot), _vp(w.type), _vp(w.param), _vp(w.query_limit), _vp(w.ang_pred),
                                        _vp(w.points), _vp(w.normals), C.c_uint64(P), _vp(pr), _vp(doff), _vp(dc), _vp(df), C.c_uint64(len(dc)),
                                        _vp(flags), _vp(w.query_kind), _vp(prox))
    assert np.array_equal(doff, z["off"]) and nc == len(z["contacts"]) and np.array_equal(prox, z["prox"])
    assert np.array_equal(_bits(dc[:nc]), _bits(z["contacts"])) and np.array_equal(df[:nc], z["feats"])
    n = len(z["s_type"])
    f, out, sf = np.zeros(n, dtype=np.uint8), np.zeros((n, 3), dtype=np.float32), np.zeros(n, dtype=np.uint32)
    st, sp, sm, sr, spts = (np.ascontiguousarray(z[k]) for k in ("s_type", "s_param", "s_pose", "s_rays", "s_points"))
    dim2_shim.shim2_ray_cast(C.c_uint64(n), _vp(st), _vp(sp), _vp(sm), _vp(spts), _vp(sr), _vp(f), _vp(out), _vp(sf))
    hit = f.astype(bool)
    assert np.array_equal(f, z["s_found"]) and np.array_equal(sf, z["s_feat"]) and np.array_equal(_bits(out[hit]), _bits(z["s_out"][hit]))
