"""CPU-only tests: the C-ABI library builds, loads and exports what include/ncb200.h declares; the product never
touches the oracle and fails loudly without a GPU; host-side mirrors (hull tables, scenes); oracle self-consistency."""
import os
import sys
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from ncollide_b200 import _ffi
    from ncollide_b200.build import build_extension

    build_extension()
    lib = _ffi.load_library()
    header = open(os.path.join(ROOT, "include", "ncb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ncb(?:2d)?_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in ncb200.h but not exported"
    assert declared == set(_ffi.EXPORTED_SYMBOLS)
    assert b"sm_100a" in lib.ncb_version()


def test_built_for_sm_100a_only():
    import subprocess

    from ncollide_b200 import _ffi

    out = subprocess.run(["cuobjdump", "-lelf", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_gpu_means_loud_failure_not_a_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ncollide_b200._ffi import NcbError
    from ncollide_b200.world import Context

    with pytest.raises(NcbError, match="no CUDA device|CPU fallback"):
        Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ncollide_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "oracle/" not in text.replace("the oracle/", ""), f


def test_contact_struct_layout_matches_header():
    from ncollide_b200._ffi import CONTACT_DTYPE, UpdateCountsC
    import ctypes

    assert CONTACT_DTYPE.itemsize == 52
    assert [CONTACT_DTYPE.fields[n][1] for n in ("world1", "world2", "normal", "depth", "f1", "f2", "pair")] == [0, 12, 24, 36, 40, 44, 48]
    assert ctypes.sizeof(UpdateCountsC) == 4 * 21  # 13 + n_proximity_pairs + n_proximity[3] + n_capsule_pairs[2] + stack_overflow + n_epa_restarts


# ---- scenes / shapes -------------------------------------------------------------------------------------------
def test_scene_generators_are_seeded_and_shaped():
    from ncollide_b200.scenes import box_side_for, config_scene, make_ray_scene

    a, b = config_scene(3, 3000), config_scene(3, 3000)
    for f in ("pos", "rot", "shape_type", "shape_param", "query_limit"):
        assert np.array_equal(getattr(a, f), getattr(b, f))
    assert a.pos.dtype == np.float32 and a.rot.shape == (3000, 4)
    assert np.allclose(np.linalg.norm(a.rot, axis=1), 1, atol=1e-6)
    assert abs(box_side_for(100_000) - 52.6) < 0.1 and abs(box_side_for(1_000_000) - 113.4) < 0.1
    assert set(np.unique(a.shape_type)) == {0, 1, 2}
    c2 = config_scene(2, 1000)
    assert c2.shape_type[-1] == 3 and c2.n == 1001
    rs = make_ray_scene("terrain", 20000, 100)
    assert abs(len(rs.tris) - 20000) < 2000 and rs.tris.max() < len(rs.verts)
    assert np.allclose(np.linalg.norm(rs.dirs, axis=1), 1, atol=1e-5)


def test_random_hulls_satisfy_try_new_invariants():
    from ncollide_b200.scenes import make_hull_library

    lib = make_hull_library(np.random.default_rng(1), 40)
    assert lib.n_hulls == 40 and lib.max_verts <= 32
    for h in lib.hulls:
        assert h.check_geometry()
        nv = int((h.vert_num_adj > 0).sum())
        assert nv + len(h.face_first) - int((~h.edge_deleted).sum()) == 2  # Euler characteristic (convex.rs:316)
        assert h.face_num.sum() == len(h.vertices_adj_to_face) == h.vert_num_adj.sum()
        assert np.allclose(np.linalg.norm(h.face_normal, axis=1), 1, atol=1e-6)


def test_degenerate_hull_inputs_are_rejected():
    from ncollide_b200.shapes import ConvexHull

    # a duplicated point -> zero-length edge -> None (convex.rs:147-158)
    assert ConvexHull.try_new([(0, 0, 0), (0, 0, 0), (1, 0, 0), (0, 1, 0)], [0, 1, 2, 0, 2, 3, 0, 3, 1, 1, 3, 2]) is None
    # an open surface: the reference either panics on its contour-walk assert! (convex.rs:228-230) or fails the
    # Euler characteristic check; the mirror raises or returns None, it never returns tables
    try:
        assert ConvexHull.try_new([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], [0, 2, 1, 0, 1, 3]) is None
    except AssertionError:
        pass


# ---- oracle self-consistency ------------------------------------------------------------------------------------
def canon(p):
    p = np.sort(np.asarray(p).reshape(-1, 2), axis=1)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


@pytest.mark.parametrize("cfg,n", [(1, 1000), (2, 3000), (3, 3000)])
def test_oracle_broad_phase_variants_agree(oracle, cfg, n):
    from ncollide_b200.scenes import config_scene

    s = config_scene(cfg, n)
    fat = oracle.compute_aabbs(s)
    dbvt = oracle.broad_phase(fat, s.groups, 0)
    assert np.all(dbvt[:, 0] > dbvt[:, 1])  # interference_started(later, earlier)
    assert np.array_equal(canon(dbvt), canon(oracle.broad_phase(fat, s.groups, 1)))
    assert np.array_equal(canon(dbvt), canon(oracle.broad_phase(fat, s.groups, 2)))
    # fat boxes: ((tight -+ query_limit) -+ margin) in f32
    tight = oracle.compute_aabbs(s, mode=0)
    ql, m = np.float32(0.02), np.float32(s.margin)
    finite = s.shape_type != 3
    assert np.array_equal(fat[finite, 3:], ((tight[finite, 3:] + ql) + m))
    assert np.array_equal(fat[finite, :3], ((tight[finite, :3] + -ql) + -m))


def test_oracle_pair_set_against_numpy_brute_force(oracle):
    """ORACLE check for an unpinned item (pair identities): the broad phase's pair set == every pair of fat boxes that intersect
    (closed intervals, AABB::intersects) and whose collision groups allow the interaction, enumerated by numpy over all N^2 / 2 pairs."""
    from ncollide_b200.scenes import config_scene

    s = config_scene(3, 4000)
    rng = np.random.default_rng(3)
    s.groups = s.groups.copy()
    pick = rng.random(s.n) < 0.3
    s.groups[pick] = (1 << 2, 0x3FFFFFFF & ~(1 << 2), 0)       # members of group 2 that do not whitelist their own group
    s.groups[rng.random(s.n) < 0.05] = (1 << 4, 0x3FFFFFFF, 1 << 2)  # group 4 blacklists group 2
    fat = oracle.compute_aabbs(s)
    got = canon(oracle.broad_phase(fat, s.groups, 0))
    lo, hi = fat[:, :3], fat[:, 3:]
    want = []
    m, w, b = (s.groups[:, k].astype(np.uint64) for k in range(3))
    for i in range(s.n):
        hit = np.all(lo[i] <= hi[i + 1 :], axis=1) & np.all(lo[i + 1 :] <= hi[i], axis=1)
        j = np.flatnonzero(hit) + i + 1
        # CollisionGroups::can_interact_with_groups (collision_groups.rs:353-359), both ways
        ok = ((m[i] & b[j]) == 0) & ((m[j] & b[i]) == 0) & ((m[i] & w[j]) != 0) & ((m[j] & w[i]) != 0)
        want += [(i, int(k)) for k in j[ok]]
    want = canon(np.array(want, dtype=np.uint32))
    assert len(want) > 3000 and np.array_equal(got, want)


def test_oracle_contact_invariants(oracle):
    from ncollide_b200.scenes import make_world_scene

    s = make_world_scene(1500, 21, (1, 1, 1), side=7.0, n_hulls=16, angular=0.03)
    fat = oracle.compute_aabbs(s)
    pairs = oracle.broad_phase(fat, s.groups, 0)
    c, off, algo, stats = oracle.narrow_phase(s, pairs)
    assert len(c) > 500 and stats[6] == 0
    assert np.allclose(np.linalg.norm(c["normal"], axis=1), 1, atol=1e-5)
    # depth == -n . (w2 - w1) (Contact::new_wo_depth) for every algorithm on the path
    d = -np.einsum("ij,ij->i", c["normal"], c["world2"] - c["world1"])
    assert np.allclose(d, c["depth"], atol=2e-5)
    # prediction: nothing farther than linear1 + linear2
    assert (c["depth"] >= -0.04 - 1e-6).all()
    cnt = np.diff(off)
    assert cnt.max() <= 16 and set(np.unique(algo)) <= {1, 4, 5}
    # swapping the pair order flips the contacts (generators are symmetric up to flip) for ball-ball
    bb = np.nonzero(algo == 1)[0][:50]
    c2, off2, _, _ = oracle.narrow_phase(s, pairs[bb][:, ::-1])
    c1 = np.concatenate([c[off[p] : off[p + 1]] for p in bb])
    assert np.allclose(c2["normal"], -c1["normal"], atol=1e-6) and np.allclose(c2["world1"], c1["world2"], atol=1e-6)


def _rotation_matrices(q):
    """nalgebra's UnitQuaternion::to_rotation_matrix for rows of (i, j, k, w)."""
    i, j, k, w = (q[:, c].astype(np.float64) for c in range(4))
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = w * w + i * i - j * j - k * k, 2 * (i * j - w * k), 2 * (w * j + i * k)
    R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = 2 * (w * k + i * j), w * w - i * i + j * j - k * k, 2 * (j * k - w * i)
    R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = 2 * (i * k - w * j), 2 * (w * i + j * k), w * w - i * i - j * j + k * k
    return R


def test_oracle_aabbs_at_arbitrary_rotations_against_numpy(oracle64):
    """ORACLE check (f64) for an unpinned item (no reference test holds AABB values of rotated shapes): the tight AABB of every ball,
    cuboid and convex hull of a randomly rotated world against numpy — centre -+ r, centre -+ |R| half_extents, min / max of the
    rotated hull vertices."""
    from ncollide_b200.scenes import make_world_scene

    s = make_world_scene(3000, 31, (1, 1, 1), side=20.0, n_hulls=32)
    s.rot = (s.rot.astype(np.float64) / np.linalg.norm(s.rot.astype(np.float64), axis=1, keepdims=True)).astype(np.float64)
    tight = oracle64.compute_aabbs(s, mode=0)
    R, t = _rotation_matrices(s.rot), s.pos.astype(np.float64)
    seen = [0, 0, 0]
    for k in range(s.n):
        typ, par = int(s.shape_type[k]), s.shape_param[k].astype(np.float64)
        if typ == 0:
            lo, hi = t[k] - par[0], t[k] + par[0]
        elif typ == 1:
            he = np.abs(R[k]) @ par[:3]
            lo, hi = t[k] - he, t[k] + he
        else:
            h = int(par[0])
            P = s.hulls.points[s.hulls.vert_off[h] : s.hulls.vert_off[h + 1]].astype(np.float64) @ R[k].T + t[k]
            lo, hi = P.min(axis=0), P.max(axis=0)
        assert np.allclose(tight[k, :3], lo, atol=1e-9) and np.allclose(tight[k, 3:], hi, atol=1e-9), (k, typ)
        seen[typ] += 1
    assert min(seen) > 500


def test_oracle_cuboid_penetration_against_separating_axes(oracle64):
    """ORACLE check (f64) for an unpinned item (manifold contents): for interpenetrating cuboids the deepest contact of the manifold is
    the EPA penetration depth, which for boxes is the smallest overlap over the 15 separating-axis candidates; and pairs the axes
    separate by more than the prediction have no contact."""
    from ncollide_b200.scenes import make_world_scene

    s = make_world_scene(1800, 33, (0, 1, 0), side=9.0, n_hulls=1)
    s.rot = s.rot.astype(np.float64) / np.linalg.norm(s.rot.astype(np.float64), axis=1, keepdims=True)
    fat = oracle64.compute_aabbs(s)
    pairs = oracle64.broad_phase(fat, s.groups, 1)
    c, off, algo, stats = oracle64.narrow_phase(s, pairs)
    R, t = _rotation_matrices(s.rot), s.pos.astype(np.float64)
    deep = apart = 0
    for p, (i1, i2) in enumerate(pairs):
        A, B, ha, hb = R[i1], R[i2], s.shape_param[i1, :3].astype(np.float64), s.shape_param[i2, :3].astype(np.float64)
        d = t[i2] - t[i1]
        axes = [A[:, k] for k in range(3)] + [B[:, k] for k in range(3)]
        for a in range(3):
            for b in range(3):
                x = np.cross(A[:, a], B[:, b])
                if np.linalg.norm(x) > 1e-6:
                    axes.append(x / np.linalg.norm(x))
        overlap = min((np.abs(A.T @ ax) @ ha) + (np.abs(B.T @ ax) @ hb) - abs(d @ ax) for ax in axes)
        depths = c["depth"][off[p] : off[p + 1]]
        if overlap > 1e-3:  # interpenetrating: min overlap == penetration depth
            assert len(depths) > 0, (p, overlap)
            assert abs(depths.max() - overlap) < 1e-6 * max(1.0, overlap), (p, depths.max(), overlap)
            deep += 1
        elif overlap < -0.04 - 1e-6:  # a separating axis with a gap beyond linear1 + linear2: nothing to report
            assert len(depths) == 0, (p, overlap)
            apart += 1
    assert deep > 150 and apart > 20, (deep, apart)


def test_oracle_hull_penetration_against_separating_axes(oracle64):
    """The same check for convex hulls (and hull x cuboid): candidate axes = the face normals of both polytopes and the cross products
    of their edge directions; the smallest overlap of the projected vertex sets == the depth query::contact (EPA) reports."""
    from ncollide_b200.scenes import WorldScene, make_world_scene

    s = make_world_scene(700, 35, (0, 1, 2), side=5.5, n_hulls=24)
    s.rot = s.rot.astype(np.float64) / np.linalg.norm(s.rot.astype(np.float64), axis=1, keepdims=True)
    fat = oracle64.compute_aabbs(s)
    pairs = oracle64.broad_phase(fat, s.groups, 1)
    c, off, algo, stats = oracle64.narrow_phase(s, pairs)
    R, t, H = _rotation_matrices(s.rot), s.pos.astype(np.float64), s.hulls
    local = {}

    def polytope(i):
        """world vertices, world face normals, world edge directions of object i"""
        if s.shape_type[i] == 1:
            he = s.shape_param[i, :3].astype(np.float64)
            V = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64) * he
            return V @ R[i].T + t[i], R[i].T.copy(), R[i].T.copy()
        h = int(s.shape_param[i, 0])
        if h not in local:  # faces and edges recomputed from the vertices by qhull in f64 (the library's own tables merge nearly
            from scipy.spatial import ConvexHull  # coplanar triangles, which moves a face normal by ~1e-4)

            V = H.points[H.vert_off[h] : H.vert_off[h + 1]].astype(np.float64)
            ch = ConvexHull(V)
            tri = ch.simplices
            E = np.concatenate([V[tri[:, 1]] - V[tri[:, 0]], V[tri[:, 2]] - V[tri[:, 1]], V[tri[:, 0]] - V[tri[:, 2]]])
            local[h] = (V, ch.equations[:, :3].copy(), E / np.linalg.norm(E, axis=1, keepdims=True))
        V, N, E = local[h]
        return V @ R[i].T + t[i], N @ R[i].T, E @ R[i].T

    deep = 0
    for p, (i1, i2) in enumerate(pairs):
        if s.shape_type[i1] == 1 and s.shape_type[i2] == 1:
            continue
        Va, Na, Ea = polytope(i1)
        Vb, Nb, Eb = polytope(i2)
        X = np.cross(Ea[:, None, :], Eb[None, :, :]).reshape(-1, 3)
        ln = np.linalg.norm(X, axis=1)
        axes = np.concatenate([Na, Nb, X[ln > 1e-6] / ln[ln > 1e-6, None]])
        pa, pb = Va @ axes.T, Vb @ axes.T
        overlap = np.minimum(pa.max(axis=0) - pb.min(axis=0), pb.max(axis=0) - pa.min(axis=0)).min()
        depths = c["depth"][off[p] : off[p + 1]]
        if overlap > 1e-3:
            # (the manifold itself holds clipped feature points measured along the EPA normal against the other shape's face plane:
            # its deepest contact equals the penetration depth for boxes, above, and only approximates it for general hulls)
            assert len(depths) > 0
            # query::contact of the same two shapes reports the EPA depth itself
            two = WorldScene(pos=s.pos[[i1, i2]], rot=s.rot[[i1, i2]], shape_type=s.shape_type[[i1, i2]], shape_param=s.shape_param[[i1, i2]],
                             groups=s.groups[[i1, i2]], query_limit=s.query_limit[[i1, i2]], ang_pred=s.ang_pred[[i1, i2]], hulls=s.hulls,
                             margin=s.margin)
            q = oracle64.query_contact(two, 0.04)
            assert q is not None and abs(q["depth"] - overlap) < 1e-7 * max(1.0, overlap), (p, q["depth"], overlap)
            deep += 1
        elif overlap < -0.04 - 1e-6:
            assert len(depths) == 0, (p, overlap)
    assert deep > 100, deep


def test_oracle_gjk_distance_against_quadratic_program(oracle64):
    """ORACLE check (f64): for separated hull / cuboid pairs query::contact(prediction = 1) reports depth = -distance (GJK); the distance
    is recomputed as the quadratic program min |A^T l - B^T m|^2 over convex weights (SLSQP), and the witness points must realise it."""
    from scipy.optimize import minimize

    from ncollide_b200.scenes import WorldScene, make_world_scene

    s = make_world_scene(500, 37, (0, 1, 2), side=7.0, n_hulls=24)
    s.rot = s.rot.astype(np.float64) / np.linalg.norm(s.rot.astype(np.float64), axis=1, keepdims=True)
    s.query_limit[:] = 0.5
    fat = oracle64.compute_aabbs(s)
    pairs = oracle64.broad_phase(fat, s.groups, 1)
    R, t, H = _rotation_matrices(s.rot), s.pos.astype(np.float64), s.hulls

    def vertices(i):
        if s.shape_type[i] == 1:
            he = s.shape_param[i, :3].astype(np.float64)
            V = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64) * he
        else:
            h = int(s.shape_param[i, 0])
            V = H.points[H.vert_off[h] : H.vert_off[h + 1]].astype(np.float64)
        return V @ R[i].T + t[i]

    checked = 0
    for i1, i2 in pairs[:600]:
        two = WorldScene(pos=s.pos[[i1, i2]], rot=s.rot[[i1, i2]], shape_type=s.shape_type[[i1, i2]], shape_param=s.shape_param[[i1, i2]],
                         groups=s.groups[[i1, i2]], query_limit=s.query_limit[[i1, i2]], ang_pred=s.ang_pred[[i1, i2]], hulls=s.hulls, margin=s.margin)
        q = oracle64.query_contact(two, 1.0)
        if q is None or q["depth"] > -1e-3:
            continue
        A, B = vertices(i1), vertices(i2)
        na, nb = len(A), len(B)

        def f(x):
            d = x[:na] @ A - x[na:] @ B
            return d @ d

        def g(x):
            d = x[:na] @ A - x[na:] @ B
            return np.concatenate([2 * A @ d, -2 * B @ d])

        cons = [{"type": "eq", "fun": lambda x: x[:na].sum() - 1, "jac": lambda x: np.concatenate([np.ones(na), np.zeros(nb)])},
                {"type": "eq", "fun": lambda x: x[na:].sum() - 1, "jac": lambda x: np.concatenate([np.zeros(na), np.ones(nb)])}]
        x0 = np.concatenate([np.full(na, 1 / na), np.full(nb, 1 / nb)])
        r = minimize(f, x0, jac=g, bounds=[(0, 1)] * (na + nb), constraints=cons, method="SLSQP", options={"ftol": 1e-15, "maxiter": 500})
        dist = np.sqrt(r.fun)
        assert abs(-q["depth"] - dist) < 2e-6 * max(1.0, dist), (i1, i2, q["depth"], dist)
        assert abs(np.linalg.norm(q["world2"] - q["world1"]) - dist) < 2e-6  # the witness points are that far apart
        checked += 1
    assert checked > 100, checked


def test_oracle_ball_hull_contact_against_point_to_polytope_distance(oracle64):
    """ORACLE check (f64): the BallConvexPolyhedron generator's depth for ball x hull pairs == radius - distance(centre, hull) with the
    distance from a quadratic program when the centre is outside, and radius + the distance to the nearest face plane (qhull) when it is
    inside (the EPA projection)."""
    from scipy.optimize import minimize
    from scipy.spatial import ConvexHull

    from ncollide_b200.scenes import make_world_scene

    s = make_world_scene(900, 39, (1, 0, 1), side=5.0, n_hulls=24)
    s.rot = s.rot.astype(np.float64) / np.linalg.norm(s.rot.astype(np.float64), axis=1, keepdims=True)
    ball = s.shape_type == 0
    s.shape_param[ball, 0] = np.random.default_rng(1).uniform(0.05, 0.5, size=int(ball.sum()))  # small balls: some end up inside a hull
    fat = oracle64.compute_aabbs(s)
    pairs = oracle64.broad_phase(fat, s.groups, 1)
    c, off, algo, stats = oracle64.narrow_phase(s, pairs)
    R, t, H = _rotation_matrices(s.rot), s.pos.astype(np.float64), s.hulls
    planes = {}
    outside = inside = 0
    for p, (i1, i2) in enumerate(pairs):
        if s.shape_type[i1] == s.shape_type[i2]:
            continue
        ib, ih = (i1, i2) if s.shape_type[i1] == 0 else (i2, i1)
        h = int(s.shape_param[ih, 0])
        V = H.points[H.vert_off[h] : H.vert_off[h + 1]].astype(np.float64)
        if h not in planes:
            planes[h] = ConvexHull(V).equations.copy()
        centre = (t[ib] - t[ih]) @ R[ih]  # the ball's centre in the hull's frame
        r = float(s.shape_param[ib, 0])
        signed = (planes[h][:, :3] @ centre + planes[h][:, 3]).max()  # < 0: inside, the distance to the nearest face plane is -signed
        if signed < -1e-6:
            want = r - signed
            inside += 1
        else:
            res = minimize(lambda x: (x @ V - centre) @ (x @ V - centre), np.full(len(V), 1 / len(V)), jac=lambda x: 2 * V @ (x @ V - centre),
                           bounds=[(0, 1)] * len(V), constraints=[{"type": "eq", "fun": lambda x: x.sum() - 1, "jac": lambda x: np.ones(len(V))}],
                           method="SLSQP", options={"ftol": 1e-15, "maxiter": 500})
            want = r - np.sqrt(res.fun)
            outside += 1
        depths = c["depth"][off[p] : off[p + 1]]
        if want < -0.04 - 1e-5:
            assert len(depths) == 0, (p, want)
        elif want > -0.04 + 1e-5:
            assert len(depths) == 1 and abs(depths[0] - want) < 3e-6, (p, depths, want)
    assert outside > 100 and inside > 10, (outside, inside)


def test_oracle_contact_points_lie_on_their_shapes(oracle64):
    """ORACLE check (f64) of the manifold contents: world1 lies on the boundary of shape 1 and world2 on the boundary of shape 2 — a
    sphere, a box surface, or the surface of the qhull polytope of the hull's vertices — for every contact of a mixed world."""
    from scipy.spatial import ConvexHull

    from ncollide_b200.scenes import make_world_scene

    s = make_world_scene(1500, 41, (1, 1, 1), side=7.0, n_hulls=24, angular=0.03)
    s.rot = s.rot.astype(np.float64) / np.linalg.norm(s.rot.astype(np.float64), axis=1, keepdims=True)
    fat = oracle64.compute_aabbs(s)
    pairs = oracle64.broad_phase(fat, s.groups, 1)
    c, off, algo, stats = oracle64.narrow_phase(s, pairs)
    R, t, H = _rotation_matrices(s.rot), s.pos.astype(np.float64), s.hulls
    planes = {}

    def surface_distance(i, p):
        loc = (p - t[i]) @ R[i]
        typ = int(s.shape_type[i])
        if typ == 0:
            return np.linalg.norm(loc) - float(s.shape_param[i, 0])
        if typ == 1:
            return float((np.abs(loc) - s.shape_param[i, :3].astype(np.float64)).max())
        h = int(s.shape_param[i, 0])
        if h not in planes:
            planes[h] = ConvexHull(H.points[H.vert_off[h] : H.vert_off[h + 1]].astype(np.float64)).equations.copy()
        return float((planes[h][:, :3] @ loc + planes[h][:, 3]).max())

    dist = {0: [], 1: [], 2: []}
    for p, (i1, i2) in enumerate(pairs):
        for k in range(off[p], off[p + 1]):
            dist[int(s.shape_type[i1])].append(abs(surface_distance(i1, c["world1"][k])))
            dist[int(s.shape_type[i2])].append(abs(surface_distance(i2, c["world2"][k])))
    assert min(len(v) for v in dist.values()) > 800
    assert max(dist[0]) < 1e-12 and max(dist[1]) < 1e-12
    # hulls: a face of the ConvexHull tables may merge nearly coplanar triangles (try_new's tolerance); a contact projected onto such a
    # face's plane then sits up to a few mm off the exact polytope — the reference's own behaviour, seen on < 1 % of the contacts
    hull = np.array(dist[2])
    assert np.percentile(hull, 99) < 1e-6 and hull.max() < 5e-3, (np.percentile(hull, 99), hull.max())


def test_oracle_ray_bvt_matches_brute_force(oracle):
    from ncollide_b200.scenes import make_ray_scene

    for kind in ("terrain", "soup"):
        rs = make_ray_scene(kind, 3000, 1500, seed=5)
        om = oracle.trimesh(rs.verts, rs.tris)
        t0, f0, n0 = om.ray_cast(rs.origins, rs.dirs, mode=0)
        t1, f1, n1 = om.ray_cast(rs.origins, rs.dirs, mode=1)
        assert (t1 >= 0).sum() > 50
        diff = np.nonzero(f0 != f1)[0]
        for i in diff:  # only ties within a few ulps may differ
            assert abs(int(t0[i].view(np.int32)) - int(t1[i].view(np.int32))) <= 4
        same = f0 == f1
        assert np.array_equal(t0[same], t1[same])
        hit = t1 >= 0
        assert np.allclose(np.linalg.norm(n1[hit], axis=1), 1, atol=1e-5)
        # hit point lies on the reported triangle's plane
        T = len(rs.tris)
        tri = rs.verts[rs.tris[f1[hit] % T]]
        p = rs.origins[hit] + rs.dirs[hit] * t1[hit, None]
        nn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        nn /= np.linalg.norm(nn, axis=1, keepdims=True)
        assert np.abs(np.einsum("ij,ij->i", p - tri[:, 0], nn)).max() < 1e-3


def test_oracle_trimesh_rays_against_moller_trumbore(oracle64):
    """ORACLE check (f64) for the part of the path the reference has no test for: TriMesh first hits against an independent
    Moller-Trumbore intersection over ALL triangles in numpy — toi, the triangle, front / back face (face + T) and the unit normal
    facing the ray; per-ray max_toi cuts hits beyond it."""
    from ncollide_b200.scenes import make_ray_scene

    checked = back = cut = 0
    for kind in ("terrain", "soup"):
        rs = make_ray_scene(kind, 900 if kind == "terrain" else 4000, 700 if kind == "terrain" else 1500, seed=11, random_pose=(kind == "soup"))
        om = oracle64.trimesh(rs.verts, rs.tris)
        pose = rs.pose.astype(np.float64) if kind == "soup" else None
        if pose is not None:  # the rays were drawn in the mesh's frame: move them along with the mesh
            pose[3:] /= np.linalg.norm(pose[3:])  # a unit quaternion in f64 (the f32 one is unit only to 1e-7)
            from ncollide_b200.scenes import transform_rays

            rs.origins, rs.dirs = transform_rays(pose, rs.origins, rs.dirs)
        rng = np.random.default_rng(3)
        limits = np.where(rng.random(len(rs.origins)) < 0.4, rng.uniform(1.0, 8.0, len(rs.origins)), 1e30)
        toi, face, normal, _ = om.ray_cast_uv(rs.origins, rs.dirs, max_toi=limits, pose=pose, mode=0)
        V = rs.verts.astype(np.float64)
        if pose is not None:  # world-space triangles; q * v evaluated like nalgebra does (the f32 quaternion is a unit one only to 1e-7)
            t, qv, qw = pose[:3].astype(np.float64), pose[3:6].astype(np.float64), float(pose[6])
            tt2 = 2 * np.cross(qv, V)
            V = V + qw * tt2 + np.cross(qv, tt2) + t
        A, B, Cc = V[rs.tris[:, 0]], V[rs.tris[:, 1]], V[rs.tris[:, 2]]
        e1, e2 = B - A, Cc - A
        nrm = np.cross(e1, e2)
        T = len(rs.tris)
        for r in range(len(rs.origins)):
            o, d = rs.origins[r].astype(np.float64), rs.dirs[r].astype(np.float64)
            pv = np.cross(d, e2)
            det = np.einsum("ij,ij->i", e1, pv)
            ok = np.abs(det) > 1e-12
            inv = 1.0 / np.where(ok, det, 1.0)
            tv = o - A
            u = np.einsum("ij,ij->i", tv, pv) * inv
            qv = np.cross(tv, e1)
            v = (qv @ d) * inv
            tt = np.einsum("ij,ij->i", e2, qv) * inv
            inside = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (tt >= 0)
            near_edge = ok & (tt >= 0) & (np.minimum(np.minimum(np.abs(u), np.abs(v)), np.abs(1 - u - v)) < 1e-7) & (u > -1e-6) & (v > -1e-6) & (u + v < 1 + 1e-6)
            if near_edge.any():
                continue
            if not inside.any():
                assert toi[r] < 0, (kind, r)
                continue
            k = int(np.argmin(np.where(inside, tt, np.inf)))
            if abs(tt[k] - limits[r]) < 1e-6 * max(1.0, limits[r]):
                continue
            if tt[k] > limits[r]:
                assert toi[r] < 0, (kind, r, tt[k], limits[r])
                cut += 1
                continue
            tol = 1e-9 if pose is None else 1e-7  # posed: the oracle works in the mesh's frame, numpy in the world's
            assert abs(toi[r] - tt[k]) < tol * max(1.0, tt[k]), (kind, r, toi[r], tt[k])
            is_back = nrm[k] @ d > 0
            assert face[r] == k + (T if is_back else 0), (kind, r, face[r], k, is_back)
            want_n = nrm[k] / np.linalg.norm(nrm[k]) * (-1.0 if is_back else 1.0)
            assert np.allclose(normal[r], want_n, atol=tol * 10), (kind, r)
            checked += 1
            back += bool(is_back)
    assert checked > 500 and back > 50 and cut > 20, (checked, back, cut)


def test_golden_fixtures_match_the_oracle(oracle):
    """tests/golden/*.npz were produced by tests/golden/make_golden.py from the oracle at commit time; they pin the
    oracle (and, under -m gpu, the device) against silent drift."""
    import glob

    from tests.golden.make_golden import scene_from_npz

    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "world_*.npz")))
    assert files
    for f in files:
        z = np.load(f)
        s = scene_from_npz(z)
        fat = oracle.compute_aabbs(s)
        assert np.array_equal(fat, z["fat_aabbs"])
        pairs = oracle.broad_phase(fat, s.groups, 0)
        assert np.array_equal(pairs, z["pairs"])
        c, off, algo, _ = oracle.narrow_phase(s, pairs)
        assert np.array_equal(off, z["manifold_off"]) and np.array_equal(algo, z["algo"])
        for name in ("world1", "world2", "normal", "depth", "f1", "f2"):
            assert np.array_equal(c[name], z["c_" + name]), (f, name)
    for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "rays_*.npz"))):
        z = np.load(f)
        om = oracle.trimesh(z["verts"], z["tris"])
        t, face, n = om.ray_cast(z["origins"], z["dirs"], mode=0)
        assert np.array_equal(t, z["toi"]) and np.array_equal(face, z["face"]) and np.array_equal(n, z["normal"])


# ---- bench.py contract (the reference arm runs on the CPU, so its JSON line can be checked here) ----------------------------------
def test_bench_reference_arm_json_contract():
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--n-objects", "3000"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly ONE JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    for k in ("metric", "value", "unit", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0
    assert d["same_config"] is True and d["config"]["n_objects_total"] == 3000
    rc = d["rays"]["cpu_baseline"]  # the Mrays/s half of the metric has its CPU baseline in both arms
    assert rc["kind"] == "port" and rc["cores"] == 1 and rc["unit"] == "Mrays/s" and rc["value"] > 0 and 0 < rc["hit_fraction"] <= 1


def test_bench_config_is_the_same_object_in_both_arms():
    sys_path = sys.path[:]
    try:
        sys.path.insert(0, ROOT)
        import bench
    finally:
        sys.path[:] = sys_path
    a, b = bench.workload_config(1_000_000, 1), bench.workload_config(1_000_000, 1)
    assert a == b and a["n_objects_total"] == 1_000_000
    assert bench.workload_config(1_000_000, 8)["n_objects_total"] == 8_000_000
    assert len(bench.source_hash()) == 16


def test_bench_native_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: the native arm must fail loudly (non-zero exit, no JSON metric line) when no CUDA device is usable."""
    import subprocess
    import sys

    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--n-objects", "2000", "--no-rays", "--no-cpu",
                        "--no-extras"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert not any(ln.startswith("{") and '"value"' in ln for ln in r.stdout.splitlines())
