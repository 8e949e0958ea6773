"""GPU parity tests: the CUDA path (through the C ABI, include/ncb200.h) against the CPU oracle on identical seeded
inputs.  Bit-exact for AABBs, pair sets, feature ids, manifold sizes and ray-hit faces; contacts / TOI within the
north-star tolerance (1e-4 relative / 1e-5 absolute).  Run with `pytest -m gpu` on a B200."""
import numpy as np
import pytest

from ncollide_b200.scenes import DEFAULT_GROUPS, config_scene, make_ray_scene, make_world_scene
from ncollide_b200.shapes import BALL, CUBOID, HULL, PLANE, ConvexHull, HullLibrary

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-5  # BASELINE.json north_star tolerance for contact points / normals / depths / TOI


@pytest.fixture(scope="module")
def ctx():
    from ncollide_b200.world import Context

    c = Context(0)
    yield c
    c.close()


def canon(pairs):
    p = np.sort(np.asarray(pairs, dtype=np.uint32).reshape(-1, 2), axis=1)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def close(a, b):
    return np.allclose(a, b, rtol=RTOL, atol=ATOL)


SCENES = [
    lambda: config_scene(1),
    lambda: config_scene(2, 4000),
    lambda: config_scene(3, 6000),
    lambda: make_world_scene(3000, 77, (1, 1, 1), side=9.0, angular=0.05, n_hulls=64, name="dense_angular"),
    lambda: make_world_scene(2000, 78, (0, 1, 1), side=7.0, n_hulls=32, linear=0.05, margin=0.01, name="dense_convex"),
]


@pytest.mark.parametrize("mk", SCENES)
def test_aabbs_bit_exact(ctx, oracle, mk):
    s = mk()
    ctx.set_scene(s)
    for mode in (0, 1, 2):
        got = ctx.compute_aabbs(s.margin, mode)
        want = oracle.compute_aabbs(s, mode=mode)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"{s.name} mode {mode}"


@pytest.mark.parametrize("mk", SCENES)
def test_broad_phase_pair_set_bit_exact(ctx, oracle, mk):
    s = mk()
    fat = oracle.compute_aabbs(s)
    got = ctx.broad_phase(fat, s.groups)
    want = oracle.broad_phase(fat, s.groups, mode=1)
    assert np.all(got[:, 0] > got[:, 1]), "pairs must be (larger handle, smaller handle)"
    assert np.array_equal(canon(got), canon(want)), s.name
    assert len(np.unique(canon(got), axis=0)) == len(got), "duplicate pairs"
    want_dbvt = oracle.broad_phase(fat, s.groups, mode=0)
    assert np.array_equal(canon(got), canon(want_dbvt))


def compare_manifolds(res, s, oracle, label):
    """res: UpdateResult for pairs res.pairs; oracle narrow phase is run on the same pairs in the same order."""
    if getattr(s, "query_kind", None) is not None:  # world with proximity sensors (tests/test_proximity.py)
        oc, ooff, oalgo, oprox = oracle.narrow_phase_kinds(s, res.pairs)
        assert res.proximity is not None and np.array_equal(res.proximity, oprox), f"{label}: proximity statuses differ"
    else:
        oc, ooff, oalgo, ostats = oracle.narrow_phase(s, res.pairs)
    assert np.array_equal(res.pair_algo, oalgo), f"{label}: dispatched algorithm differs"
    ocount = np.diff(ooff)
    bad = np.nonzero(res.manifold_count != ocount)[0]
    assert len(bad) == 0, f"{label}: manifold sizes differ on {len(bad)} pairs, first {bad[:5]} algo {oalgo[bad[:5]]}"
    # gather the device contacts in oracle order
    idx = np.concatenate([np.arange(st, st + c) for st, c in zip(res.manifold_start, res.manifold_count)]) if len(oc) else np.zeros(0, int)
    dc = res.contacts[idx.astype(np.int64)]
    assert len(dc) == len(oc)
    if len(oc) == 0:
        return
    assert np.array_equal(dc["f1"], oc["f1"]) and np.array_equal(dc["f2"], oc["f2"]), f"{label}: feature ids differ"
    pair_of = np.repeat(np.arange(len(res.pairs)), res.manifold_count)
    assert np.array_equal(dc["pair"], pair_of), f"{label}: contact.pair back-references differ"
    for name in ("world1", "world2", "normal", "depth"):
        ok = np.isclose(dc[name], oc[name], rtol=RTOL, atol=ATOL)
        if not ok.all():
            w = np.nonzero(~ok.reshape(len(oc), -1).all(axis=1))[0]
            raise AssertionError(f"{label}: {name} differs on {len(w)} contacts, e.g. pair {pair_of[w[0]]} algo {oalgo[pair_of[w[0]]]}: {dc[name][w[0]]} vs {oc[name][w[0]]}")
    exact = sum(np.array_equal(dc[n].view(np.uint32), oc[n].view(np.uint32)) for n in ("world1", "world2", "normal", "depth"))
    return exact == 4


@pytest.mark.parametrize("mk", SCENES)
def test_generate_contacts_parity(ctx, oracle, mk):
    s = mk()
    ctx.set_scene(s)
    fat = oracle.compute_aabbs(s)
    pairs = oracle.broad_phase(fat, s.groups, mode=0)  # reference callback order and orientation
    res = ctx.generate_contacts(pairs)
    compare_manifolds(res, s, oracle, s.name)


@pytest.mark.parametrize("mk", SCENES)
def test_world_update_parity(ctx, oracle, mk):
    s = mk()
    ctx.set_hulls(s.hulls)
    res = ctx.world_update(s)
    assert res.counts["epa_overflow"] == 0 and res.counts["stack_overflow"] == 0
    fat = oracle.compute_aabbs(s)
    want = oracle.broad_phase(fat, s.groups, mode=1)
    assert np.array_equal(canon(res.pairs), canon(want)), "pair set"
    assert np.all(res.pairs[:, 0] > res.pairs[:, 1])
    compare_manifolds(res, s, oracle, s.name)
    assert res.counts["n_contact_pairs"] == int((res.manifold_count > 0).sum())
    assert sum(res.counts["n_algo"].values()) == len(res.pairs)


def test_world_update_poses_equals_world_update(ctx, oracle):
    """ncb_world_update_poses (objects persist, only the poses travel) returns what ncb_world_update returns for the same world."""
    s = config_scene(3, 5000)
    ctx.set_hulls(s.hulls)
    a = ctx.world_update(s)
    rng = np.random.default_rng(5)
    s2 = config_scene(3, 5000)
    s2.pos = (s.pos + rng.normal(0, 0.05, size=s.pos.shape)).astype(np.float32)
    b = ctx.world_update_poses(s2.pos, s2.rot, s2.margin)
    want = ctx.world_update(s2)
    # pair order within a key segment depends on the order of atomic slot allocations: compare by pair
    def by_pair(r):
        order = np.lexsort((r.pairs[:, 1], r.pairs[:, 0]))
        return r.pairs[order], r.pair_algo[order], r.manifold_count[order]

    for x, y in zip(by_pair(b), by_pair(want)):
        assert np.array_equal(x, y)
    assert len(b.contacts) == len(want.contacts) and b.counts["n_contact_pairs"] == want.counts["n_contact_pairs"]
    assert not np.array_equal(canon(a.pairs), canon(b.pairs))
    compare_manifolds(b, s2, oracle, "poses")


@pytest.mark.parametrize("mk", [SCENES[1], SCENES[3], SCENES[4]])
def test_contact_kinematics_match_oracle(ctx, oracle, mk):
    """ncb_set_kinematics: every contact of a fresh-world update comes with its ContactKinematic (local1 / local2, NeighborhoodGeometry
    kind + direction per side, dilations); compared with the oracle's restatement of each generator's kinematic.  The contacts
    themselves must not change when kinematics are requested."""
    s = mk()
    ctx.set_hulls(s.hulls)
    plain = ctx.world_update(s)
    ctx.set_kinematics(True)
    try:
        res = ctx.world_update(s)
        kin = ctx.fetch_kinematics(len(res.contacts))
    finally:
        ctx.set_kinematics(False)
    assert len(res.contacts) == len(plain.contacts) and res.counts["n_contact_pairs"] == plain.counts["n_contact_pairs"]
    oc, ok, ooff, oalgo = oracle.narrow_phase_kinematics(s, res.pairs)
    assert np.array_equal(res.manifold_count, np.diff(ooff)) and np.array_equal(res.pair_algo, oalgo)
    idx = np.concatenate([np.arange(a, a + c) for a, c in zip(res.manifold_start, res.manifold_count)]).astype(np.int64)
    dk, dc = kin[idx], res.contacts[idx]
    assert np.array_equal(dc["f1"], oc["f1"]) and np.array_equal(dc["f2"], oc["f2"])
    assert np.array_equal(dk["g1"], ok["g1"]) and np.array_equal(dk["g2"], ok["g2"]), "NeighborhoodGeometry kinds"
    for name in ("local1", "local2", "dir1", "dir2", "dil1", "dil2"):
        assert np.allclose(dk[name], ok[name], rtol=RTOL, atol=ATOL), name
    assert (ok["g1"] == 1).any() and (ok["g1"] == 2).any()
    with pytest.raises(Exception, match="ncb_set_kinematics"):
        ctx.world_update(s)
        ctx.fetch_kinematics(4)


def test_deep_tree_coincident_boxes(ctx, oracle):
    """120 000 boxes of which 100 000 share a few dozen distinct positions (identical Morton codes: the LBVH splits them on the tie-break
    bits, its deepest shape) and 20 000 sit within a few ulps of them.  The pair search must neither lose a pair nor run out of its
    traversal stack; groups keep the pair count finite (every object only collides with its own small group)."""
    rng = np.random.default_rng(99)
    n, n_sites = 120_000, 40
    sites = rng.uniform(0, 50, size=(n_sites, 3)).astype(np.float32)
    which = rng.integers(0, n_sites, size=n)
    lo = sites[which].copy()
    near = np.arange(n) >= 100_000
    lo[near] += (rng.integers(1, 5, size=(int(near.sum()), 3)) * np.spacing(lo[near])).astype(np.float32)  # 1..4 ulps away
    fat = np.concatenate([lo, lo + np.float32(0.25)], axis=1).astype(np.float32)
    # 30 collision groups, membership = whitelist = one group
    g = (np.arange(n) % 30).astype(np.uint32)
    groups = np.stack([1 << g, 1 << g, np.zeros(n, np.uint32)], axis=1).astype(np.uint32)
    # thin out: only every 8th object keeps its group, the others get a private "no collision" whitelist
    lonely = (np.arange(n) // 30) % 8 != 0
    groups[lonely, 1] = 0
    got = ctx.broad_phase(fat, groups)
    want = oracle.broad_phase(fat, groups, mode=1)
    assert len(want) > 20_000
    assert np.array_equal(canon(got), canon(want))
    assert ctx.traversal_overflows() == 0, "a BVH walk ran out of its traversal stack"


def _adversarial_scenes():
    """Configurations that stress degenerate branches: exact coincidence, axis-aligned face contact, touching at depth 0,
    far-from-origin coordinates, a dense clump, tiny / huge shapes side by side."""
    F = np.float32
    out = []
    # 1. everything at one point (all pairs; GJK starts from a zero direction, EPA from degenerate simplices)
    s = make_world_scene(60, 301, (1, 1, 1), side=1.0, n_hulls=8, name="coincident")
    s.pos[:] = F(0.25)
    out.append(s)
    # 2. an axis-aligned lattice of unit cubes and balls exactly touching (depth 0, parallel faces, shared edges / corners)
    s = make_world_scene(343, 302, (1, 1, 0), side=1.0, name="lattice")
    g = np.stack(np.meshgrid(*[np.arange(7)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(F)
    s.pos[:] = g
    s.rot[:] = (0, 0, 0, 1)
    s.shape_param[:, :3] = F(0.5)
    out.append(s)
    # 3. the same lattice slightly compressed (everything interpenetrates by the same amount)
    s2 = make_world_scene(343, 302, (1, 1, 0), side=1.0, name="lattice_compressed")
    s2.pos[:] = g * F(0.9)
    s2.rot[:] = (0, 0, 0, 1)
    s2.shape_param[:, :3] = F(0.5)
    out.append(s2)
    # 4. far from the origin (cancellation in every difference)
    s = make_world_scene(1500, 303, (1, 1, 1), side=6.0, n_hulls=16, name="far_away")
    s.pos[:] = (s.pos + np.array([4096.0, -2048.0, 8192.0], dtype=F)).astype(F)
    out.append(s)
    # 5. a dense clump: ~40 neighbours per object, deep penetrations, many EPA expansions
    out.append(make_world_scene(1200, 304, (1, 1, 1), side=2.2, n_hulls=32, name="clump"))
    # 6. mixed scales: 1e-2 .. 5 (prediction and the 0.02 manifold threshold are absolute)
    s = make_world_scene(1500, 305, (1, 1, 0), side=8.0, name="scales")
    k = np.random.default_rng(305).choice([0.02, 0.2, 1.0, 5.0], size=s.n).astype(F)
    s.shape_param[:, :3] = (s.shape_param[:, :3] * k[:, None]).astype(F)
    out.append(s)
    return out


@pytest.mark.parametrize("k", range(6))
def test_adversarial_scenes_parity(ctx, oracle, k):
    s = _adversarial_scenes()[k]
    ctx.set_hulls(s.hulls)
    res = ctx.world_update(s)
    assert res.counts["epa_overflow"] == 0, s.name
    want = oracle.broad_phase(oracle.compute_aabbs(s), s.groups, mode=1)
    assert np.array_equal(canon(res.pairs), canon(want)), s.name
    compare_manifolds(res, s, oracle, s.name)


def _prism(k, r=1.0, h=0.25):
    a = np.arange(k) * 2 * np.pi / k
    ring = np.stack([r * np.cos(a), r * np.sin(a)], 1)
    pts = np.concatenate([np.c_[ring, np.full(k, -h)], np.c_[ring, np.full(k, h)]]).astype(np.float32)
    return ConvexHull.try_from_points(pts)


@pytest.mark.parametrize("k", [8, 12, 16])
def test_large_faces_and_many_contacts(ctx, oracle, k):
    """Capacity edges: k-gon prisms stacked face to face (16-vertex faces = the device maximum; up to 2k + k^2 clip candidates,
    16-19 distinct contacts per manifold), also against a cuboid and a ball; and a stepping world on the same bodies, whose
    manifold cache holds the previous contacts as well."""
    from test_oracle_kat import scene_of

    lib = HullLibrary([_prism(k)])
    q = (0, 0, float(np.sin(0.1)), float(np.cos(0.1)))
    s = scene_of([(HULL, [0], (0, 0, 0)), (HULL, [0], (0.05, 0.02, 0.45), q), (CUBOID, [0.8, 0.8, 0.2], (0.1, 0, -0.4), q), (BALL, [0.5], (0.2, 0.1, 1.1))],
                 hulls=lib)
    ctx.set_hulls(s.hulls)
    res = ctx.world_update(s)
    assert res.counts["epa_overflow"] == 0
    assert len(res.pairs) >= 3 and res.manifold_count.max() >= 12
    compare_manifolds(res, s, oracle, f"prism{k}")
    # stepping: small motions keep the stacks in contact; overflow of the persistent cache would be counted
    from ncollide_b200.world import SteppingWorld

    dev, orc = SteppingWorld(ctx, s), oracle.sim(s)
    pos, rot = s.pos.copy(), s.rot.copy()
    for t in range(4):
        if t:
            pos[1] += np.array([0.004, -0.003, 0.001], dtype=np.float32)
            for w in (dev, orc):
                w.set_positions([1], pos[1:2], rot[1:2])
        a, b = dev.step(), orc.step()
        assert a["counts"]["epa_overflow"] == 0, "persistent manifold cache overflow"
        keep = a["algo"] != 0
        assert np.array_equal(a["pairs"][keep], b["pairs"]) and np.array_equal(a["count"][keep], np.diff(b["off"]))
        assert np.array_equal(a["ids"], b["ids"])
        for f in ("world1", "world2", "normal", "depth"):
            assert np.allclose(a["contacts"][f], b["contacts"][f], rtol=RTOL, atol=ATOL)


def test_world_update_is_deterministic_as_a_set(ctx):
    s = config_scene(3, 5000)
    ctx.set_hulls(s.hulls)
    a = ctx.world_update(s)
    b = ctx.world_update(s)
    assert np.array_equal(canon(a.pairs), canon(b.pairs))

    def key(r):
        out = {}
        for i, p in enumerate(map(tuple, r.pairs.tolist())):
            c = r.contacts_of(i)
            out[p] = b"".join(np.ascontiguousarray(c[n]).tobytes() for n in ("world1", "world2", "normal", "depth", "f1", "f2"))
        return out

    assert key(a) == key(b)


def mini_scene(objs, linear=0.02, angular=0.0, margin=0.02, hulls=None, groups=None):
    from ncollide_b200.scenes import WorldScene

    n = len(objs)
    return WorldScene(
        pos=np.array([o[2] for o in objs], dtype=np.float32).reshape(n, 3),
        rot=np.array([o[3] if len(o) > 3 else (0, 0, 0, 1) for o in objs], dtype=np.float32).reshape(n, 4),
        shape_type=np.array([o[0] for o in objs], dtype=np.uint32),
        shape_param=np.array([list(o[1]) + [0] * (4 - len(o[1])) for o in objs], dtype=np.float32).reshape(n, 4),
        groups=np.asarray(groups, dtype=np.uint32) if groups is not None else np.tile(np.array(DEFAULT_GROUPS, dtype=np.uint32), (n, 1)),
        query_limit=np.full(n, linear, dtype=np.float32),
        ang_pred=np.full(n, angular, dtype=np.float32),
        hulls=hulls or HullLibrary([]),
        margin=margin,
    )


def test_edge_cases_empty_single_and_coincident(ctx, oracle):
    empty = mini_scene([])
    ctx.set_hulls(empty.hulls)
    r = ctx.world_update(empty)
    assert len(r.pairs) == 0 and len(r.contacts) == 0
    one = mini_scene([(BALL, [0.5], (0, 0, 0))])
    r = ctx.world_update(one)
    assert len(r.pairs) == 0
    # many coincident objects: identical Morton codes, every pair overlaps (ties broken by position in the LBVH)
    k = 40
    same = mini_scene([(BALL, [0.5], (1, 2, 3))] * k)
    r = ctx.world_update(same)
    assert len(r.pairs) == k * (k - 1) // 2
    compare_manifolds(r, same, oracle, "coincident")
    # the reference's dbvt_broad_phase3d example: 4 balls -> 6 pairs; without two of them -> 1
    ex = mini_scene([(BALL, [0.5], p) for p in [(0, 0, 0), (0, 0.5, 0), (0.5, 0, 0), (0.5, 0.5, 0)]], linear=0.0, margin=0.2)
    tight = oracle.compute_aabbs(ex, fat=False)
    assert len(ctx.broad_phase(tight)) == 6
    assert len(ctx.broad_phase(tight[2:])) == 1


def test_planes_and_groups(ctx, oracle):
    rng = np.random.default_rng(5)
    objs = [(PLANE, [0, 1, 0], (0, 0, 0)), (PLANE, [1, 0, 0], (0, 0, 0))]
    for i in range(300):
        p = tuple(rng.uniform(-0.5, 3, size=3))
        if i % 2:
            objs.append((BALL, [0.4], p))
        else:
            q = rng.standard_normal(4)
            objs.append((CUBOID, list(rng.uniform(0.2, 0.5, 3)), p, tuple(q / np.linalg.norm(q))))
    n = len(objs)
    groups = np.tile(np.array(DEFAULT_GROUPS, dtype=np.uint32), (n, 1))
    # objects 10..60 are members of group 3 only and blacklist group 5; 60..120 are members of group 5
    groups[10:60] = (1 << 3, 0x3FFFFFFF, 1 << 5)
    groups[60:120] = (1 << 5, 0x3FFFFFFF, 0)
    s = mini_scene(objs, groups=groups)
    ctx.set_hulls(s.hulls)
    r = ctx.world_update(s)
    fat = oracle.compute_aabbs(s)
    want = oracle.broad_phase(fat, s.groups, mode=2)
    assert np.array_equal(canon(r.pairs), canon(want))
    assert (1, 0) in set(map(tuple, r.pairs.tolist())), "plane x plane stays a broad-phase pair"
    compare_manifolds(r, s, oracle, "planes+groups")


def test_reference_kats_on_device(ctx, oracle):
    # tests/geometry/contact.rs:8-26 (issue #182): just-touching cuboids must give finite results
    s = mini_scene([(CUBOID, [0.5, 0.5, 0.1], (0, 0, 0)), (CUBOID, [0.5, 0.5, 0.1], (0, 1, 0))], linear=0.0, margin=0.02)
    ctx.set_hulls(s.hulls)
    r = ctx.world_update(s)
    assert len(r.pairs) == 1
    for name in ("world1", "world2", "normal", "depth"):
        assert np.all(np.isfinite(r.contacts[name]))
    compare_manifolds(r, s, oracle, "issue182")
    # tests/geometry/epa3.rs:7-22 first case (f32): depth 0.5 along -x
    s = mini_scene([(CUBOID, [2, 1, 1], (0, 0, 0)), (CUBOID, [2, 1, 1], (3.5, 0, 0))], linear=10.0, margin=0.0)
    r = ctx.world_update(s)
    assert len(r.pairs) == 1 and r.manifold_count[0] >= 1
    c = r.contacts_of(0)
    assert np.max(c["depth"]) == np.float32(0.5) and tuple(c["normal"][0]) == (-1.0, 0.0, 0.0)
    compare_manifolds(r, s, oracle, "epa3")
    # examples/contact_query3d.rs:8-35 through the world (ball r=1 vs unit cuboid)
    for p, sign in (((1, 1, 1), 1), ((2, 2, 2), -1), ((3, 3, 3), 0)):
        s = mini_scene([(CUBOID, [1, 1, 1], (0, 0, 0)), (BALL, [1.0], p)], linear=0.5, margin=2.0)
        r = ctx.world_update(s)
        compare_manifolds(r, s, oracle, "contact_query3d")
        if sign == 0:
            assert len(r.contacts) == 0
        else:
            assert np.sign(r.contacts["depth"][0]) == sign


def test_hull_library_kats(ctx, oracle):
    cube = ConvexHull.try_from_points(np.array([[x, y, z] for x in (-0.5, 0.5) for y in (-0.5, 0.5) for z in (-0.5, 0.5)], dtype=np.float32))
    octa = ConvexHull.try_new([(0, 0, 1), (0, 0, -1), (0, 1, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0)],
                              [0, 4, 2, 0, 3, 4, 5, 0, 2, 5, 3, 0, 1, 5, 2, 1, 3, 5, 4, 1, 2, 4, 3, 1])
    lib = HullLibrary([cube, octa])
    rng = np.random.default_rng(9)
    objs = []
    for i in range(400):
        q = rng.standard_normal(4)
        q = tuple(q / np.linalg.norm(q)) if i % 3 else (0, 0, 0, 1)
        objs.append((HULL, [i % 2], tuple(rng.uniform(0, 4, 3)), q))
    objs.append((PLANE, [0, 0, 1], (0, 0, 0.2)))
    s = mini_scene(objs, hulls=lib, angular=0.02)
    ctx.set_hulls(lib)
    r = ctx.world_update(s)
    assert r.counts["epa_overflow"] == 0
    compare_manifolds(r, s, oracle, "hull kats")


# ---- ray casting -----------------------------------------------------------------------------------------------
def ulp_diff(a, b):
    ia = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    ib = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def check_rays(ctx, oracle, rs, brute, pose=None):
    mesh = ctx.trimesh(rs.verts, rs.tris)
    om = oracle.trimesh(rs.verts, rs.tris)
    toi, face, normal = mesh.toi_and_normal_with_ray(pose, rs.origins, rs.dirs)
    T = len(rs.tris)
    # reference-faithful BVT best-first answer
    rtoi, rface, rnormal = om.ray_cast(rs.origins, rs.dirs, pose=pose, mode=0)
    refs = [("bvt", rtoi, rface, rnormal)]
    if brute:
        btoi, bface, bnormal = om.ray_cast(rs.origins, rs.dirs, pose=pose, mode=1)
        # device semantics == brute-force definition, bit for bit
        assert np.array_equal(face, bface), f"{rs.name}: {(face != bface).sum()} faces differ from the brute-force definition"
        assert np.array_equal(toi.view(np.uint32), btoi.view(np.uint32))
        hit = btoi >= 0
        assert close(normal[hit], bnormal[hit])
        refs.append(("brute", btoi, bface, bnormal))
    # against the reference-faithful traversal: identical outside the tie class
    diff = np.nonzero(face != rface)[0]
    for i in diff:
        assert toi[i] >= 0 and rtoi[i] >= 0, f"{rs.name}: ray {i} hit/miss disagreement"
        assert ulp_diff(toi[i], rtoi[i]) <= 4, f"{rs.name}: ray {i} face {face[i]} vs {rface[i]} toi {toi[i]} vs {rtoi[i]}"
    same = face == rface
    assert close(toi[same], rtoi[same])
    hit = same & (rtoi >= 0)
    assert close(normal[hit], rnormal[hit])
    assert (face[toi >= 0] % T < T).all()
    assert ctx.traversal_overflows() == 0, "a ray walk ran out of its stack"
    mesh.close()
    return len(diff), int((toi >= 0).sum())


@pytest.mark.parametrize("kind", ["terrain", "soup"])
def test_ray_cast_small_vs_brute_force(ctx, oracle, kind):
    rs = make_ray_scene(kind, 5000, 3000, seed=11)
    nties, nhits = check_rays(ctx, oracle, rs, brute=True)
    assert nhits > 100


@pytest.mark.parametrize("kind", ["terrain", "soup"])
def test_ray_cast_medium_vs_bvt(ctx, oracle, kind):
    rs = make_ray_scene(kind, 200_000, 100_000, seed=12)
    nties, nhits = check_rays(ctx, oracle, rs, brute=False)
    assert nhits > 1000
    assert nties <= 10


def test_ray_cast_with_pose_and_max_toi(ctx, oracle):
    rs = make_ray_scene("terrain", 20_000, 5000, seed=13, random_pose=True)
    # move the rays into world space so that they still hit the posed mesh
    t, q = rs.pose[:3].astype(np.float64), rs.pose[3:].astype(np.float64)

    def rot(v):
        qv = q[:3]
        tt = 2 * np.cross(qv, v)
        return v + q[3] * tt + np.cross(qv, tt)

    rs.origins = (rot(rs.origins.astype(np.float64)) + t).astype(np.float32)
    rs.dirs = rot(rs.dirs.astype(np.float64)).astype(np.float32)
    check_rays(ctx, oracle, rs, brute=True, pose=rs.pose)
    mesh = ctx.trimesh(rs.verts, rs.tris)
    om = oracle.trimesh(rs.verts, rs.tris)
    toi, face, _ = mesh.toi_and_normal_with_ray(rs.pose, rs.origins, rs.dirs, max_toi=4.0)
    btoi, bface, _ = om.ray_cast(rs.origins, rs.dirs, max_toi=4.0, pose=rs.pose, mode=1)
    assert np.array_equal(face, bface) and np.array_equal(toi.view(np.uint32), btoi.view(np.uint32))
    assert (toi[toi >= 0] <= 4.0).all() and (toi < 0).any()


def test_ray_cast_edge_cases(ctx, oracle):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float32)
    one = ctx.trimesh(v, np.array([[0, 1, 2]], dtype=np.uint32))
    toi, face, n = one.toi_and_normal_with_ray(None, [[0.2, 0.2, 1]], [[0, 0, -1]])
    assert toi[0] == 1.0 and face[0] == 0 and tuple(n[0]) == (0, 0, 1)
    toi, face, n = one.toi_and_normal_with_ray(None, [[0.2, 0.2, -1]], [[0, 0, 1]])
    assert toi[0] == 1.0 and face[0] == 1 and tuple(n[0]) == (0, 0, -1)  # back face -> i + n_tris
    toi, face, n = one.toi_and_normal_with_ray(None, [[2, 2, 1]], [[0, 0, -1]])
    assert toi[0] < 0
    toi, face, n = one.toi_and_normal_with_ray(None, np.zeros((0, 3)), np.zeros((0, 3)))
    assert len(toi) == 0
    # duplicated triangles: ties go to the smallest face index
    two = ctx.trimesh(v, np.array([[0, 1, 2], [0, 1, 2], [0, 1, 3]], dtype=np.uint32))
    toi, face, n = two.toi_and_normal_with_ray(None, [[0.2, 0.2, 1]], [[0, 0, -1]])
    assert face[0] == 0
    om = oracle.trimesh(v, np.array([[0, 1, 2], [0, 1, 2], [0, 1, 3]], dtype=np.uint32))
    bt, bf, bn = om.ray_cast(np.array([[0.2, 0.2, 1]], dtype=np.float32), np.array([[0, 0, -1]], dtype=np.float32), mode=1)
    assert bf[0] == face[0]


def test_query_slices_partition_the_pair_set(ctx, oracle):
    """Multi-GPU sharding rule on one device: every Morton-order query slice emits its own pairs, the union is the
    full pair set, no pair is emitted twice, and per-pair manifolds do not depend on the slicing."""
    import ctypes as C

    from ncollide_b200 import _ffi

    s = config_scene(3, 7001)
    ctx.set_scene(s)
    full = ctx.world_fetch(ctx.world_update_device(s.margin))
    n = s.n
    cuts = [0, 1500, 1501, 4000, n]
    got_pairs = []
    per_pair = {}
    for b, e in zip(cuts, cuts[1:]):
        c = _ffi.UpdateCountsC()
        ctx.check(ctx.lib.ncb_world_update_stage(ctx.h, 0, C.c_float(s.margin), C.c_uint32(0), C.c_uint32(n), None), "stage 0")
        ctx.check(ctx.lib.ncb_world_update_stage(ctx.h, 1, C.c_float(s.margin), C.c_uint32(b), C.c_uint32(e), C.byref(c)), "stage 1")
        r = ctx.world_fetch(ctx._counts(c))
        got_pairs.append(r.pairs.copy())
        for i, p in enumerate(map(tuple, r.pairs.tolist())):
            per_pair[p] = r.contacts_of(i)[["world1", "world2", "normal", "depth", "f1", "f2"]].copy()
    allp = np.concatenate(got_pairs)
    assert len(allp) == len(full.pairs), "a pair was emitted by two slices or by none"
    assert np.array_equal(canon(allp), canon(full.pairs))
    for i, p in enumerate(map(tuple, full.pairs.tolist())):
        a = full.contacts_of(i)[["world1", "world2", "normal", "depth", "f1", "f2"]]
        b = per_pair[p]
        assert len(a) == len(b)
        for name in ("world1", "world2", "normal", "depth", "f1", "f2"):
            assert np.array_equal(a[name], b[name])


@pytest.mark.parametrize("world,mk", [(2, lambda: config_scene(3, 7001)), (4, lambda: config_scene(2, 9000)), (8, lambda: config_scene(5, 30000)),
                                      (3, lambda: make_world_scene(5000, 91, (1, 1, 1), side=4.0, n_hulls=16, name="tiny_dense"))])
def test_spatial_shards_partition_the_pair_set(ctx, world, mk):
    """Multi-GPU spatial sharding on one device: rank r of `world` selects owned + ghost objects, builds its own LBVH and
    reports its share; the union over the ranks is the full pair set, no pair twice, manifolds identical."""
    s = mk()
    ctx.set_scene(s)
    full = ctx.world_fetch(ctx.world_update_device(s.margin))
    per_pair, total, owned = {}, 0, []
    for rank in range(world):
        r = ctx.world_fetch(ctx.world_update_sharded(s.margin, rank, world))
        assert np.all(r.pairs[:, 0] > r.pairs[:, 1])
        total += len(r.pairs)
        for i, p in enumerate(map(tuple, r.pairs.tolist())):
            assert p not in per_pair, "pair reported by two ranks"
            per_pair[p] = r.contacts_of(i)[["world1", "world2", "normal", "depth", "f1", "f2"]].copy()
    assert total == len(full.pairs), "a pair was reported by two ranks or by none"
    assert np.array_equal(canon(np.array(list(per_pair), dtype=np.uint32)), canon(full.pairs))
    for i, p in enumerate(map(tuple, full.pairs.tolist())):
        a = full.contacts_of(i)[["world1", "world2", "normal", "depth", "f1", "f2"]]
        b = per_pair[p]
        assert len(a) == len(b)
        for name in ("world1", "world2", "normal", "depth", "f1", "f2"):
            assert np.array_equal(a[name], b[name])


def test_early_fetch_with_staged_updates(ctx):
    """ncb_world_fetch_early + a staged / sharded device update + ncb_world_fetch gives the same bytes as a plain fetch."""
    import ctypes as C

    from ncollide_b200 import _ffi

    s = config_scene(3, 30011)
    ctx.set_scene(s)
    lib, h = ctx.lib, ctx.h
    ref = ctx.world_fetch(ctx.world_update_device(s.margin))
    for mode in ("stage", "sharded"):
        bufs = ctx.alloc_result_buffers(len(ref.pairs) + 64, len(ref.contacts) + 64)
        ctx.check(lib.ncb_world_fetch_early(h, _ffi.ptr(bufs["pairs"]), C.c_uint32(len(bufs["pairs"])), _ffi.ptr(bufs["algo"]),
                                            _ffi.ptr(bufs["contacts"]), C.c_uint32(len(bufs["contacts"]))), "fetch_early")
        c = _ffi.UpdateCountsC()
        ctx.check(lib.ncb_world_update_stage(h, 0, C.c_float(s.margin), C.c_uint32(0), C.c_uint32(s.n), None), "stage 0")
        if mode == "stage":
            ctx.check(lib.ncb_world_update_stage(h, 1, C.c_float(s.margin), C.c_uint32(0), C.c_uint32(0xFFFFFFFF), C.byref(c)), "stage 1")
        else:
            ctx.check(lib.ncb_world_update_sharded(h, C.c_float(s.margin), C.c_int(0), C.c_int(1), C.byref(c)), "sharded")
        ctx.check(lib.ncb_world_fetch(h, _ffi.ptr(bufs["pairs"]), C.c_uint32(len(bufs["pairs"])), _ffi.ptr(bufs["algo"]), _ffi.ptr(bufs["start"]),
                                      _ffi.ptr(bufs["count"]), _ffi.ptr(bufs["contacts"]), C.c_uint32(len(bufs["contacts"]))), "fetch")
        P, Cn = c.n_pairs, c.n_contacts
        assert P == len(ref.pairs) and Cn == len(ref.contacts)
        # the pair order after the key sort is deterministic up to atomics inside one key segment: compare per pair
        def key(c):  # per-field bytes: a multi-field view still carries the `pair` back-reference, which depends on the pair order
            return b"".join(np.ascontiguousarray(c[f]).tobytes() for f in ("world1", "world2", "normal", "depth", "f1", "f2"))

        got = {tuple(p): key(bufs["contacts"][st : st + k])
               for p, st, k in zip(bufs["pairs"][:P].tolist(), bufs["start"][:P].tolist(), bufs["count"][:P].tolist())}
        want = {tuple(p): key(ref.contacts_of(i)) for i, p in enumerate(ref.pairs.tolist())}
        assert got == want, mode
        assert np.array_equal(np.sort(bufs["algo"][:P]), np.sort(ref.pair_algo))


def test_golden_fixtures_on_device(ctx):
    """tests/golden/*.npz (oracle outputs, see make_golden.py) reproduced by the device."""
    import glob
    import os

    from tests.golden.make_golden import scene_from_npz

    root = os.path.dirname(os.path.abspath(__file__))
    for f in sorted(glob.glob(os.path.join(root, "golden", "world_*.npz"))):
        z = np.load(f)
        s = scene_from_npz(z)
        ctx.set_hulls(s.hulls)
        r = ctx.world_update(s)
        assert np.array_equal(canon(r.pairs), canon(z["pairs"])), f
        order = {tuple(p): i for i, p in enumerate(map(tuple, z["pairs"].tolist()))}
        off = z["manifold_off"]
        for i, p in enumerate(map(tuple, r.pairs.tolist())):
            j = order[p]
            c = r.contacts_of(i)
            sl = slice(off[j], off[j + 1])
            assert len(c) == off[j + 1] - off[j], (f, p)
            assert np.array_equal(c["f1"], z["c_f1"][sl]) and np.array_equal(c["f2"], z["c_f2"][sl])
            for name in ("world1", "world2", "normal", "depth"):
                assert np.allclose(c[name], z["c_" + name][sl], rtol=RTOL, atol=ATOL), (f, p, name)
    for f in sorted(glob.glob(os.path.join(root, "golden", "rays_*.npz"))):
        z = np.load(f)
        mesh = ctx.trimesh(z["verts"], z["tris"])
        toi, face, normal = mesh.toi_and_normal_with_ray(None, z["origins"], z["dirs"])
        diff = np.nonzero(face != z["face"])[0]
        for i in diff:
            assert ulp_diff(toi[i], z["toi"][i]) <= 4
        same = face == z["face"]
        assert np.array_equal(toi[same], z["toi"][same])
        mesh.close()


def test_full_size_cfg1_cfg2_parity(ctx, oracle):
    """BASELINE.json configs[0] (1,000 balls) and configs[1] (100,000 balls / cuboids + one plane) at their full sizes."""
    for cfg, n in ((1, 1000), (2, 100_000)):
        s = config_scene(cfg)
        assert s.n in (n, n + 1)
        ctx.set_hulls(s.hulls)
        res = ctx.world_update(s)
        assert res.counts["epa_overflow"] == 0 and res.counts["ref_panics"] == 0
        want = oracle.broad_phase(oracle.compute_aabbs(s), s.groups, mode=0)
        assert np.array_equal(canon(res.pairs), canon(want)), f"cfg{cfg} pair set"
        compare_manifolds(res, s, oracle, f"cfg{cfg} full size")


def test_full_size_1M_world_update_parity(ctx, oracle):
    """BASELINE.json configs[2] at its full size: 1,000,000 mixed balls / cuboids / hulls.  The canonical pair set must be
    bit-exact against the reference-faithful DBVT broad phase of the oracle, and every manifold within tolerance."""
    s = config_scene(3)
    assert s.n == 1_000_000
    ctx.set_hulls(s.hulls)
    res = ctx.world_update(s)
    assert res.counts["epa_overflow"] == 0 and res.counts["ref_panics"] == 0
    fat = oracle.compute_aabbs(s)
    got = ctx.compute_aabbs(s.margin, 2)
    assert np.array_equal(got.view(np.uint32), fat.view(np.uint32)), "fat AABBs differ at 1M"
    want = oracle.broad_phase(fat, s.groups, mode=0)
    assert len(want) == len(res.pairs)
    assert np.array_equal(canon(res.pairs), canon(want)), "pair set differs at 1M"
    compare_manifolds(res, s, oracle, "cfg3 1M")
    # size-independent properties
    c = res.contacts
    assert np.allclose(np.linalg.norm(c["normal"], axis=1), 1, atol=1e-5)
    d = -np.einsum("ij,ij->i", c["normal"], c["world2"] - c["world1"])
    assert np.allclose(d, c["depth"], atol=3e-5)
    assert int(res.manifold_count.sum()) == len(c)


def test_full_size_1M_ray_cast_parity(ctx, oracle):
    """BASELINE.json configs[3] at full size: 1M rays vs a 1M-triangle terrain TriMesh, against the reference-faithful BVT
    best-first search of the oracle (face ids identical outside the tie class, TOI within tolerance)."""
    rs = make_ray_scene("terrain", 1_000_000, 1_000_000, seed=1004)
    nties, nhits = check_rays(ctx, oracle, rs, brute=False)
    assert nhits > 500_000
    assert nties <= 50
