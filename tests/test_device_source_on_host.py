"""The per-pair DEVICE functions of csrc/proximity.cu (+ the GJK / simplex code of gjk.cuh they call), compiled for the host
by g++ through tests/host_shim/cuda_runtime.h, against the oracle.  This checks the device SOURCE (logic, f32 operation
order, no FMA contraction) where no GPU exists; the compiled kernels themselves are checked by tests/test_proximity.py -m gpu.
Test infrastructure only: nothing in ncollide_b200/ uses the shim."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ncollide_b200 import _ffi
from ncollide_b200.scenes import make_world_scene, random_unit_quaternions

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
F = np.float32


def _build_shim(name, src):
    out = os.path.join(HERE, "host_shim", "_build", name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call([
        "g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-DNCB_HOST_SHIM",
        "-I", os.path.join(HERE, "host_shim"), "-I", os.path.join(ROOT, "ncollide_b200", "csrc"),
        "-shared", "-o", out, os.path.join(HERE, "host_shim", src),
    ])
    return C.CDLL(out)


@pytest.fixture(scope="module")
def shim():
    return _build_shim("libprox_host.so", "proximity_host.cpp")


@pytest.fixture(scope="module")
def ray_shim():
    return _build_shim("libray_host.so", "ray_host.cpp")


@pytest.fixture(scope="module")
def query_shim():
    return _build_shim("libquery_host.so", "query_host.cpp")


@pytest.fixture(scope="module")
def narrow_shim():
    return _build_shim("libnarrow_host.so", "narrow_host.cpp")


@pytest.fixture(scope="module")
def gjk_shim():
    return _build_shim("libgjk_host.so", "gjk_host.cpp")


def shim_proximity(lib, scene, pairs, margins=None, axis_io=None):
    oc, keep = _ffi.pack_objects(scene)
    hc, keep2 = _ffi.pack_hull_library(scene.hulls)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    out = np.zeros(len(pairs), dtype=np.uint8)
    m = None if margins is None else np.ascontiguousarray(margins, dtype=F)
    lib.shim_proximity(C.byref(oc), C.byref(hc), C.c_uint64(len(pairs)), _ffi.ptr(pairs), _ffi.ptr(m), _ffi.ptr(axis_io), _ffi.ptr(out))
    return out


SCENES = [
    lambda: make_world_scene(3000, 61, (1, 1, 1), side=8.0, plane=True, n_hulls=48, linear=0.1),
    lambda: make_world_scene(2000, 62, (0, 1, 1), side=4.0, n_hulls=24, linear=0.0),
    lambda: make_world_scene(2000, 63, (1, 0, 0), side=6.0, linear=0.2),
]


@pytest.mark.parametrize("mk", SCENES)
def test_device_proximity_source_matches_oracle(shim, oracle, mk):
    s = mk()
    fat = oracle.compute_aabbs(s)
    bp = oracle.broad_phase(fat, s.groups)
    rng = np.random.default_rng(17)
    extra = rng.integers(0, s.n, size=(20000, 2)).astype(np.uint32)
    pairs = np.concatenate([bp, bp[:, ::-1], extra[extra[:, 0] != extra[:, 1]]])
    got, want = shim_proximity(shim, s, pairs), oracle.proximity(s, pairs)
    assert np.array_equal(got, want), int((got != want).sum())
    margins = rng.uniform(0, 2.5, size=len(pairs)).astype(F)
    got, want = shim_proximity(shim, s, pairs, margins), oracle.proximity(s, pairs, margins)
    assert np.array_equal(got, want), int((got != want).sum())
    assert set(want.tolist()) >= {0, 1, 2}


def test_device_proximity_source_warm_start_matches_oracle(shim, oracle):
    """The detector's sep_axis carried over several updates while the objects move (the stepping-world use)."""
    s = make_world_scene(1500, 64, (1, 1, 1), side=5.0, n_hulls=24, linear=0.15)
    fat = oracle.compute_aabbs(s)
    pairs = oracle.broad_phase(fat, s.groups)
    rng = np.random.default_rng(23)
    ax_d = np.zeros((len(pairs), 4), dtype=F)
    ax_o = np.zeros((len(pairs), 4), dtype=F)
    seen = set()
    for step in range(5):
        got, want = shim_proximity(shim, s, pairs, None, ax_d), oracle.proximity_warm(s, pairs, None, ax_o)
        assert np.array_equal(got, want), (step, int((got != want).sum()))
        assert np.array_equal(ax_d.view(np.uint32), ax_o.view(np.uint32)), step
        seen |= set(want.tolist())
        s.pos = np.ascontiguousarray((s.pos + rng.normal(0, 0.08, size=s.pos.shape)).astype(F))
        s.rot = random_unit_quaternions(rng, s.n) if step == 2 else s.rot
    assert seen >= {0, 1, 2} and ax_o[:, 3].any()


@pytest.mark.parametrize("k", range(6))
def test_device_proximity_source_on_adversarial_scenes(shim, oracle, k):
    """Degenerate placements (everything coincident, exactly touching lattices, far from the origin, a dense clump, mixed scales):
    zero start directions, origin-on-simplex exits, ties.  All pairs of the broad phase, margins 0 / pair margins / large."""
    from test_gpu_parity import _adversarial_scenes

    s = _adversarial_scenes()[k]
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), s.groups)
    pairs = np.concatenate([pairs, pairs[:, ::-1]])
    for margins in (None, np.zeros(len(pairs), dtype=F), np.full(len(pairs), 0.5, dtype=F)):
        got, want = shim_proximity(shim, s, pairs, margins), oracle.proximity(s, pairs, margins)
        assert np.array_equal(got, want), (s.name, int((got != want).sum()))


def test_device_proximity_source_reproduces_the_golden_fixture(shim):
    """tests/golden/prox_mixed_plane_400.npz (oracle-made, committed): the device source must give the same statuses."""
    from test_proximity import _load_prox_golden

    z, s = _load_prox_golden()
    sel = z["algo"] == 6
    assert np.array_equal(shim_proximity(shim, s, z["pairs"][sel]), z["prox"][sel])
    assert np.array_equal(shim_proximity(shim, s, z["batch_pairs"], z["batch_margins"]), z["batch_prox"])


# ---- the device GJK / EPA of gjk.cuh (what k_cc_gjk / k_cc_epa run per pair) -----------------------------------------------------
def shim_contact_sm_sm(lib, scene, pairs, predictions=None, compact=False):
    oc, keep = _ffi.pack_objects(scene)
    hc, keep2 = _ffi.pack_hull_library(scene.hulls)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    out = np.zeros((len(pairs), 10), dtype=F)
    flags = np.zeros(4, dtype=np.uint32)
    m = None if predictions is None else np.ascontiguousarray(predictions, dtype=F)
    fn = lib.shim_contact_sm_sm_compact if compact else lib.shim_contact_sm_sm
    fn(C.byref(oc), C.byref(hc), C.c_uint64(len(pairs)), _ffi.ptr(pairs), _ffi.ptr(m), _ffi.ptr(out), _ffi.ptr(flags))
    return out, flags


def _convex_pairs(oracle, s):
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), s.groups)
    t = s.shape_type
    keep = (t[pairs[:, 0]] != 0) & (t[pairs[:, 1]] != 0) & (t[pairs[:, 0]] != 3) & (t[pairs[:, 1]] != 3)  # cuboid / hull operands
    return pairs[keep]


GJK_SCENES = [
    lambda: make_world_scene(3000, 71, (0, 1, 1), side=7.0, n_hulls=48, linear=0.02),
    lambda: make_world_scene(1500, 72, (0, 1, 1), side=3.0, n_hulls=24, linear=0.05),   # dense: deep penetrations, long EPA runs
    lambda: make_world_scene(2000, 73, (0, 1, 0), side=5.0, linear=0.02),               # cuboids only: parallel faces, degenerate simplices
]


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("mk", GJK_SCENES)
def test_device_gjk_epa_source_matches_oracle(gjk_shim, oracle, mk, compact):
    """contact_support_map_support_map through the device's gjk_closest_points + epa_init / epa_step (fixed-capacity polytope,
    packed topology, deferred heap pushes) against the oracle's std-container restatement: same found / not found, same points and
    normals BIT FOR BIT — the device follows the reference's iteration path, not just its maths.
    compact: EPA on the shared-memory polytope store of k_cc_epa_s (16 / 48 / 24 capacities, one packed word per face, face normals
    recomputed, slim operands), pairs beyond its capacities restarted on the big store like the kernel's overflow queue does."""
    s = mk()
    pairs = _convex_pairs(oracle, s)
    pairs = np.concatenate([pairs, pairs[:, ::-1]])
    assert len(pairs) > 2000
    got, flags = shim_contact_sm_sm(gjk_shim, s, pairs, compact=compact)
    want, stats = oracle.contact_sm_sm(s, pairs)
    assert flags[0] == 0 and flags[1] == 0, "EPA capacity overflow / reference panic"
    if compact:
        assert flags[3] < 0.1 * flags[2], "the compact store is sized for (almost) every pair"
    assert np.array_equal(got[:, 9], want[:, 9]), int((got[:, 9] != want[:, 9]).sum())
    assert stats[2] > 100 and stats[3] == 0  # EPA ran and never failed
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), int((got.view(np.uint32) != want.view(np.uint32)).any(axis=1).sum())


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("k", [0, 2, 3, 4])
def test_device_gjk_epa_source_on_adversarial_scenes(gjk_shim, oracle, k, compact):
    from test_gpu_parity import _adversarial_scenes

    s = _adversarial_scenes()[k]
    pairs = _convex_pairs(oracle, s)
    if len(pairs) == 0:
        pytest.skip("no convex pairs")
    got, flags = shim_contact_sm_sm(gjk_shim, s, pairs, compact=compact)
    want, stats = oracle.contact_sm_sm(s, pairs)
    assert flags[0] == 0
    assert np.array_equal(got[:, 9], want[:, 9])
    assert np.allclose(got, want, rtol=1e-4, atol=1e-5)


# ---- the whole fresh-world narrow phase of narrow.cu, one pair after the other --------------------------------------------------
def shim_narrow_phase(lib, scene, pairs):
    oc, keep = _ffi.pack_objects(scene)
    hc, keep2 = _ffi.pack_hull_library(scene.hulls)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    P = len(pairs)
    off = np.zeros(P + 1, dtype=np.uint32)
    algo = np.zeros(P, dtype=np.uint8)
    flags = np.zeros(4, dtype=np.uint32)
    lib.shim_narrow_phase.restype = C.c_uint64
    cap = max(4 * P, 64)
    while True:
        out = np.zeros(cap, dtype=_ffi.CONTACT_DTYPE)
        flags[:] = 0
        nc = lib.shim_narrow_phase(C.byref(oc), C.byref(hc), C.c_uint64(P), _ffi.ptr(pairs), _ffi.ptr(out), C.c_uint64(cap), _ffi.ptr(off), _ffi.ptr(algo),
                                   _ffi.ptr(flags))
        if nc <= cap:
            return out[:nc], off, algo, flags
        cap = int(nc)


def compare_narrow(got, want, label):
    (dc, doff, dalgo, flags), (oc, ooff, oalgo, _) = got, want
    assert flags[0] == 0 and flags[1] == 0, f"{label}: capacity overflow / reference panic"
    assert np.array_equal(dalgo, oalgo), f"{label}: dispatched algorithm"
    assert np.array_equal(doff, ooff), f"{label}: manifold sizes differ on {int((np.diff(doff) != np.diff(ooff)).sum())} pairs"
    assert np.array_equal(dc["f1"], oc["f1"]) and np.array_equal(dc["f2"], oc["f2"]), f"{label}: feature ids"
    inexact = 0
    for name in ("world1", "world2", "normal", "depth"):
        assert np.allclose(dc[name], oc[name], rtol=1e-4, atol=1e-5), f"{label}: {name}"
        inexact += int((dc[name].view(np.uint32) != oc[name].view(np.uint32)).sum())
    return inexact


NARROW_SCENES = [
    lambda: make_world_scene(4000, 81, (1, 1, 1), side=9.5, plane=True, n_hulls=48),
    lambda: make_world_scene(2500, 82, (1, 1, 1), side=6.0, n_hulls=32, angular=0.05, linear=0.05),
    lambda: make_world_scene(2000, 83, (0, 1, 1), side=4.5, n_hulls=24, angular=0.02),
    lambda: make_world_scene(2000, 84, (1, 1, 0), side=5.0, plane=True),
]


@pytest.mark.parametrize("mk", NARROW_SCENES)
def test_device_narrow_phase_source_matches_oracle(narrow_shim, oracle, mk):
    """All five contact generators (features, clipping, manifold with the 0.02 dedup) from the device source vs the oracle on the
    broad-phase pairs in the reference's callback orientation: algorithm, manifold sizes, feature ids exact; contacts bit-exact."""
    s = mk()
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), s.groups, mode=0)
    assert len(pairs) > 3000
    inexact = compare_narrow(shim_narrow_phase(narrow_shim, s, pairs), oracle.narrow_phase(s, pairs), "scene")
    assert inexact == 0, f"{inexact} contact fields within tolerance but not bit-exact"


def shim_narrow_phase_kinematics(lib, scene, pairs):
    """The harness with ncb_set_kinematics on: contacts + ContactKinematic records (capsule worlds included)."""
    oc, keep = _ffi.pack_objects(scene)
    hc, keep2 = _ffi.pack_hull_library(scene.hulls)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    seg = np.zeros((scene.n, 6), dtype=F)
    cap = scene.shape_type == 4
    seg[cap, 1] = scene.shape_param[cap, 0]
    seg[cap, 4] = -scene.shape_param[cap, 0]
    P = len(pairs)
    off = np.zeros(P + 1, dtype=np.uint32)
    algo = np.zeros(P, dtype=np.uint8)
    flags = np.zeros(4, dtype=np.uint32)
    lib.shim_narrow_phase_kin.restype = C.c_uint64
    cap_c = max(4 * P, 64)
    while True:
        out = np.zeros(cap_c, dtype=_ffi.CONTACT_DTYPE)
        kin = np.zeros(cap_c, dtype=_ffi.KINEMATIC_DTYPE)
        flags[:] = 0
        nc = lib.shim_narrow_phase_kin(C.byref(oc), C.byref(hc), _ffi.ptr(seg), C.c_uint64(P), _ffi.ptr(pairs), _ffi.ptr(out), _ffi.ptr(kin),
                                       C.c_uint64(cap_c), _ffi.ptr(off), _ffi.ptr(algo), _ffi.ptr(flags))
        if nc <= cap_c:
            return out[:nc], kin[:nc], off, algo
        cap_c = int(nc)


def compare_kinematics(dk, ok, label):
    """ContactKinematic records: geometry kinds and dilations exact, local points / directions bit for bit."""
    assert len(dk) == len(ok), label
    assert np.array_equal(dk["g1"], ok["g1"]) and np.array_equal(dk["g2"], ok["g2"]), f"{label}: NeighborhoodGeometry kinds"
    for name in ("local1", "local2", "dir1", "dir2", "dil1", "dil2"):
        a, b = np.ascontiguousarray(dk[name]), np.ascontiguousarray(ok[name])
        assert np.allclose(a, b, rtol=1e-4, atol=1e-5), f"{label}: {name}"
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{label}: {name} within tolerance but not bit-exact"


@pytest.mark.parametrize("mk", NARROW_SCENES + [lambda: _capsule_scene(2400, 141, (1, 1, 1), 8.0, True), lambda: _capsule_scene(1800, 142, (0, 1, 1), 5.5, False, 0.03)])
def test_device_contact_kinematics_match_oracle(narrow_shim, oracle, mk):
    """ContactKinematic (local1 / local2, NeighborhoodGeometry per side, dilations: contact_kinematic.rs:57-66) as the device source
    produces it with ncb_set_kinematics on, against the oracle's restatement of every generator's kinematic (ball-ball, plane-ball,
    plane-polyhedron, ball-polyhedron, add_contact_to_manifold for polyhedron pairs, the capsule preprocessor's dilation): bit for
    bit, both operand orders; the contacts themselves are unchanged by the request."""
    s = mk()
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), s.groups, mode=0)
    both = np.concatenate([pairs, pairs[:, ::-1]])
    dc, dk, doff, dalgo = shim_narrow_phase_kinematics(narrow_shim, s, both)
    oc, ok, ooff, oalgo = oracle.narrow_phase_kinematics(s, both)
    assert np.array_equal(doff, ooff) and np.array_equal(dalgo, oalgo)
    for name in ("world1", "world2", "normal", "depth", "f1", "f2"):
        assert np.array_equal(np.ascontiguousarray(dc[name]).view(np.uint32), np.ascontiguousarray(oc[name]).view(np.uint32)), name
    compare_kinematics(dk, ok, "kinematics")
    assert (ok["g1"] == 1).any() and (ok["g1"] == 2).any() and (ok["g2"] == 1).any() and (ok["g2"] == 2).any()
    # sanity of the restatement itself: Plane / Line directions are unit vectors, and the tracked local points, moved to world space
    # and dilated along the normal, are the contact points: world1 = m1 * local1 + n * dilation1, world2 = m2 * local2 - n * dilation2
    for g, d in (("g1", "dir1"), ("g2", "dir2")):
        sel = ok[g] != 0
        assert np.allclose(np.linalg.norm(ok[d][sel], axis=1), 1.0, atol=1e-5)
    pair_of = np.repeat(np.arange(len(both)), np.diff(ooff))

    def to_world(obj, local):
        q, t = s.rot[obj].astype(np.float64), s.pos[obj].astype(np.float64)
        qv = q[:, :3]
        tt = 2 * np.cross(qv, local)
        return local + q[:, 3:4] * tt + np.cross(qv, tt) + t

    n = oc["normal"].astype(np.float64)
    w1 = to_world(both[pair_of, 0], ok["local1"].astype(np.float64)) + n * ok["dil1"][:, None]
    w2 = to_world(both[pair_of, 1], ok["local2"].astype(np.float64)) - n * ok["dil2"][:, None]
    assert np.allclose(w1, oc["world1"], atol=2e-5) and np.allclose(w2, oc["world2"], atol=2e-5)
    # a Plane geometry is (within the angular tolerances of the support features) the contact normal seen from the object
    def to_world_vec(obj, v):
        q = s.rot[obj].astype(np.float64)
        qv = q[:, :3]
        tt = 2 * np.cross(qv, v)
        return v + q[:, 3:4] * tt + np.cross(qv, tt)

    p1 = ok["g1"] == 2
    assert (np.einsum("ij,ij->i", to_world_vec(both[pair_of, 0], ok["dir1"].astype(np.float64))[p1], n[p1]) > 0.3).all()  # the support FACE toward the normal (cuboid: >= 1 / sqrt 3)
    p2 = ok["g2"] == 2
    assert (np.einsum("ij,ij->i", to_world_vec(both[pair_of, 1], ok["dir2"].astype(np.float64))[p2], n[p2]) < -0.3).all()


@pytest.mark.parametrize("k", range(6))
def test_device_narrow_phase_source_on_adversarial_scenes(narrow_shim, oracle, k):
    from test_gpu_parity import _adversarial_scenes

    s = _adversarial_scenes()[k]
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), s.groups, mode=0)
    compare_narrow(shim_narrow_phase(narrow_shim, s, pairs), oracle.narrow_phase(s, pairs), s.name)


def test_device_narrow_phase_source_reproduces_the_golden_fixtures(narrow_shim):
    import glob

    from golden.make_golden import scene_from_npz

    files = sorted(glob.glob(os.path.join(HERE, "golden", "world_*.npz")))
    assert files
    for f in files:
        z = np.load(f)
        s = scene_from_npz(z)
        dc, doff, dalgo, flags = shim_narrow_phase(narrow_shim, s, z["pairs"])
        assert np.array_equal(doff, z["manifold_off"]) and np.array_equal(dalgo, z["algo"]), f
        assert np.array_equal(dc["f1"], z["c_f1"]) and np.array_equal(dc["f2"], z["c_f2"]), f
        for name in ("world1", "world2", "normal", "depth"):
            assert np.allclose(dc[name], z["c_" + name], rtol=1e-4, atol=1e-5), (f, name)


# ---- stepping world, per pair: persistent manifold cache + GJK warm start (pm_load_and_age / manifold_push<true> / pm_store) ----
class ShimEdges:
    """The device's per-pair persistent state (slot = pair index) driven through tests/host_shim/narrow_host.cpp."""

    def __init__(self, lib, pairs):
        self.lib, self.pairs = lib, np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        P = len(self.pairs)
        self.dir = np.zeros((P, 4), dtype=F)
        self.hdr = np.zeros((P, 8), dtype=np.uint32)         # PM_HDR_WORDS
        self.entry = np.zeros((P, 24 * 4, 4), dtype=F)       # PM_CAP * PM_ENTRY_F4 float4
        self.flags = np.zeros(4, dtype=np.uint32)
        self.overflow = np.zeros(1, dtype=np.uint32)

    def update(self, scene, which):
        oc, keep = _ffi.pack_objects(scene)
        hc, keep2 = _ffi.pack_hull_library(scene.hulls)
        which = np.ascontiguousarray(which, dtype=np.uint32)
        pr = np.ascontiguousarray(self.pairs[which])
        ev = np.zeros(len(which) + 1, dtype=np.uint64)
        nev = np.zeros(1, dtype=np.uint32)
        seg = np.zeros((scene.n, 6), dtype=F)  # capsules (shape_type 4): the 2-point hull [b, a] of the segment
        cap = scene.shape_type == 4
        seg[cap, 1] = scene.shape_param[cap, 0]
        seg[cap, 4] = -scene.shape_param[cap, 0]
        self.lib.shim_persist_update_ex(C.byref(oc), C.byref(hc), _ffi.ptr(seg), C.c_uint64(len(which)), _ffi.ptr(pr), _ffi.ptr(which), _ffi.ptr(self.dir),
                                        _ffi.ptr(self.hdr), _ffi.ptr(self.entry), _ffi.ptr(ev), _ffi.ptr(nev), C.c_uint32(len(ev)), _ffi.ptr(self.overflow),
                                        _ffi.ptr(self.flags))
        ev = ev[: int(nev[0])]
        return np.stack([(ev >> np.uint64(32)) & np.uint64(0x7FFFFFFF), ev & np.uint64(0xFFFFFFFF), ev >> np.uint64(63)], axis=1).astype(np.uint32)

    def fetch(self):
        P = len(self.pairs)
        slots = np.arange(P, dtype=np.uint32)
        off = np.zeros(P + 1, dtype=np.uint32)
        self.lib.shim_persist_export.restype = C.c_uint64
        cap = max(8 * P, 64)
        while True:
            c = np.zeros(cap, dtype=_ffi.CONTACT_DTYPE)
            ids = np.zeros(cap, dtype=np.uint32)
            nc = self.lib.shim_persist_export(C.c_uint64(P), _ffi.ptr(slots), _ffi.ptr(self.hdr), _ffi.ptr(self.entry), _ffi.ptr(off), _ffi.ptr(c), _ffi.ptr(ids),
                                              C.c_uint64(cap))
            if nc <= cap:
                return c[:nc], off, ids[:nc], self.dir
            cap = int(nc)


def _sorted_rows(a):
    a = np.asarray(a).reshape(-1, 3)
    return a[np.lexsort(a.T[::-1])] if len(a) else a


@pytest.mark.parametrize("n,kinds,side,plane,ang,seed", [(1500, (1, 1, 1), 6.5, True, 0.0, 3), (1500, (0, 1, 1), 5.5, False, 0.02, 5), (1200, (1, 1, 0), 5.0, True, 0.0, 7)])
def test_device_persistent_manifold_source_matches_oracle(narrow_shim, oracle, n, kinds, side, plane, ang, seed):
    """Seven updates of a fixed edge set while the objects jitter / jump / rotate: after every update the device's persistent
    manifolds (contacts in slab order, contact ids, feature ids), the generators' warm-start directions and the contact events
    equal the oracle's ContactManifold cache — values bit for bit."""
    from sim_scenario import step_poses

    s = make_world_scene(n, 90 + seed, kinds, side=side, n_hulls=24, plane=plane, angular=ang)
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), s.groups, mode=0)
    t = s.shape_type
    pairs = pairs[~((t[pairs[:, 0]] == 3) & (t[pairs[:, 1]] == 3))]
    dev, orc = ShimEdges(narrow_shim, pairs), oracle.edges(pairs)
    rng = np.random.default_rng(seed)
    n_events = n_kept_ids = 0
    prev_ids = None
    for step in range(7):
        if step == 0:
            which = np.arange(len(pairs), dtype=np.uint32)
        elif step == 3:
            which = np.zeros(0, dtype=np.uint32)  # an idle update
        else:
            idx = step_poses(s, s.pos, s.rot, rng, 0.4)
            moved = np.zeros(s.n, dtype=bool)
            moved[idx] = True
            which = np.nonzero(moved[pairs[:, 0]] | moved[pairs[:, 1]])[0].astype(np.uint32)
        ed, eo = dev.update(s, which), orc.update(s, which)
        assert dev.flags[0] == 0 and dev.flags[1] == 0 and dev.overflow[0] == 0
        assert np.array_equal(_sorted_rows(ed), _sorted_rows(eo)), f"step {step}: contact events"
        n_events += len(eo) if step else 0
        (dc, doff, dids, ddir), (oc, ooff, oids, odir) = dev.fetch(), orc.fetch()
        assert np.array_equal(doff, ooff), f"step {step}: manifold sizes differ on {int((np.diff(doff) != np.diff(ooff)).sum())} edges"
        assert np.array_equal(dids, oids), f"step {step}: contact ids"
        for name in ("f1", "f2", "world1", "world2", "normal", "depth"):
            assert np.array_equal(dc[name].view(np.uint32), oc[name].view(np.uint32)), f"step {step}: {name}"
        has = odir[:, 3] != 0
        assert np.array_equal(ddir[:, 3] != 0, has) and np.array_equal(ddir[has].view(np.uint32), odir[has].view(np.uint32)), f"step {step}: last_gjk_dir"
        if prev_ids is not None:
            n_kept_ids += len(np.intersect1d(prev_ids, (np.repeat(np.arange(len(pairs)), np.diff(ooff)).astype(np.uint64) << np.uint64(32)) | oids))
        prev_ids = (np.repeat(np.arange(len(pairs)), np.diff(ooff)).astype(np.uint64) << np.uint64(32)) | oids
    assert n_events > 5 and n_kept_ids > 100  # contacts started / stopped, and contacts that kept their id across updates


# ---- world queries: per-shape ray casts and point containment of query.cu -----------------------------------------------------------
def _rays_at_objects(s, rng, k):
    """k rays per call aimed at random objects: origins on a shell around the object (some inside it), directions towards it with
    jitter (some pointing away), a mix of max_toi values."""
    which = rng.integers(0, s.n, size=k).astype(np.uint32)
    c = s.pos[which]
    off = rng.normal(size=(k, 3))
    off /= np.linalg.norm(off, axis=1, keepdims=True)
    r = rng.choice([0.0, 0.2, 0.8, 2.0, 6.0], size=k)[:, None]
    o = (c + off * r).astype(F)
    d = (-off + rng.normal(0, 0.35, size=(k, 3))).astype(F)
    d[rng.random(k) < 0.1] *= F(-1)
    d[rng.random(k) < 0.3] *= F(3.7)  # the reference does not require unit directions
    t = rng.choice([0.5, 3.0, 1e3, np.finfo(np.float32).max], size=k).astype(F)
    return which, np.ascontiguousarray(np.concatenate([o, d, t[:, None]], axis=1), dtype=F)


@pytest.mark.parametrize("seed,plane", [(1, True), (2, False), (3, True)])
def test_device_shape_ray_casts_and_point_queries_match_oracle(query_shim, oracle, seed, plane):
    """RayCast::toi_and_normal_with_ray(solid = true) per shape (ball, cuboid with the reference's face ids, plane, hull through the
    GJK ray cast) and PointQuery::contains_point: hit / miss, feature ids and inside flags exact, toi and normals bit for bit."""
    s = make_world_scene(600, 100 + seed, (1, 1, 1), side=6.0, n_hulls=32, plane=plane)
    rng = np.random.default_rng(seed)
    which, rays = _rays_at_objects(s, rng, 30000)
    if plane:
        which[:2000] = s.n - 1  # the plane
    oc, keep = _ffi.pack_objects(s)
    hc, keep2 = _ffi.pack_hull_library(s.hulls)
    out = np.zeros((len(which), 4), dtype=F)
    feat = np.zeros(len(which), dtype=np.uint32)
    hit = np.zeros(len(which), dtype=np.uint8)
    query_shim.shim_shape_ray_cast(C.byref(oc), C.byref(hc), C.c_uint64(len(which)), _ffi.ptr(which), _ffi.ptr(rays), _ffi.ptr(out), _ffi.ptr(feat), _ffi.ptr(hit))
    ohit, oout, ofeat = oracle.shape_ray_cast_batch(s, which, rays)
    assert np.array_equal(hit, ohit), int((hit != ohit).sum())
    assert 0.2 < hit.mean() < 0.95
    assert np.array_equal(feat, ofeat)
    assert np.array_equal(out.view(np.uint32), oout.view(np.uint32)), int((out.view(np.uint32) != oout.view(np.uint32)).any(axis=1).sum())
    for t in range(4):
        assert hit[s.shape_type[which] == t].any() or (t == 3 and not plane)
    # points: near / inside the objects
    pts = (s.pos[which] + rng.normal(0, 0.3, size=(len(which), 3))).astype(F)
    inside = np.zeros(len(which), dtype=np.uint8)
    query_shim.shim_shape_contains_point(C.byref(oc), C.byref(hc), C.c_uint64(len(which)), _ffi.ptr(which), _ffi.ptr(pts), _ffi.ptr(inside))
    want = oracle.shape_contains_point_batch(s, which, pts)
    assert np.array_equal(inside, want) and 0.1 < want.mean() < 0.9


# ---- TriMesh ray casting: slab_toi / ray_triangle of ray.cu with the device's hit semantics, without the tree ------------------------
@pytest.mark.parametrize("kind,posed", [("terrain", False), ("soup", False), ("soup", True)])
def test_device_trimesh_ray_primitives_match_oracle(ray_shim, oracle, kind, posed):
    """Minimum toi over {triangle AABB entered, triangle hit, toi <= max_toi}, ties -> smallest face, back faces as face + T, normalised
    normal: the device's per-ray arithmetic against the oracle's brute-force mode — faces exact, toi and normals bit for bit."""
    from ncollide_b200.scenes import make_ray_scene

    rs = make_ray_scene(kind, 3000, 2500, seed=1204, random_pose=posed)
    pose = rs.pose if posed else None
    if posed:  # the generator aims the rays at the mesh in its LOCAL frame: move them with the mesh
        t, (qi, qj, qk, qw) = rs.pose[:3].astype(np.float64), rs.pose[3:].astype(np.float64)
        R = np.array([[1 - 2 * (qj * qj + qk * qk), 2 * (qi * qj - qk * qw), 2 * (qi * qk + qj * qw)],
                      [2 * (qi * qj + qk * qw), 1 - 2 * (qi * qi + qk * qk), 2 * (qj * qk - qi * qw)],
                      [2 * (qi * qk - qj * qw), 2 * (qj * qk + qi * qw), 1 - 2 * (qi * qi + qj * qj)]])
        rs.origins = np.ascontiguousarray((rs.origins @ R.T + t).astype(F))
        rs.dirs = np.ascontiguousarray((rs.dirs @ R.T).astype(F))
    for max_toi in (np.finfo(np.float32).max, 12.0):
        n = len(rs.origins)
        toi = np.zeros(n, dtype=F)
        face = np.zeros(n, dtype=np.uint32)
        normal = np.zeros((n, 3), dtype=F)
        ray_shim.shim_trimesh_ray_cast(C.c_uint32(len(rs.tris)), _ffi.ptr(rs.verts), _ffi.ptr(rs.tris), _ffi.ptr(np.ascontiguousarray(pose, dtype=F)) if posed else None,
                                       C.c_uint64(n), _ffi.ptr(rs.origins), _ffi.ptr(rs.dirs), C.c_float(max_toi), _ffi.ptr(toi), _ffi.ptr(face), _ffi.ptr(normal))
        otoi, oface, onormal = oracle.trimesh(rs.verts, rs.tris).ray_cast(rs.origins, rs.dirs, max_toi=max_toi, pose=pose, mode=1)
        hit = otoi >= 0
        assert hit.sum() >= 100
        assert np.array_equal(face[hit], oface[hit]) and np.array_equal(toi >= 0, hit)
        assert np.array_equal(toi[hit].view(np.uint32), otoi[hit].view(np.uint32))
        assert np.array_equal(normal[hit].view(np.uint32), onormal[hit].view(np.uint32))


def test_device_gjk_epa_handles_segments_as_two_point_hulls(gjk_shim, oracle):
    """Plan check for the capsule device path: the segment of a capsule is given to the EXISTING device GJK / EPA as a
    2-point hull [b, a] — in that order the hull scan's first-maximum rule is the segment's own support rule — and must reproduce the
    oracle's contact_support_map_support_map on (segment, segment) and (segment, cuboid) pairs bit for bit."""
    from ncollide_b200.scenes import WorldScene

    rng = np.random.default_rng(31)
    n = 400
    s = make_world_scene(n, 131, (0, 1, 0), side=3.2, linear=0.05)  # cuboids; every second one becomes a capsule
    cap = np.arange(n) % 2 == 0
    hh = rng.uniform(0.2, 0.6, size=n).astype(F)
    s_or = WorldScene(pos=s.pos, rot=s.rot, shape_type=np.where(cap, 4, s.shape_type).astype(np.uint32), shape_param=s.shape_param.copy(),
                      groups=None, query_limit=s.query_limit, ang_pred=s.ang_pred, hulls=s.hulls, margin=s.margin)
    s_or.shape_param[cap, 0] = hh[cap]
    s_or.shape_param[cap, 1] = F(0.1)
    s_or.shape_param[cap, 2:] = 0

    class Lib:  # only the vertex tables are read by the support function
        FIELDS = s.hulls.FIELDS

    lib = Lib()
    ids = np.cumsum(cap) - 1
    lib.n_hulls = int(cap.sum())
    for f in Lib.FIELDS:
        setattr(lib, f, np.zeros(1, dtype=np.float32 if f in ("points", "face_normal", "edge_dir") else np.uint32))
    lib.vert_off = (2 * np.arange(lib.n_hulls + 1)).astype(np.uint32)
    for f in ("face_off", "edge_off", "fadj_off", "vadj_off"):
        setattr(lib, f, np.zeros(lib.n_hulls + 1, dtype=np.uint32))
    pts = np.zeros((lib.n_hulls, 2, 3), dtype=F)
    pts[:, 0, 1] = hh[cap]    # b first
    pts[:, 1, 1] = -hh[cap]   # then a
    lib.points = np.ascontiguousarray(pts.reshape(-1, 3))
    s_dev = WorldScene(pos=s.pos, rot=s.rot, shape_type=np.where(cap, 2, s.shape_type).astype(np.uint32), shape_param=s.shape_param.copy(),
                       groups=None, query_limit=s.query_limit, ang_pred=s.ang_pred, hulls=lib, margin=s.margin)
    s_dev.shape_param[cap, 0] = ids[cap].astype(F)
    pairs = rng.integers(0, n, size=(30000, 2)).astype(np.uint32)
    pairs = pairs[(pairs[:, 0] != pairs[:, 1]) & (cap[pairs[:, 0]] | cap[pairs[:, 1]])]
    d = np.linalg.norm(s.pos[pairs[:, 0]] - s.pos[pairs[:, 1]], axis=1)
    pairs = pairs[d < 1.0]
    assert len(pairs) > 800
    got, flags = shim_contact_sm_sm(gjk_shim, s_dev, pairs)
    want, stats = oracle.contact_sm_sm(s_or, pairs)
    assert flags[0] == 0 and flags[1] == 0
    assert np.array_equal(got[:, 9], want[:, 9]) and 0.1 < want[:, 9].mean() < 1.0 and stats[2] > 50
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), int((got.view(np.uint32) != want.view(np.uint32)).any(axis=1).sum())


# ---- capsules: the STAGED device functions of csrc/capsule.cuh (no kernel calls them yet) --------------------------------------------
def shim_narrow_phase_capsules(lib, scene, pairs):
    """scene.shape_type may contain 4 (capsule: shape_param = half_height, radius)."""
    oc, keep = _ffi.pack_objects(scene)
    hc, keep2 = _ffi.pack_hull_library(scene.hulls)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    seg = np.zeros((scene.n, 6), dtype=F)  # the 2-point hull [b, a] of every capsule's segment
    cap = scene.shape_type == 4
    seg[cap, 1] = scene.shape_param[cap, 0]
    seg[cap, 4] = -scene.shape_param[cap, 0]
    P = len(pairs)
    off = np.zeros(P + 1, dtype=np.uint32)
    algo = np.zeros(P, dtype=np.uint8)
    flags = np.zeros(4, dtype=np.uint32)
    lib.shim_narrow_phase_ex.restype = C.c_uint64
    cap_c = max(4 * P, 64)
    while True:
        out = np.zeros(cap_c, dtype=_ffi.CONTACT_DTYPE)
        flags[:] = 0
        nc = lib.shim_narrow_phase_ex(C.byref(oc), C.byref(hc), _ffi.ptr(seg), C.c_uint64(P), _ffi.ptr(pairs), _ffi.ptr(out), C.c_uint64(cap_c), _ffi.ptr(off),
                                      _ffi.ptr(algo), _ffi.ptr(flags))
        if nc <= cap_c:
            return out[:nc], off, algo, flags
        cap_c = int(nc)


def _capsule_scene(n, seed, kinds, side, plane, ang=0.0):
    s = make_world_scene(n, seed, kinds, side=side, n_hulls=24, plane=plane, angular=ang)
    rng = np.random.default_rng(seed + 1)
    cap = np.zeros(s.n, dtype=bool)
    cap[:n] = np.arange(n) % 3 == 1
    s.shape_type[cap] = 4
    s.shape_param[cap, 0] = rng.uniform(0.2, 0.5, size=int(cap.sum())).astype(F)
    s.shape_param[cap, 1] = rng.uniform(0.15, 0.3, size=int(cap.sum())).astype(F)
    s.shape_param[cap, 2:] = 0
    return s


@pytest.mark.parametrize("n,kinds,side,plane,ang,seed", [(2400, (1, 1, 1), 8.0, True, 0.0, 141), (1800, (0, 1, 1), 5.5, False, 0.03, 142), (1500, (1, 1, 0), 5.0, True, 0.0, 143)])
def test_capsule_device_source_matches_oracle(narrow_shim, oracle, n, kinds, side, plane, ang, seed):
    """Worlds with capsules through csrc/capsule.cuh (capsule_pair, the function k_capsule runs per thread: segment features, the capsule's contact preprocessor, the capsule generators on top
    of the existing GJK / EPA / clipping / manifold code) against the oracle: algorithm, manifold sizes, feature ids exact, contacts
    bit for bit — for capsule x {capsule, ball, cuboid, hull, plane} in both orders, and unchanged results for the other pairs."""
    s = _capsule_scene(n, seed, kinds, side, plane, ang)
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), s.groups, mode=0)
    both = np.concatenate([pairs, pairs[:, ::-1]])
    got, want = shim_narrow_phase_capsules(narrow_shim, s, both), oracle.narrow_phase(s, both)
    assert (want[2] == 7).sum() > 100 and (want[2] == 8).sum() > 500
    inexact = compare_narrow(got, want, "capsules")
    assert inexact == 0, f"{inexact} contact fields within tolerance but not bit-exact"
    t = s.shape_type
    for other in set(t[t != 4].tolist()):
        sel = ((t[both[:, 0]] == 4) & (t[both[:, 1]] == other)) | ((t[both[:, 0]] == other) & (t[both[:, 1]] == 4))
        assert np.diff(want[1])[sel].sum() > 0, f"no contact between a capsule and shape {other}"


def test_capsule_device_source_reproduces_the_golden_fixture(narrow_shim):
    from golden.make_golden import scene_from_npz

    z = np.load(os.path.join(HERE, "golden", "capsule_mixed_plane_300.npz"))
    s = scene_from_npz(z)
    dc, doff, dalgo, flags = shim_narrow_phase_capsules(narrow_shim, s, z["pairs"])
    assert flags[0] == 0 and flags[1] == 0
    assert np.array_equal(doff, z["manifold_off"]) and np.array_equal(dalgo, z["algo"])
    for name in ("f1", "f2", "world1", "world2", "normal", "depth"):
        assert np.array_equal(dc[name].view(np.uint32), z["c_" + name].view(np.uint32)), name


def test_capsule_aabb_matches_oracle(narrow_shim, oracle):
    s = _capsule_scene(3000, 151, (1, 1, 1), 9.0, False)
    oc, keep = _ffi.pack_objects(s)
    out = np.zeros((s.n, 6), dtype=F)
    narrow_shim.shim_capsule_aabbs(C.byref(oc), _ffi.ptr(out))
    cap = s.shape_type == 4
    want = oracle.compute_aabbs(s, mode=0)
    assert cap.sum() == 1000 and np.array_equal(out[cap].view(np.uint32), want[cap].view(np.uint32))


def test_capsule_persistent_state_matches_oracle(narrow_shim, oracle):
    """Stepping-world state per pair with capsules: the capsule generators instantiated with the persistent manifold (load + age, warm
    started GJK on the segment hulls, store, contact events) over six updates against the oracle's caller-driven edges."""
    from sim_scenario import step_poses

    s = _capsule_scene(1200, 161, (1, 1, 1), 5.8, True)
    pairs = oracle.broad_phase(oracle.compute_aabbs(s), s.groups, mode=0)
    t = s.shape_type
    pairs = pairs[~((t[pairs[:, 0]] == 3) & (t[pairs[:, 1]] == 3))]
    dev, orc = ShimEdges(narrow_shim, pairs), oracle.edges(pairs)
    rng = np.random.default_rng(5)
    n_events = 0
    for step in range(6):
        if step == 0:
            which = np.arange(len(pairs), dtype=np.uint32)
        else:
            idx = step_poses(s, s.pos, s.rot, rng, 0.4)
            moved = np.zeros(s.n, dtype=bool)
            moved[idx] = True
            which = np.nonzero(moved[pairs[:, 0]] | moved[pairs[:, 1]])[0].astype(np.uint32)
        ed, eo = dev.update(s, which), orc.update(s, which)
        assert dev.flags[0] == 0 and dev.flags[1] == 0 and dev.overflow[0] == 0
        assert np.array_equal(_sorted_rows(ed), _sorted_rows(eo)), f"step {step}: contact events"
        n_events += len(eo) if step else 0
        (dc, doff, dids, ddir), (oc, ooff, oids, odir) = dev.fetch(), orc.fetch()
        assert np.array_equal(doff, ooff) and np.array_equal(dids, oids), f"step {step}: manifold sizes / contact ids"
        for name in ("f1", "f2", "world1", "world2", "normal", "depth"):
            assert np.array_equal(dc[name].view(np.uint32), oc[name].view(np.uint32)), f"step {step}: {name}"
        has = odir[:, 3] != 0
        assert np.array_equal(ddir[:, 3] != 0, has) and np.array_equal(ddir[has].view(np.uint32), odir[has].view(np.uint32)), f"step {step}: last_gjk_dir"
    capsule_edges = (t[pairs[:, 0]] == 4) | (t[pairs[:, 1]] == 4)
    assert n_events > 5 and np.diff(ooff)[capsule_edges].sum() > 100
