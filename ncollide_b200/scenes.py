"""Synthetic scene generators for the BASELINE.json configs.

All arrays are f32 / u32, generated with numpy's PCG64 from a fixed seed so that the CUDA path, the oracle
and the CPU baseline see bit-identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .shapes import BALL, CUBOID, HULL, PLANE, ConvexHull, HullLibrary

F32 = np.float32

DEFAULT_GROUPS = (0x3FFFFFFF, 0x3FFFFFFF, 0)  # collision_groups.rs:54-60 (membership, whitelist, blacklist)


@dataclass
class WorldScene:
    pos: np.ndarray  # [N,3] f32
    rot: np.ndarray  # [N,4] f32 (i,j,k,w)
    shape_type: np.ndarray  # [N] u32
    shape_param: np.ndarray  # [N,4] f32
    groups: np.ndarray  # [N,3] u32
    query_limit: np.ndarray  # [N] f32  GeometricQueryType::Contacts(linear, _)
    ang_pred: np.ndarray  # [N] f32  GeometricQueryType::Contacts(_, angular)
    hulls: HullLibrary
    margin: float = 0.02
    name: str = ""
    # GeometricQueryType per object (pipeline/object/query_type.rs:8-37): None / 0 = Contacts(query_limit, ang_pred),
    # 1 = Proximity(query_limit) — a sensor: its pairs get a Proximity status instead of a contact manifold.
    query_kind: np.ndarray | None = None  # [N] u8

    @property
    def n(self):
        return len(self.pos)


def with_sensors(scene, fraction, seed, margin=None):
    """Marks a seeded random `fraction` of the objects as GeometricQueryType::Proximity(margin) (margin None: keep query_limit)."""
    rng = np.random.default_rng(seed)
    kind = (rng.random(scene.n) < fraction).astype(np.uint8)
    scene.query_kind = np.ascontiguousarray(kind)
    if margin is not None:
        ql = scene.query_limit.copy()
        ql[kind != 0] = F32(margin)
        scene.query_limit = np.ascontiguousarray(ql)
    return scene


def random_unit_quaternions(rng, n):
    q = rng.standard_normal((n, 4)).astype(F32)
    nrm = np.sqrt((q.astype(F32) ** 2).sum(axis=1, dtype=F32), dtype=F32)
    nrm[nrm == 0] = 1
    q = (q / nrm[:, None]).astype(F32)
    # re-normalise once more in f32
    nrm = np.sqrt((q * q).sum(axis=1, dtype=F32), dtype=F32)
    return (q / nrm[:, None]).astype(F32)


def make_hull_library(rng, n_hulls=1024, min_pts=8, max_pts=32):
    hulls = []
    while len(hulls) < n_hulls:
        k = int(rng.integers(min_pts, max_pts + 1))
        p = rng.standard_normal((k, 3))
        r = rng.uniform(0.25, 0.5)
        p = p / np.linalg.norm(p, axis=1).max() * r
        h = ConvexHull.try_from_points(p.astype(F32))
        if h is not None:
            hulls.append(h)
    return HullLibrary(hulls)


def box_side_for(n, mean_half_extent=0.45, neighbours=4.0):
    """L such that per-object AABB neighbours ~= N (4h)^3 / L^3 = `neighbours`."""
    return float((n * (4.0 * mean_half_extent) ** 3 / neighbours) ** (1.0 / 3.0))


def make_world_scene(
    n,
    seed,
    fractions=(1.0, 0.0, 0.0),
    side=None,
    plane=False,
    n_hulls=1024,
    linear=0.02,
    angular=0.0,
    margin=0.02,
    ball_radius=None,
    hull_library=None,
    name="",
):
    """N objects: `fractions` = (balls, cuboids, hulls); centres ~ U[0, side)^3; random rotations."""
    rng = np.random.default_rng(seed)
    fb, fc, fh = fractions
    nb = int(round(n * fb / (fb + fc + fh)))
    nc = min(int(round(n * fc / (fb + fc + fh))), n - nb)
    if fh == 0:  # no hulls asked for: rounding must not create one
        if fc == 0:
            nb = n
        nc = n - nb
    nh = n - nb - nc
    types = np.concatenate([np.full(nb, BALL), np.full(nc, CUBOID), np.full(nh, HULL)]).astype(np.uint32)
    rng.shuffle(types)
    if side is None:
        side = box_side_for(n)
    pos = (rng.random((n, 3)) * side).astype(F32)
    rot = random_unit_quaternions(rng, n)
    param = np.zeros((n, 4), dtype=F32)
    isb, isc, ish = types == BALL, types == CUBOID, types == HULL
    if ball_radius is not None:
        param[isb, 0] = F32(ball_radius)
    else:
        param[isb, 0] = rng.uniform(0.25, 0.5, size=isb.sum()).astype(F32)
    param[isc, :3] = rng.uniform(0.25, 0.5, size=(isc.sum(), 3)).astype(F32)
    if nh:
        lib = hull_library if hull_library is not None else make_hull_library(rng, n_hulls)
        param[ish, 0] = rng.integers(0, lib.n_hulls, size=ish.sum()).astype(F32)
    else:
        lib = hull_library if hull_library is not None else HullLibrary([])
    if plane:
        types = np.concatenate([types, np.array([PLANE], dtype=np.uint32)])
        pos = np.concatenate([pos, np.zeros((1, 3), dtype=F32)])
        rot = np.concatenate([rot, np.array([[0, 0, 0, 1]], dtype=F32)])
        param = np.concatenate([param, np.array([[0, 1, 0, 0]], dtype=F32)])
    n_all = len(types)
    groups = np.tile(np.array(DEFAULT_GROUPS, dtype=np.uint32), (n_all, 1))
    return WorldScene(
        pos=np.ascontiguousarray(pos),
        rot=np.ascontiguousarray(rot),
        shape_type=np.ascontiguousarray(types),
        shape_param=np.ascontiguousarray(param),
        groups=np.ascontiguousarray(groups),
        query_limit=np.full(n_all, linear, dtype=F32),
        ang_pred=np.full(n_all, angular, dtype=F32),
        hulls=lib,
        margin=margin,
        name=name,
    )


def config_scene(cfg, n=None, seed=None):
    """The BASELINE.json configs 1-3 and 5, optionally at a reduced object count."""
    if cfg == 1:
        n = n or 1000
        side = 10.0 * (n / 1000.0) ** (1.0 / 3.0)
        return make_world_scene(n, seed or 1001, (1, 0, 0), side=side, ball_radius=0.5, name=f"cfg1_balls_{n}")
    if cfg == 2:
        n = n or 100_000
        return make_world_scene(n, seed or 1002, (1, 1, 0), plane=True, name=f"cfg2_balls_cuboids_plane_{n}")
    if cfg == 3:
        n = n or 1_000_000
        return make_world_scene(n, seed or 1003, (1, 1, 1), name=f"cfg3_mixed_{n}")
    if cfg == 5:
        n = n or 16_000_000
        return make_world_scene(n, seed or 1005, (0, 1, 1), name=f"cfg5_convex_{n}")
    raise ValueError(cfg)


# ------------------------------------------------------------------------------------------------
# Ray-casting scenes (config 4)
# ------------------------------------------------------------------------------------------------
@dataclass
class RayScene:
    verts: np.ndarray  # [V,3] f32
    tris: np.ndarray  # [T,3] u32
    origins: np.ndarray  # [R,3] f32
    dirs: np.ndarray  # [R,3] f32
    pose: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0, 0, 0, 0, 1], dtype=F32))
    name: str = ""


def make_terrain(nx, ny, seed, size_x=100.0, size_y=50.0):
    """nx x ny quads -> 2 nx ny triangles; height = seeded sum of 8 sinusoids in [0, 5]."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(0.0, size_x, nx + 1)
    ys = np.linspace(0.0, size_y, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    Z = np.zeros_like(X)
    for _ in range(8):
        fx, fy = rng.uniform(0.02, 0.6, size=2)
        ph = rng.uniform(0, 2 * np.pi)
        Z += np.sin(fx * X + fy * Y + ph)
    Z = (Z - Z.min()) / max(Z.max() - Z.min(), 1e-9) * 5.0
    verts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(F32)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    v00 = (i * (ny + 1) + j).ravel()
    v10 = ((i + 1) * (ny + 1) + j).ravel()
    v01 = (i * (ny + 1) + j + 1).ravel()
    v11 = ((i + 1) * (ny + 1) + j + 1).ravel()
    tris = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], axis=0).astype(np.uint32)
    return np.ascontiguousarray(verts), np.ascontiguousarray(tris)


def make_soup(t, seed, side=100.0, jitter=0.5):
    rng = np.random.default_rng(seed)
    c = rng.random((t, 1, 3)) * side
    off = rng.uniform(-jitter, jitter, size=(t, 3, 3))
    verts = (c + off).reshape(-1, 3).astype(F32)
    tris = np.arange(3 * t, dtype=np.uint32).reshape(t, 3)
    return np.ascontiguousarray(verts), np.ascontiguousarray(tris)


def transform_rays(pose, origins, dirs):
    """Moves rays given in the mesh's local frame into world space with the mesh pose (t xyz, q ijkw): a posed cast of the moved
    rays answers the same question as the identity-pose cast of the original ones."""
    t, q = np.asarray(pose[:3], dtype=np.float64), np.asarray(pose[3:], dtype=np.float64)

    def rot(v):
        qv = q[:3]
        tt = 2 * np.cross(qv, v)
        return v + q[3] * tt + np.cross(qv, tt)

    o = (rot(np.asarray(origins, dtype=np.float64)) + t).astype(F32)
    d = rot(np.asarray(dirs, dtype=np.float64)).astype(F32)
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


def make_ray_scene(kind, n_tris, n_rays, seed=1004, random_pose=False):
    rng = np.random.default_rng(seed + 17)
    if kind == "terrain":
        ny = max(1, int(round((n_tris / 4.0) ** 0.5)))
        nx = max(1, n_tris // (2 * ny))
        verts, tris = make_terrain(nx, ny, seed)
        lo, hi = verts.min(0), verts.max(0)
        o = rng.random((n_rays, 3)) * (hi - lo) + lo
        o[:, 2] += 6.0
        d = rng.standard_normal((n_rays, 3))
        d[:, 2] = -np.abs(d[:, 2])
    else:
        side = 100.0 * (n_tris / 1_000_000.0) ** (1.0 / 3.0)
        verts, tris = make_soup(n_tris, seed, side=side)
        lo, hi = verts.min(0), verts.max(0)
        o = rng.random((n_rays, 3)) * (hi - lo) + lo
        d = rng.standard_normal((n_rays, 3))
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    pose = np.array([0, 0, 0, 0, 0, 0, 1], dtype=F32)
    if random_pose:
        q = random_unit_quaternions(rng, 1)[0]
        t = rng.uniform(-5, 5, size=3).astype(F32)
        pose = np.concatenate([t, q]).astype(F32)
    return RayScene(
        verts=verts,
        tris=tris,
        origins=np.ascontiguousarray(o.astype(F32)),
        dirs=np.ascontiguousarray(d.astype(F32)),
        pose=pose,
        name=f"cfg4_{kind}_{len(tris)}tris_{n_rays}rays",
    )


def make_polyline_scene(kind, n_edges, n_rays, seed=2004):
    """ncollide2d Polyline ray scenes: (points [p, 2], edges [m, 2] or None for the line strip, origins [r, 2], dirs [r, 2]).
    "terrain": one noisy height profile as a line strip, rays from above pointing down; "soup": short random segments scattered over a
    square (explicit edge list), rays in every direction; a share of the rays is axis-aligned (the slab test's dir == 0 branch)."""
    rng = np.random.default_rng(seed)
    if kind == "terrain":
        x = np.arange(n_edges + 1, dtype=np.float64) * 0.25
        y = 2.0 * np.sin(x * 0.05) + 0.7 * np.sin(x * 0.31 + 1.0) + 0.15 * rng.standard_normal(n_edges + 1)
        pts, edges = np.stack([x, y], axis=1), None
        o = np.stack([rng.uniform(x[0], x[-1], n_rays), rng.uniform(4.0, 9.0, n_rays)], axis=1)
        d = rng.standard_normal((n_rays, 2))
        d[:, 1] = -np.abs(d[:, 1])
    else:
        side = 40.0 * (n_edges / 10_000.0) ** 0.5
        a = rng.uniform(0, side, size=(n_edges, 2))
        b = a + rng.uniform(-0.6, 0.6, size=(n_edges, 2))
        flat = rng.random(n_edges) < 0.1  # axis-aligned segments: degenerate boxes, rays parallel to them
        b[flat, 1] = a[flat, 1]
        pts = np.concatenate([a, b])
        edges = np.stack([np.arange(n_edges), np.arange(n_edges) + n_edges], axis=1).astype(np.uint32)
        o = rng.uniform(0, side, size=(n_rays, 2))
        d = rng.standard_normal((n_rays, 2))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    axis = rng.random(n_rays) < 0.05
    d[axis] = np.where(rng.random((int(axis.sum()), 1)) < 0.5, [[0.0, -1.0]], [[1.0, 0.0]])
    return (np.ascontiguousarray(pts, dtype=F32), edges, np.ascontiguousarray(o, dtype=F32), np.ascontiguousarray(d, dtype=F32))
