"""Host mirror of the first ncollide2d slice on the device (``ncb2d_contact``, csrc/dim2.cu): ``ncollide2d::query::contact`` for
batches of 2-D shape pairs.  Shapes mirror ``Ball::new(radius)``, ``Cuboid::new(half_extents)`` and ``ConvexPolygon::try_new(points)``
of the 2-D crate; a pose is ``Isometry2::new(translation, angle)``: (x, y, cos(angle), sin(angle)) with cos / sin from the host libm,
like the reference's ``UnitComplex::new``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import as_f32, as_u32, ptr

BALL, CUBOID, POLYGON, PLANE, SEGMENT = 0, 1, 2, 3, 4
F32 = np.float32
EPS = np.finfo(np.float32).eps


def isometry2(translation, angle):
    """``Isometry2::new(translation, angle)`` rows (x, y, re, im); angle in radians (scalar or array)."""
    t = as_f32(translation).reshape(-1, 2)
    a = np.broadcast_to(np.asarray(angle, dtype=np.float32).reshape(-1), (len(t),))
    return np.ascontiguousarray(np.stack([t[:, 0], t[:, 1], np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)], axis=1), dtype=np.float32)


class Shapes2D:
    """A batch of 2-D shapes: ``type`` [n] and ``param`` [n, 4]; polygons index into a shared point array."""

    def __init__(self):
        self.type, self.param, self.points, self.normals = [], [], [], []

    def ball(self, radius):
        self.type.append(BALL), self.param.append((radius, 0, 0, 0))
        return self

    def cuboid(self, hx, hy):
        self.type.append(CUBOID), self.param.append((hx, hy, 0, 0))
        return self

    def plane(self, normal):
        """``Plane::new(Unit::new_normalize(normal))`` (a half-space in 2-D); the normal is normalised in f32."""
        v = as_f32(normal).reshape(2)
        nrm = np.sqrt(F32(F32(v[0] * v[0]) + F32(v[1] * v[1])), dtype=F32)
        self.type.append(PLANE), self.param.append((F32(v[0] / nrm), F32(v[1] / nrm), 0, 0))
        return self

    def segment(self, a, b):
        """``Segment::new(a, b)`` (shape/segment.rs): the two end points in the shape's frame, a != b."""
        a, b = as_f32(a).reshape(2), as_f32(b).reshape(2)
        if a[0] == b[0] and a[1] == b[1]:
            raise ValueError("a segment needs two different end points")
        self.type.append(SEGMENT), self.param.append((a[0], a[1], b[0], b[1]))
        return self

    def polygon(self, points):
        """``ConvexPolygon::try_new(points)`` (shape/convex_polygon.rs:29-71): vertices of a counter-clockwise convex polyline;
        edge normals are computed and vertices between (nearly) collinear edges are removed, in f32 like the reference.
        Raises ValueError where the reference returns None."""
        pts = [np.asarray(p, dtype=F32) for p in as_f32(points).reshape(-1, 2)]
        n = len(pts)
        eps = F32(np.sqrt(EPS))
        normals = []
        for i1 in range(n):
            ab = pts[(i1 + 1) % n] - pts[i1]
            res = np.array([ab[1], -ab[0]], dtype=F32)  # ccw_face_normal (dim2)
            sq = F32(F32(res[0] * res[0]) + F32(res[1] * res[1]))
            if not sq > F32(EPS * EPS):
                raise ValueError("ConvexPolygon::try_new: consecutive points are identical")
            normals.append((res / np.sqrt(sq, dtype=F32)).astype(F32))

        def dot(a, b):
            return F32(F32(a[0] * b[0]) + F32(a[1] * b[1]))

        removed = 1 if dot(normals[0], normals[-1]) > F32(1) - eps else 0
        for i2 in range(1, n):
            if dot(normals[i2 - 1], normals[i2]) > F32(1) - eps:
                removed += 1
            else:
                pts[i2 - removed] = pts[i2]
                normals[i2 - removed] = normals[i2]
        keep = n - removed
        if keep == 0:
            raise ValueError("ConvexPolygon::try_new: no vertex left")
        self.type.append(POLYGON), self.param.append((len(self.points), keep, 0, 0))
        self.points.extend(p.tolist() for p in pts[:keep])
        self.normals.extend(q.tolist() for q in normals[:keep])
        return self

    def arrays(self):
        """(type [n], param [n, 4], points [m, 2], normals [m, 2])"""
        return (as_u32(self.type), as_f32(self.param).reshape(-1, 4), as_f32(self.points if self.points else [[0, 0]]).reshape(-1, 2),
                as_f32(self.normals if self.normals else [[0, 0]]).reshape(-1, 2))


def contact(ctx, type1, param1, pose1, type2, param2, pose2, poly_points=None, prediction=0.0, poly_normals=None):
    """``query::contact`` for every pair k: (type1[k], param1[k]) at pose1[k] against (type2[k], param2[k]) at pose2[k].
    Returns (found [n] bool, contacts [n, 7] = world1, world2, normal, depth, info dict)."""
    t1, t2 = as_u32(type1).reshape(-1), as_u32(type2).reshape(-1)
    p1, p2 = as_f32(param1).reshape(-1, 4), as_f32(param2).reshape(-1, 4)
    m1, m2 = as_f32(pose1).reshape(-1, 4), as_f32(pose2).reshape(-1, 4)
    n = len(t1)
    if not (len(t2) == len(p1) == len(p2) == len(m1) == len(m2) == n):
        raise ValueError("one type / param / pose row per pair and side")
    pts = as_f32(poly_points).reshape(-1, 2) if poly_points is not None else None
    nrm = as_f32(poly_normals).reshape(-1, 2) if poly_normals is not None else None
    if nrm is not None and (pts is None or len(nrm) != len(pts)):
        raise ValueError("one normal per polygon point")
    found = np.zeros(n, dtype=np.uint8)
    out = np.zeros((n, 7), dtype=np.float32)
    panics, over = C.c_uint32(0), C.c_uint32(0)
    ctx.check(ctx.lib.ncb2d_contact(ctx.h, C.c_uint32(n), ptr(t1), ptr(p1), ptr(m1), ptr(t2), ptr(p2), ptr(m2), ptr(pts), ptr(nrm),
                                    C.c_uint32(0 if pts is None else len(pts)), C.c_float(prediction), ptr(found), ptr(out), C.byref(panics),
                                    C.byref(over)), "ncb2d_contact")
    return found.astype(bool), out, {"ref_panics": panics.value, "epa_overflow": over.value}


def proximity(ctx, type1, param1, pose1, type2, param2, pose2, poly_points=None, margins=0.0):
    """``ncollide2d::query::proximity`` per pair (``ncb2d_proximity``): 0 Intersecting, 1 WithinMargin, 2 Disjoint."""
    t1, t2 = as_u32(type1).reshape(-1), as_u32(type2).reshape(-1)
    p1, p2 = as_f32(param1).reshape(-1, 4), as_f32(param2).reshape(-1, 4)
    m1, m2 = as_f32(pose1).reshape(-1, 4), as_f32(pose2).reshape(-1, 4)
    n = len(t1)
    pts = as_f32(poly_points).reshape(-1, 2) if poly_points is not None else None
    mg = np.ascontiguousarray(np.broadcast_to(np.asarray(margins, dtype=np.float32).reshape(-1), (n,)), dtype=np.float32)
    out = np.zeros(n, dtype=np.uint8)
    ctx.check(ctx.lib.ncb2d_proximity(ctx.h, C.c_uint32(n), ptr(t1), ptr(p1), ptr(m1), ptr(t2), ptr(p2), ptr(m2), ptr(pts),
                                      C.c_uint32(0 if pts is None else len(pts)), ptr(mg), ptr(out)), "ncb2d_proximity")
    return out


def ray_cast(ctx, types, params, poses, rays, poly_points=None):
    """``RayCast::toi_and_normal_with_ray(m, ray, max_toi, solid = true)`` of shape k for ray k (``ncb2d_ray_cast``).
    rays [n, 5] = origin, dir, max_toi.  Returns (found, out [n, 3] = toi, normal, feature)."""
    t = as_u32(types).reshape(-1)
    n = len(t)
    p, m, q = as_f32(params).reshape(-1, 4), as_f32(poses).reshape(-1, 4), as_f32(rays).reshape(-1, 5)
    pts = as_f32(poly_points).reshape(-1, 2) if poly_points is not None else None
    found, out, feat = np.zeros(n, dtype=np.uint8), np.zeros((n, 3), dtype=np.float32), np.zeros(n, dtype=np.uint32)
    ctx.check(ctx.lib.ncb2d_ray_cast(ctx.h, C.c_uint32(n), ptr(t), ptr(p), ptr(m), ptr(pts), C.c_uint32(0 if pts is None else len(pts)), ptr(q),
                                     ptr(found), ptr(out), ptr(feat)), "ncb2d_ray_cast")
    return found.astype(bool), out, feat


def world_ray_cast(ctx, rays, groups=None, first_only=False):
    """``CollisionWorld::interferences_with_ray`` / ``first_interference_with_ray`` of the 2-D world of the last ``world_update``
    (``ncb2d_world_ray_cast``).  rays [n, 5] = origin, dir, max_toi; groups = (membership, whitelist, blacklist) or None.
    Returns (idx [k, 2] = (ray, handle), val [k, 3] = (toi, normal), feature [k]), sorted by (ray, handle)."""
    q = as_f32(rays).reshape(-1, 5)
    g = as_u32(groups).reshape(3) if groups is not None else None
    cap = max(4 * len(q), 1024)
    while True:
        idx, val, feat = np.zeros((cap, 2), dtype=np.uint32), np.zeros((cap, 3), dtype=np.float32), np.zeros(cap, dtype=np.uint32)
        n_out = C.c_uint32(0)
        r = ctx.check(ctx.lib.ncb2d_world_ray_cast(ctx.h, C.c_uint32(len(q)), ptr(q), ptr(g), C.c_int(1 if first_only else 0), ptr(idx), ptr(val),
                                                   ptr(feat), C.c_uint32(cap), C.byref(n_out)), "ncb2d_world_ray_cast")
        if r == 0:
            k = n_out.value
            return idx[:k], val[:k], feat[:k]
        cap = n_out.value + 1024


def world_query(ctx, kind, queries, groups=None):
    """``interferences_with_aabb`` (kind "aabb": rows of mins x y, maxs x y) / ``interferences_with_point`` (kind "point": x y) of the
    2-D world of the last ``world_update`` (``ncb2d_world_query``).  Returns idx [k, 2] = (query, handle), sorted."""
    k = {"aabb": 0, "point": 2}[kind]
    q = as_f32(queries).reshape(-1, 4 if k == 0 else 2)
    g = as_u32(groups).reshape(3) if groups is not None else None
    cap = max(8 * len(q), 1024)
    while True:
        idx = np.zeros((cap, 2), dtype=np.uint32)
        n_out = C.c_uint32(0)
        r = ctx.check(ctx.lib.ncb2d_world_query(ctx.h, C.c_int(k), C.c_uint32(len(q)), ptr(q), ptr(g), ptr(idx), C.c_uint32(cap), C.byref(n_out)),
                      "ncb2d_world_query")
        if r == 0:
            return idx[: n_out.value]
        cap = n_out.value + 1024


class Polyline:
    """``ncollide2d::shape::Polyline::new(points, indices)`` with ``RayCast::toi_and_normal_with_ray`` for a batch of rays
    (``ncb2d_polyline_create`` / ``ncb2d_polyline_ray_cast``).  ``edges`` None = the line strip."""

    def __init__(self, ctx, points, edges=None):
        self.ctx = ctx
        self.points = as_f32(points).reshape(-1, 2)
        self.edges = as_u32(edges).reshape(-1, 2) if edges is not None else None
        self.n_edges = len(self.edges) if self.edges is not None else max(len(self.points) - 1, 0)
        h = C.c_void_p()
        ctx.check(ctx.lib.ncb2d_polyline_create(ctx.h, C.c_uint32(len(self.points)), ptr(self.points), C.c_uint32(self.n_edges), ptr(self.edges),
                                                C.byref(h)), "ncb2d_polyline_create")
        self.h = h

    def toi_and_normal_with_ray(self, pose, origins, dirs, max_toi=None, want_normals=True, out=None):
        """pose: None or (x, y, re, im).  max_toi: None, one value, or one per ray.  Returns (toi [-1 = None], feature [edge, or
        edge + n_edges for the segment's Face(1)], normals [the scaled segment normal, as in the reference])."""
        o, d = as_f32(origins).reshape(-1, 2), as_f32(dirs).reshape(-1, 2)
        n = len(o)
        out = out or {}
        toi = out.get("toi") if out.get("toi") is not None else np.zeros(n, dtype=np.float32)
        feat = out.get("feature") if out.get("feature") is not None else np.zeros(n, dtype=np.uint32)
        normal = (out.get("normal") if out.get("normal") is not None else np.zeros((n, 2), dtype=np.float32)) if want_normals else None
        p = as_f32(pose) if pose is not None else None
        per_ray = None
        if max_toi is None:
            max_toi = np.finfo(np.float32).max
        elif np.ndim(max_toi) > 0:
            per_ray = as_f32(max_toi).reshape(-1)
            if len(per_ray) != n:
                raise ValueError("one max_toi per ray")
            max_toi = 0.0
        self.ctx.check(self.ctx.lib.ncb2d_polyline_ray_cast(self.h, ptr(p), C.c_uint32(n), ptr(o), ptr(d), C.c_float(max_toi), ptr(per_ray),
                                                            ptr(toi), ptr(feat), ptr(normal)), "ncb2d_polyline_ray_cast")
        return toi, feat, normal

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.ncb2d_polyline_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class World2D:
    """A fresh ``ncollide2d::CollisionWorld``: objects = (shape, Isometry2, CollisionGroups, GeometricQueryType::Contacts(linear, angular)).
    ``shapes`` is a Shapes2D batch (one entry per object); ``pos`` [n, 2], ``angle`` [n]."""

    def __init__(self, shapes: Shapes2D, pos, angle, margin=0.02, linear=0.02, angular=0.0, groups=None):
        self.type, self.param, self.points, self.normals = shapes.arrays()
        self.n = len(self.type)
        self.pos = as_f32(pos).reshape(-1, 2)
        a = np.broadcast_to(np.asarray(angle, dtype=np.float32).reshape(-1), (self.n,))
        self.rot = np.ascontiguousarray(np.stack([np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)], axis=1), dtype=np.float32)
        self.margin = float(margin)
        self.query_limit = np.ascontiguousarray(np.broadcast_to(np.asarray(linear, dtype=np.float32).reshape(-1), (self.n,)), dtype=np.float32)
        self.ang_pred = np.ascontiguousarray(np.broadcast_to(np.asarray(angular, dtype=np.float32).reshape(-1), (self.n,)), dtype=np.float32)
        self.groups = None if groups is None else as_u32(groups).reshape(-1, 3)
        self.query_kind = None  # or uint8 [n]: 1 = GeometricQueryType::Proximity(query_limit), a sensor (set_sensors)
        if len(self.pos) != self.n:
            raise ValueError("one position per shape")

    def set_sensors(self, mask):
        """Objects with mask != 0 become ``GeometricQueryType::Proximity(query_limit)``: pairs with one get a proximity status instead of a manifold."""
        self.query_kind = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
        if len(self.query_kind) != self.n:
            raise ValueError("one flag per object")
        return self


    @classmethod
    def from_library(cls, shapes: Shapes2D, pick, pos, angle, **kw):
        """A world whose object k is a copy of library shape ``pick[k]`` (large synthetic worlds without a Python loop per object)."""
        w = cls.__new__(cls)
        typ, par, w.points, w.normals = shapes.arrays()
        pick = np.asarray(pick, dtype=np.int64)
        w.type, w.param = np.ascontiguousarray(typ[pick]), np.ascontiguousarray(par[pick])
        w.n = len(pick)
        w.pos = as_f32(pos).reshape(-1, 2)
        a = np.broadcast_to(np.asarray(angle, dtype=np.float32).reshape(-1), (w.n,))
        w.rot = np.ascontiguousarray(np.stack([np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)], axis=1), dtype=np.float32)
        w.margin = float(kw.get("margin", 0.02))
        w.query_limit = np.full(w.n, kw.get("linear", 0.02), dtype=np.float32)
        w.ang_pred = np.full(w.n, kw.get("angular", 0.0), dtype=np.float32)
        w.groups = None
        w.query_kind = None
        return w


class _Objects2DC(C.Structure):
    _fields_ = [("n", C.c_uint32), ("pos", C.c_void_p), ("rot", C.c_void_p), ("shape_type", C.c_void_p), ("shape_param", C.c_void_p),
                ("groups", C.c_void_p), ("query_limit", C.c_void_p), ("ang_pred", C.c_void_p), ("poly_points", C.c_void_p),
                ("poly_normals", C.c_void_p), ("n_poly_points", C.c_uint32), ("query_kind", C.c_void_p)]


def world_update(ctx, w: World2D, bufs=None):
    """``CollisionWorld::update`` of a fresh 2-D world (``ncb2d_world_update``).  Returns a dict: pairs [P, 2] (object1 = larger handle),
    manifold_start / manifold_count [P], contacts [C, 7] (world1, world2, normal, depth), features [C, 2], diag.
    bufs: optional dict of preallocated (e.g. page-locked) result arrays "pairs" [cap, 2] u32, "start" [cap] u32, "count" [cap] u8,
    "contacts" [cap_c, 7] f32, "features" [cap_c, 2] u32 — used as they are when large enough."""
    o = _Objects2DC()
    o.n = w.n
    o.pos, o.rot, o.shape_type, o.shape_param = (a.ctypes.data for a in (w.pos, w.rot, w.type, w.param))
    o.groups = w.groups.ctypes.data if w.groups is not None else None
    o.query_limit, o.ang_pred = w.query_limit.ctypes.data, w.ang_pred.ctypes.data
    o.poly_points, o.poly_normals, o.n_poly_points = w.points.ctypes.data, w.normals.ctypes.data, len(w.points)
    o.query_kind = w.query_kind.ctypes.data if getattr(w, "query_kind", None) is not None else None
    cap_p, cap_c = max(8 * w.n, 1024), max(8 * w.n, 1024)
    while True:
        if bufs is not None:
            pairs, start, count, contacts, feats = (bufs[k] for k in ("pairs", "start", "count", "contacts", "features"))
            cap_p, cap_c = min(len(pairs), len(start), len(count)), min(len(contacts), len(feats))
            bufs = None  # a retry (too small) falls back to fresh arrays
        else:
            pairs = np.zeros((cap_p, 2), dtype=np.uint32)
            start, count = np.zeros(cap_p, dtype=np.uint32), np.zeros(cap_p, dtype=np.uint8)
            contacts, feats = np.zeros((cap_c, 7), dtype=np.float32), np.zeros((cap_c, 2), dtype=np.uint32)
        npairs, ncont = C.c_uint32(0), C.c_uint32(0)
        diag = np.zeros(4, dtype=np.uint32)
        r = ctx.check(ctx.lib.ncb2d_world_update(ctx.h, C.byref(o), C.c_float(w.margin), ptr(pairs), C.c_uint32(cap_p), ptr(start), ptr(count),
                                                 ptr(contacts), ptr(feats), C.c_uint32(cap_c), C.byref(npairs), C.byref(ncont), ptr(diag)),
                      "ncb2d_world_update")
        if r == 0:
            P, Cn = npairs.value, ncont.value
            prox = None
            if o.query_kind:
                prox = np.zeros(P, dtype=np.uint8)
                ctx.check(ctx.lib.ncb2d_world_fetch_proximity(ctx.h, ptr(prox), C.c_uint32(P)), "ncb2d_world_fetch_proximity")
            return {"proximity": prox, "pairs": pairs[:P], "manifold_start": start[:P], "manifold_count": count[:P], "contacts": contacts[:Cn], "features": feats[:Cn],
                    "diag": dict(zip(("ref_panics", "epa_overflow", "manifold_overflow", "stack_overflow"), diag.tolist()))}
        cap_p, cap_c = max(cap_p, npairs.value + 1024), max(cap_c, ncont.value + 1024)
