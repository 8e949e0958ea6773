"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_build/liboracle_f{32,64}.so.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
The product package (ncollide_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    """Compile the oracle (g++, seconds).  Building the checker is not using it."""
    out = os.path.join(_HERE, "_build", "liboracle_f32.so")
    if force or not os.path.exists(out) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(out)
        for f in os.listdir(_HERE)
        if f.endswith((".cpp", ".hpp", ".h"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return out


class _HullLib(C.Structure):
    _fields_ = [("n_hulls", C.c_uint32)] + [
        (n, C.c_void_p)
        for n in (
            "vert_off face_off edge_off fadj_off vadj_off points vert_first_adj vert_num_adj face_first face_num "
            "face_normal vertices_adj_to_face edges_adj_to_face edge_vertices edge_faces edge_dir "
            "faces_adj_to_vertex edges_adj_to_vertex"
        ).split()
    ]


class _Objects(C.Structure):
    _fields_ = [
        ("n", C.c_uint32),
        ("pos", C.c_void_p),
        ("rot", C.c_void_p),
        ("shape_type", C.c_void_p),
        ("shape_param", C.c_void_p),
        ("groups", C.c_void_p),
        ("query_limit", C.c_void_p),
        ("ang_pred", C.c_void_p),
        ("hulls", C.c_void_p),
        ("query_kind", C.c_void_p),
    ]


FLOAT_HULL_FIELDS = {"points", "face_normal", "edge_dir"}


class Oracle:
    def __init__(self, dtype=np.float32):
        build()
        self.dtype = np.dtype(dtype)
        name = "liboracle_f32.so" if self.dtype == np.float32 else "liboracle_f64.so"
        self.lib = C.CDLL(os.path.join(_HERE, "_build", name))
        self.creal = C.c_float if self.dtype == np.float32 else C.c_double
        L = self.lib
        L.orc_broad_phase.restype = C.c_uint64
        L.orc_narrow_phase.restype = C.c_uint64
        L.orc_trimesh_create.restype = C.c_void_p
        L.orc_query_contact.restype = C.c_int
        self.contact_dtype = np.dtype(
            [("world1", self.dtype, 3), ("world2", self.dtype, 3), ("normal", self.dtype, 3), ("depth", self.dtype), ("f1", np.uint32), ("f2", np.uint32)],
            align=True,
        )

    # -- marshalling ---------------------------------------------------------------------------
    def _objects(self, scene):
        keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data

        hl = _HullLib()
        lib = scene.hulls
        hl.n_hulls = lib.n_hulls
        for f in lib.FIELDS:
            setattr(hl, f, arr(getattr(lib, f), self.dtype if f in FLOAT_HULL_FIELDS else np.uint32))
        keep.append(hl)
        o = _Objects()
        o.n = scene.n
        o.pos = arr(scene.pos, self.dtype)
        o.rot = arr(scene.rot, self.dtype)
        o.shape_type = arr(scene.shape_type, np.uint32)
        o.shape_param = arr(scene.shape_param, self.dtype)
        o.groups = arr(scene.groups, np.uint32) if scene.groups is not None else None
        o.query_limit = arr(scene.query_limit, self.dtype)
        o.ang_pred = arr(scene.ang_pred, self.dtype)
        o.hulls = C.addressof(hl)
        qk = getattr(scene, "query_kind", None)
        o.query_kind = arr(qk, np.uint8) if qk is not None else None
        return o, keep

    # -- API -----------------------------------------------------------------------------------
    def compute_aabbs(self, scene, fat=True, mode=None):
        """mode 0: shape AABB, 1: + query_limit, 2: + margin (fat=True -> 2, fat=False -> 0)."""
        o, keep = self._objects(scene)
        out = np.zeros((scene.n, 6), dtype=self.dtype)
        if mode is None:
            mode = 2 if fat else 0
        self.lib.orc_compute_aabbs(C.byref(o), self.creal(scene.margin), C.c_int(mode), C.c_void_p(out.ctypes.data))
        return out

    def broad_phase(self, aabbs, groups=None, mode=1):
        """Returns [P,2] u32 pairs (larger handle, smaller handle). mode: 0 DBVT-faithful, 1 sweep, 2 brute force."""
        aabbs = np.ascontiguousarray(aabbs, dtype=self.dtype)
        n = len(aabbs)
        g = np.ascontiguousarray(groups, dtype=np.uint32) if groups is not None else None
        gp = C.c_void_p(g.ctypes.data) if g is not None else None
        cap = max(16 * n, 1024)
        while True:
            out = np.zeros((cap, 2), dtype=np.uint32)
            np_ = self.lib.orc_broad_phase(C.c_uint32(n), C.c_void_p(aabbs.ctypes.data), gp, C.c_int(mode), C.c_void_p(out.ctypes.data), C.c_uint64(cap))
            if np_ <= cap:
                return out[:np_].copy()
            cap = int(np_)

    def narrow_phase(self, scene, pairs):
        """pairs: [P,2] (object1, object2).  Returns (contacts structured array, manifold_off[P+1], algo[P], stats[8])."""
        o, keep = self._objects(scene)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        P = len(pairs)
        cap = max(4 * P, 64)
        off = np.zeros(P + 1, dtype=np.uint32)
        algo = np.zeros(P, dtype=np.uint8)
        stats = np.zeros(8, dtype=np.uint32)
        while True:
            out = np.zeros(cap, dtype=self.contact_dtype)
            nc = self.lib.orc_narrow_phase(
                C.byref(o), C.c_uint64(P), C.c_void_p(pairs.ctypes.data), C.c_void_p(out.ctypes.data), C.c_uint64(cap),
                C.c_void_p(off.ctypes.data), C.c_void_p(algo.ctypes.data), C.c_void_p(stats.ctypes.data),
            )
            if nc <= cap:
                return out[:nc].copy(), off, algo, stats
            cap = int(nc)

    def narrow_phase_kinematics(self, scene, pairs):
        """narrow_phase + the ContactKinematic of every contact (local1, local2, NeighborhoodGeometry kind + direction per side, dilations).
        Returns (contacts, kinematics, manifold_off, algo)."""
        o, keep = self._objects(scene)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        P = len(pairs)
        cap = max(4 * P, 64)
        off = np.zeros(P + 1, dtype=np.uint32)
        algo = np.zeros(P, dtype=np.uint8)
        kd = self.kinematic_dtype()
        self.lib.orc_narrow_phase_kin.restype = C.c_uint64
        while True:
            out = np.zeros(cap, dtype=self.contact_dtype)
            kin = np.zeros(cap, dtype=kd)
            nc = self.lib.orc_narrow_phase_kin(
                C.byref(o), C.c_uint64(P), C.c_void_p(pairs.ctypes.data), C.c_void_p(out.ctypes.data), C.c_void_p(kin.ctypes.data), C.c_uint64(cap),
                C.c_void_p(off.ctypes.data), C.c_void_p(algo.ctypes.data), None,
            )
            if nc <= cap:
                return out[:nc].copy(), kin[:nc].copy(), off, algo
            cap = int(nc)

    def kinematic_dtype(self):
        r = self.dtype
        return np.dtype([("local1", r, 3), ("local2", r, 3), ("dir1", r, 3), ("dir2", r, 3), ("dil1", r), ("dil2", r), ("g1", np.uint32), ("g2", np.uint32)])

    def contact_sm_sm(self, scene, pairs, predictions=None):
        """contact_support_map_support_map for cuboid / hull pairs -> (out[P,10] = p1, p2, normal, found; stats[4])."""
        o, keep = self._objects(scene)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        out = np.zeros((len(pairs), 10), dtype=self.dtype)
        stats = np.zeros(4, dtype=np.uint32)
        m = None if predictions is None else np.ascontiguousarray(predictions, dtype=self.dtype)
        self.lib.orc_contact_sm_sm(C.byref(o), C.c_uint64(len(pairs)), C.c_void_p(pairs.ctypes.data), C.c_void_p(m.ctypes.data) if m is not None else None,
                                   C.c_void_p(out.ctypes.data), C.c_void_p(stats.ctypes.data))
        return out, stats

    def kat_cylinder_cuboid(self, half_height, radius, t1, he, t2, margin, prediction):
        """(distance, proximity status, contact found) of the reference's cylinder_cuboid_contact test."""
        a = [np.ascontiguousarray(x, dtype=self.dtype) for x in (t1, he, t2)]
        out = np.zeros(3, dtype=self.dtype)
        self.lib.orc_kat_cylinder_cuboid(self.creal(half_height), self.creal(radius), C.c_void_p(a[0].ctypes.data), C.c_void_p(a[1].ctypes.data),
                                         C.c_void_p(a[2].ctypes.data), self.creal(margin), self.creal(prediction), C.c_void_p(out.ctypes.data))
        return float(out[0]), int(out[1]), bool(out[2])

    def proximity(self, scene, pairs, margins=None):
        """ProximityDetector::update with fresh detectors per (object1, object2) pair -> u8 status (0 Intersecting, 1 WithinMargin,
        2 Disjoint, 255 no detector).  margins None: query_limit[o1] + query_limit[o2]."""
        o, keep = self._objects(scene)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        out = np.zeros(len(pairs), dtype=np.uint8)
        m = None if margins is None else np.ascontiguousarray(margins, dtype=self.dtype)
        self.lib.orc_proximity(C.byref(o), C.c_uint64(len(pairs)), C.c_void_p(pairs.ctypes.data), C.c_void_p(m.ctypes.data) if m is not None else None,
                               C.c_void_p(out.ctypes.data))
        return out

    def proximity_warm(self, scene, pairs, margins, axis_io):
        """proximity() with the detectors' sep_axis carried in axis_io[P,4] (xyz + Some flag), updated in place."""
        o, keep = self._objects(scene)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        out = np.zeros(len(pairs), dtype=np.uint8)
        m = None if margins is None else np.ascontiguousarray(margins, dtype=self.dtype)
        assert axis_io.dtype == self.dtype and axis_io.flags.c_contiguous and axis_io.shape == (len(pairs), 4)
        self.lib.orc_proximity_warm(C.byref(o), C.c_uint64(len(pairs)), C.c_void_p(pairs.ctypes.data), C.c_void_p(m.ctypes.data) if m is not None else None,
                                    C.c_void_p(axis_io.ctypes.data), C.c_void_p(out.ctypes.data))
        return out

    def query_proximity(self, scene, margin):
        """query::proximity between objects 0 and 1."""
        o, keep = self._objects(scene)
        return int(self.lib.orc_query_proximity(C.byref(o), self.creal(margin)))

    def narrow_phase_kinds(self, scene, pairs):
        """Narrow phase honouring scene.query_kind.  Returns (contacts, manifold_off[P+1], algo[P], prox[P])."""
        o, keep = self._objects(scene)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        P = len(pairs)
        cap = max(4 * P, 64)
        off = np.zeros(P + 1, dtype=np.uint32)
        algo = np.zeros(P, dtype=np.uint8)
        prox = np.zeros(P, dtype=np.uint8)
        self.lib.orc_narrow_phase_kinds.restype = C.c_uint64
        while True:
            out = np.zeros(cap, dtype=self.contact_dtype)
            nc = self.lib.orc_narrow_phase_kinds(
                C.byref(o), C.c_uint64(P), C.c_void_p(pairs.ctypes.data), C.c_void_p(out.ctypes.data), C.c_uint64(cap),
                C.c_void_p(off.ctypes.data), C.c_void_p(algo.ctypes.data), C.c_void_p(prox.ctypes.data),
            )
            if nc <= cap:
                return out[:nc].copy(), off, algo, prox
            cap = int(nc)

    def world_update_timed(self, scene):
        o, keep = self._objects(scene)
        times = (C.c_double * 3)()
        counts = (C.c_uint64 * 3)()
        self.lib.orc_world_update_timed(C.byref(o), self.creal(scene.margin), times, counts)
        return list(times), list(counts)

    def query_contact(self, scene, prediction):
        o, keep = self._objects(scene)
        out = np.zeros(1, dtype=self.contact_dtype)
        r = self.lib.orc_query_contact(C.byref(o), self.creal(prediction), C.c_void_p(out.ctypes.data))
        return out[0] if r else None

    def trimesh(self, verts, tris):
        return OracleTriMesh(self, verts, tris)

    def ray_cast2d(self, types, params, poses, rays, poly_points=None):
        """ncollide2d ``RayCast::toi_and_normal_with_ray`` (solid) of shape k for ray k [origin, dir, max_toi]:
        (found u8, out [toi, normal], feature)."""
        dt = self.dtype
        t = np.ascontiguousarray(types, dtype=np.uint32)
        n = len(t)
        p, m, q = (np.ascontiguousarray(a, dtype=dt) for a in (params, poses, rays))
        pts = np.ascontiguousarray(poly_points if poly_points is not None else np.zeros((1, 2)), dtype=dt)
        found, out, feat = np.zeros(n, dtype=np.uint8), np.zeros((n, 3), dtype=dt), np.zeros(n, dtype=np.uint32)
        vp = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        self.lib.orc2_ray_cast(C.c_uint64(n), vp(t), vp(p), vp(m), vp(pts), vp(q), vp(found), vp(out), vp(feat))
        return found, out, feat

    def polyline(self, points, edges=None):
        return OraclePolyline(self, points, edges)

    def segment_ray_cast(self, a, b, origin, direction, pose=None):
        """ncollide2d ``Segment::toi_and_normal_with_ray``: None or (toi, normal, (feature kind [1 Face, 2 Vertex], id))."""
        dt = self.dtype
        ab = np.ascontiguousarray(list(a) + list(b), dtype=dt)
        o, d = np.ascontiguousarray(origin, dtype=dt), np.ascontiguousarray(direction, dtype=dt)
        p = np.ascontiguousarray(pose, dtype=dt) if pose is not None else None
        toi, normal, feat = np.zeros(1, dtype=dt), np.zeros(2, dtype=dt), C.c_uint32(0)
        vp = lambda x: C.c_void_p(x.ctypes.data) if x is not None else None  # noqa: E731
        if not self.lib.orc2_segment_ray_cast(vp(ab), vp(p), vp(o), vp(d), vp(toi), vp(normal), C.byref(feat)):
            return None
        return toi[0], normal, (feat.value >> 30, feat.value & 0x3FFFFFFF)

    def contact2d(self, type1, param1, pose1, type2, param2, pose2, poly_points=None, prediction=0.0, poly_normals=None):
        """ncollide2d ``query::contact`` for n pairs (oracle/dim2.cpp).  Returns (found u8 [1 Some, 0 None, 2 not restated],
        out[n,7] = world1, world2, normal, depth, panics)."""
        dt = self.dtype
        t1, t2 = np.ascontiguousarray(type1, dtype=np.uint32), np.ascontiguousarray(type2, dtype=np.uint32)
        p1, p2 = np.ascontiguousarray(param1, dtype=dt).reshape(-1, 4), np.ascontiguousarray(param2, dtype=dt).reshape(-1, 4)
        m1, m2 = np.ascontiguousarray(pose1, dtype=dt).reshape(-1, 4), np.ascontiguousarray(pose2, dtype=dt).reshape(-1, 4)
        pts = np.ascontiguousarray(poly_points if poly_points is not None else np.zeros((1, 2)), dtype=dt).reshape(-1, 2)
        n = len(t1)
        found = np.zeros(n, dtype=np.uint8)
        out = np.zeros((n, 7), dtype=dt)
        panics = C.c_uint32(0)
        vp = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        nrm = np.ascontiguousarray(poly_normals, dtype=dt).reshape(-1, 2) if poly_normals is not None else None
        self.lib.orc2_contact(C.c_uint64(n), vp(t1), vp(p1), vp(m1), vp(t2), vp(p2), vp(m2), vp(pts), vp(nrm) if nrm is not None else None,
                              self.creal(prediction), vp(found), vp(out), C.byref(panics))
        return found, out, panics.value

    def proximity2d(self, type1, param1, pose1, type2, param2, pose2, poly_points, margins):
        """ncollide2d ``query::proximity`` per pair: 0 Intersecting, 1 WithinMargin, 2 Disjoint."""
        dt = self.dtype
        t1, t2 = np.ascontiguousarray(type1, dtype=np.uint32), np.ascontiguousarray(type2, dtype=np.uint32)
        arrs = [np.ascontiguousarray(a, dtype=dt) for a in (param1, pose1, param2, pose2, poly_points if poly_points is not None else np.zeros((1, 2)))]
        n = len(t1)
        mg = np.ascontiguousarray(np.broadcast_to(np.asarray(margins, dtype=dt).reshape(-1), (n,)), dtype=dt)
        out = np.zeros(n, dtype=np.uint8)
        vp = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        self.lib.orc2_proximity(C.c_uint64(n), vp(t1), vp(arrs[0]), vp(arrs[1]), vp(t2), vp(arrs[2]), vp(arrs[3]), vp(arrs[4]), vp(mg), vp(out))
        return out

    def world_update2d(self, w):
        """ncollide2d fresh-world update for a ncollide_b200.dim2.World2D: (pairs [P,2] canonical, offsets [P+1], contacts [C,7], features [C,2],
        panics).  Fat AABBs (oracle/dim2.cpp) -> the 3-D oracle's broad phase on boxes with z = 0 -> the 2-D generators."""
        dt = self.dtype

        class O2(C.Structure):
            _fields_ = [("n", C.c_uint32), ("pos", C.c_void_p), ("rot", C.c_void_p), ("type", C.c_void_p), ("param", C.c_void_p),
                        ("query_limit", C.c_void_p), ("ang_pred", C.c_void_p), ("poly_points", C.c_void_p), ("poly_normals", C.c_void_p),
                        ("query_kind", C.c_void_p)]

        keep = [np.ascontiguousarray(a, dtype=dt) for a in (w.pos, w.rot, w.param, w.query_limit, w.ang_pred, w.points, w.normals)]
        typ = np.ascontiguousarray(w.type, dtype=np.uint32)
        qk = np.ascontiguousarray(w.query_kind, dtype=np.uint8) if getattr(w, "query_kind", None) is not None else None
        o = O2(w.n, keep[0].ctypes.data, keep[1].ctypes.data, typ.ctypes.data, keep[2].ctypes.data, keep[3].ctypes.data, keep[4].ctypes.data,
               keep[5].ctypes.data, keep[6].ctypes.data, qk.ctypes.data if qk is not None else None)
        fat = np.zeros((w.n, 6), dtype=dt)
        self.lib.orc2_compute_aabbs(C.byref(o), self.creal(w.margin), C.c_void_p(fat.ctypes.data))
        pairs = self.broad_phase(fat, w.groups, mode=1)
        P = len(pairs)
        cap = max(4 * P, 64)
        off = np.zeros(P + 1, dtype=np.uint32)
        contacts = np.zeros((cap, 7), dtype=dt)
        feats = np.zeros((cap, 2), dtype=np.uint32)
        panics = C.c_uint32(0)
        self.lib.orc2_narrow_phase.restype = C.c_uint64
        pr = np.ascontiguousarray(pairs, dtype=np.uint32)
        prox = np.full(P, 255, dtype=np.uint8)
        nc = self.lib.orc2_narrow_phase(C.byref(o), C.c_uint64(P), C.c_void_p(pr.ctypes.data), C.c_void_p(off.ctypes.data),
                                        C.c_void_p(contacts.ctypes.data), C.c_void_p(feats.ctypes.data), C.c_uint64(cap), C.byref(panics),
                                        C.c_void_p(prox.ctypes.data))
        assert nc <= cap
        self.last_proximity2d = prox  # per pair: 0 / 1 / 2 for pairs with a sensor, 255 otherwise
        return pairs, off, contacts[:nc], feats[:nc], panics.value, fat

    def world_ray_cast2d(self, w, rays, groups=None, first_only=False):
        """ncollide2d ``interferences_with_ray`` / ``first_interference_with_ray`` over a dim2.World2D (after a fresh update): rows
        (idx [k, 2] = (ray, handle), val [k, 3] = (toi, normal), feature [k]) in (ray, handle) order."""
        dt = self.dtype

        class O2(C.Structure):
            _fields_ = [("n", C.c_uint32), ("pos", C.c_void_p), ("rot", C.c_void_p), ("type", C.c_void_p), ("param", C.c_void_p),
                        ("query_limit", C.c_void_p), ("ang_pred", C.c_void_p), ("poly_points", C.c_void_p), ("poly_normals", C.c_void_p),
                        ("query_kind", C.c_void_p)]

        keep = [np.ascontiguousarray(a, dtype=dt) for a in (w.pos, w.rot, w.param, w.query_limit, w.ang_pred, w.points, w.normals)]
        typ = np.ascontiguousarray(w.type, dtype=np.uint32)
        o = O2(w.n, keep[0].ctypes.data, keep[1].ctypes.data, typ.ctypes.data, keep[2].ctypes.data, keep[3].ctypes.data, keep[4].ctypes.data,
               keep[5].ctypes.data, keep[6].ctypes.data, None)
        fat = np.zeros((w.n, 6), dtype=dt)
        self.lib.orc2_compute_aabbs(C.byref(o), self.creal(w.margin), C.c_void_p(fat.ctypes.data))
        q = np.ascontiguousarray(rays, dtype=dt).reshape(-1, 5)
        og = np.ascontiguousarray(w.groups, dtype=np.uint32) if w.groups is not None else None
        g = np.ascontiguousarray(groups, dtype=np.uint32) if groups is not None else None
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None  # noqa: E731
        cap = max(64 * len(q), 1024)
        idx, val, feat = np.zeros((cap, 2), dtype=np.uint32), np.zeros((cap, 3), dtype=dt), np.zeros(cap, dtype=np.uint32)
        self.lib.orc2_world_ray_cast.restype = C.c_uint64
        k = self.lib.orc2_world_ray_cast(C.byref(o), vp(fat), vp(og), C.c_uint64(len(q)), vp(q), vp(g), C.c_int(1 if first_only else 0), vp(idx),
                                         vp(val), vp(feat), C.c_uint64(cap))
        assert k <= cap
        return idx[:k], val[:k], feat[:k]

    def contains_point2d(self, types, params, poses, pts, poly_points=None):
        """ncollide2d ``PointQuery::contains_point`` of shape k for point k -> bool [n]."""
        dt = self.dtype
        t = np.ascontiguousarray(types, dtype=np.uint32)
        p, m, q = (np.ascontiguousarray(a, dtype=dt) for a in (params, poses, pts))
        pp = np.ascontiguousarray(poly_points if poly_points is not None else np.zeros((1, 2)), dtype=dt)
        out = np.zeros(len(t), dtype=np.uint8)
        vp = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        self.lib.orc2_contains_point(C.c_uint64(len(t)), vp(t), vp(p), vp(m), vp(pp), vp(q), vp(out))
        return out.astype(bool)

    def world_query2d(self, w, kind, queries, groups=None):
        """ncollide2d ``interferences_with_aabb`` ("aabb") / ``interferences_with_point`` ("point") over a dim2.World2D: idx [k, 2]."""
        dt = self.dtype

        class O2(C.Structure):
            _fields_ = [("n", C.c_uint32), ("pos", C.c_void_p), ("rot", C.c_void_p), ("type", C.c_void_p), ("param", C.c_void_p),
                        ("query_limit", C.c_void_p), ("ang_pred", C.c_void_p), ("poly_points", C.c_void_p), ("poly_normals", C.c_void_p),
                        ("query_kind", C.c_void_p)]

        keep = [np.ascontiguousarray(a, dtype=dt) for a in (w.pos, w.rot, w.param, w.query_limit, w.ang_pred, w.points, w.normals)]
        typ = np.ascontiguousarray(w.type, dtype=np.uint32)
        o = O2(w.n, keep[0].ctypes.data, keep[1].ctypes.data, typ.ctypes.data, keep[2].ctypes.data, keep[3].ctypes.data, keep[4].ctypes.data,
               keep[5].ctypes.data, keep[6].ctypes.data, None)
        fat = np.zeros((w.n, 6), dtype=dt)
        self.lib.orc2_compute_aabbs(C.byref(o), self.creal(w.margin), C.c_void_p(fat.ctypes.data))
        k = {"aabb": 0, "point": 2}[kind]
        q = np.ascontiguousarray(queries, dtype=dt).reshape(-1, 4 if k == 0 else 2)
        og = np.ascontiguousarray(w.groups, dtype=np.uint32) if w.groups is not None else None
        g = np.ascontiguousarray(groups, dtype=np.uint32) if groups is not None else None
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None  # noqa: E731
        cap = max(64 * len(q), 1024)
        idx = np.zeros((cap, 2), dtype=np.uint32)
        self.lib.orc2_world_query.restype = C.c_uint64
        n = self.lib.orc2_world_query(C.byref(o), vp(fat), vp(og), C.c_int(k), C.c_uint64(len(q)), vp(q), vp(g), vp(idx), C.c_uint64(cap))
        assert n <= cap
        return idx[:n]

    def broad_phase_persistent(self, margin):
        return OracleBroadPhase(self, margin)

    def sim(self, scene):
        return OracleSim(self, scene)

    def edges(self, pairs):
        return OracleEdges(self, pairs)

    def shape_ray_cast(self, scene, i, origin, direction, max_toi):
        """RayCast::toi_and_normal_with_ray(position_i, ray, max_toi, solid = true) -> (toi, normal, feature) or None."""
        o, keep = self._objects(scene)
        ro = np.ascontiguousarray(origin, dtype=self.dtype)
        rd = np.ascontiguousarray(direction, dtype=self.dtype)
        out = np.zeros(4, dtype=self.dtype)
        feat = C.c_uint32()
        r = self.lib.orc_shape_ray_cast(C.byref(o), C.c_uint32(i), C.c_void_p(ro.ctypes.data), C.c_void_p(rd.ctypes.data), self.creal(max_toi),
                                        C.c_void_p(out.ctypes.data), C.byref(feat))
        return (out[0], out[1:4].copy(), feat.value) if r else None

    def shape_ray_cast_batch(self, scene, which, rays):
        """rays[K,7] (origin, dir, max_toi) against objects which[K] -> (hit[K] u8, out[K,4] = toi + normal, feature[K])."""
        o, keep = self._objects(scene)
        which = np.ascontiguousarray(which, dtype=np.uint32)
        rays = np.ascontiguousarray(rays, dtype=self.dtype).reshape(-1, 7)
        out = np.zeros((len(which), 4), dtype=self.dtype)
        feat = np.zeros(len(which), dtype=np.uint32)
        hit = np.zeros(len(which), dtype=np.uint8)
        self.lib.orc_shape_ray_cast_batch(C.byref(o), C.c_uint64(len(which)), C.c_void_p(which.ctypes.data), C.c_void_p(rays.ctypes.data),
                                          C.c_void_p(out.ctypes.data), C.c_void_p(feat.ctypes.data), C.c_void_p(hit.ctypes.data))
        return hit, out, feat

    def shape_contains_point_batch(self, scene, which, pts):
        o, keep = self._objects(scene)
        which = np.ascontiguousarray(which, dtype=np.uint32)
        pts = np.ascontiguousarray(pts, dtype=self.dtype).reshape(-1, 3)
        inside = np.zeros(len(which), dtype=np.uint8)
        self.lib.orc_shape_contains_point_batch(C.byref(o), C.c_uint64(len(which)), C.c_void_p(which.ctypes.data), C.c_void_p(pts.ctypes.data),
                                                C.c_void_p(inside.ctypes.data))
        return inside

    def aabb_toi_with_ray(self, minmax, origin, direction, max_toi, solid):
        mm = np.ascontiguousarray(minmax, dtype=self.dtype)
        o = np.ascontiguousarray(origin, dtype=self.dtype)
        d = np.ascontiguousarray(direction, dtype=self.dtype)
        t = self.creal(0)
        self.lib.orc_aabb_toi_with_ray(C.c_void_p(mm.ctypes.data), C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data), self.creal(max_toi), C.c_int(int(solid)), C.byref(t))
        return None if t.value < 0 else t.value


class OracleEdges:
    """A caller-driven set of contact edges (orc_edges_*): update_contact on chosen edges, state kept per edge."""

    def __init__(self, oracle, pairs):
        self.o = oracle
        L = oracle.lib
        L.orc_edges_create.restype = C.c_void_p
        L.orc_edges_update.restype = C.c_uint64
        L.orc_edges_fetch.restype = C.c_uint64
        self.pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        self.h = C.c_void_p(L.orc_edges_create(C.c_uint64(len(self.pairs)), C.c_void_p(self.pairs.ctypes.data)))

    def __del__(self):
        try:
            self.o.lib.orc_edges_destroy(self.h)
        except Exception:
            pass

    def update(self, scene, which):
        o, keep = self.o._objects(scene)
        which = np.ascontiguousarray(which, dtype=np.uint32)
        ev = np.zeros((max(len(which), 1), 3), dtype=np.uint32)
        ne = self.o.lib.orc_edges_update(self.h, C.byref(o), C.c_uint64(len(which)), C.c_void_p(which.ctypes.data), C.c_void_p(ev.ctypes.data),
                                         C.c_uint64(len(ev)))
        return ev[:ne]

    def fetch(self):
        P = len(self.pairs)
        off = np.zeros(P + 1, dtype=np.uint32)
        dirs = np.zeros((P, 4), dtype=self.o.dtype)
        cap = max(8 * P, 64)
        while True:
            c = np.zeros(cap, dtype=self.o.contact_dtype)
            ids = np.zeros(cap, dtype=np.uint32)
            nc = self.o.lib.orc_edges_fetch(self.h, C.c_void_p(off.ctypes.data), C.c_void_p(c.ctypes.data), C.c_void_p(ids.ctypes.data), C.c_uint64(cap),
                                            C.c_void_p(dirs.ctypes.data))
            if nc <= cap:
                return c[:nc], off, ids[:nc], dirs
            cap = int(nc)


class OracleSim:
    """Reference-faithful stepping CollisionWorld (oracle/narrow.cpp, orc_sim_*)."""

    def __init__(self, oracle, scene):
        self.o = oracle
        L = oracle.lib
        L.orc_sim_create.restype = C.c_void_p
        for f in ("orc_sim_num_pairs", "orc_sim_num_contacts", "orc_sim_events", "orc_sim_bp_num_interferences", "orc_sim_proximity_events"):
            getattr(L, f).restype = C.c_uint64
        self._objs, self._keep = oracle._objects(scene)
        self.h = C.c_void_p(L.orc_sim_create(C.byref(self._objs), oracle.creal(scene.margin)))

    def set_positions(self, handles, pos, rot):
        hs = None if handles is None else np.ascontiguousarray(handles, dtype=np.uint32)
        p = np.ascontiguousarray(pos, dtype=self.o.dtype)
        r = np.ascontiguousarray(rot, dtype=self.o.dtype)
        self.o.lib.orc_sim_set_positions(self.h, C.c_uint32(len(p)), C.c_void_p(hs.ctypes.data) if hs is not None else None,
                                         C.c_void_p(p.ctypes.data), C.c_void_p(r.ctypes.data))

    def set_collision_groups(self, handles, groups):
        hs = np.ascontiguousarray(handles, dtype=np.uint32).reshape(-1)
        g = np.ascontiguousarray(groups, dtype=np.uint32).reshape(-1, 3)
        assert len(g) == len(hs)
        r = self.o.lib.orc_sim_set_collision_groups(self.h, C.c_uint32(len(hs)), C.c_void_p(hs.ctypes.data), C.c_void_p(g.ctypes.data))
        if r != 0:
            raise ValueError("unknown object handle")

    def remove(self, handles):
        hs = np.ascontiguousarray(handles, dtype=np.uint32)
        if self.o.lib.orc_sim_remove(self.h, C.c_uint32(len(hs)), C.c_void_p(hs.ctypes.data)) != 0:
            raise RuntimeError("orc_sim_remove: unknown handle")

    def add(self, scene):
        """CollisionWorld::add for every object of `scene` (same hull library); returns the handles."""
        objs, keep = self.o._objects(scene)
        out = np.zeros(scene.n, dtype=np.uint32)
        if self.o.lib.orc_sim_add(self.h, C.byref(objs), C.c_void_p(out.ctypes.data)) != 0:
            raise RuntimeError("orc_sim_add failed")
        return out

    def step(self):
        """One CollisionWorld::update.  Returns dict(pairs[P,2] (h1, h2), algo[P], off[P+1], contacts, ids, events[E,3])."""
        L = self.o.lib
        L.orc_sim_step(self.h)
        P, Cn = int(L.orc_sim_num_pairs(self.h)), int(L.orc_sim_num_contacts(self.h))
        pairs = np.zeros((P, 2), dtype=np.uint32)
        algo = np.zeros(P, dtype=np.uint8)
        off = np.zeros(P + 1, dtype=np.uint32)
        contacts = np.zeros(max(Cn, 1), dtype=self.o.contact_dtype)
        ids = np.zeros(max(Cn, 1), dtype=np.uint32)
        L.orc_sim_fetch(self.h, C.c_void_p(pairs.ctypes.data), C.c_void_p(algo.ctypes.data), C.c_void_p(off.ctypes.data),
                        C.c_void_p(contacts.ctypes.data), C.c_void_p(ids.ctypes.data))
        ne = int(L.orc_sim_events(self.h, None, C.c_uint64(0)))
        ev = np.zeros((max(ne, 1), 3), dtype=np.uint32)
        L.orc_sim_events(self.h, C.c_void_p(ev.ctypes.data), C.c_uint64(ne))
        prox = np.zeros(max(P, 1), dtype=np.uint8)
        L.orc_sim_fetch_proximity(self.h, C.c_void_p(prox.ctypes.data))
        npe = int(L.orc_sim_proximity_events(self.h, None, C.c_uint64(0)))
        pev = np.zeros((max(npe, 1), 4), dtype=np.uint32)
        L.orc_sim_proximity_events(self.h, C.c_void_p(pev.ctypes.data), C.c_uint64(npe))
        return {"pairs": pairs, "algo": algo, "off": off, "contacts": contacts[:Cn], "ids": ids[:Cn], "events": ev[:ne],
                "prox": prox[:P], "prox_events": pev[:npe], "bp_pairs": int(L.orc_sim_bp_num_interferences(self.h))}

    def query(self, kind, q, groups=None):
        """kind 0: interferences_with_aabb (q[n,6]); kind 2: interferences_with_point (q[n,3]).  Rows (query, handle)."""
        q = np.ascontiguousarray(q, dtype=self.o.dtype)
        g = None if groups is None else np.ascontiguousarray(groups, dtype=np.uint32)
        self.o.lib.orc_sim_query.restype = C.c_uint64
        cap = max(64 * len(q), 4096)
        while True:
            idx = np.zeros((cap, 2), dtype=np.uint32)
            n = self.o.lib.orc_sim_query(self.h, C.c_int(kind), C.c_uint64(len(q)), C.c_void_p(q.ctypes.data),
                                         C.c_void_p(g.ctypes.data) if g is not None else None, C.c_void_p(idx.ctypes.data), C.c_uint64(cap))
            if n <= cap:
                return idx[:n]
            cap = int(n)

    def ray_cast(self, origins, dirs, max_toi, groups=None, first_only=False):
        """glue::interferences_with_ray / first_interference_with_ray.  Returns (idx[K,2] (ray, handle), toi[K], normal[K,3], feature[K])."""
        o = np.ascontiguousarray(origins, dtype=self.o.dtype).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, dtype=self.o.dtype).reshape(-1, 3)
        t = np.broadcast_to(np.asarray(max_toi, dtype=self.o.dtype).reshape(-1, 1), (len(o), 1))
        rays = np.ascontiguousarray(np.concatenate([o, d, t], axis=1))
        g = None if groups is None else np.ascontiguousarray(groups, dtype=np.uint32)
        self.o.lib.orc_sim_ray_cast.restype = C.c_uint64
        cap = max(64 * len(o), 4096)
        while True:
            idx = np.zeros((cap, 2), dtype=np.uint32)
            val = np.zeros((cap, 4), dtype=self.o.dtype)
            feat = np.zeros(cap, dtype=np.uint32)
            n = self.o.lib.orc_sim_ray_cast(self.h, C.c_uint64(len(o)), C.c_void_p(rays.ctypes.data), C.c_void_p(g.ctypes.data) if g is not None else None,
                                            C.c_int(int(first_only)), C.c_void_p(idx.ctypes.data), C.c_void_p(val.ctypes.data),
                                            C.c_void_p(feat.ctypes.data), C.c_uint64(cap))
            if n <= cap:
                return idx[:n], val[:n, 0].copy(), val[:n, 1:4].copy(), feat[:n]
            cap = int(n)

    def __del__(self):
        try:
            self.o.lib.orc_sim_destroy(self.h)
        except Exception:
            pass


class OraclePolyline:
    """ncollide2d ``Polyline`` (points [n, 2], edges [m, 2] or None for the line strip) with ``toi_and_normal_with_ray`` for a batch."""

    def __init__(self, oracle, points, edges=None):
        self.o = oracle
        p = np.ascontiguousarray(points, dtype=oracle.dtype).reshape(-1, 2)
        e = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2) if edges is not None else None
        self.nedges = len(e) if e is not None else max(len(p) - 1, 0)
        oracle.lib.orc2_polyline_create.restype = C.c_void_p
        self.h = C.c_void_p(oracle.lib.orc2_polyline_create(C.c_uint32(len(p)), C.c_void_p(p.ctypes.data), C.c_uint32(self.nedges),
                                                            C.c_void_p(e.ctypes.data) if e is not None else None))

    def ray_cast(self, origins, dirs, max_toi=None, pose=None, mode=0):
        """(toi, feature, normal): toi < 0 = None; feature = edge or edge + n_edges (the segment's Face(1)); max_toi scalar or per ray."""
        dt = self.o.dtype
        o = np.ascontiguousarray(origins, dtype=dt).reshape(-1, 2)
        d = np.ascontiguousarray(dirs, dtype=dt).reshape(-1, 2)
        n = len(o)
        toi, feat, normal = np.zeros(n, dtype=dt), np.zeros(n, dtype=np.uint32), np.zeros((n, 2), dtype=dt)
        per = None
        if max_toi is None:
            max_toi = np.finfo(dt).max
        elif np.ndim(max_toi) > 0:
            per = np.ascontiguousarray(max_toi, dtype=dt)
            max_toi = 0.0
        p = np.ascontiguousarray(pose, dtype=dt) if pose is not None else None
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None  # noqa: E731
        self.o.lib.orc2_polyline_ray_cast(self.h, vp(p), C.c_uint64(n), vp(o), vp(d), self.o.creal(max_toi), vp(per), C.c_int(mode), vp(toi),
                                          vp(feat), vp(normal))
        return toi, feat, normal

    def __del__(self):
        try:
            self.o.lib.orc2_polyline_destroy(self.h)
        except Exception:
            pass


class OracleTriMesh:
    def __init__(self, oracle, verts, tris):
        self.o = oracle
        v = np.ascontiguousarray(verts, dtype=oracle.dtype)
        t = np.ascontiguousarray(tris, dtype=np.uint32)
        self.ntris = len(t)
        self.h = C.c_void_p(oracle.lib.orc_trimesh_create(C.c_uint32(len(v)), C.c_void_p(v.ctypes.data), C.c_uint32(len(t)), C.c_void_p(t.ctypes.data)))

    def ray_cast(self, origins, dirs, max_toi=None, pose=None, mode=0):
        dt = self.o.dtype
        o = np.ascontiguousarray(origins, dtype=dt)
        d = np.ascontiguousarray(dirs, dtype=dt)
        n = len(o)
        toi = np.zeros(n, dtype=dt)
        face = np.zeros(n, dtype=np.uint32)
        normal = np.zeros((n, 3), dtype=dt)
        if max_toi is None:
            max_toi = np.finfo(dt).max
        p = np.ascontiguousarray(pose, dtype=dt) if pose is not None else None
        self.o.lib.orc_trimesh_ray_cast(
            self.h, C.c_void_p(p.ctypes.data) if p is not None else None, C.c_uint64(n), C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data),
            self.o.creal(max_toi), C.c_int(mode), C.c_void_p(toi.ctypes.data), C.c_void_p(face.ctypes.data), C.c_void_p(normal.ctypes.data),
        )
        return toi, face, normal

    def ray_cast_uv(self, origins, dirs, uvs=None, max_toi=None, pose=None, mode=0):
        """toi_and_normal_and_uv_with_ray; max_toi: None, a scalar, or one per ray.  Returns (toi, face, normal, uv)."""
        dt = self.o.dtype
        o = np.ascontiguousarray(origins, dtype=dt)
        d = np.ascontiguousarray(dirs, dtype=dt)
        n = len(o)
        toi = np.zeros(n, dtype=dt)
        face = np.zeros(n, dtype=np.uint32)
        normal = np.zeros((n, 3), dtype=dt)
        uv = np.zeros((n, 2), dtype=dt)
        per = None
        if max_toi is None:
            max_toi = np.finfo(dt).max
        elif np.ndim(max_toi) > 0:
            per = np.ascontiguousarray(max_toi, dtype=dt)
            max_toi = 0.0
        p = np.ascontiguousarray(pose, dtype=dt) if pose is not None else None
        u = np.ascontiguousarray(uvs, dtype=dt) if uvs is not None else None
        vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None  # noqa: E731
        self.o.lib.orc_trimesh_ray_cast_uv(self.h, vp(p), C.c_uint64(n), vp(o), vp(d), self.o.creal(max_toi), vp(per), vp(u), C.c_int(mode),
                                           vp(toi), vp(face), vp(normal), vp(uv))
        return toi, face, normal, uv

    def __del__(self):
        try:
            self.o.lib.orc_trimesh_destroy(self.h)
        except Exception:
            pass


class OracleBroadPhase:
    """Reference-faithful multi-step DBVTBroadPhase (oracle/bp_persistent.cpp)."""

    def __init__(self, oracle, margin):
        self.o = oracle
        L = oracle.lib
        L.orc_bp_create.restype = C.c_void_p
        L.orc_bp_create_proxy.restype = C.c_uint32
        L.orc_bp_num_interferences.restype = C.c_uint64
        L.orc_bp_pairs.restype = C.c_uint64
        self.h = C.c_void_p(L.orc_bp_create(oracle.creal(margin)))

    def create_proxy(self, bv):
        a = np.ascontiguousarray(bv, dtype=self.o.dtype)
        return int(self.o.lib.orc_bp_create_proxy(self.h, C.c_void_p(a.ctypes.data)))

    def deferred_set_bounding_volume(self, handle, bv):
        a = np.ascontiguousarray(bv, dtype=self.o.dtype)
        r = self.o.lib.orc_bp_set_bounding_volume(self.h, C.c_uint32(handle), C.c_void_p(a.ctypes.data))
        if r != 0:
            raise RuntimeError("Attempting to set the bounding volume of an object that does not exist.")

    def remove(self, handles):
        hs = np.ascontiguousarray(handles, dtype=np.uint32)
        cap = max(16 * len(hs) + 1024, 4096)
        while True:
            out = np.zeros((cap, 2), dtype=np.uint32)
            n = C.c_uint64()
            # the call mutates state: size the buffer generously up front instead of retrying
            r = self.o.lib.orc_bp_remove(self.h, C.c_uint32(len(hs)), C.c_void_p(hs.ctypes.data), C.c_void_p(out.ctypes.data), C.c_uint64(cap), C.byref(n))
            if r != 0:
                raise RuntimeError("Attempting to remove an object that does not exist.")
            return out[: min(n.value, cap)].copy()

    def update(self, groups=None, cap=None):
        cap = cap or 1 << 22
        st = np.zeros((cap, 2), dtype=np.uint32)
        sp = np.zeros((cap, 2), dtype=np.uint32)
        ns, np_ = C.c_uint64(), C.c_uint64()
        g = np.ascontiguousarray(groups, dtype=np.uint32) if groups is not None else None
        self.o.lib.orc_bp_update(self.h, C.c_void_p(g.ctypes.data) if g is not None else None, C.c_void_p(st.ctypes.data), C.c_uint64(cap), C.byref(ns),
                                 C.c_void_p(sp.ctypes.data), C.c_uint64(cap), C.byref(np_))
        assert ns.value <= cap and np_.value <= cap
        return st[: ns.value].copy(), sp[: np_.value].copy()

    def deferred_recompute_all_proximities_with(self, handle):
        self.o.lib.orc_bp_recompute_with(self.h, C.c_uint32(handle))

    def deferred_recompute_all_proximities(self):
        self.o.lib.orc_bp_recompute_all(self.h)

    def _query(self, kind, q):
        q = np.ascontiguousarray(q, dtype=self.o.dtype)
        self.o.lib.orc_bp_query.restype = C.c_uint64
        cap = 4096
        while True:
            out = np.zeros(cap, dtype=np.uint32)
            n = self.o.lib.orc_bp_query(self.h, C.c_int(kind), C.c_void_p(q.ctypes.data), C.c_void_p(out.ctypes.data), C.c_uint64(cap))
            if n <= cap:
                return out[:n].copy()
            cap = int(n)

    def interferences_with_bounding_volume(self, bv):
        return self._query(0, bv)

    def interferences_with_ray(self, origin, direction, max_toi):
        return self._query(1, np.concatenate([np.asarray(origin).reshape(3), np.asarray(direction).reshape(3), [max_toi]]))

    def interferences_with_point(self, point):
        return self._query(2, point)

    def num_interferences(self):
        return int(self.o.lib.orc_bp_num_interferences(self.h))

    def proxy(self, handle):
        out = np.zeros(6, dtype=self.o.dtype)
        r = self.o.lib.orc_bp_proxy(self.h, C.c_uint32(handle), C.c_void_p(out.ctypes.data))
        return out if r else None

    def pairs(self):
        n = self.num_interferences()
        out = np.zeros((max(n, 1), 2), dtype=np.uint32)
        self.o.lib.orc_bp_pairs(self.h, C.c_void_p(out.ctypes.data), C.c_uint64(n))
        return out[:n]

    def __del__(self):
        try:
            self.o.lib.orc_bp_destroy(self.h)
        except Exception:
            pass
