"""Host-side shape mirror of the reference's shape types that are on the hot path.

Mirrors (names, argument meaning, failure behaviour) of
  * ``Ball::new(radius)``                      shape/ball.rs:9-23
  * ``Cuboid::new(half_extents)``              shape/cuboid.rs:24-31
  * ``Plane::new(normal)``                     shape/plane.rs:8-23
  * ``ConvexHull::try_new(points, indices)``   shape/convex.rs:109-335  (adjacency tables, merged coplanar faces)
  * ``TriMesh::new(points, indices, None)``    shape/trimesh.rs:100-197 (see ray.py)

The convex-hull tables are *setup-time* data (built once per hull, on the host, exactly like the Rust side
does in ``ConvexHull::try_new``); the device consumes them through ``ncb_set_hulls`` (include/ncb200.h).
All arithmetic that ends up in the tables is done in f32, operation by operation, in the reference's order.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
EPS = np.finfo(np.float32).eps

BALL, CUBOID, HULL, PLANE = 0, 1, 2, 3
CAPSULE = 4  # CapsuleCapsule / CapsuleShape contact generators on the device (csrc/capsule.cuh)


def _dot(a, b):
    return F32(F32(F32(a[0] * b[0]) + F32(a[1] * b[1])) + F32(a[2] * b[2]))


def _cross(a, b):
    return np.array(
        [
            F32(F32(a[1] * b[2]) - F32(a[2] * b[1])),
            F32(F32(a[2] * b[0]) - F32(a[0] * b[2])),
            F32(F32(a[0] * b[1]) - F32(a[1] * b[0])),
        ],
        dtype=F32,
    )


def _unit_try_new(v, min_norm):
    """Unit::try_new: succeeds iff norm_squared > min_norm^2; divides by sqrt(norm_squared)."""
    sq = _dot(v, v)
    if sq > F32(min_norm) * F32(min_norm):
        n = np.sqrt(sq, dtype=F32)
        return (v / n).astype(F32)
    return None


class Ball:
    type_id = BALL

    def __init__(self, radius):
        self.radius = F32(radius)

    def param(self):
        return np.array([self.radius, 0, 0, 0], dtype=F32)


class Capsule:
    """``Capsule::new(half_height, radius)`` (shape/capsule.rs:19-30), principal axis = local y.  Contact generation runs on the
    device (k_capsule); sensors and world ray / point queries in a world with capsules are refused loudly (NCB_ERR_UNSUPPORTED)."""

    type_id = CAPSULE

    def __init__(self, half_height, radius):
        self.half_height, self.radius = F32(half_height), F32(radius)

    def param(self):
        return np.array([self.half_height, self.radius, 0, 0], dtype=F32)


class Cuboid:
    type_id = CUBOID

    def __init__(self, half_extents):
        self.half_extents = np.asarray(half_extents, dtype=F32).reshape(3)

    def param(self):
        return np.array([*self.half_extents, 0], dtype=F32)


class Plane:
    type_id = PLANE

    def __init__(self, normal):
        self.normal = np.asarray(normal, dtype=F32).reshape(3)

    def param(self):
        return np.array([*self.normal, 0], dtype=F32)


class ConvexHull:
    """Convex polyhedron tables; ``try_new`` follows shape/convex.rs:109-335 step by step."""

    type_id = HULL

    def __init__(self):
        self.points = None
        self.vert_first_adj = None
        self.vert_num_adj = None
        self.face_first = None
        self.face_num = None
        self.face_normal = None
        self.vertices_adj_to_face = None
        self.edges_adj_to_face = None
        self.edge_vertices = None
        self.edge_faces = None
        self.edge_dir = None
        self.edge_deleted = None
        self.faces_adj_to_vertex = None
        self.edges_adj_to_vertex = None

    @staticmethod
    def try_new(points, indices):
        """Returns a ConvexHull or None (degenerate edge/triangle, or Euler characteristic violated)."""
        pts = np.ascontiguousarray(points, dtype=F32).reshape(-1, 3)
        idx = np.asarray(indices, dtype=np.int64).reshape(-1, 3)
        eps = np.sqrt(EPS, dtype=F32)

        edges = []  # dict(vertices, faces, dir, deleted)
        triangles = []  # dict(vertices, edges, normal, parent_face)
        edge_map = {}
        faces = []  # [first, num, normal]
        edges_adj_to_face = []
        vertices_adj_to_face = []

        for vtx in idx:
            edges_id = [0, 0, 0]
            face_id = len(triangles)
            for i1 in range(3):
                i2 = (i1 + 1) % 3
                a, b = int(vtx[i1]), int(vtx[i2])
                key = (min(a, b), max(a, b))
                if key in edge_map:
                    edges_id[i1] = edge_map[key]
                    edges[edge_map[key]]["faces"][1] = face_id
                else:
                    edge_map[key] = len(edges)
                    edges_id[i1] = len(edges)
                    d = _unit_try_new((pts[b] - pts[a]).astype(F32), EPS)
                    if d is None:
                        return None
                    edges.append({"vertices": [a, b], "faces": [face_id, 0], "dir": d, "deleted": False})
            # utils::ccw_face_normal
            ab = (pts[vtx[1]] - pts[vtx[0]]).astype(F32)
            ac = (pts[vtx[2]] - pts[vtx[0]]).astype(F32)
            normal = _unit_try_new(_cross(ab, ac), EPS)
            if normal is None:
                return None
            triangles.append({"vertices": [int(v) for v in vtx], "edges": edges_id, "normal": normal, "parent_face": None})

        num_valid_edges = 0
        for e in edges:
            n1 = triangles[e["faces"][0]]["normal"]
            n2 = triangles[e["faces"][1]]["normal"]
            if _dot(n1, n2) > F32(1.0) - eps:
                e["deleted"] = True
            else:
                num_valid_edges += 1

        for i in range(len(triangles)):
            if triangles[i]["parent_face"] is None:
                for j1 in range(3):
                    if not edges[triangles[i]["edges"][j1]]["deleted"]:
                        new_face_id = len(faces)
                        new_face = [len(edges_adj_to_face), 1, triangles[i]["normal"]]
                        edges_adj_to_face.append(triangles[i]["edges"][j1])
                        vertices_adj_to_face.append(triangles[i]["vertices"][j1])
                        j2 = (j1 + 1) % 3
                        start_vertex = triangles[i]["vertices"][j1]
                        curr_triangle = i
                        curr_edge_id = j2
                        guard = 0
                        while triangles[curr_triangle]["vertices"][curr_edge_id] != start_vertex:
                            guard += 1
                            if guard > 100000:
                                return None
                            curr_edge = triangles[curr_triangle]["edges"][curr_edge_id]
                            curr_vertex = triangles[curr_triangle]["vertices"][curr_edge_id]
                            triangles[curr_triangle]["parent_face"] = new_face_id
                            if not edges[curr_edge]["deleted"]:
                                edges_adj_to_face.append(curr_edge)
                                vertices_adj_to_face.append(curr_vertex)
                                new_face[1] += 1
                                curr_edge_id = (curr_edge_id + 1) % 3
                            else:
                                f = edges[curr_edge]["faces"]
                                curr_triangle = f[1] if curr_triangle == f[0] else f[0]
                                curr_edge_id = (triangles[curr_triangle]["edges"].index(curr_edge) + 1) % 3
                                assert triangles[curr_triangle]["vertices"][curr_edge_id] == curr_vertex
                        if new_face[1] > 2:
                            faces.append(new_face)
                        break

        for e in edges:
            for k in range(2):
                fidx = triangles[e["faces"][k]]["parent_face"]
                if fidx is not None:
                    e["faces"][k] = fidx

        nv = len(pts)
        num_adj = [0] * nv
        for first, num, _ in faces:
            for v in vertices_adj_to_face[first : first + num]:
                num_adj[v] += 1
        first_adj = [0] * nv
        total = 0
        for v in range(nv):
            first_adj[v] = total
            total += num_adj[v]
        faces_adj_to_vertex = [0] * total
        edges_adj_to_vertex = [0] * total
        fill = [0] * nv
        for face_id, (first, num, _) in enumerate(faces):
            for vid in range(first, first + num):
                v = vertices_adj_to_face[vid]
                faces_adj_to_vertex[first_adj[v] + fill[v]] = face_id
                edges_adj_to_vertex[first_adj[v] + fill[v]] = edges_adj_to_face[vid]
                fill[v] += 1
        num_valid_vertices = sum(1 for v in range(nv) if fill[v] != 0)
        if num_valid_vertices + len(faces) - num_valid_edges != 2:
            return None

        h = ConvexHull()
        h.points = pts
        h.vert_first_adj = np.asarray(first_adj, dtype=np.uint32)
        h.vert_num_adj = np.asarray(fill, dtype=np.uint32)
        h.face_first = np.asarray([f[0] for f in faces], dtype=np.uint32)
        h.face_num = np.asarray([f[1] for f in faces], dtype=np.uint32)
        h.face_normal = np.asarray([f[2] for f in faces], dtype=F32).reshape(-1, 3)
        h.vertices_adj_to_face = np.asarray(vertices_adj_to_face, dtype=np.uint32)
        h.edges_adj_to_face = np.asarray(edges_adj_to_face, dtype=np.uint32)
        h.edge_vertices = np.asarray([e["vertices"] for e in edges], dtype=np.uint32).reshape(-1, 2)
        h.edge_faces = np.asarray([e["faces"] for e in edges], dtype=np.uint32).reshape(-1, 2)
        h.edge_dir = np.asarray([e["dir"] for e in edges], dtype=F32).reshape(-1, 3)
        h.edge_deleted = np.asarray([e["deleted"] for e in edges], dtype=bool)
        h.faces_adj_to_vertex = np.asarray(faces_adj_to_vertex, dtype=np.uint32)
        h.edges_adj_to_vertex = np.asarray(edges_adj_to_vertex, dtype=np.uint32)
        return h

    @staticmethod
    def try_from_points(points):
        """ConvexHull::try_from_points (shape/convex.rs:90-99): the hull triangulation comes from
        scipy's qhull here (the reference's own quickhull, transformation/convex_hull3.rs, is setup code
        outside the hot path); only hull vertices are kept and triangles are oriented outward."""
        from scipy.spatial import ConvexHull as QHull

        pts = np.ascontiguousarray(points, dtype=F32).reshape(-1, 3)
        q = QHull(pts.astype(np.float64))
        used = np.unique(q.simplices)
        remap = -np.ones(len(pts), dtype=np.int64)
        remap[used] = np.arange(len(used))
        hp = pts[used]
        tris = remap[q.simplices]
        c = hp.astype(np.float64).mean(axis=0)
        for t in tris:
            a, b, cc = hp[t[0]].astype(np.float64), hp[t[1]].astype(np.float64), hp[t[2]].astype(np.float64)
            if np.dot(np.cross(b - a, cc - a), a - c) < 0:
                t[1], t[2] = t[2], t[1]
        return ConvexHull.try_new(hp, tris)

    def check_geometry(self):
        """ConvexHull::check_geometry (shape/convex.rs:338-347)."""
        for f in range(len(self.face_first)):
            p0 = self.points[self.vertices_adj_to_face[self.face_first[f]]]
            for v in self.points:
                if not _dot((v - p0).astype(F32), self.face_normal[f]) <= EPS:
                    return False
        return True


class HullLibrary:
    """Flat, offset-indexed pack of ConvexHull tables (layout of ncb_hull_library in include/ncb200.h)."""

    def __init__(self, hulls):
        self.hulls = list(hulls)
        n = len(self.hulls)

        def offs(lens):
            o = np.zeros(n + 1, dtype=np.uint32)
            if n:
                o[1:] = np.cumsum(lens)
            return o

        def cat(arrs, dtype, width=None):
            if not arrs:
                return np.zeros((0,) if width is None else (0, width), dtype=dtype)
            return np.ascontiguousarray(np.concatenate(arrs).astype(dtype))

        H = self.hulls
        self.n_hulls = n
        self.vert_off = offs([len(h.points) for h in H])
        self.face_off = offs([len(h.face_first) for h in H])
        self.edge_off = offs([len(h.edge_vertices) for h in H])
        self.fadj_off = offs([len(h.vertices_adj_to_face) for h in H])
        self.vadj_off = offs([len(h.faces_adj_to_vertex) for h in H])
        self.points = cat([h.points for h in H], F32, 3)
        self.vert_first_adj = cat([h.vert_first_adj for h in H], np.uint32)
        self.vert_num_adj = cat([h.vert_num_adj for h in H], np.uint32)
        self.face_first = cat([h.face_first for h in H], np.uint32)
        self.face_num = cat([h.face_num for h in H], np.uint32)
        self.face_normal = cat([h.face_normal for h in H], F32, 3)
        self.vertices_adj_to_face = cat([h.vertices_adj_to_face for h in H], np.uint32)
        self.edges_adj_to_face = cat([h.edges_adj_to_face for h in H], np.uint32)
        self.edge_vertices = cat([h.edge_vertices for h in H], np.uint32, 2)
        self.edge_faces = cat([h.edge_faces for h in H], np.uint32, 2)
        self.edge_dir = cat([h.edge_dir for h in H], F32, 3)
        self.faces_adj_to_vertex = cat([h.faces_adj_to_vertex for h in H], np.uint32)
        self.edges_adj_to_vertex = cat([h.edges_adj_to_vertex for h in H], np.uint32)

    FIELDS = (
        "vert_off face_off edge_off fadj_off vadj_off points vert_first_adj vert_num_adj face_first face_num "
        "face_normal vertices_adj_to_face edges_adj_to_face edge_vertices edge_faces edge_dir faces_adj_to_vertex "
        "edges_adj_to_vertex"
    ).split()

    @property
    def max_verts(self):
        return int(np.diff(self.vert_off).max()) if self.n_hulls else 0

    @property
    def max_face_verts(self):
        return int(self.face_num.max()) if len(self.face_num) else 0
