// ORACLE — TEST INFRASTRUCTURE ONLY (see na.hpp header).
// GJK + EPA restated from query/algorithms/gjk.rs:26-177,367-388, query/algorithms/epa3.rs:13-454,
// utils/ccw_face_normal.rs:21-27, utils/triangle.rs:85-102, shape/support_map.rs:26-29,
// shape/cuboid.rs:137-145, utils/point_cloud_support_point.rs:6-24, and Rust's std BinaryHeap
// (push = sift_up, pop = swap-remove root + sift_down_to_bottom + sift_up).
#pragma once
#include <vector>
#include "scene.hpp"
#include "simplex.hpp"

namespace orc {

// A support-mapped operand: cuboid, convex hull, ball (for one-shot queries) or ConstantOrigin.
struct Support {
    // S_CYLINDER: only for the reference's cylinder / cuboid KAT.  S_SEGMENT / S_CAPSULE: half_height in he.x (capsule radius in radius).
    enum Kind { S_CUBOID, S_HULL, S_BALL, S_ORIGIN, S_CYLINDER, S_SEGMENT, S_CAPSULE } kind;
    V3 he;
    real radius;
    Hull hull;
    V3 local_support_point(V3 dir) const {
        switch (kind) {
            case S_CUBOID:  // cuboid.rs:137-145
                return v3(std::copysign(he.x, dir.x), std::copysign(he.y, dir.y), std::copysign(he.z, dir.z));
            case S_HULL: {  // point_cloud_support_point.rs:6-24 (first maximum wins)
                uint32_t best = 0;
                real best_dot = dot(hull.pt(0), dir);
                for (uint32_t i = 1; i < hull.nv; ++i) {
                    real d = dot(hull.pt(i), dir);
                    if (d > best_dot) {
                        best_dot = d;
                        best = i;
                    }
                }
                return hull.pt(best);
            }
            case S_CYLINDER: {  // cylinder.rs:47-62 (half_height in he.x)
                V3 vres = v3(dir.x, 0, dir.z);
                real n = norm(vres);
                vres = n == real(0) ? v3(0, 0, 0) : (vres / n) * radius;
                vres.y = std::copysign(he.x, dir.y);
                return vres;
            }
            case S_SEGMENT: {  // segment.rs:182-190 with a = (0, -hh, 0), b = (0, hh, 0)
                V3 a = v3(0, -he.x, 0), b = v3(0, he.x, 0);
                return dot(a, dir) > dot(b, dir) ? a : b;
            }
            case S_CAPSULE: {  // capsule.rs:72-85: local_support_point_toward(normalize(dir))
                V3 d = normalize(dir);
                return v3(0, std::copysign(he.x, d.y), 0) + d * radius;
            }
            case S_BALL:  // ball.rs:41-48: local_support_point_toward(normalize(dir)) = dir * radius
                return normalize(dir) * radius;
            default:
                return v3(0, 0, 0);
        }
    }
    V3 support_point(const Iso& m, V3 dir) const {  // support_map.rs:26-29
        if (kind == S_ORIGIN) return v3(0, 0, 0);
        if (kind == S_BALL) return m.t + normalize(dir) * radius;  // Ball overrides it (ball.rs:31-38): the rotation is not applied
        V3 ld = iso_inv_vec(m, dir);
        return iso_mul_point(m, local_support_point(ld));
    }
    V3 support_point_toward(const Iso& m, V3 unit_dir) const {  // support_map.rs:32-35; ball.rs:36-38 (no normalisation)
        if (kind == S_BALL) return m.t + unit_dir * radius;
        if (kind == S_CAPSULE) {  // capsule.rs:79-84 through the default support_point_toward: the local direction is not re-normalised
            V3 ld = iso_inv_vec(m, unit_dir);
            return iso_mul_point(m, v3(0, std::copysign(he.x, ld.y), 0) + ld * radius);
        }
        return support_point(m, unit_dir);
    }
};

static inline CSOPoint cso_from_shapes(const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, V3 dir) {
    V3 sp1 = g1.support_point(m1, dir);
    V3 sp2 = g2.support_point(m2, -dir);
    return cso_new(sp1, sp2);
}

static inline real gjk_eps_tol() { return EPS * real(10); }

enum GJKKind { GJK_INTERSECTION, GJK_CLOSEST_POINTS, GJK_PROXIMITY, GJK_NO_INTERSECTION };
struct GJKResult {
    GJKKind kind;
    V3 p1, p2, dir;
};

static inline void gjk_result(const VoronoiSimplex& s, bool prev, V3* p1, V3* p2) {  // gjk.rs:367-388
    V3 r0 = v3(0, 0, 0), r1 = v3(0, 0, 0);
    if (prev) {
        for (int i = 0; i < s.prev_dim + 1; ++i) {
            real coord = s.prev_proj[i];
            const CSOPoint& pt = s.prev_point(i);
            r0 = r0 + pt.orig1 * coord;
            r1 = r1 + pt.orig2 * coord;
        }
    } else {
        for (int i = 0; i < s.dim + 1; ++i) {
            real coord = s.proj[i];
            const CSOPoint& pt = s.vertices[i];
            r0 = r0 + pt.orig1 * coord;
            r1 = r1 + pt.orig2 * coord;
        }
    }
    *p1 = r0;
    *p2 = r1;
}

struct GJKStats {
    uint32_t gjk_iters = 0, epa_iters = 0, epa_max_verts = 0, epa_max_faces = 0, epa_max_heap = 0, epa_calls = 0, epa_fail = 0;
};

// gjk.rs:76-177.  exact_dist = true is what the contact generators use; exact_dist = false (the proximity detectors,
// proximity_support_map_support_map.rs:68) stops as soon as the origin is proven closer than max_dist but outside.
static inline GJKResult gjk_closest_points(const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, real max_dist,
                                           VoronoiSimplex& simplex, GJKStats* st, bool exact_dist = true) {
    const real _eps_tol = gjk_eps_tol();
    const real _eps_rel = std::sqrt(_eps_tol);
    GJKResult res;
    V3 proj = simplex.project_origin_and_reduce();
    V3 old_dir;
    {
        V3 pd;
        if (unit_try_new(proj, real(0), &pd))
            old_dir = -pd;
        else
            return {GJK_INTERSECTION, {}, {}, {}};
    }
    real max_bound = FMAX;
    V3 dir;
    int niter = 0;
    for (;;) {
        real old_max_bound = max_bound;
        real dist;
        if (unit_try_new_and_get(-proj, _eps_tol, &dir, &dist))
            max_bound = dist;
        else
            return {GJK_INTERSECTION, {}, {}, {}};

        if (max_bound >= old_max_bound) {
            if (!exact_dist) return {GJK_PROXIMITY, {}, {}, old_dir};
            res.kind = GJK_CLOSEST_POINTS;
            gjk_result(simplex, true, &res.p1, &res.p2);
            res.dir = old_dir;
            return res;
        }
        CSOPoint cso_point = cso_from_shapes(m1, g1, m2, g2, dir);
        real min_bound = -dot(dir, cso_point.point);
        if (min_bound > max_dist) {
            return {GJK_NO_INTERSECTION, {}, {}, dir};
        } else if (!exact_dist && min_bound > real(0) && max_bound <= max_dist) {
            return {GJK_PROXIMITY, {}, {}, old_dir};
        } else if (max_bound - min_bound <= _eps_rel * max_bound) {
            if (!exact_dist) return {GJK_PROXIMITY, {}, {}, dir};
            res.kind = GJK_CLOSEST_POINTS;
            gjk_result(simplex, false, &res.p1, &res.p2);
            res.dir = dir;
            return res;
        }
        if (!simplex.add_point(cso_point, _eps_tol)) {
            if (!exact_dist) return {GJK_PROXIMITY, {}, {}, dir};
            res.kind = GJK_CLOSEST_POINTS;
            gjk_result(simplex, false, &res.p1, &res.p2);
            res.dir = dir;
            return res;
        }
        old_dir = dir;
        proj = simplex.project_origin_and_reduce();
        if (simplex.dim == 3) {
            if (min_bound >= _eps_tol) {
                if (!exact_dist) return {GJK_PROXIMITY, {}, {}, old_dir};
                res.kind = GJK_CLOSEST_POINTS;
                gjk_result(simplex, true, &res.p1, &res.p2);
                res.dir = old_dir;
                return res;
            }
            return {GJK_INTERSECTION, {}, {}, {}};
        }
        niter += 1;
        if (st) st->gjk_iters++;
        if (niter == 10000) return {GJK_NO_INTERSECTION, {}, {}, v3(1, 0, 0)};
    }
}

// ------------------------------------------------------------------------------------------------
// gjk::cast_ray -> minkowski_ray_cast (query/algorithms/gjk.rs:180-365) with g2 = ConstantOrigin, m2 = identity.
// Returns true + (toi, normal) on a hit.  The simplex must have been reset by the caller (ray_support_map.rs:28-29).
// ------------------------------------------------------------------------------------------------
static inline bool ray_toi_with_plane(V3 center, V3 normal, V3 origin, V3 dir, real* t_out) {  // ray_plane.rs:9-42
    V3 dpos = center - origin;
    real denom = dot(normal, dir);
    if (relative_eq(denom, real(0))) return false;
    real t = dot(normal, dpos) / denom;
    if (t >= real(0)) {
        *t_out = t;
        return true;
    }
    return false;
}
static inline bool minkowski_ray_cast(const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, V3 ray_origin, V3 ray_dir, real max_toi,
                                      VoronoiSimplex& simplex, real* toi_out, V3* normal_out) {
    const real _eps_tol = EPS * real(10);
    const real _eps_rel = std::sqrt(_eps_tol);
    real ray_length = norm(ray_dir);
    if (relative_eq(ray_length, real(0))) return false;
    real ltoi = 0;
    V3 curr_origin = ray_origin, curr_dir = ray_dir / ray_length;
    V3 dir0 = -curr_dir;
    V3 ldir = dir0;
    CSOPoint sp0 = cso_from_shapes(m1, g1, m2, g2, dir0);
    sp0.point = sp0.point + (-curr_origin);
    simplex.reset(sp0);
    V3 proj = simplex.project_origin_and_reduce();
    real max_bound = FMAX;
    V3 dir;
    int niter = 0;
    bool last_chance = false;
    for (;;) {
        real old_max_bound = max_bound;
        real dist;
        if (unit_try_new_and_get(-proj, _eps_tol, &dir, &dist))
            max_bound = dist;
        else {
            *toi_out = ltoi / ray_length, *normal_out = ldir;
            return true;
        }
        CSOPoint support_point;
        if (max_bound >= old_max_bound) {
            last_chance = true;
            V3 p = proj + curr_origin;  // CSOPoint::single_point
            support_point = CSOPoint{p, p, v3(0, 0, 0)};
        } else {
            support_point = cso_from_shapes(m1, g1, m2, g2, dir);
        }
        if (last_chance && ltoi > real(0)) {
            *toi_out = ltoi / ray_length, *normal_out = ldir;
            return true;
        }
        real t;
        if (ray_toi_with_plane(support_point.point, dir, curr_origin, curr_dir, &t)) {
            if (dot(dir, curr_dir) < real(0) && t > real(0)) {
                ldir = dir;
                ltoi += t;
                if (ltoi / ray_length > max_toi) return false;
                V3 shift = curr_dir * t;
                curr_origin = curr_origin + shift;
                max_bound = FMAX;
                for (int i = 0; i < simplex.dim + 1; ++i) simplex.vertices[i].point = simplex.vertices[i].point + (-shift);
                last_chance = false;
            }
        } else if (dot(dir, curr_dir) > _eps_tol) {
            return false;
        }
        if (last_chance) return false;
        real min_bound = -dot(dir, support_point.point - curr_origin);
        if (max_bound - min_bound <= _eps_rel * max_bound) return false;  // feature improved_fixed_point_support is off
        CSOPoint tp = support_point;
        tp.point = tp.point + (-curr_origin);
        (void)simplex.add_point(tp, _eps_tol);
        proj = simplex.project_origin_and_reduce();
        if (simplex.dim == 3) {
            if (min_bound >= _eps_tol) return false;
            *toi_out = ltoi / ray_length, *normal_out = ldir;
            return true;
        }
        niter += 1;
        if (niter == 10000) return false;
    }
}

// ------------------------------------------------------------------------------------------------
// EPA (epa3.rs)
// ------------------------------------------------------------------------------------------------
struct EpaFaceId {
    uint32_t id;
    real neg_dist;
};
static inline bool fid_le(const EpaFaceId& a, const EpaFaceId& b) { return a.neg_dist <= b.neg_dist; }  // PartialOrd `<=`

// Rust std::collections::BinaryHeap<FaceId> (max-heap)
struct RustBinaryHeap {
    std::vector<EpaFaceId> data;
    void sift_up(size_t start, size_t pos) {
        EpaFaceId elt = data[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (fid_le(elt, data[parent])) break;
            data[pos] = data[parent];
            pos = parent;
        }
        data[pos] = elt;
    }
    void push(EpaFaceId item) {
        size_t old_len = data.size();
        data.push_back(item);
        sift_up(0, old_len);
    }
    bool pop(EpaFaceId* out) {
        if (data.empty()) return false;
        EpaFaceId item = data.back();
        data.pop_back();
        if (!data.empty()) {
            std::swap(item, data[0]);
            // sift_down_to_bottom(0)
            size_t end = data.size(), pos = 0;
            EpaFaceId elt = data[0];
            size_t child = 1;
            while (child <= (end >= 2 ? end - 2 : 0) && end >= 2) {
                if (fid_le(data[child], data[child + 1])) child += 1;
                data[pos] = data[child];
                pos = child;
                child = 2 * pos + 1;
            }
            if (child == end - 1) {
                data[pos] = data[child];
                pos = child;
            }
            data[pos] = elt;
            sift_up(0, pos);
        }
        *out = item;
        return true;
    }
};

struct EpaFace {
    uint32_t pts[3], adj[3];
    V3 normal;
    real bcoords[3];
    bool deleted;
};

static inline bool ccw_face_normal(V3 a, V3 b, V3 c, V3* n) {  // ccw_face_normal.rs:21-27
    return unit_try_new(cross(b - a, c - a), EPS, n);
}
static inline bool is_affinely_dependent_triangle(V3 p1, V3 p2, V3 p3) {  // utils/triangle.rs:85-102
    V3 p1p2 = p2 - p1, p1p3 = p3 - p1;
    real eps_tol = EPS * real(100);
    return relative_eq(norm_squared(cross(p1p2, p1p3)), real(0), eps_tol * eps_tol);
}

struct EPA {
    std::vector<CSOPoint> vertices;
    std::vector<EpaFace> faces;
    std::vector<std::pair<uint32_t, uint32_t>> silhouette;  // (face_id, opp_pt_id)
    RustBinaryHeap heap;
    bool panicked = false;  // an assert!/unwrap of the reference would have fired

    EpaFace face_new(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t a0, uint32_t a1, uint32_t a2, bool* proj_inside) {
        Location loc;
        project_on_triangle(vertices[p0].point, vertices[p1].point, vertices[p2].point, v3(0, 0, 0), true, &loc);
        EpaFace f;
        f.pts[0] = p0, f.pts[1] = p1, f.pts[2] = p2;
        f.adj[0] = a0, f.adj[1] = a1, f.adj[2] = a2;
        if (!ccw_face_normal(vertices[p0].point, vertices[p1].point, vertices[p2].point, &f.normal)) f.normal = v3(0, 0, 0);
        f.deleted = false;
        if (loc.kind == ON_FACE) {
            f.bcoords[0] = loc.bc[0], f.bcoords[1] = loc.bc[1], f.bcoords[2] = loc.bc[2];
            *proj_inside = true;
        } else {
            f.bcoords[0] = f.bcoords[1] = f.bcoords[2] = 0;
            *proj_inside = false;
        }
        return f;
    }
    void face_closest_points(const EpaFace& f, V3* p1, V3* p2) const {
        *p1 = vertices[f.pts[0]].orig1 * f.bcoords[0] + vertices[f.pts[1]].orig1 * f.bcoords[1] +
              vertices[f.pts[2]].orig1 * f.bcoords[2];
        *p2 = vertices[f.pts[0]].orig2 * f.bcoords[0] + vertices[f.pts[1]].orig2 * f.bcoords[1] +
              vertices[f.pts[2]].orig2 * f.bcoords[2];
    }
    uint32_t next_ccw_pt_id(const EpaFace& f, uint32_t id) {
        if (f.pts[0] == id) return 1;
        if (f.pts[1] == id) return 2;
        if (f.pts[2] != id) panicked = true;  // assert_eq!
        return 0;
    }
    bool can_be_seen_by(const EpaFace& f, uint32_t point, uint32_t opp) const {
        V3 p0 = vertices[f.pts[opp]].point;
        V3 p1 = vertices[f.pts[(opp + 1) % 3]].point;
        V3 p2 = vertices[f.pts[(opp + 2) % 3]].point;
        V3 pt = vertices[point].point;
        return dot(pt - p0, f.normal) >= -gjk_eps_tol() || is_affinely_dependent_triangle(p1, p2, pt);
    }
    void compute_silhouette(uint32_t point, uint32_t id, uint32_t opp) {  // epa3.rs:432-454 (recursive)
        if (panicked) return;
        if (!faces[id].deleted) {
            if (!can_be_seen_by(faces[id], point, opp)) {
                silhouette.push_back({id, opp});
            } else {
                faces[id].deleted = true;
                uint32_t adj_pt_id1 = (opp + 2) % 3, adj_pt_id2 = opp;
                uint32_t adj1 = faces[id].adj[adj_pt_id1], adj2 = faces[id].adj[adj_pt_id2];
                uint32_t o1 = next_ccw_pt_id(faces[adj1], faces[id].pts[adj_pt_id1]);
                uint32_t o2 = next_ccw_pt_id(faces[adj2], faces[id].pts[adj_pt_id2]);
                compute_silhouette(point, adj1, o1);
                compute_silhouette(point, adj2, o2);
            }
        }
    }

    // epa3.rs:219-430.  Returns false for `None`.
    bool closest_points(const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, const VoronoiSimplex& simplex, V3* out1,
                        V3* out2, V3* out_n, GJKStats* st) {
        const real _eps_tol = EPS * real(100);
        vertices.clear();
        faces.clear();
        heap.data.clear();
        silhouette.clear();
        for (int i = 0; i < simplex.dim + 1; ++i) vertices.push_back(simplex.vertices[i]);

#define EPA_PUSH(ID, ND)                           \
    {                                              \
        real nd__ = (ND);                          \
        if (nd__ > gjk_eps_tol()) return false;    \
        heap.push({(uint32_t)(ID), nd__});         \
    }
        if (simplex.dim == 0) {
            *out1 = v3(0, 0, 0);
            *out2 = v3(0, 0, 0);
            *out_n = v3(0, 1, 0);
            return true;
        } else if (simplex.dim == 3) {
            V3 dp1 = vertices[1].point - vertices[0].point;
            V3 dp2 = vertices[2].point - vertices[0].point;
            V3 dp3 = vertices[3].point - vertices[0].point;
            if (dot(cross(dp1, dp2), dp3) > real(0)) std::swap(vertices[1], vertices[2]);
            bool in1, in2, in3, in4;
            EpaFace f1 = face_new(0, 1, 2, 3, 1, 2, &in1);
            EpaFace f2 = face_new(1, 3, 2, 3, 2, 0, &in2);
            EpaFace f3 = face_new(0, 2, 3, 0, 1, 3, &in3);
            EpaFace f4 = face_new(0, 3, 1, 2, 1, 0, &in4);
            faces.push_back(f1);
            faces.push_back(f2);
            faces.push_back(f3);
            faces.push_back(f4);
            if (in1) EPA_PUSH(0, -dot(faces[0].normal, vertices[0].point));
            if (in2) EPA_PUSH(1, -dot(faces[1].normal, vertices[1].point));
            if (in3) EPA_PUSH(2, -dot(faces[2].normal, vertices[2].point));
            if (in4) EPA_PUSH(3, -dot(faces[3].normal, vertices[3].point));
        } else {
            if (simplex.dim == 1) {
                V3 dpt = vertices[1].point - vertices[0].point;
                V3 first, second;
                orthonormal_basis(dpt, &first, &second);
                vertices.push_back(cso_from_shapes(m1, g1, m2, g2, first));
            }
            bool in;
            EpaFace f1 = face_new(0, 1, 2, 1, 1, 1, &in);
            EpaFace f2 = face_new(0, 2, 1, 0, 0, 0, &in);
            faces.push_back(f1);
            faces.push_back(f2);
            EPA_PUSH(0, real(0));
            EPA_PUSH(1, real(0));
        }

        int niter = 0;
        real max_dist = FMAX;
        if (heap.data.empty()) {  // heap.peek().unwrap() panics in the reference
            panicked = true;
            return false;
        }
        EpaFaceId best_face_id = heap.data[0];
        EpaFaceId face_id;
        while (heap.pop(&face_id)) {
            EpaFace face = faces[face_id.id];
            if (face.deleted) continue;
            CSOPoint cso_point = cso_from_shapes(m1, g1, m2, g2, face.normal);
            uint32_t support_point_id = (uint32_t)vertices.size();
            vertices.push_back(cso_point);
            real candidate_max_dist = dot(cso_point.point, face.normal);
            if (candidate_max_dist < max_dist) {
                best_face_id = face_id;
                max_dist = candidate_max_dist;
            }
            real curr_dist = -face_id.neg_dist;
            if (max_dist - curr_dist < _eps_tol) {
                const EpaFace& bf = faces[best_face_id.id];
                face_closest_points(bf, out1, out2);
                *out_n = bf.normal;
                return true;
            }
            faces[face_id.id].deleted = true;
            uint32_t o1 = next_ccw_pt_id(faces[face.adj[0]], face.pts[0]);
            uint32_t o2 = next_ccw_pt_id(faces[face.adj[1]], face.pts[1]);
            uint32_t o3 = next_ccw_pt_id(faces[face.adj[2]], face.pts[2]);
            compute_silhouette(support_point_id, face.adj[0], o1);
            compute_silhouette(support_point_id, face.adj[1], o2);
            compute_silhouette(support_point_id, face.adj[2], o3);
            if (panicked) return false;
            uint32_t first_new_face_id = (uint32_t)faces.size();
            if (silhouette.empty()) return false;
            for (auto& edge : silhouette) {
                if (!faces[edge.first].deleted) {
                    uint32_t new_face_id = (uint32_t)faces.size();
                    EpaFace& face_adj = faces[edge.first];
                    uint32_t pt_id1 = face_adj.pts[(edge.second + 2) % 3];
                    uint32_t pt_id2 = face_adj.pts[(edge.second + 1) % 3];
                    bool inside;
                    EpaFace nf = face_new(pt_id1, pt_id2, support_point_id, edge.first, new_face_id + 1, new_face_id - 1, &inside);
                    faces[edge.first].adj[(edge.second + 1) % 3] = new_face_id;
                    faces.push_back(nf);
                    if (inside) {
                        V3 pt = vertices[faces[new_face_id].pts[0]].point;
                        real dist = dot(faces[new_face_id].normal, pt);
                        if (dist < curr_dist) {
                            face_closest_points(face, out1, out2);
                            *out_n = face.normal;
                            return true;
                        }
                        EPA_PUSH(new_face_id, -dist);
                    }
                }
            }
            if (first_new_face_id == faces.size()) return false;
            faces[first_new_face_id].adj[2] = (uint32_t)faces.size() - 1;
            faces.back().adj[1] = first_new_face_id;
            silhouette.clear();
            niter += 1;
            if (st) {
                st->epa_iters++;
                if (vertices.size() > st->epa_max_verts) st->epa_max_verts = (uint32_t)vertices.size();
                if (faces.size() > st->epa_max_faces) st->epa_max_faces = (uint32_t)faces.size();
                if (heap.data.size() > st->epa_max_heap) st->epa_max_heap = (uint32_t)heap.data.size();
            }
            if (niter > 10000) return false;
        }
#undef EPA_PUSH
        const EpaFace& bf = faces[best_face_id.id];
        face_closest_points(bf, out1, out2);
        *out_n = bf.normal;
        return true;
    }
};

// contact_support_map_support_map.rs:38-79
static inline GJKResult contact_support_map_support_map_with_params(const Iso& m1, const Support& g1, const Iso& m2, const Support& g2,
                                                                    real prediction, VoronoiSimplex& simplex, const V3* init_dir,
                                                                    GJKStats* st) {
    V3 dir;
    if (init_dir)
        dir = *init_dir;
    else if (!unit_try_new(m2.t - m1.t, EPS, &dir))
        dir = v3(1, 0, 0);
    simplex.reset(cso_from_shapes(m1, g1, m2, g2, dir));
    GJKResult cpts = gjk_closest_points(m1, g1, m2, g2, prediction, simplex, st);
    if (cpts.kind != GJK_INTERSECTION) return cpts;
    EPA epa;
    if (st) st->epa_calls++;
    V3 p1, p2, n;
    if (epa.closest_points(m1, g1, m2, g2, simplex, &p1, &p2, &n, st)) return {GJK_CLOSEST_POINTS, p1, p2, n};
    if (st) st->epa_fail++;
    return {GJK_NO_INTERSECTION, {}, {}, v3(1, 0, 0)};
}

}  // namespace orc
