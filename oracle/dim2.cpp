// ORACLE — TEST INFRASTRUCTURE ONLY (see na.hpp header).
// ncollide2d: `query::contact` between 2-D balls, cuboids and convex polygons, restated from the reference
// (the crate is the same source tree built with feature "dim2", build/ncollide2d/Cargo.toml):
//   query/contact/contact_shape_shape.rs:15-60 (dispatch), contact_ball_ball.rs:8-38, contact_ball_convex_polyhedron.rs:12-75,
//   contact_support_map_support_map.rs:9-79, query/algorithms/gjk.rs:76-177,367-388 (DIM = 2), voronoi_simplex2.rs:20-168,
//   epa2.rs:15-378, cso_point.rs:70-85, query/point/point_segment.rs:52-91, point_triangle.rs:60-305 (dim2 branches),
//   point_aabb.rs:14-135 + point_cuboid.rs:17-26, shape/cuboid.rs:137-145,469-503 (support point, 2-D feature normals),
//   shape/ball.rs:29-48, shape/convex_polygon.rs (support point = utils/point_cloud_support_point.rs:6-24), utils/ccw_face_normal.rs:8-14.
// nalgebra's Isometry2 = Translation2 * UnitComplex: rotation * v = (re x - im y, im x + re y) (unit_complex_ops.rs); not vendored, restated.
// PINNED on the reference's own 2-D known-answer tests: build/ncollide2d/tests/geometry/epa2.rs (cuboid / cuboid: depth == 0.5 / 1.8 and
// normal == -x / -y exactly; issue #181 does not panic) and ball_cuboid_contact.rs (f32 and f64) — tests/test_dim2.py.
#include <algorithm>
#include <vector>
#include "na.hpp"
#include "oracle.h"

namespace orc {
namespace d2 {

struct P2 {
    real x, y;
};
static inline P2 p2(real x, real y) { return P2{x, y}; }
static inline P2 operator+(P2 a, P2 b) { return {a.x + b.x, a.y + b.y}; }
static inline P2 operator-(P2 a, P2 b) { return {a.x - b.x, a.y - b.y}; }
static inline P2 operator-(P2 a) { return {-a.x, -a.y}; }
static inline P2 operator*(P2 a, real s) { return {a.x * s, a.y * s}; }
static inline P2 operator/(P2 a, real s) { return {a.x / s, a.y / s}; }
static inline real dot(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }
static inline real perp(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
static inline real nsq(P2 a) { return dot(a, a); }
static inline bool unit_try_new_and_get(P2 a, real min_norm, P2* out, real* n_out) {
    real sq = nsq(a);
    if (sq > min_norm * min_norm) {
        real n = std::sqrt(sq);
        *out = a / n;
        *n_out = n;
        return true;
    }
    return false;
}
static inline bool unit_try_new(P2 a, real min_norm, P2* out) {
    real n;
    return unit_try_new_and_get(a, min_norm, out, &n);
}
static inline P2 normalize(P2 a) { return a / std::sqrt(nsq(a)); }

struct Iso2 {
    P2 t;
    real re, im;
};
static inline P2 rot(const Iso2& m, P2 v) { return {m.re * v.x - m.im * v.y, m.im * v.x + m.re * v.y}; }
static inline P2 inv_rot(const Iso2& m, P2 v) { return {m.re * v.x + m.im * v.y, -m.im * v.x + m.re * v.y}; }  // conjugate: im -> -im
static inline P2 mul_point(const Iso2& m, P2 p) { return rot(m, p) + m.t; }
static inline P2 inv_point(const Iso2& m, P2 p) { return inv_rot(m, p - m.t); }

enum { BALL2 = 0, CUBOID2 = 1, POLYGON2 = 2, PLANE2 = 3, SEGMENT2 = 4, ORIGIN2 = 7 };  // SEGMENT2: param = a.x a.y b.x b.y (shape/segment.rs)  // PLANE2: param = unit normal; ORIGIN2: special_support_maps::ConstantOrigin
struct Shape2 {
    uint32_t type;
    real radius;
    P2 he;
    P2 sb = {0, 0};  // SEGMENT2: a = he, b = sb
    const real* pts;
    const real* normals;  // ConvexPolygon::normals (one per edge i -> i + 1), from try_new
    uint32_t npts;
};

// SupportMap::support_point
static P2 support_point(const Shape2& g, const Iso2& m, P2 dir) {
    if (g.type == ORIGIN2) return m.t;  // ConstantOrigin: m * Point::origin()
    if (g.type == BALL2) {  // ball.rs:29-48: support_point_toward(m, Unit::new_normalize(dir)) = translation + dir * radius
        P2 d = normalize(dir);
        return m.t + d * g.radius;
    }
    P2 ld = inv_rot(m, dir);
    P2 lp;
    if (g.type == CUBOID2) {
        lp = p2(std::copysign(g.he.x, ld.x), std::copysign(g.he.y, ld.y));
    } else if (g.type == SEGMENT2) {  // segment.rs:182-191: a if a . dir > b . dir, else b
        lp = dot(g.he, ld) > dot(g.sb, ld) ? g.he : g.sb;
    } else {  // point_cloud_support_point: first maximum
        uint32_t best = 0;
        real best_dot = g.pts[0] * ld.x + g.pts[1] * ld.y;
        for (uint32_t i = 1; i < g.npts; ++i) {
            real d = g.pts[2 * i] * ld.x + g.pts[2 * i + 1] * ld.y;
            if (d > best_dot) best_dot = d, best = i;
        }
        lp = p2(g.pts[2 * best], g.pts[2 * best + 1]);
    }
    return mul_point(m, lp);
}

struct CSO {
    P2 point, orig1, orig2;
};
static CSO cso_from_shapes(const Iso2& m1, const Shape2& g1, const Iso2& m2, const Shape2& g2, P2 dir) {
    CSO c;
    c.orig1 = support_point(g1, m1, dir);
    c.orig2 = support_point(g2, m2, -dir);
    c.point = c.orig1 - c.orig2;
    return c;
}

static const real EPS_TOL = EPS * real(10);  // gjk::eps_tol()

// ---- VoronoiSimplex (voronoi_simplex2.rs) -----------------------------------------------------------------------------------
struct Simplex2 {
    int prev_vertices[3] = {0, 1, 2};
    int prev_dim = 0;
    real prev_proj[2] = {0, 0};
    CSO vertices[3];
    real proj[2] = {0, 0};
    int dim = 0;
    Simplex2() {
        for (auto& v : vertices) v.point = v.orig1 = v.orig2 = p2(0, 0);
    }
    void swap(int a, int b) {
        std::swap(vertices[a], vertices[b]);
        std::swap(prev_vertices[a], prev_vertices[b]);
    }
    void reset(const CSO& pt) {
        prev_dim = 0, dim = 0;
        vertices[0] = pt;
    }
    bool add_point(const CSO& pt) {
        prev_dim = dim;
        prev_proj[0] = proj[0], prev_proj[1] = proj[1];
        prev_vertices[0] = 0, prev_vertices[1] = 1, prev_vertices[2] = 2;
        for (int i = 0; i < dim + 1; ++i)
            if (nsq(vertices[i].point - pt.point) < EPS_TOL) return false;
        dim += 1;
        vertices[dim] = pt;
        return true;
    }
    P2 project_origin_and_reduce() {
        const P2 O = p2(0, 0);
        if (dim == 0) {
            proj[0] = 1;
            return vertices[0].point;
        }
        if (dim == 1) {  // Segment::project_point_with_location (point_segment.rs:52-91), identity isometry
            P2 a = vertices[0].point, b = vertices[1].point;
            P2 ab = b - a, ap = O - a;
            real ab_ap = dot(ab, ap), sqnab = nsq(ab);
            if (ab_ap <= 0) {
                proj[0] = 1, dim = 0;
                return a;
            }
            if (ab_ap >= sqnab) {
                proj[0] = 1;
                swap(0, 1);
                dim = 0;
                return b;
            }
            real u = ab_ap / sqnab;
            proj[0] = real(1) - u, proj[1] = u;
            return a + ab * u;
        }
        // Triangle::project_point_with_location, dim2 (point_triangle.rs:60-250), solid = true
        P2 a = vertices[0].point, b = vertices[1].point, c = vertices[2].point;
        P2 ab = b - a, ac = c - a, ap = O - a;
        real ab_ap = dot(ab, ap), ac_ap = dot(ac, ap);
        if (ab_ap <= 0 && ac_ap <= 0) {
            swap(0, 0), proj[0] = 1, dim = 0;
            return a;
        }
        P2 bp = O - b;
        real ab_bp = dot(ab, bp), ac_bp = dot(ac, bp);
        if (ab_bp >= 0 && ac_bp <= ab_bp) {
            swap(0, 1), proj[0] = 1, dim = 0;
            return b;
        }
        P2 cp = O - c;
        real ab_cp = dot(ab, cp), ac_cp = dot(ac, cp);
        if (ac_cp >= 0 && ab_cp <= ac_cp) {
            swap(0, 2), proj[0] = 1, dim = 0;
            return c;
        }
        P2 bc = c - b;
        real n = perp(ab, ac);
        real vc = n * perp(ab, ap);
        if (vc < 0 && ab_ap >= 0 && ab_bp <= 0) {  // OnEdge(0)
            real v = ab_ap / nsq(ab);
            proj[0] = real(1) - v, proj[1] = v;
            dim = 1;
            return a + ab * v;
        }
        real vb = -n * perp(ac, cp);
        if (vb < 0 && ac_ap >= 0 && ac_cp <= 0) {  // OnEdge(2): swap(1, 2), proj = coords
            real w = ac_ap / nsq(ac);
            swap(1, 2);
            proj[0] = real(1) - w, proj[1] = w;
            dim = 1;
            return a + ac * w;
        }
        real va = n * perp(bc, bp);
        if (va < 0 && ac_bp - ab_bp >= 0 && ab_cp - ac_cp >= 0) {  // OnEdge(1): swap(0, 2), proj = [coords[1], coords[0]]
            real w = dot(bc, bp) / nsq(bc);
            swap(0, 2);
            proj[0] = w, proj[1] = real(1) - w;
            dim = 1;
            return b + bc * w;
        }
        return O;  // OnFace in 2-D + solid: the point itself (OnSolid); the simplex keeps dimension 2
    }
};

// gjk.rs:367-388
static void gjk_result(const Simplex2& s, bool prev, P2* p1, P2* p2_) {
    P2 r0 = p2(0, 0), r1 = p2(0, 0);
    if (prev) {
        for (int i = 0; i < s.prev_dim + 1; ++i) {
            real coord = s.prev_proj[i];
            const CSO& pt = s.vertices[s.prev_vertices[i]];
            r0 = r0 + pt.orig1 * coord;
            r1 = r1 + pt.orig2 * coord;
        }
    } else {
        for (int i = 0; i < s.dim + 1; ++i) {
            real coord = s.proj[i];
            r0 = r0 + s.vertices[i].orig1 * coord;
            r1 = r1 + s.vertices[i].orig2 * coord;
        }
    }
    *p1 = r0, *p2_ = r1;
}

enum { R_INTERSECTION = 0, R_CLOSEST = 1, R_NONE = 3 };
// gjk::closest_points, exact_dist = true (gjk.rs:76-177), DIM = 2
static int gjk_closest_points(const Iso2& m1, const Shape2& g1, const Iso2& m2, const Shape2& g2, real max_dist, Simplex2& s, P2* p1, P2* p2_,
                              P2* out_dir) {
    const real eps_rel = std::sqrt(EPS_TOL);
    P2 proj = s.project_origin_and_reduce();
    P2 old_dir, pd;
    if (!unit_try_new(proj, 0, &pd)) return R_INTERSECTION;
    old_dir = -pd;
    real max_bound = FMAX;
    P2 dir;
    int niter = 0;
    for (;;) {
        real old_max_bound = max_bound, dist;
        if (!unit_try_new_and_get(-proj, EPS_TOL, &dir, &dist)) return R_INTERSECTION;
        max_bound = dist;
        if (max_bound >= old_max_bound) {
            gjk_result(s, true, p1, p2_);
            *out_dir = old_dir;
            return R_CLOSEST;
        }
        CSO cso = cso_from_shapes(m1, g1, m2, g2, dir);
        real min_bound = -dot(dir, cso.point);
        if (min_bound > max_dist) {
            *out_dir = dir;
            return R_NONE;
        } else if (max_bound - min_bound <= eps_rel * max_bound) {
            gjk_result(s, false, p1, p2_);
            *out_dir = dir;
            return R_CLOSEST;
        }
        if (!s.add_point(cso)) {
            gjk_result(s, false, p1, p2_);
            *out_dir = dir;
            return R_CLOSEST;
        }
        old_dir = dir;
        proj = s.project_origin_and_reduce();
        if (s.dim == 2) {
            if (min_bound >= EPS_TOL) {
                gjk_result(s, true, p1, p2_);
                *out_dir = old_dir;
                return R_CLOSEST;
            }
            return R_INTERSECTION;
        }
        if (++niter == 10000) {
            *out_dir = p2(1, 0);
            return R_NONE;
        }
    }
}

// ---- EPA (epa2.rs) ----------------------------------------------------------------------------------------------------------
struct FaceId2 {
    size_t id;
    real neg_dist;
};
struct Heap2 {  // Rust std BinaryHeap<FaceId> (max-heap on neg_dist; Ord::cmp via <, >)
    std::vector<FaceId2> data;
    static bool le(const FaceId2& a, const FaceId2& b) { return a.neg_dist <= b.neg_dist; }  // PartialOrd `<=` (sift_up / sift_down_to_bottom)
    void sift_up(size_t start, size_t pos) {
        FaceId2 elt = data[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (le(elt, data[parent])) break;
            data[pos] = data[parent];
            pos = parent;
        }
        data[pos] = elt;
    }
    void push(FaceId2 f) {
        data.push_back(f);
        sift_up(0, data.size() - 1);
    }
    bool pop(FaceId2* out) {
        if (data.empty()) return false;
        FaceId2 item = data.back();
        data.pop_back();
        if (!data.empty()) {
            std::swap(item, data[0]);
            size_t end = data.size(), pos = 0, child = 1;
            FaceId2 elt = data[0];
            while (end >= 2 && child <= end - 2) {
                if (le(data[child], data[child + 1])) child += 1;
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
struct Face2 {
    size_t pts[2];
    P2 normal, proj;
    real bcoords[2];
    bool deleted;
};
// epa2.rs:352-378
static bool project_origin_seg(P2 a, P2 b, P2* res, real bc[2]) {
    P2 ab = b - a, ap = -a;
    real ab_ap = dot(ab, ap), sqnab = nsq(ab);
    if (sqnab == 0) return false;
    if (ab_ap < -EPS_TOL || ab_ap > sqnab + EPS_TOL) return false;
    real pos = ab_ap / sqnab;
    *res = a + ab * pos;
    bc[0] = real(1) - pos, bc[1] = pos;
    return true;
}
static Face2 face_new_with_proj(const std::vector<CSO>& v, P2 proj, const real bc[2], size_t p0, size_t p1) {
    Face2 f;
    f.pts[0] = p0, f.pts[1] = p1, f.proj = proj, f.bcoords[0] = bc[0], f.bcoords[1] = bc[1];
    P2 ab = v[p1].point - v[p0].point;  // ccw_face_normal (dim2): (ab.y, -ab.x) normalised
    if (unit_try_new(p2(ab.y, -ab.x), EPS, &f.normal)) {
        f.deleted = false;
    } else {
        f.normal = p2(0, 0);
        f.deleted = true;
    }
    return f;
}
static Face2 face_new(const std::vector<CSO>& v, size_t p0, size_t p1, bool* inside) {
    P2 proj;
    real bc[2];
    if (project_origin_seg(v[p0].point, v[p1].point, &proj, bc)) {
        *inside = true;
        return face_new_with_proj(v, proj, bc, p0, p1);
    }
    real z[2] = {0, 0};
    *inside = false;
    return face_new_with_proj(v, p2(0, 0), z, p0, p1);
}
static void face_closest_points(const Face2& f, const std::vector<CSO>& v, P2* a, P2* b) {
    *a = v[f.pts[0]].orig1 * f.bcoords[0] + v[f.pts[1]].orig1 * f.bcoords[1];
    *b = v[f.pts[0]].orig2 * f.bcoords[0] + v[f.pts[1]].orig2 * f.bcoords[1];
}
// FaceId::new: None when neg_dist > eps_tol (the `?` in the callers returns None from closest_points)
#define PUSH_OR_FAIL(ID, ND)              \
    do {                                  \
        if ((ND) > EPS_TOL) return false; \
        heap.push(FaceId2{(ID), (ND)});   \
    } while (0)

static bool epa_closest_points(const Iso2& m1, const Shape2& g1, const Iso2& m2, const Shape2& g2, const Simplex2& simplex, P2* o1, P2* o2,
                               P2* on, int* panicked) {
    const real eps_tol = EPS * real(100);
    std::vector<CSO> vertices;
    std::vector<Face2> faces;
    Heap2 heap;
    for (int i = 0; i < simplex.dim + 1; ++i) vertices.push_back(simplex.vertices[i]);
    if (simplex.dim == 0) {
        P2 n = p2(0, 1);
        P2 orig1 = vertices[0].orig1;
        for (int it = 0; it < 100; ++it) {
            P2 supp1 = support_point(g1, m1, n), tangent;
            if (unit_try_new(supp1 - orig1, eps_tol, &tangent)) {
                if (dot(n, tangent) < eps_tol) break;
                n = p2(-tangent.y, tangent.x);
            } else
                break;
        }
        P2 orig2 = vertices[0].orig2;
        for (int it = 0; it < 100; ++it) {
            P2 supp2 = support_point(g2, m2, -n), tangent;
            if (unit_try_new(supp2 - orig2, eps_tol, &tangent)) {
                if (dot(-n, tangent) < eps_tol) break;
                n = p2(-tangent.y, tangent.x);
            } else
                break;
        }
        *o1 = p2(0, 0), *o2 = p2(0, 0), *on = n;
        return true;
    } else if (simplex.dim == 2) {
        P2 dp1 = vertices[1].point - vertices[0].point, dp2 = vertices[2].point - vertices[0].point;
        if (perp(dp1, dp2) < 0) std::swap(vertices[1], vertices[2]);
        bool in1, in2, in3;
        Face2 f1 = face_new(vertices, 0, 1, &in1), f2 = face_new(vertices, 1, 2, &in2), f3 = face_new(vertices, 2, 0, &in3);
        faces.push_back(f1), faces.push_back(f2), faces.push_back(f3);
        if (in1) PUSH_OR_FAIL(0, -dot(faces[0].normal, vertices[0].point));
        if (in2) PUSH_OR_FAIL(1, -dot(faces[1].normal, vertices[1].point));
        if (in3) PUSH_OR_FAIL(2, -dot(faces[2].normal, vertices[2].point));
    } else {
        real one_zero[2] = {1, 0};
        faces.push_back(face_new_with_proj(vertices, p2(0, 0), one_zero, 0, 1));
        faces.push_back(face_new_with_proj(vertices, p2(0, 0), one_zero, 1, 0));
        real dist1 = dot(faces[0].normal, vertices[0].point), dist2 = dot(faces[1].normal, vertices[1].point);
        PUSH_OR_FAIL(0, dist1);
        PUSH_OR_FAIL(1, dist2);
    }
    int niter = 0;
    real max_dist = FMAX;
    if (heap.data.empty()) {  // heap.peek().unwrap() on an empty heap: the reference panics here
        *panicked = 1;
        return false;
    }
    FaceId2 best_face_id = heap.data[0];
    FaceId2 face_id;
    while (heap.pop(&face_id)) {
        Face2 face = faces[face_id.id];
        if (face.deleted) continue;
        CSO cso = cso_from_shapes(m1, g1, m2, g2, face.normal);
        size_t support_point_id = vertices.size();
        vertices.push_back(cso);
        real candidate_max_dist = dot(cso.point, face.normal);
        if (candidate_max_dist < max_dist) best_face_id = face_id, max_dist = candidate_max_dist;
        real curr_dist = -face_id.neg_dist;
        if (max_dist - curr_dist < eps_tol) {
            const Face2& bf = faces[best_face_id.id];
            face_closest_points(bf, vertices, o1, o2);
            *on = bf.normal;
            return true;
        }
        bool in[2];
        Face2 nf[2] = {face_new(vertices, face.pts[0], support_point_id, &in[0]), face_new(vertices, support_point_id, face.pts[1], &in[1])};
        for (int k = 0; k < 2; ++k) {
            if (in[k]) {
                real dist = dot(nf[k].normal, nf[k].proj);
                if (dist < curr_dist) {
                    face_closest_points(nf[k], vertices, o1, o2);
                    *on = nf[k].normal;
                    return true;
                }
                if (!nf[k].deleted) PUSH_OR_FAIL(faces.size(), -dist);
            }
            faces.push_back(nf[k]);
        }
        if (++niter > 10000) return false;
    }
    const Face2& bf = faces[best_face_id.id];
    face_closest_points(bf, vertices, o1, o2);
    *on = bf.normal;
    return true;
}

struct Contact2 {
    P2 w1, w2, n;
    real depth;
};

// contact_support_map_support_map (contact_support_map_support_map.rs:9-79)
static bool contact_sm_sm(const Iso2& m1, const Shape2& g1, const Iso2& m2, const Shape2& g2, real prediction, Contact2* c, int* panicked) {
    P2 dir;
    if (!unit_try_new(m2.t - m1.t, EPS, &dir)) dir = p2(1, 0);
    Simplex2 s;
    s.reset(cso_from_shapes(m1, g1, m2, g2, dir));
    P2 p1, p2_, n;
    int r = gjk_closest_points(m1, g1, m2, g2, prediction, s, &p1, &p2_, &n);
    if (r == R_NONE) return false;
    if (r == R_INTERSECTION) {
        if (!epa_closest_points(m1, g1, m2, g2, s, &p1, &p2_, &n, panicked)) return false;  // NoIntersection(x axis)
    }
    c->w1 = p1, c->w2 = p2_, c->n = n;
    c->depth = -dot(n, p2_ - p1);  // Contact::new_wo_depth
    return true;
}

// contact_ball_ball.rs:8-38
static bool contact_ball_ball(P2 c1, real r1, P2 c2, real r2, real prediction, Contact2* c) {
    P2 delta = c2 - c1;
    real dsq = nsq(delta), sum = r1 + r2, sum_err = sum + prediction;
    if (dsq < sum_err * sum_err) {
        P2 n = dsq != 0 ? normalize(delta) : p2(1, 0);
        c->w1 = c1 + n * r1, c->w2 = c2 + n * (-r2), c->n = n, c->depth = sum - std::sqrt(dsq);
        return true;
    }
    return false;
}

// contact_plane_support_map (contact_plane_support_map.rs:8-28): `other.support_point_toward(mother, -plane_normal)`
static bool contact_plane_sm(const Iso2& mp, P2 plane_n_local, const Iso2& mo, const Shape2& other, real prediction, Contact2* c) {
    P2 n = rot(mp, plane_n_local);
    P2 deepest = other.type == BALL2 ? mo.t + (-n) * other.radius  // Ball::support_point_toward: the unit direction is used as is
                                     : support_point(other, mo, -n);
    real distance = dot(n, mp.t - deepest);
    if (distance > -prediction) {
        c->w1 = deepest + n * distance, c->w2 = deepest, c->n = n, c->depth = distance;
        return true;
    }
    return false;
}

enum { F_UNKNOWN = 0xffffffffu, F_FACE = 0x40000000u, F_VERTEX = 0x80000000u };
// Cuboid::project_point_with_feature -> AABB (point_aabb.rs:14-135), dim2
static P2 cuboid_project(const Shape2& g, const Iso2& m, P2 pt, bool* inside_out, uint32_t* feature) {
    P2 mins = -g.he, maxs = g.he;
    P2 ls = inv_point(m, pt);
    real mp[2] = {mins.x - ls.x, mins.y - ls.y}, pm[2] = {ls.x - maxs.x, ls.y - maxs.y};
    real shift[2];
    for (int i = 0; i < 2; ++i) shift[i] = std::fmax(mp[i], real(0)) - std::fmax(pm[i], real(0));
    bool inside = shift[0] == 0 && shift[1] == 0;
    real lp[2] = {ls.x, ls.y};
    if (!inside) {
        lp[0] += shift[0], lp[1] += shift[1];
    } else {  // solid = false
        real best = -FMAX;
        bool is_mins = false;
        int best_id = 0;
        for (int i = 0; i < 2; ++i) {
            if (mp[i] < pm[i]) {
                if (pm[i] > best) best_id = i, is_mins = false, best = pm[i];
            } else if (mp[i] > best) {
                best_id = i, is_mins = true, best = mp[i];
            }
        }
        shift[0] = shift[1] = 0;
        shift[best_id] = is_mins ? best : -best;
        lp[0] += shift[0], lp[1] += shift[1];
    }
    *inside_out = inside;
    P2 proj = mul_point(m, p2(lp[0], lp[1]));
    int nzero = 0, last_not_zero = 0;
    for (int i = 0; i < 2; ++i) {
        if (shift[i] == 0)
            nzero++;
        else
            last_not_zero = i;
    }
    real mn[2] = {mins.x, mins.y}, mx[2] = {maxs.x, maxs.y};
    if (nzero == 2) {
        *feature = F_UNKNOWN;
        for (int i = 0; i < 2; ++i) {
            if (lp[i] > mx[i] - EPS) {
                *feature = F_FACE | (uint32_t)i;
                break;
            }
            if (lp[i] <= mn[i] + EPS) {
                *feature = F_FACE | (uint32_t)(i + 2);
                break;
            }
        }
    } else if (nzero == 1) {
        real center = (mn[last_not_zero] + mx[last_not_zero]) * real(0.5);  // na::center(mins, maxs)
        *feature = F_FACE | (uint32_t)(lp[last_not_zero] < center ? last_not_zero + 2 : last_not_zero);
    } else {
        uint32_t id = 0;
        for (int i = 0; i < 2; ++i) {
            real center = (mn[i] + mx[i]) * real(0.5);
            if (lp[i] < center) id |= 1u << i;
        }
        *feature = F_VERTEX | id;
    }
    return proj;
}
// cuboid.rs:469-503 (dim2)
static P2 cuboid_feature_normal(uint32_t f) {
    uint32_t id = f & 0xffffu;
    if (f & F_FACE) {
        real d[2] = {0, 0};
        if (id < 2)
            d[id] = 1;
        else
            d[id - 2] = -1;
        return p2(d[0], d[1]);
    }
    P2 d = p2(0, 0);
    switch (id) {
        case 0: d = p2(1, 1); break;
        case 1: d = p2(-1, 1); break;
        case 3: d = p2(-1, -1); break;
        default: d = p2(1, -1); break;
    }
    return normalize(d);
}
// Segment::project_point_with_feature (query/point/point_segment.rs:14-91, dim2): is_inside = relative_eq!(proj, pt)
static P2 segment_project(const Shape2& g, const Iso2& m, P2 pt, bool* inside, uint32_t* feature) {
    P2 ls = inv_point(m, pt), a = g.he, b = g.sb;
    P2 ab = b - a, ap = ls - a;
    real ab_ap = dot(ab, ap), sqnab = nsq(ab);
    P2 proj;
    bool on_edge = false;
    if (ab_ap <= 0) {
        *feature = F_VERTEX | 0u, proj = mul_point(m, a);
    } else if (ab_ap >= sqnab) {
        *feature = F_VERTEX | 1u, proj = mul_point(m, b);
    } else {
        real u = ab_ap / sqnab;
        proj = mul_point(m, a + ab * u);
        on_edge = true;
    }
    *inside = relative_eq(proj.x, pt.x) && relative_eq(proj.y, pt.y);
    if (on_edge) {  // dim2: the side of the segment the point is on
        P2 dpt = pt - proj;
        *feature = perp(dpt, ab) >= 0 ? (F_FACE | 0u) : (F_FACE | 1u);
    }
    return proj;
}
// Segment::feature_normal (segment.rs:237-284, dim2); no direction (a == b): the y axis
static P2 segment_feature_normal(const Shape2& g, uint32_t f) {
    P2 dir;
    if (!unit_try_new(g.sb - g.he, EPS, &dir)) return p2(0, 1);
    uint32_t id = f & 0xffffu;
    if (f & F_VERTEX) return id == 0 ? dir : -dir;
    return id == 0 ? p2(dir.y, -dir.x) : p2(-dir.y, dir.x);
}
// contact_ball_convex_polyhedron.rs:12-62 with a segment
static bool contact_ball_segment(P2 center, real radius, const Iso2& m2, const Shape2& g2, real prediction, Contact2* c) {
    bool inside;
    uint32_t f2;
    P2 world2 = segment_project(g2, m2, center, &inside, &f2);
    P2 dpt = world2 - center, dir, normal;
    real dist, depth;
    if (unit_try_new_and_get(dpt, EPS, &dir, &dist)) {
        if (inside)
            depth = dist + radius, normal = -dir;
        else
            depth = -dist + radius, normal = dir;
    } else {
        depth = radius;
        normal = -segment_feature_normal(g2, f2);
    }
    if (depth >= -prediction) {
        c->w1 = center + normal * radius, c->w2 = world2, c->n = normal, c->depth = depth;
        return true;
    }
    return false;
}

// contact_ball_convex_polyhedron.rs:12-62 with a cuboid
static bool contact_ball_cuboid(P2 center, real radius, const Iso2& m2, const Shape2& g2, real prediction, Contact2* c) {
    bool inside;
    uint32_t f2;
    P2 world2 = cuboid_project(g2, m2, center, &inside, &f2);
    P2 dpt = world2 - center, dir, normal;
    real dist, depth;
    if (unit_try_new_and_get(dpt, EPS, &dir, &dist)) {
        if (inside)
            depth = dist + radius, normal = -dir;
        else
            depth = -dist + radius, normal = dir;
    } else {
        if (f2 == F_UNKNOWN) return false;
        depth = radius;
        normal = -cuboid_feature_normal(f2);  // as in the reference: the LOCAL feature normal is used as is (:50), not rotated by m2
    }
    if (depth >= -prediction) {
        c->w1 = center + normal * radius, c->w2 = world2, c->n = normal, c->depth = depth;
        return true;
    }
    return false;
}

// point_projection_on_support_map (point_support_map.rs:14-55) with solid = false, for a convex polygon
static P2 polygon_project(const Shape2& g, const Iso2& m_in, P2 point, bool* inside, int* panicked) {
    Iso2 m = m_in;
    m.t = (-point) + m_in.t;  // Translation::from(-point.coords) * m
    Iso2 id = {p2(0, 0), 1, 0};
    Shape2 origin;
    origin.type = ORIGIN2, origin.radius = 0, origin.he = p2(0, 0), origin.pts = origin.normals = nullptr, origin.npts = 0;
    P2 dir;
    if (!unit_try_new(-m.t, EPS, &dir)) dir = p2(1, 0);
    Simplex2 s;
    s.reset(cso_from_shapes(m, g, id, origin, dir));
    P2 p1, p2_, n;
    int r = gjk_closest_points(m, g, id, origin, FMAX, s, &p1, &p2_, &n);  // gjk::project_origin
    if (r == R_CLOSEST) {
        *inside = false;
        return p1 + point;
    }
    *inside = true;
    if (epa_closest_points(m, g, id, origin, s, &p1, &p2_, &n, panicked)) return p1 + point;  // EPA::project_origin
    return point;
}
// ConvexPolygon::support_feature_id_toward (convex_polygon.rs:186-203) / feature_normal (:139-152)
static uint32_t polygon_feature_toward(const Shape2& g, P2 local_dir) {
    const real ceps = std::cos(real(3.14159265358979323846 / 180.0));
    for (uint32_t i = 0; i < g.npts; ++i)
        if (g.normals[2 * i] * local_dir.x + g.normals[2 * i + 1] * local_dir.y >= ceps) return F_FACE | i;
    uint32_t best = 0;
    real best_dot = g.pts[0] * local_dir.x + g.pts[1] * local_dir.y;
    for (uint32_t i = 1; i < g.npts; ++i) {
        real d = g.pts[2 * i] * local_dir.x + g.pts[2 * i + 1] * local_dir.y;
        if (d > best_dot) best_dot = d, best = i;
    }
    return F_VERTEX | best;
}
static P2 polygon_feature_normal(const Shape2& g, uint32_t f) {
    uint32_t id = f & 0xffffu;
    if (f & F_FACE) return p2(g.normals[2 * id], g.normals[2 * id + 1]);
    uint32_t id1 = id == 0 ? g.npts - 1 : id - 1;
    return normalize(p2(g.normals[2 * id1], g.normals[2 * id1 + 1]) + p2(g.normals[2 * id], g.normals[2 * id + 1]));
}
// contact_ball_convex_polyhedron.rs:12-62 with a convex polygon (ConvexPolygon::project_point_with_feature, point_support_map.rs:120-146)
static bool contact_ball_polygon(P2 center, real radius, const Iso2& m2, const Shape2& g2, real prediction, Contact2* c, int* panicked) {
    bool inside;
    P2 world2 = polygon_project(g2, m2, center, &inside, panicked);
    P2 dpt_f = center - world2;
    P2 local_dir = inv_rot(m2, inside ? -dpt_f : dpt_f), ld;
    uint32_t f2 = F_UNKNOWN;
    if (unit_try_new(local_dir, EPS, &ld)) f2 = polygon_feature_toward(g2, ld);
    P2 dpt = world2 - center, dir, normal;
    real dist, depth;
    if (unit_try_new_and_get(dpt, EPS, &dir, &dist)) {
        if (inside)
            depth = dist + radius, normal = -dir;
        else
            depth = -dist + radius, normal = dir;
    } else {
        if (f2 == F_UNKNOWN) return false;
        depth = radius;
        normal = -polygon_feature_normal(g2, f2);
    }
    if (depth >= -prediction) {
        c->w1 = center + normal * radius, c->w2 = world2, c->n = normal, c->depth = depth;
        return true;
    }
    return false;
}

// ---- query::proximity in 2-D (query/proximity/proximity_shape_shape.rs:8-33) -----------------------------------------------------
enum { PX_INTERSECTING = 0, PX_WITHIN_MARGIN = 1, PX_DISJOINT = 2 };
// proximity_ball_ball.rs:8-36
static int proximity_ball_ball(P2 c1, real r1, P2 c2, real r2, real margin) {
    real dsq = nsq(c2 - c1), sum = r1 + r2, sum_err = sum + margin;
    if (dsq <= sum_err * sum_err) return dsq <= sum * sum ? PX_INTERSECTING : PX_WITHIN_MARGIN;
    return PX_DISJOINT;
}
// proximity_plane_support_map.rs:9-47
static int proximity_plane_sm(const Iso2& mp, P2 plane_n_local, const Iso2& mo, const Shape2& other, real margin) {
    P2 n = rot(mp, plane_n_local);
    P2 deepest = other.type == BALL2 ? mo.t + (-n) * other.radius : support_point(other, mo, -n);
    real distance = dot(n, mp.t - deepest);
    if (distance >= -margin) return distance >= 0 ? PX_INTERSECTING : PX_WITHIN_MARGIN;
    return PX_DISJOINT;
}
// proximity_support_map_support_map (proximity_support_map_support_map.rs:12-75) = gjk::closest_points with exact_dist = false
static int proximity_sm_sm(const Iso2& m1, const Shape2& g1, const Iso2& m2, const Shape2& g2, real max_dist) {
    const real eps_rel = std::sqrt(EPS_TOL);
    P2 dir;
    if (!unit_try_new(m2.t - m1.t, EPS, &dir)) dir = p2(1, 0);
    Simplex2 s;
    s.reset(cso_from_shapes(m1, g1, m2, g2, dir));
    P2 proj = s.project_origin_and_reduce(), pd;
    if (!unit_try_new(proj, 0, &pd)) return PX_INTERSECTING;
    real max_bound = FMAX;
    int niter = 0;
    for (;;) {
        real old_max_bound = max_bound, dist;
        if (!unit_try_new_and_get(-proj, EPS_TOL, &dir, &dist)) return PX_INTERSECTING;
        max_bound = dist;
        if (max_bound >= old_max_bound) return PX_WITHIN_MARGIN;
        CSO cso = cso_from_shapes(m1, g1, m2, g2, dir);
        real min_bound = -dot(dir, cso.point);
        if (min_bound > max_dist) return PX_DISJOINT;
        if (min_bound > 0 && max_bound <= max_dist) return PX_WITHIN_MARGIN;
        if (max_bound - min_bound <= eps_rel * max_bound) return PX_WITHIN_MARGIN;
        if (!s.add_point(cso)) return PX_WITHIN_MARGIN;
        proj = s.project_origin_and_reduce();
        if (s.dim == 2) return min_bound >= EPS_TOL ? PX_WITHIN_MARGIN : PX_INTERSECTING;
        if (++niter == 10000) return PX_DISJOINT;
    }
}

// =============================================================================================================================
// World update in 2-D: AABBs (bounding_volume/aabb_ball.rs:8-13, aabb_cuboid.rs:9-14, aabb_convex_polygon.rs + aabb_utils.rs:59-79,
// loosened like pipeline/object/collision_object.rs:89-93 + dbvt_broad_phase.rs:341) and the contact-manifold generators
// (ball_ball_manifold_generator.rs, ball_convex_polyhedron_manifold_generator.rs, convex_polyhedron_convex_polyhedron_manifold_generator.rs
// with shape/convex_polygonal_feature2.rs and the dim2 branches of shape/cuboid.rs / shape/convex_polygon.rs), pushed into a fresh
// ContactManifold with DistanceBased(0.02) tracking (query/contact/contact_manifold.rs:165-236).
// =============================================================================================================================
struct Feature2 {  // ConvexPolygonalFeature (2-D)
    P2 v[2];
    int nv = 0;
    bool has_normal = false;
    P2 normal = {0, 0};
    uint32_t fid = F_UNKNOWN, vid[2] = {F_UNKNOWN, F_UNKNOWN};
    void clear() { nv = 0, has_normal = false, fid = F_UNKNOWN; }
    void push(P2 p, uint32_t id) { v[nv] = p, vid[nv] = id, nv++; }
    void transform_by(const Iso2& m) {
        for (auto& p : v) p = mul_point(m, p);  // both slots, like the reference's loop over the array
        if (has_normal) normal = rot(m, normal);
    }
};
// Cuboid::face (cuboid.rs:186-226, dim2)
static void cuboid_face(const Shape2& g, uint32_t i, Feature2& out) {
    out.clear();
    uint32_t i1 = i < 2 ? i : i - 2;
    real sign = i < 2 ? real(1) : real(-1);
    uint32_t i2 = (i1 + 1) % 2;
    real vertex[2] = {g.he.x, g.he.y};
    vertex[i1] *= sign;
    vertex[i2] *= (i1 == 0) ? -sign : sign;
    P2 p1 = p2(vertex[0], vertex[1]);
    vertex[i2] = -vertex[i2];
    P2 p2_ = p2(vertex[0], vertex[1]);
    uint32_t vid1 = sign < 0 ? (1u << i1) : 0u, vid2 = vid1;
    real p1_i2 = i2 == 0 ? p1.x : p1.y;
    if (p1_i2 < 0)
        vid1 |= 1u << i2;
    else
        vid2 |= 1u << i2;
    out.push(p1, F_VERTEX | vid1);
    out.push(p2_, F_VERTEX | vid2);
    real nrm[2] = {0, 0};
    nrm[i1] = sign;
    out.normal = p2(nrm[0], nrm[1]), out.has_normal = true;
    out.fid = F_FACE | i;
}
static void polygon_face(const Shape2& g, uint32_t ia, Feature2& out) {  // convex_polygon.rs:123-135
    out.clear();
    uint32_t ib = (ia + 1) % g.npts;
    out.push(p2(g.pts[2 * ia], g.pts[2 * ia + 1]), F_VERTEX | ia);
    out.push(p2(g.pts[2 * ib], g.pts[2 * ib + 1]), F_VERTEX | ib);
    out.normal = p2(g.normals[2 * ia], g.normals[2 * ia + 1]), out.has_normal = true;
    out.fid = F_FACE | ia;
}
// Segment::face (segment.rs:212-235, dim2); a degenerate segment is the single vertex a
static void segment_face(P2 a, P2 b, uint32_t id, Feature2& out) {
    out.clear();
    P2 ab = b - a, nrm;
    if (unit_try_new(p2(ab.y, -ab.x), EPS, &nrm)) {  // utils::ccw_face_normal
        out.fid = F_FACE | id;
        if (id == 0) {
            out.push(a, F_VERTEX | 0u), out.push(b, F_VERTEX | 1u);
            out.normal = nrm, out.has_normal = true;
        } else {
            out.push(b, F_VERTEX | 1u), out.push(a, F_VERTEX | 0u);
            out.normal = -nrm, out.has_normal = true;
        }
    } else {
        out.push(a, F_VERTEX | 0u);
        out.fid = F_VERTEX | 0u;
    }
}
static void support_face_toward(const Shape2& g, const Iso2& m, P2 dir, Feature2& out) {
    if (g.type == SEGMENT2) {  // segment.rs:286-299 (dim2): `dir` is NOT brought into the segment's frame (as in the reference)
        segment_face(g.he, g.sb, perp(dir, g.sb - g.he) >= 0 ? 0u : 1u, out);
        out.transform_by(m);
        return;
    }
    P2 ld = inv_rot(m, dir);
    if (g.type == CUBOID2) {  // cuboid.rs:279-308
        real l[2] = {ld.x, ld.y};
        int iamax = 0;
        real amax = std::fabs(l[0]);
        if (std::fabs(l[1]) > amax) iamax = 1;
        cuboid_face(g, l[iamax] > 0 ? iamax : iamax + 2, out);
    } else {  // convex_polygon.rs:154-172
        uint32_t best = 0;
        real max_dot = g.normals[0] * ld.x + g.normals[1] * ld.y;
        for (uint32_t i = 1; i < g.npts; ++i) {
            real d = g.normals[2 * i] * ld.x + g.normals[2 * i + 1] * ld.y;
            if (d > max_dot) max_dot = d, best = i;
        }
        polygon_face(g, best, out);
    }
    out.transform_by(m);
}
static real signum(real x) { return std::isnan(x) ? x : (std::signbit(x) ? real(-1) : real(1)); }  // f32::signum: -0.0 -> -1.0
static void support_feature_toward(const Shape2& g, const Iso2& m, P2 dir, real cang, real sang, Feature2& out) {
    if (g.type == SEGMENT2) {  // segment.rs:315-345 (dim2): eps.sin() is the angular prediction's sine
        out.clear();
        P2 a = mul_point(m, g.he), b = mul_point(m, g.sb), sd;  // self.transformed(transform)
        if (unit_try_new(b - a, EPS, &sd)) {
            real c = dot(dir, sd);
            if (c > sang)
                out.fid = F_VERTEX | 1u, out.push(b, F_VERTEX | 1u);
            else if (c < -sang)
                out.fid = F_VERTEX | 0u, out.push(a, F_VERTEX | 0u);
            else
                segment_face(a, b, perp(dir, sd) >= 0 ? 0u : 1u, out);
        }
        return;
    }
    if (g.type != CUBOID2) {  // convex_polygon.rs:174-184: the support face
        support_face_toward(g, m, dir, out);
        return;
    }
    P2 ld = inv_rot(m, dir);  // cuboid.rs:310-350 (dim2)
    real l[2] = {ld.x, ld.y}, sp[2] = {g.he.x, g.he.y};
    out.clear();
    uint32_t spid = 0;
    for (int i1 = 0; i1 < 2; ++i1) {
        real sign = signum(l[i1]);
        if (sign * l[i1] >= cang) {
            cuboid_face(g, sign > 0 ? i1 : i1 + 2, out);
            out.transform_by(m);
            return;
        }
        if (sign < 0) spid |= 1u << i1;
        sp[i1] *= sign;
    }
    out.push(mul_point(m, p2(sp[0], sp[1])), F_VERTEX | spid);
    out.fid = F_VERTEX | spid;
}

struct Cand2 {
    Contact2 c;
    uint32_t f1, f2;
};
// ConvexPolygonalFeature::clip (convex_polygonal_feature2.rs:95-180)
static void clip2(const Feature2& self, const Feature2& other, P2 normal, real prediction, std::vector<Cand2>& out) {
    if (self.nv <= 1 || other.nv <= 1) return;
    P2 ortho = p2(-normal.y, normal.x);
    P2 s1a = self.v[0], s1b = self.v[1], s2a = other.v[0], s2b = other.v[1];
    P2 ref = s1a;
    real r1[2] = {dot(s1a - ref, ortho), dot(s1b - ref, ortho)}, r2[2] = {dot(s2a - ref, ortho), dot(s2b - ref, ortho)};
    uint32_t f1[2] = {self.vid[0], self.vid[1]}, f2[2] = {other.vid[0], other.vid[1]};
    if (r1[1] < r1[0]) std::swap(r1[0], r1[1]), std::swap(f1[0], f1[1]), std::swap(s1a, s1b);
    if (r2[1] < r2[0]) std::swap(r2[0], r2[1]), std::swap(f2[0], f2[1]), std::swap(s2a, s2b);
    if (r2[0] > r1[1] || r1[0] > r2[1]) return;
    real len1 = r1[1] - r1[0], len2 = r2[1] - r2[0];
    auto point_at = [](P2 a, P2 b, real bc) { return p2(a.x * (real(1) - bc) + b.x * bc, a.y * (real(1) - bc) + b.y * bc); };  // a * c0 + b.coords * c1
    auto emit = [&](P2 w1, P2 w2, uint32_t fa, uint32_t fb) {
        Contact2 c;
        c.w1 = w1, c.w2 = w2, c.n = normal, c.depth = -dot(normal, w2 - w1);
        if (-c.depth <= prediction) out.push_back({c, fa, fb});
    };
    if (r2[0] > r1[0])
        emit(point_at(s1a, s1b, (r2[0] - r1[0]) / len1), s2a, self.fid, f2[0]);
    else
        emit(s1a, point_at(s2a, s2b, (r1[0] - r2[0]) / len2), f1[0], other.fid);
    if (r2[1] < r1[1])
        emit(point_at(s1a, s1b, (r2[1] - r1[0]) / len1), s2b, self.fid, f2[1]);
    else
        emit(s1b, point_at(s2a, s2b, (r1[1] - r2[0]) / len2), f1[1], other.fid);
}

struct Manifold2 {  // a fresh ContactManifold, DistanceBased(0.02)
    std::vector<Cand2> c;
    std::vector<P2> track;
    void push(const Contact2& ct, uint32_t f1, uint32_t f2, P2 tracking_pt) {
        const real threshold = real(0.02);
        size_t closest = c.size();
        real closest_dist = threshold * threshold;
        for (size_t i = 0; i < c.size(); ++i) {
            real d = nsq(tracking_pt - track[i]);
            if (d < closest_dist) closest_dist = d, closest = i;
        }
        if (closest == c.size()) {
            c.push_back({ct, f1, f2});
            track.push_back(tracking_pt);
        } else if (ct.depth > c[closest].c.depth) {  // a contact of this update is matched: the deeper one stays (:203-216)
            c[closest] = {ct, f1, f2};
            track[closest] = tracking_pt;
        }
    }
};

struct Box2 {
    P2 lo, hi;
};
static Box2 shape_aabb2(const Shape2& g, const Iso2& m) {
    P2 lo, hi;
    if (g.type == PLANE2) {  // aabb_plane.rs:13-21
        real mx = FMAX * real(0.5);
        lo = p2(-mx, -mx), hi = p2(mx, mx);
    } else if (g.type == BALL2) {
        lo = p2(m.t.x + (-g.radius), m.t.y + (-g.radius)), hi = p2(m.t.x + g.radius, m.t.y + g.radius);
    } else if (g.type == CUBOID2) {
        real are = std::fabs(m.re), aim = std::fabs(m.im);
        P2 w = p2(are * g.he.x + aim * g.he.y, aim * g.he.x + are * g.he.y);
        lo = m.t - w, hi = m.t + w;
    } else if (g.type == SEGMENT2) {  // aabb_segment.rs -> support_map_aabb (aabb_utils.rs:9-31): one support point per axis direction
        hi = p2(support_point(g, m, p2(1, 0)).x, support_point(g, m, p2(0, 1)).y);
        lo = p2(support_point(g, m, p2(-1, 0)).x, support_point(g, m, p2(0, -1)).y);
    } else {
        P2 wp = mul_point(m, p2(g.pts[0], g.pts[1]));
        lo = hi = wp;
        for (uint32_t i = 1; i < g.npts; ++i) {
            wp = mul_point(m, p2(g.pts[2 * i], g.pts[2 * i + 1]));
            lo = p2(std::fmin(lo.x, wp.x), std::fmin(lo.y, wp.y)), hi = p2(std::fmax(hi.x, wp.x), std::fmax(hi.y, wp.y));
        }
    }
    return Box2{lo, hi};
}

// one pair through its generator; returns the manifold in push order
static void generate_contacts2(const Shape2& g1, const Iso2& m1, const Shape2& g2, const Iso2& m2, real linear, real cang1, real cang2, real sang1,
                               real sang2, Manifold2& mf, int* panicked) {
    const uint32_t FACE0 = F_FACE | 0u;
    if (g1.type == PLANE2 && g2.type == PLANE2) return;  // no contact algorithm: no interaction edge
    if (g1.type == BALL2 && g2.type == BALL2) {
        Contact2 c;
        if (contact_ball_ball(m1.t, g1.radius, m2.t, g2.radius, linear, &c)) mf.push(c, FACE0, FACE0, p2(0, 0));
        return;
    }
    if (g1.type == PLANE2 || g2.type == PLANE2) {
        bool flip = g1.type != PLANE2;
        const Shape2& pl = flip ? g2 : g1;
        const Shape2& ot = flip ? g1 : g2;
        const Iso2& mp = flip ? m2 : m1;
        const Iso2& mo = flip ? m1 : m2;
        P2 n = rot(mp, pl.he), center = mp.t;
        if (ot.type == BALL2) {  // PlaneBallManifoldGenerator (plane_ball_manifold_generator.rs:40-77)
            real dist = dot(mo.t - center, n), depth = -dist + ot.radius;
            if (depth > -linear) {
                P2 world1 = mo.t + n * (-dist), world2 = mo.t + n * (-ot.radius);
                Contact2 c;
                if (!flip) {
                    c.w1 = world1, c.w2 = world2, c.n = n, c.depth = depth;
                    mf.push(c, FACE0, FACE0, p2(0, 0));
                } else {
                    c.w1 = world2, c.w2 = world1, c.n = -n, c.depth = depth;
                    mf.push(c, FACE0, FACE0, p2(0, 0));
                }
            }
            return;
        }
        Feature2 f;  // PlaneConvexPolyhedronManifoldGenerator (plane_convex_polyhedron_manifold_generator.rs:40-85): both array slots
        support_face_toward(ot, mo, -n, f);
        for (int i = 0; i < 2; ++i) {
            P2 world2 = f.v[i];
            real dist = dot(world2 - center, n);
            if (dist <= linear) {
                P2 world1 = world2 + (-n) * dist;
                P2 local2 = inv_point(mo, world2);
                Contact2 c;
                if (!flip) {
                    c.w1 = world1, c.w2 = world2, c.n = n, c.depth = -dist;
                    mf.push(c, FACE0, f.vid[i], local2);
                } else {
                    c.w1 = world2, c.w2 = world1, c.n = -n, c.depth = -dist;
                    mf.push(c, f.vid[i], FACE0, local2);
                }
            }
        }
        return;
    }
    if (g1.type == BALL2 || g2.type == BALL2) {  // BallConvexPolyhedronManifoldGenerator::new(flip = ball is second)
        bool flip = g1.type != BALL2;
        const Shape2& ball = flip ? g2 : g1;
        const Shape2& cp = flip ? g1 : g2;
        const Iso2& mb = flip ? m2 : m1;
        const Iso2& mc = flip ? m1 : m2;
        bool inside;
        uint32_t f2 = F_UNKNOWN;
        P2 world2;
        if (cp.type == CUBOID2) {
            world2 = cuboid_project(cp, mc, mb.t, &inside, &f2);
        } else if (cp.type == SEGMENT2) {
            world2 = segment_project(cp, mc, mb.t, &inside, &f2);
        } else {
            world2 = polygon_project(cp, mc, mb.t, &inside, panicked);
            P2 back = mb.t - world2, ld;
            if (unit_try_new(inv_rot(mc, inside ? -back : back), EPS, &ld)) f2 = polygon_feature_toward(cp, ld);
        }
        P2 dpt = world2 - mb.t, dir, normal;
        real dist, depth;
        if (unit_try_new_and_get(dpt, EPS, &dir, &dist)) {
            depth = inside ? dist + ball.radius : -dist + ball.radius;
            normal = inside ? -dir : dir;
        } else {
            if (f2 == F_UNKNOWN) return;
            depth = ball.radius;
            normal = -(cp.type == CUBOID2 ? cuboid_feature_normal(f2) : cp.type == SEGMENT2 ? segment_feature_normal(cp, f2) : polygon_feature_normal(cp, f2));
        }
        if (depth >= -linear) {
            if (f2 == F_UNKNOWN) {  // "Feature id cannot be unknown."
                *panicked += 1;
                return;
            }
            P2 world1 = mb.t + normal * ball.radius;
            Contact2 c;
            if (!flip) {
                c.w1 = world1, c.w2 = world2, c.n = normal, c.depth = depth;
                mf.push(c, FACE0, f2, p2(0, 0));
            } else {
                c.w1 = world2, c.w2 = world1, c.n = -normal, c.depth = depth;
                mf.push(c, f2, FACE0, p2(0, 0));
            }
        }
        return;
    }
    // ConvexPolyhedronConvexPolyhedronManifoldGenerator (fresh: last_gjk_dir = None)
    P2 dir0;
    if (!unit_try_new(m2.t - m1.t, EPS, &dir0)) dir0 = p2(1, 0);
    Simplex2 s;
    s.reset(cso_from_shapes(m1, g1, m2, g2, dir0));
    P2 w1, w2, n;
    int r = gjk_closest_points(m1, g1, m2, g2, linear, s, &w1, &w2, &n);
    if (r == R_NONE) return;
    if (r == R_INTERSECTION && !epa_closest_points(m1, g1, m2, g2, s, &w1, &w2, &n, panicked)) return;
    Contact2 ct;
    ct.w1 = w1, ct.w2 = w2, ct.n = n, ct.depth = -dot(n, w2 - w1);
    Feature2 fa, fb;
    if (ct.depth > 0) {
        support_face_toward(g1, m1, n, fa);
        support_face_toward(g2, m2, -n, fb);
    } else {
        support_feature_toward(g1, m1, n, cang1, sang1, fa);
        support_feature_toward(g2, m2, -n, cang2, sang2, fb);
    }
    std::vector<Cand2> fresh;
    clip2(fa, fb, n, linear, fresh);
    if (fresh.empty()) fresh.push_back({ct, fa.fid, fb.fid});
    for (auto& k : fresh) {  // add_contact_to_manifold (:182-236): Unknown features are dropped; tracking point = local1
        if (k.f1 == F_UNKNOWN || k.f2 == F_UNKNOWN) continue;
        mf.push(k.c, k.f1, k.f2, inv_point(m1, k.c.w1));
    }
}

// ---- RayCast for the 2-D shapes (solid = true, what the world queries and the reference's tests use) ----------------------------------
//   query/ray/ray_ball.rs:15-142 (Ball), ray_cuboid.rs + ray_aabb.rs:52-75,183-300 (Cuboid = its local AABB; note the `+ 3` of the
//   far-side face id, hard-coded for both dimensions), ray_plane.rs:9-79, ray_support_map.rs:15-60,165-189 (ConvexPolygon through
//   gjk::cast_ray), query/algorithms/gjk.rs:180-365 (minkowski_ray_cast, DIM = 2), query/ray/ray.rs:36-41.
struct RayHit2 {
    bool hit = false;
    real toi = 0;
    P2 n = {0, 0};
    uint32_t feature = 0xffffffffu;  // kind << 30 | id (1 Face), 0xffffffff Unknown
};
static const uint32_t FACE2 = 0x40000000u;

static RayHit2 ray2_ball(P2 center, real radius, P2 o, P2 d, real max_toi) {
    RayHit2 h;
    P2 dcenter = o - center;
    real a = nsq(d), b = dot(dcenter, d), c = nsq(dcenter) - radius * radius;
    bool inside = false;
    real t = 0;
    if (a == real(0)) {
        if (c > real(0)) return h;
        inside = true;
    } else if (c > real(0) && b > real(0)) {
        return h;
    } else {
        real delta = b * b - a * c;
        if (delta < real(0)) return h;
        t = (-b - std::sqrt(delta)) / a;
        if (t <= real(0)) inside = true, t = 0;  // solid
    }
    if (!(t <= max_toi)) return h;
    P2 pos = (o + d * t) - center;
    P2 normal = normalize(pos);
    h.hit = true, h.toi = t, h.n = inside ? -normal : normal, h.feature = FACE2;
    return h;
}

static RayHit2 ray2_cuboid(P2 he, const Iso2& m, P2 o_w, P2 d_w, real max_toi) {
    RayHit2 h;
    P2 o = inv_point(m, o_w), d = inv_rot(m, d_w);
    const real oo[2] = {o.x, o.y}, dd[2] = {d.x, d.y}, mn[2] = {-he.x, -he.y}, mx[2] = {he.x, he.y};
    real tmax = FMAX, tmin = -FMAX;  // clip_line (ray_aabb.rs:183-279)
    int near_side = 0, far_side = 0;
    bool near_diag = false;
    for (int i = 0; i < 2; ++i) {
        if (dd[i] == real(0)) {
            if (oo[i] < mn[i] || oo[i] > mx[i]) return h;
        } else {
            real denom = real(1) / dd[i];
            real near = (mn[i] - oo[i]) * denom, far = (mx[i] - oo[i]) * denom;
            bool flip = false;
            if (near > far) flip = true, std::swap(near, far);
            if (near > tmin)
                tmin = near, near_side = flip ? -(i + 1) : (i + 1), near_diag = false;
            else if (near == tmin)
                near_diag = true;
            if (far < tmax) tmax = far, far_side = !flip ? -(i + 1) : (i + 1);
            if (tmax < real(0) || tmin > tmax) return h;
        }
    }
    P2 near_n = p2(0, 0);
    if (near_diag)
        near_n = -normalize(d);
    else if (near_side != 0) {
        real* c = &near_n.x;
        if (near_side < 0)
            c[-near_side - 1] = real(1);
        else
            c[near_side - 1] = -real(1);
    }
    real t;
    P2 n;
    int side;
    if (tmin < real(0))
        t = 0, n = p2(0, 0), side = far_side;  // ray_aabb (:282-300), solid
    else if (tmin <= max_toi)
        t = tmin, n = near_n, side = near_side;
    else
        return h;
    h.hit = true, h.toi = t, h.n = rot(m, n);
    h.feature = FACE2 | ((uint32_t)(side < 0 ? (-side - 1 + 3) : (side - 1)) & 0x3fffffffu);
    return h;
}

static RayHit2 ray2_plane(P2 pn, const Iso2& m, P2 o_w, P2 d_w, real max_toi) {
    RayHit2 h;
    P2 o = inv_point(m, o_w), d = inv_rot(m, d_w);
    P2 dpos = -o;
    real dot_normal_dpos = dot(pn, dpos);
    if (dot_normal_dpos > real(0)) {  // solid: the origin is inside the half-space
        h.hit = true, h.toi = 0, h.n = p2(0, 0), h.feature = FACE2;
        return h;
    }
    real t = dot_normal_dpos / dot(pn, d);
    if (t >= real(0) && t <= max_toi) h.hit = true, h.toi = t, h.n = rot(m, pn), h.feature = FACE2;
    return h;
}

static bool ray2_toi_with_plane(P2 center, P2 normal, P2 origin, P2 dir, real* t_out) {  // ray_plane.rs:9-42
    P2 dpos = center - origin;
    real denom = dot(normal, dir);
    if (relative_eq(denom, real(0))) return false;
    real t = dot(normal, dpos) / denom;
    if (t >= real(0)) {
        *t_out = t;
        return true;
    }
    return false;
}

// gjk::cast_ray = minkowski_ray_cast(m1, g1, identity, ConstantOrigin, ...) (gjk.rs:180-365), DIM = 2
static bool minkowski_ray_cast2(const Iso2& m1, const Shape2& g1, const Iso2& m2, const Shape2& g2, P2 ray_origin, P2 ray_dir, real max_toi,
                                Simplex2& simplex, real* toi_out, P2* normal_out) {
    const real eps_rel = std::sqrt(EPS_TOL);
    real ray_length = std::sqrt(nsq(ray_dir));
    if (relative_eq(ray_length, real(0))) return false;
    real ltoi = 0;
    P2 curr_origin = ray_origin, curr_dir = ray_dir / ray_length;
    P2 dir0 = -curr_dir, ldir = dir0;
    CSO sp0 = cso_from_shapes(m1, g1, m2, g2, dir0);
    sp0.point = sp0.point + (-curr_origin);  // translate(&-origin.coords): only `point` moves
    simplex.reset(sp0);
    P2 proj = simplex.project_origin_and_reduce();
    real max_bound = FMAX;
    P2 dir;
    int niter = 0;
    bool last_chance = false;
    for (;;) {
        real old_max_bound = max_bound, dist;
        if (unit_try_new_and_get(-proj, EPS_TOL, &dir, &dist))
            max_bound = dist;
        else {
            *toi_out = ltoi / ray_length, *normal_out = ldir;
            return true;
        }
        CSO support_point;
        if (max_bound >= old_max_bound) {
            last_chance = true;
            P2 p = proj + curr_origin;
            support_point = CSO{p, p, p2(0, 0)};  // CSOPoint::single_point
        } else {
            support_point = cso_from_shapes(m1, g1, m2, g2, dir);
        }
        if (last_chance && ltoi > real(0)) {
            *toi_out = ltoi / ray_length, *normal_out = ldir;
            return true;
        }
        real t;
        if (ray2_toi_with_plane(support_point.point, dir, curr_origin, curr_dir, &t)) {
            if (dot(dir, curr_dir) < real(0) && t > real(0)) {
                ldir = dir;
                ltoi += t;
                if (ltoi / ray_length > max_toi) return false;
                P2 shift = curr_dir * t;
                curr_origin = curr_origin + shift;
                max_bound = FMAX;
                for (int i = 0; i < simplex.dim + 1; ++i) simplex.vertices[i].point = simplex.vertices[i].point + (-shift);
                last_chance = false;
            }
        } else if (dot(dir, curr_dir) > EPS_TOL) {
            return false;
        }
        if (last_chance) return false;
        real min_bound = -dot(dir, support_point.point - curr_origin);
        if (max_bound - min_bound <= eps_rel * max_bound) return false;  // improved_fixed_point_support is off
        CSO tp = support_point;
        tp.point = tp.point + (-curr_origin);
        (void)simplex.add_point(tp);
        proj = simplex.project_origin_and_reduce();
        if (simplex.dim == 2) {
            if (min_bound >= EPS_TOL) return false;
            *toi_out = ltoi / ray_length, *normal_out = ldir;
            return true;
        }
        if (++niter == 10000) return false;
    }
}

// RayCast for ConvexPolygon (ray_support_map.rs:165-189 -> :15-60), solid = true
static RayHit2 ray2_polygon(const Shape2& g, const Iso2& m, P2 o_w, P2 d_w, real max_toi) {
    RayHit2 h;
    P2 o = inv_point(m, o_w), d = inv_rot(m, d_w);
    Shape2 origin;
    origin.type = ORIGIN2, origin.radius = 0, origin.he = p2(0, 0), origin.pts = origin.normals = nullptr, origin.npts = 0;
    Iso2 id = {p2(0, 0), real(1), real(0)};
    Simplex2 simplex;
    P2 supp = support_point(g, id, -d);
    P2 p = supp - o;
    simplex.reset(CSO{p, p, p2(0, 0)});  // replaced by minkowski_ray_cast's own reset, like in the reference
    real toi;
    P2 normal;
    if (!minkowski_ray_cast2(id, g, id, origin, o, d, max_toi, simplex, &toi, &normal)) return h;
    h.hit = true, h.toi = toi, h.n = rot(m, normal), h.feature = 0xffffffffu;
    return h;
}

// RayCast for Segment, dim2 (query/ray/ray_support_map.rs:219-293): the segment moved by m, line / line parameters
// (closest_points_line_line.rs:27-70), the collinear cases; the normal is the SCALED normal and max_toi is never applied
static RayHit2 ray2_segment(const Shape2& g, const Iso2& m, P2 o, P2 d) {
    RayHit2 h;
    P2 a = mul_point(m, g.he), b = mul_point(m, g.sb);
    P2 sd = b - a, r = o - a;
    real aa = nsq(d), e = nsq(sd), f = dot(sd, r), s, t;
    bool parallel = false;
    if (aa <= EPS && e <= EPS) {
        s = 0, t = 0;
    } else if (aa <= EPS) {
        s = 0, t = f / e;
    } else {
        real c = dot(d, r);
        if (e <= EPS) {
            s = -c / aa, t = 0;
        } else {
            real bq = dot(d, sd), ae = aa * e, bb = bq * bq, denom = ae - bb;
            parallel = denom <= EPS || ulps_eq(ae, bb);
            s = !parallel ? (bq * f - c * e) / denom : real(0);
            t = (bq * s + f) / e;
        }
    }
    P2 nrm = p2(sd.y, -sd.x);
    if (parallel) {
        P2 dpos = a - o;
        if (std::fabs(dot(dpos, nrm)) < EPS) {
            real dist1 = dot(dpos, d), dist2 = dist1 + dot(sd, d);
            if (dist1 >= 0 && dist2 >= 0) {
                h.hit = true, h.n = nrm;
                if (dist1 <= dist2)
                    h.toi = dist1 / nsq(d), h.feature = 0x80000000u | 0u;
                else
                    h.toi = dist2 / nsq(d), h.feature = 0x80000000u | 1u;
            } else if (dist1 >= 0 || dist2 >= 0) {
                h.hit = true, h.toi = 0, h.n = nrm, h.feature = FACE2 | 0u;
            }
        }
    } else if (s >= 0 && t >= 0 && t <= 1) {
        h.hit = true, h.toi = s;
        if (dot(nrm, d) > 0)
            h.n = -nrm, h.feature = FACE2 | 1u;
        else
            h.n = nrm, h.feature = FACE2 | 0u;
    }
    return h;
}

static RayHit2 shape_ray_cast2(const Shape2& g, const Iso2& m, P2 o, P2 d, real max_toi) {
    if (g.type == SEGMENT2) return ray2_segment(g, m, o, d);
    switch (g.type) {
        case BALL2: return ray2_ball(m.t, g.radius, o, d, max_toi);
        case CUBOID2: return ray2_cuboid(g.he, m, o, d, max_toi);
        case POLYGON2: return ray2_polygon(g, m, o, d, max_toi);
        default: return ray2_plane(g.he, m, o, d, max_toi);
    }
}

}  // namespace d2
}  // namespace orc

using namespace orc;
using namespace orc::d2;

extern "C" {

// query::contact for n pairs.  type: 0 ball, 1 cuboid, 2 convex polygon; param (4 reals per shape): ball (radius), cuboid (hx, hy),
// polygon (first point, point count) into poly_points (x, y); pose (4 reals): translation x, y, rotation re, im (UnitComplex).
// poly_normals: ConvexPolygon::normals aligned with poly_points (needed for ball x polygon).  found[p]: 1 Some, 0 None, 2 = ball x polygon without normals.  out (7 reals): world1, world2, normal, depth.
// panics (optional): number of pairs on which the reference would panic (heap.peek().unwrap() on an empty heap, epa2.rs:279).
void orc2_contact(uint64_t n, const uint32_t* type1, const real* param1, const real* pose1, const uint32_t* type2, const real* param2,
                  const real* pose2, const real* poly_points, const real* poly_normals, real prediction, uint8_t* found, real* out,
                  uint32_t* panics) {
    auto shape = [&](uint32_t t, const real* p) {
        Shape2 g;
        g.type = t, g.radius = p[0], g.he = p2(p[0], p[1]), g.sb = p2(p[2], p[3]), g.pts = g.normals = nullptr, g.npts = 0;
        if (t == POLYGON2) {
            g.pts = poly_points + 2 * (size_t)p[0], g.npts = (uint32_t)p[1];
            g.normals = poly_normals ? poly_normals + 2 * (size_t)p[0] : nullptr;
        }
        return g;
    };
    uint32_t np = 0;
    for (uint64_t k = 0; k < n; ++k) {
        Shape2 g1 = shape(type1[k], param1 + 4 * k), g2 = shape(type2[k], param2 + 4 * k);
        Iso2 m1 = {p2(pose1[4 * k], pose1[4 * k + 1]), pose1[4 * k + 2], pose1[4 * k + 3]};
        Iso2 m2 = {p2(pose2[4 * k], pose2[4 * k + 1]), pose2[4 * k + 2], pose2[4 * k + 3]};
        Contact2 c = {p2(0, 0), p2(0, 0), p2(0, 0), 0};
        bool ok = false;
        int panicked = 0;
        uint8_t code = 0;
        if (g1.type == PLANE2 && g2.type == PLANE2) {
            code = 2;  // the reference panics: "No algorithm known to compute a contact point between the given pair of shapes."
        } else if (g1.type == BALL2 && g2.type == BALL2) {
            ok = contact_ball_ball(m1.t, g1.radius, m2.t, g2.radius, prediction, &c);
        } else if (g1.type == PLANE2) {
            ok = contact_plane_sm(m1, g1.he, m2, g2, prediction, &c);
        } else if (g2.type == PLANE2) {  // contact_support_map_plane: flipped
            ok = contact_plane_sm(m2, g2.he, m1, g1, prediction, &c);
            if (ok) {
                std::swap(c.w1, c.w2);
                c.n = -c.n;
            }
        } else if (g1.type == BALL2 && g2.type == CUBOID2) {
            ok = contact_ball_cuboid(m1.t, g1.radius, m2, g2, prediction, &c);
        } else if (g1.type == CUBOID2 && g2.type == BALL2) {  // contact_convex_polyhedron_ball: flip
            ok = contact_ball_cuboid(m2.t, g2.radius, m1, g1, prediction, &c);
            if (ok) {
                std::swap(c.w1, c.w2);
                c.n = -c.n;
            }
        } else if (g1.type == BALL2 && g2.type == SEGMENT2) {
            ok = contact_ball_segment(m1.t, g1.radius, m2, g2, prediction, &c);
        } else if (g1.type == SEGMENT2 && g2.type == BALL2) {
            ok = contact_ball_segment(m2.t, g2.radius, m1, g1, prediction, &c);
            if (ok) {
                std::swap(c.w1, c.w2);
                c.n = -c.n;
            }
        } else if (g1.type == BALL2 && g2.type == POLYGON2 && g2.normals) {
            ok = contact_ball_polygon(m1.t, g1.radius, m2, g2, prediction, &c, &panicked);
        } else if (g1.type == POLYGON2 && g2.type == BALL2 && g1.normals) {
            ok = contact_ball_polygon(m2.t, g2.radius, m1, g1, prediction, &c, &panicked);
            if (ok) {
                std::swap(c.w1, c.w2);
                c.n = -c.n;
            }
        } else if (g1.type == BALL2 || g2.type == BALL2) {
            code = 2;  // ball x polygon without the polygon's normals
        } else {
            ok = contact_sm_sm(m1, g1, m2, g2, prediction, &c, &panicked);
        }
        np += panicked;
        found[k] = code ? code : (ok ? 1 : 0);
        real* o = out + 7 * k;
        o[0] = c.w1.x, o[1] = c.w1.y, o[2] = c.w2.x, o[3] = c.w2.y, o[4] = c.n.x, o[5] = c.n.y, o[6] = c.depth;
    }
    if (panics) *panics = np;
}

// query::proximity for n pairs (margin per pair); out: 0 Intersecting, 1 WithinMargin, 2 Disjoint, 255 = plane x plane (the reference panics)
void orc2_proximity(uint64_t n, const uint32_t* type1, const real* param1, const real* pose1, const uint32_t* type2, const real* param2,
                    const real* pose2, const real* poly_points, const real* margins, uint8_t* out) {
    for (uint64_t k = 0; k < n; ++k) {
        auto shape = [&](uint32_t t, const real* p) {
            Shape2 g;
            g.type = t, g.radius = p[0], g.he = p2(p[0], p[1]), g.sb = p2(p[2], p[3]), g.pts = g.normals = nullptr, g.npts = 0;
            if (t == POLYGON2) g.pts = poly_points + 2 * (size_t)p[0], g.npts = (uint32_t)p[1];
            return g;
        };
        Shape2 g1 = shape(type1[k], param1 + 4 * k), g2 = shape(type2[k], param2 + 4 * k);
        Iso2 m1 = {p2(pose1[4 * k], pose1[4 * k + 1]), pose1[4 * k + 2], pose1[4 * k + 3]};
        Iso2 m2 = {p2(pose2[4 * k], pose2[4 * k + 1]), pose2[4 * k + 2], pose2[4 * k + 3]};
        real margin = margins[k];
        int r;
        if (g1.type == PLANE2 && g2.type == PLANE2)
            r = 255;
        else if (g1.type == BALL2 && g2.type == BALL2)
            r = proximity_ball_ball(m1.t, g1.radius, m2.t, g2.radius, margin);
        else if (g1.type == PLANE2)
            r = proximity_plane_sm(m1, g1.he, m2, g2, margin);
        else if (g2.type == PLANE2)
            r = proximity_plane_sm(m2, g2.he, m1, g1, margin);
        else
            r = proximity_sm_sm(m1, g1, m2, g2, margin);
        out[k] = (uint8_t)r;
    }
}

// RayCast::toi_and_normal_with_ray(m, ray, max_toi, solid = true) of shape k for ray k; rays: origin x y, dir x y, max_toi;
// found 1 Some / 0 None; out: toi, normal x y; feature: kind << 30 | id (1 Face) or 0xffffffff (Unknown).
void orc2_ray_cast(uint64_t n, const uint32_t* type, const real* param, const real* pose, const real* poly_points, const real* rays,
                   uint8_t* found, real* out, uint32_t* feature) {
    for (uint64_t k = 0; k < n; ++k) {
        const real* p = param + 4 * k;
        Shape2 g;
        g.type = type[k], g.radius = p[0], g.he = p2(p[0], p[1]), g.sb = p2(p[2], p[3]), g.pts = g.normals = nullptr, g.npts = 0;
        if (g.type == POLYGON2) g.pts = poly_points + 2 * (size_t)p[0], g.npts = (uint32_t)p[1];
        Iso2 m = {p2(pose[4 * k], pose[4 * k + 1]), pose[4 * k + 2], pose[4 * k + 3]};
        const real* q = rays + 5 * k;
        RayHit2 h = shape_ray_cast2(g, m, p2(q[0], q[1]), p2(q[2], q[3]), q[4]);
        found[k] = h.hit ? 1 : 0;
        out[3 * k] = h.toi, out[3 * k + 1] = h.n.x, out[3 * k + 2] = h.n.y;
        feature[k] = h.hit ? h.feature : 0xffffffffu;
    }
}

// PointQuery::contains_point (point_ball.rs:45-47, point_cuboid.rs:34-38, point_plane.rs:44-48; ConvexPolygon: the trait's default,
// project_point(m, pt, false).is_inside, point_query.rs:50-52 + point_support_map.rs:14-55)
static bool contains_point2(const Shape2& g, const Iso2& m, P2 pt) {
    if (g.type == BALL2) return nsq(inv_point(m, pt)) <= g.radius * g.radius;
    if (g.type == CUBOID2) {
        P2 l = inv_point(m, pt);
        return !(l.x < -g.he.x || l.x > g.he.x || l.y < -g.he.y || l.y > g.he.y);
    }
    if (g.type == PLANE2) return dot(g.he, inv_point(m, pt)) <= real(0);
    if (g.type == SEGMENT2) {  // the trait's default: project_point(m, pt, false).is_inside, i.e. relative_eq!(proj, pt) (point_segment.rs:88)
        bool inside;
        uint32_t f;
        (void)segment_project(g, m, pt, &inside, &f);
        return inside;
    }
    Iso2 ms = m;
    ms.t = (-pt) + m.t;
    Iso2 id = {p2(0, 0), 1, 0};
    Shape2 origin;
    origin.type = ORIGIN2, origin.radius = 0, origin.he = p2(0, 0), origin.pts = origin.normals = nullptr, origin.npts = 0;
    P2 dir;
    if (!unit_try_new(-ms.t, EPS, &dir)) dir = p2(1, 0);
    Simplex2 s;
    s.reset(cso_from_shapes(ms, g, id, origin, dir));
    P2 p1, p2_, n;
    return gjk_closest_points(ms, g, id, origin, FMAX, s, &p1, &p2_, &n) != R_CLOSEST;
}

// AABB::toi_with_ray's hit test, DIM = 2 (ray_aabb.rs:13-50)
static bool box_hit_by_ray2(const real* mm, P2 o, P2 d, real max_toi) {
    real tmin = 0, tmax = max_toi;
    const real oo[2] = {o.x, o.y}, dd[2] = {d.x, d.y};
    for (int i = 0; i < 2; ++i) {
        if (dd[i] == real(0)) {
            if (oo[i] < mm[i] || oo[i] > mm[3 + i]) return false;
        } else {
            real denom = real(1) / dd[i];
            real near = (mm[i] - oo[i]) * denom, far = (mm[3 + i] - oo[i]) * denom;
            if (near > far) std::swap(near, far);
            tmin = std::fmax(tmin, near), tmax = std::fmin(tmax, far);
            if (tmin > tmax) return false;
        }
    }
    return true;
}

// ---- 2-D world -------------------------------------------------------------------------------------------------------------
struct orc2_objects {
    uint32_t n;
    const real* pos;          // 2 per object
    const real* rot;          // re, im per object
    const uint32_t* type;
    const real* param;        // 4 per object
    const real* query_limit;
    const real* ang_pred;
    const real* poly_points;
    const real* poly_normals;
    const uint8_t* query_kind;  // NULL, or per object 1 = GeometricQueryType::Proximity
};
static Shape2 obj_shape(const orc2_objects* o, uint32_t i) {
    Shape2 g;
    const real* p = o->param + 4 * (size_t)i;
    g.type = o->type[i], g.radius = p[0], g.he = p2(p[0], p[1]), g.sb = p2(p[2], p[3]), g.pts = g.normals = nullptr, g.npts = 0;
    if (g.type == POLYGON2) g.pts = o->poly_points + 2 * (size_t)p[0], g.normals = o->poly_normals + 2 * (size_t)p[0], g.npts = (uint32_t)p[1];
    return g;
}
static Iso2 obj_iso(const orc2_objects* o, uint32_t i) { return Iso2{p2(o->pos[2 * i], o->pos[2 * i + 1]), o->rot[2 * i], o->rot[2 * i + 1]}; }

// fat AABBs as 6 reals (mins xyz, maxs xyz, z = 0): ((shape AABB -/+ query_limit) -/+ margin), usable with orc_broad_phase
void orc2_compute_aabbs(const orc2_objects* o, real margin, real* out) {
    for (uint32_t i = 0; i < o->n; ++i) {
        Box2 a = shape_aabb2(obj_shape(o, i), obj_iso(o, i));
        real ql = o->query_limit[i];
        real mm[4] = {a.lo.x + (-ql), a.lo.y + (-ql), a.hi.x + ql, a.hi.y + ql};
        out[6 * i] = mm[0] + (-margin), out[6 * i + 1] = mm[1] + (-margin), out[6 * i + 2] = 0;
        out[6 * i + 3] = mm[2] + margin, out[6 * i + 4] = mm[3] + margin, out[6 * i + 5] = 0;
    }
}
// Contact manifolds of the given pairs (object1, object2): manifold_off[P + 1], contacts = 9 reals (world1, world2, normal, depth, f1, f2
// as reals holding the 32-bit feature codes: kind << 30 | id with kind 1 = face, 2 = vertex).  Returns the number of contacts.
// prox (optional): per pair the status of the proximity detector for pairs with a sensor (narrow_phase.rs:138-167; a fresh detector:
// BallBall / PlaneSupportMap / SupportMapSupportMap with sep_axis = None), 255 for contact pairs and for plane x plane (no detector).
uint64_t orc2_narrow_phase(const orc2_objects* o, uint64_t n_pairs, const uint32_t* pairs, uint32_t* manifold_off, real* contacts, uint32_t* feats,
                           uint64_t cap, uint32_t* panics, uint8_t* prox) {
    uint64_t nc = 0;
    int panicked = 0;
    for (uint64_t k = 0; k < n_pairs; ++k) {
        uint32_t i1 = pairs[2 * k], i2 = pairs[2 * k + 1];
        Manifold2 mf;
        real linear = o->query_limit[i1] + o->query_limit[i2];
        bool sensor = o->query_kind && (o->query_kind[i1] || o->query_kind[i2]);
        if (prox) prox[k] = 255;
        if (sensor) {
            Shape2 g1 = obj_shape(o, i1), g2 = obj_shape(o, i2);
            Iso2 m1 = obj_iso(o, i1), m2 = obj_iso(o, i2);
            int r = 255;
            if (g1.type == PLANE2 && g2.type == PLANE2)
                r = 255;
            else if (g1.type == BALL2 && g2.type == BALL2)
                r = proximity_ball_ball(m1.t, g1.radius, m2.t, g2.radius, linear);
            else if (g1.type == PLANE2)
                r = proximity_plane_sm(m1, g1.he, m2, g2, linear);
            else if (g2.type == PLANE2)
                r = proximity_plane_sm(m2, g2.he, m1, g1, linear);
            else
                r = proximity_sm_sm(m1, g1, m2, g2, linear);
            if (prox) prox[k] = (uint8_t)r;
        } else
        generate_contacts2(obj_shape(o, i1), obj_iso(o, i1), obj_shape(o, i2), obj_iso(o, i2), linear, std::cos(o->ang_pred[i1]),
                           std::cos(o->ang_pred[i2]), std::sin(o->ang_pred[i1]), std::sin(o->ang_pred[i2]), mf, &panicked);
        manifold_off[k] = (uint32_t)nc;
        for (auto& c : mf.c) {
            if (nc < cap) {
                real* q = contacts + 7 * nc;
                q[0] = c.c.w1.x, q[1] = c.c.w1.y, q[2] = c.c.w2.x, q[3] = c.c.w2.y, q[4] = c.c.n.x, q[5] = c.c.n.y, q[6] = c.c.depth;
                feats[2 * nc] = c.f1, feats[2 * nc + 1] = c.f2;
            }
            nc++;
        }
    }
    manifold_off[n_pairs] = (uint32_t)nc;
    if (panics) *panics = (uint32_t)panicked;
    return nc;
}


// glue::interferences_with_ray / first_interference_with_ray (pipeline/glue/query.rs:13-77,183-224) over a 2-D world by brute force:
// boxes = the broad phase's stored boxes (6 reals per object: mins xyz, maxs xyz, z unused), obj_groups = 3 words per object or NULL
// (defaults), groups = the query's or NULL.  Rows in (ray, handle) order; first_only: the smallest toi per ray, ties -> smallest handle.
uint64_t orc2_world_ray_cast(const orc2_objects* o, const real* boxes, const uint32_t* obj_groups, uint64_t n_rays, const real* rays,
                             const uint32_t* groups, int first_only, uint32_t* idx, real* val, uint32_t* feat, uint64_t cap) {
    uint64_t k = 0;
    for (uint64_t r = 0; r < n_rays; ++r) {
        const real* q = rays + 5 * r;
        P2 ro = p2(q[0], q[1]), rd = p2(q[2], q[3]);
        RayHit2 best;
        uint32_t best_h = 0;
        for (uint32_t h = 0; h < o->n; ++h) {
            if (!box_hit_by_ray2(boxes + 6 * (size_t)h, ro, rd, q[4])) continue;
            if (groups) {
                uint32_t m1 = 0x3fffffffu, w1 = 0x3fffffffu, b1 = 0;
                if (obj_groups) m1 = obj_groups[3 * h], w1 = obj_groups[3 * h + 1], b1 = obj_groups[3 * h + 2];
                if (!((m1 & groups[2]) == 0 && (groups[0] & b1) == 0 && (m1 & groups[1]) != 0 && (groups[0] & w1) != 0)) continue;
            }
            RayHit2 hit = shape_ray_cast2(obj_shape(o, h), obj_iso(o, h), ro, rd, q[4]);
            if (!hit.hit) continue;
            if (first_only) {
                if (!best.hit || hit.toi < best.toi) best = hit, best_h = h;
            } else {
                if (k < cap) idx[2 * k] = (uint32_t)r, idx[2 * k + 1] = h, val[3 * k] = hit.toi, val[3 * k + 1] = hit.n.x, val[3 * k + 2] = hit.n.y, feat[k] = hit.feature;
                k++;
            }
        }
        if (first_only && best.hit) {
            if (k < cap) idx[2 * k] = (uint32_t)r, idx[2 * k + 1] = best_h, val[3 * k] = best.toi, val[3 * k + 1] = best.n.x, val[3 * k + 2] = best.n.y, feat[k] = best.feature;
            k++;
        }
    }
    return k;
}

void orc2_contains_point(uint64_t n, const uint32_t* type, const real* param, const real* pose, const real* poly_points, const real* pts,
                         uint8_t* out) {
    for (uint64_t k = 0; k < n; ++k) {
        const real* p = param + 4 * k;
        Shape2 g;
        g.type = type[k], g.radius = p[0], g.he = p2(p[0], p[1]), g.sb = p2(p[2], p[3]), g.pts = g.normals = nullptr, g.npts = 0;
        if (g.type == POLYGON2) g.pts = poly_points + 2 * (size_t)p[0], g.npts = (uint32_t)p[1];
        Iso2 m = {p2(pose[4 * k], pose[4 * k + 1]), pose[4 * k + 2], pose[4 * k + 3]};
        out[k] = contains_point2(g, m, p2(pts[2 * k], pts[2 * k + 1])) ? 1 : 0;
    }
}

// glue::interferences_with_aabb (kind 0: 4 reals) / interferences_with_point (kind 2: 2 reals) by brute force; rows (query, handle) sorted
uint64_t orc2_world_query(const orc2_objects* o, const real* boxes, const uint32_t* obj_groups, int kind, uint64_t n_queries, const real* queries,
                          const uint32_t* groups, uint32_t* idx, uint64_t cap) {
    uint64_t k = 0;
    const int W = kind == 0 ? 4 : 2;
    for (uint64_t qi = 0; qi < n_queries; ++qi) {
        const real* q = queries + W * qi;
        for (uint32_t h = 0; h < o->n; ++h) {
            const real* b = boxes + 6 * (size_t)h;
            if (kind == 0) {
                if (!(b[0] <= q[2] && b[1] <= q[3] && b[3] >= q[0] && b[4] >= q[1])) continue;
            } else {
                if (q[0] < b[0] || q[0] > b[3] || q[1] < b[1] || q[1] > b[4]) continue;
            }
            if (groups) {
                uint32_t m1 = 0x3fffffffu, w1 = 0x3fffffffu, b1 = 0;
                if (obj_groups) m1 = obj_groups[3 * h], w1 = obj_groups[3 * h + 1], b1 = obj_groups[3 * h + 2];
                if (!((m1 & groups[2]) == 0 && (groups[0] & b1) == 0 && (m1 & groups[1]) != 0 && (groups[0] & w1) != 0)) continue;
            }
            if (kind == 2 && !contains_point2(obj_shape(o, h), obj_iso(o, h), p2(q[0], q[1]))) continue;
            if (k < cap) idx[2 * k] = (uint32_t)qi, idx[2 * k + 1] = h;
            k++;
        }
    }
    return k;
}

}  // extern "C"
