// ncollide2d on the device, first slice (the 2-D build): batched `query::contact` between 2-D balls, cuboids and
// convex polygons — one thread per pair, fixed capacities, no recursion, no heap allocation.
//
// Replaces (reference, file:line; the 2-D crate is the same tree built with feature "dim2"):
//   query::contact dispatch               query/contact/contact_shape_shape.rs:15-60
//   contact_ball_ball                     query/contact/contact_ball_ball.rs:8-38
//   contact_ball_convex_polyhedron (+ flipped)  query/contact/contact_ball_convex_polyhedron.rs:12-92 with Cuboid::project_point_with_feature
//                                         (query/point/point_cuboid.rs:17-26 -> point_aabb.rs:14-135) and the 2-D feature normals (shape/cuboid.rs:469-503)
//   contact_support_map_support_map       query/contact/contact_support_map_support_map.rs:9-79
//   gjk::closest_points (DIM = 2)         query/algorithms/gjk.rs:76-177,367-388
//   VoronoiSimplex (2-D)                  query/algorithms/voronoi_simplex2.rs:20-168, query/point/point_segment.rs:52-91,
//                                         query/point/point_triangle.rs:60-250 (dim2 branches)
//   EPA (2-D)                             query/algorithms/epa2.rs:15-378 (Vec + std BinaryHeap -> fixed arrays, same sift order)
//   support maps                          shape/ball.rs:29-48, shape/cuboid.rs:137-145, shape/convex_polygon.rs + utils/point_cloud_support_point.rs:6-24
// Same arithmetic contract as the 3-D path: --fmad=false, IEEE division / sqrt, nalgebra's evaluation order
// (UnitComplex * v = (re x - im y, im x + re y); Isometry2 * p = rotation * p + translation).
//   ball x convex polygon                 query/point/point_support_map.rs:14-55,120-146 (GJK / EPA projection of the ball centre),
//                                         shape/convex_polygon.rs:139-152,186-203 (feature normal, support feature)
// Not in this slice: the manifold generators / ConvexPolygonalFeature2, the 2-D broad phase and world.
#ifndef NCB_HOST_SHIM
#include <cub/cub.cuh>
#endif
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "ncb_internal.h"
#include "vec.cuh"  // NCB_EPS / NCB_FMAX only: the 2-D types are this file's own

namespace ncb {
namespace d2 {

struct W2 {
    float x, y;
};
__device__ __forceinline__ W2 w2(float x, float y) { return W2{x, y}; }
__device__ __forceinline__ W2 operator+(W2 a, W2 b) { return W2{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ W2 operator-(W2 a, W2 b) { return W2{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ W2 operator-(W2 a) { return W2{-a.x, -a.y}; }
__device__ __forceinline__ W2 operator*(W2 a, float s) { return W2{a.x * s, a.y * s}; }
__device__ __forceinline__ W2 operator/(W2 a, float s) { return W2{a.x / s, a.y / s}; }
__device__ __forceinline__ float dot(W2 a, W2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float perp(W2 a, W2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float nsq(W2 a) { return dot(a, a); }
// Unit::try_new_and_get: succeeds iff norm_squared > min_norm^2
__device__ __forceinline__ bool unit_get(W2 a, float min_norm, W2& u, float& n) {
    float sq = nsq(a);
    if (!(sq > min_norm * min_norm)) return false;
    n = sqrtf(sq);
    u = a / n;
    return true;
}
__device__ __forceinline__ bool unit(W2 a, float min_norm, W2& u) {
    float n;
    return unit_get(a, min_norm, u, n);
}
__device__ __forceinline__ W2 normalized(W2 a) { return a / sqrtf(nsq(a)); }

struct Pose2 {
    W2 t;
    float re, im;
};
__device__ __forceinline__ W2 rotate(const Pose2& m, W2 v) { return W2{m.re * v.x - m.im * v.y, m.im * v.x + m.re * v.y}; }
__device__ __forceinline__ W2 unrotate(const Pose2& m, W2 v) { return W2{m.re * v.x + m.im * v.y, -m.im * v.x + m.re * v.y}; }
__device__ __forceinline__ W2 to_world(const Pose2& m, W2 p) { return rotate(m, p) + m.t; }
__device__ __forceinline__ W2 to_local(const Pose2& m, W2 p) { return unrotate(m, p - m.t); }

#define D2_BALL 0u
#define D2_CUBOID 1u
#define D2_POLYGON 2u
#define D2_PLANE 3u   // shape/plane.rs: (a, b) = the unit normal; the code of the 3-D path's plane, so the broad phase treats its box as an outlier
#define D2_SEGMENT 4u  // shape/segment.rs: (a, b) = point a, (c, d) = point b
#define D2_ORIGIN 7u  // special_support_maps::ConstantOrigin
struct Operand2 {
    uint32_t kind;
    float a, b;          // radius | half extents | segment point a
    float c, d;          // segment point b
    const float* pts;    // polygon vertices (x, y)
    const float* nrm;    // polygon edge normals (ConvexPolygon::normals), may be null when no ball meets a polygon
    uint32_t npts;
    Pose2 m;
};

__device__ W2 support(const Operand2& g, W2 dir) {
    if (g.kind == D2_ORIGIN) return g.m.t;
    if (g.kind == D2_BALL) return g.m.t + normalized(dir) * g.a;  // support_point_toward(m, Unit::new_normalize(dir)): the rotation plays no part
    W2 ld = unrotate(g.m, dir), lp;
    if (g.kind == D2_CUBOID) {
        lp = w2(copysignf(g.a, ld.x), copysignf(g.b, ld.y));
    } else if (g.kind == D2_SEGMENT) {  // segment.rs:182-191
        lp = dot(w2(g.a, g.b), ld) > dot(w2(g.c, g.d), ld) ? w2(g.a, g.b) : w2(g.c, g.d);
    } else {
        uint32_t arg = 0;
        float best = __ldg(g.pts) * ld.x + __ldg(g.pts + 1) * ld.y;
        for (uint32_t i = 1; i < g.npts; ++i) {
            float d = __ldg(g.pts + 2 * i) * ld.x + __ldg(g.pts + 2 * i + 1) * ld.y;
            if (d > best) best = d, arg = i;  // first maximum
        }
        lp = w2(__ldg(g.pts + 2 * arg), __ldg(g.pts + 2 * arg + 1));
    }
    return to_world(g.m, lp);
}

// CSOPoint::from_shapes as three parallel arrays (point = orig1 - orig2)
struct MinkowskiPt {
    W2 p, o1, o2;
};
__device__ __forceinline__ MinkowskiPt minkowski(const Operand2& g1, const Operand2& g2, W2 dir) {
    MinkowskiPt c;
    c.o1 = support(g1, dir);
    c.o2 = support(g2, -dir);
    c.p = c.o1 - c.o2;
    return c;
}

#define TOL10 (NCB_EPS * 10.0f)    // gjk::eps_tol()
#define TOL100 (NCB_EPS * 100.0f)  // epa2's _eps_tol

// VoronoiSimplex of dimension <= 2
struct Tri2 {
    MinkowskiPt v[3];
    float bary[2], old_bary[2];
    int old_idx[3];
    int dim, old_dim;
};
__device__ __forceinline__ void tri_swap(Tri2& s, int a, int b) {
    MinkowskiPt t = s.v[a];
    s.v[a] = s.v[b];
    s.v[b] = t;
    int u = s.old_idx[a];
    s.old_idx[a] = s.old_idx[b];
    s.old_idx[b] = u;
}
__device__ bool tri_add(Tri2& s, const MinkowskiPt& c) {
    s.old_dim = s.dim;
    s.old_bary[0] = s.bary[0], s.old_bary[1] = s.bary[1];
    s.old_idx[0] = 0, s.old_idx[1] = 1, s.old_idx[2] = 2;
    for (int i = 0; i <= s.dim; ++i)
        if (nsq(s.v[i].p - c.p) < TOL10) return false;
    s.dim += 1;
    s.v[s.dim] = c;
    return true;
}
// project_origin_and_reduce: projection of the origin on the simplex, which shrinks to the sub-simplex that carries it
__device__ W2 tri_project(Tri2& s) {
    const W2 O = w2(0.f, 0.f);
    if (s.dim == 0) {
        s.bary[0] = 1.f;
        return s.v[0].p;
    }
    W2 a = s.v[0].p, b = s.v[1].p;
    W2 ab = b - a, ap = O - a;
    float ab_ap = dot(ab, ap);
    if (s.dim == 1) {
        float len2 = nsq(ab);
        if (ab_ap <= 0.f) {
            s.bary[0] = 1.f, s.dim = 0;
            return a;
        }
        if (ab_ap >= len2) {
            s.bary[0] = 1.f;
            tri_swap(s, 0, 1);
            s.dim = 0;
            return b;
        }
        float u = ab_ap / len2;
        s.bary[0] = 1.f - u, s.bary[1] = u;
        return a + ab * u;
    }
    W2 c = s.v[2].p, ac = c - a;
    float ac_ap = dot(ac, ap);
    if (ab_ap <= 0.f && ac_ap <= 0.f) {
        s.bary[0] = 1.f, s.dim = 0;
        return a;
    }
    W2 bp = O - b;
    float ab_bp = dot(ab, bp), ac_bp = dot(ac, bp);
    if (ab_bp >= 0.f && ac_bp <= ab_bp) {
        tri_swap(s, 0, 1);
        s.bary[0] = 1.f, s.dim = 0;
        return b;
    }
    W2 cp = O - c;
    float ab_cp = dot(ab, cp), ac_cp = dot(ac, cp);
    if (ac_cp >= 0.f && ab_cp <= ac_cp) {
        tri_swap(s, 0, 2);
        s.bary[0] = 1.f, s.dim = 0;
        return c;
    }
    W2 bc = c - b;
    float n = perp(ab, ac);
    if (n * perp(ab, ap) < 0.f && ab_ap >= 0.f && ab_bp <= 0.f) {  // edge ab
        float v = ab_ap / nsq(ab);
        s.bary[0] = 1.f - v, s.bary[1] = v, s.dim = 1;
        return a + ab * v;
    }
    if (-n * perp(ac, cp) < 0.f && ac_ap >= 0.f && ac_cp <= 0.f) {  // edge ac: vertices (a, c) stay, in that order
        float w = ac_ap / nsq(ac);
        tri_swap(s, 1, 2);
        s.bary[0] = 1.f - w, s.bary[1] = w, s.dim = 1;
        return a + ac * w;
    }
    if (n * perp(bc, bp) < 0.f && ac_bp - ab_bp >= 0.f && ab_cp - ac_cp >= 0.f) {  // edge bc: becomes (c, b)
        float w = dot(bc, bp) / nsq(bc);
        tri_swap(s, 0, 2);
        s.bary[0] = w, s.bary[1] = 1.f - w, s.dim = 1;
        return b + bc * w;
    }
    return O;  // inside the triangle (solid): dimension stays 2
}
__device__ void witness(const Tri2& s, bool old, W2& p1, W2& p2) {
    W2 r1 = w2(0.f, 0.f), r2 = w2(0.f, 0.f);
    int n = old ? s.old_dim : s.dim;
    for (int i = 0; i <= n; ++i) {
        float k = old ? s.old_bary[i] : s.bary[i];
        const MinkowskiPt& q = s.v[old ? s.old_idx[i] : i];
        r1 = r1 + q.o1 * k;
        r2 = r2 + q.o2 * k;
    }
    p1 = r1, p2 = r2;
}

enum { G_INSIDE = 0, G_POINTS = 1, G_APART = 3 };
__device__ int gjk2(const Operand2& g1, const Operand2& g2, float max_dist, Tri2& s, W2& p1, W2& p2, W2& axis) {
    const float rel = sqrtf(TOL10);
    W2 proj = tri_project(s), u;
    if (!unit(proj, 0.f, u)) return G_INSIDE;
    W2 prev_dir = -u, dir;
    float upper = NCB_FMAX;
    for (int it = 0;; ++it) {
        float prev_upper = upper, len;
        if (!unit_get(-proj, TOL10, dir, len)) return G_INSIDE;
        upper = len;
        if (upper >= prev_upper) {
            witness(s, true, p1, p2);
            axis = prev_dir;
            return G_POINTS;
        }
        MinkowskiPt c = minkowski(g1, g2, dir);
        float lower = -dot(dir, c.p);
        if (lower > max_dist) {
            axis = dir;
            return G_APART;
        }
        if (upper - lower <= rel * upper || !tri_add(s, c)) {
            witness(s, false, p1, p2);
            axis = dir;
            return G_POINTS;
        }
        prev_dir = dir;
        proj = tri_project(s);
        if (s.dim == 2) {
            if (lower >= TOL10) {
                witness(s, true, p1, p2);
                axis = prev_dir;
                return G_POINTS;
            }
            return G_INSIDE;
        }
        if (it + 1 == 10000) {
            axis = w2(1.f, 0.f);
            return G_APART;
        }
    }
}

// ---- EPA in the plane: the polytope is a polygon, its faces are edges ---------------------------------------------------------
#define E2_VERTS 72
#define E2_FACES 140
#define E2_HEAP 140
struct Poly2 {
    MinkowskiPt v[E2_VERTS];
    // face f: end points fa[f] -> fb[f], outward normal, projection of the origin, its barycentric coordinates, deleted flag
    uint8_t fa[E2_FACES], fb[E2_FACES], dead[E2_FACES];
    W2 fn[E2_FACES], fproj[E2_FACES];
    float fbary[E2_FACES][2];
    int nv, nf;
    // Rust BinaryHeap<FaceId> (max-heap on neg_dist)
    uint8_t hid[E2_HEAP];
    float hkey[E2_HEAP];
    int nh;
    bool overflow;
};
__device__ void heap_up(Poly2& e, int start, int pos) {
    uint8_t id = e.hid[pos];
    float key = e.hkey[pos];
    while (pos > start) {
        int parent = (pos - 1) / 2;
        if (key <= e.hkey[parent]) break;
        e.hid[pos] = e.hid[parent], e.hkey[pos] = e.hkey[parent];
        pos = parent;
    }
    e.hid[pos] = id, e.hkey[pos] = key;
}
__device__ void heap_push(Poly2& e, int id, float key) {
    if (e.nh >= E2_HEAP) {
        e.overflow = true;
        return;
    }
    e.hid[e.nh] = (uint8_t)id, e.hkey[e.nh] = key;
    e.nh++;
    heap_up(e, 0, e.nh - 1);
}
__device__ bool heap_pop(Poly2& e, int& id, float& key) {
    if (e.nh == 0) return false;
    e.nh--;
    uint8_t lid = e.hid[e.nh];
    float lkey = e.hkey[e.nh];
    if (e.nh > 0) {
        id = e.hid[0], key = e.hkey[0];
        e.hid[0] = lid, e.hkey[0] = lkey;
        int end = e.nh, pos = 0, child = 1;  // sift_down_to_bottom(0), then sift_up
        while (end >= 2 && child <= end - 2) {
            if (e.hkey[child] <= e.hkey[child + 1]) child += 1;
            e.hid[pos] = e.hid[child], e.hkey[pos] = e.hkey[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            e.hid[pos] = e.hid[child], e.hkey[pos] = e.hkey[child];
            pos = child;
        }
        e.hid[pos] = lid, e.hkey[pos] = lkey;
        heap_up(e, 0, pos);
    } else {
        id = lid, key = lkey;
    }
    return true;
}
// Face::new / new_with_proj: appends nothing, fills slot f; returns whether the origin projects inside the segment
__device__ bool face_make(Poly2& e, int f, int a, int b, bool project) {
    W2 pa = e.v[a].p, pb = e.v[b].p, ab = pb - pa;
    bool inside = false;
    W2 pr = w2(0.f, 0.f);
    float c0 = 0.f, c1 = 0.f;
    if (project) {
        float t = dot(ab, -pa), len2 = nsq(ab);
        if (len2 != 0.f && !(t < -TOL10 || t > len2 + TOL10)) {
            float k = t / len2;
            pr = pa + ab * k;
            c0 = 1.f - k, c1 = k;
            inside = true;
        }
    } else {
        c0 = 1.f;  // the segment simplex case: proj = origin, bcoords = [1, 0]
    }
    e.fa[f] = (uint8_t)a, e.fb[f] = (uint8_t)b;
    e.fproj[f] = pr;
    e.fbary[f][0] = c0, e.fbary[f][1] = c1;
    W2 nrm;
    if (unit(w2(ab.y, -ab.x), NCB_EPS, nrm)) {
        e.fn[f] = nrm, e.dead[f] = 0;
    } else {
        e.fn[f] = w2(0.f, 0.f), e.dead[f] = 1;
    }
    return inside;
}
__device__ void face_points(const Poly2& e, int f, W2& p1, W2& p2) {
    const MinkowskiPt &a = e.v[e.fa[f]], &b = e.v[e.fb[f]];
    p1 = a.o1 * e.fbary[f][0] + b.o1 * e.fbary[f][1];
    p2 = a.o2 * e.fbary[f][0] + b.o2 * e.fbary[f][1];
}
// returns 1 found, 0 None, -1 the reference would panic (peek on an empty heap), -2 capacity exceeded
__device__ int epa2(const Operand2& g1, const Operand2& g2, const Tri2& s, Poly2& e, W2& p1, W2& p2, W2& n_out) {
    e.nv = e.nf = e.nh = 0;
    e.overflow = false;
    for (int i = 0; i <= s.dim; ++i) e.v[e.nv++] = s.v[i];
    if (s.dim == 0) {
        // vertex-vertex: a normal inside both vertices' normal cones, by rotating towards the tangents (at most 100 turns each)
        W2 n = w2(0.f, 1.f), tg;
        for (int it = 0; it < 100; ++it) {
            if (!unit(support(g1, n) - e.v[0].o1, TOL100, tg) || dot(n, tg) < TOL100) break;
            n = w2(-tg.y, tg.x);
        }
        for (int it = 0; it < 100; ++it) {
            if (!unit(support(g2, -n) - e.v[0].o2, TOL100, tg) || dot(-n, tg) < TOL100) break;
            n = w2(-tg.y, tg.x);
        }
        p1 = p2 = w2(0.f, 0.f), n_out = n;
        return 1;
    }
    if (s.dim == 2) {
        if (perp(e.v[1].p - e.v[0].p, e.v[2].p - e.v[0].p) < 0.f) {
            MinkowskiPt t = e.v[1];
            e.v[1] = e.v[2];
            e.v[2] = t;
        }
        bool in0 = face_make(e, 0, 0, 1, true), in1 = face_make(e, 1, 1, 2, true), in2 = face_make(e, 2, 2, 0, true);
        e.nf = 3;
        const bool ins[3] = {in0, in1, in2};
        for (int f = 0; f < 3; ++f)
            if (ins[f]) {
                float nd = -dot(e.fn[f], e.v[f].p);
                if (nd > TOL10) return 0;  // FaceId::new -> None -> `?`
                heap_push(e, f, nd);
            }
    } else {
        face_make(e, 0, 0, 1, false);
        face_make(e, 1, 1, 0, false);
        e.nf = 2;
        float d0 = dot(e.fn[0], e.v[0].p), d1 = dot(e.fn[1], e.v[1].p);
        if (d0 > TOL10) return 0;
        heap_push(e, 0, d0);
        if (d1 > TOL10) return 0;
        heap_push(e, 1, d1);
    }
    if (e.nh == 0) return -1;
    int best = e.hid[0];
    float upper = NCB_FMAX;
    int fid;
    float key;
    for (int it = 0; heap_pop(e, fid, key);) {
        if (e.dead[fid]) continue;
        W2 fnorm = e.fn[fid];
        int fa = e.fa[fid], fb = e.fb[fid];
        if (e.nv >= E2_VERTS || e.nf + 2 > E2_FACES) return -2;
        MinkowskiPt c = minkowski(g1, g2, fnorm);
        int nid = e.nv;
        e.v[e.nv++] = c;
        float cand = dot(c.p, fnorm);
        if (cand < upper) best = fid, upper = cand;
        float cur = -key;
        if (upper - cur < TOL100) {
            face_points(e, best, p1, p2);
            n_out = e.fn[best];
            return 1;
        }
        const int ea[2] = {fa, nid}, eb[2] = {nid, fb};
        // both new faces are built before either is examined (the reference fills `new_faces` first)
        bool inside[2];
        inside[0] = face_make(e, e.nf, ea[0], eb[0], true);
        inside[1] = face_make(e, e.nf + 1, ea[1], eb[1], true);
        for (int k = 0; k < 2; ++k) {
            int f = e.nf;
            if (inside[k]) {
                float d = dot(e.fn[f], e.fproj[f]);
                if (d < cur) {  // numerical trouble: take this face
                    face_points(e, f, p1, p2);
                    n_out = e.fn[f];
                    return 1;
                }
                if (!e.dead[f]) {
                    if (-d > TOL10) return 0;
                    heap_push(e, f, -d);
                    if (e.overflow) return -2;
                }
            }
            e.nf++;
        }
        if (++it > 10000) return 0;
    }
    face_points(e, best, p1, p2);
    n_out = e.fn[best];
    return 1;
}

struct Hit2 {
    W2 w1, w2, n;
    float depth;
};

__device__ bool ball_ball(W2 c1, float r1, W2 c2, float r2, float prediction, Hit2& h) {
    W2 d = c2 - c1;
    float d2 = nsq(d), sum = r1 + r2, lim = sum + prediction;
    if (!(d2 < lim * lim)) return false;
    W2 n = d2 != 0.f ? normalized(d) : w2(1.f, 0.f);
    h.w1 = c1 + n * r1, h.w2 = c2 + n * (-r2), h.n = n, h.depth = sum - sqrtf(d2);
    return true;
}

// AABB::project_point_with_feature for the cuboid's box, then contact_ball_convex_polyhedron
#define FEAT2_UNKNOWN 0xffffffffu
#define FEAT2_FACE 0x40000000u    // | face index
#define FEAT2_VERTEX 0x80000000u  // | vertex index
__device__ bool ball_cuboid(W2 center, float radius, const Operand2& box, float prediction, Hit2& h, uint32_t* feature = nullptr) {
    float he[2] = {box.a, box.b};
    W2 l = to_local(box.m, center);
    float lp[2] = {l.x, l.y}, below[2], above[2], shift[2];
    for (int i = 0; i < 2; ++i) {
        below[i] = -he[i] - lp[i], above[i] = lp[i] - he[i];
        shift[i] = fmaxf(below[i], 0.f) - fmaxf(above[i], 0.f);
    }
    bool inside = shift[0] == 0.f && shift[1] == 0.f;
    if (inside) {  // solid = false: move to the nearest side
        float best = -NCB_FMAX;
        bool low = false;
        int axis = 0;
        for (int i = 0; i < 2; ++i) {
            if (below[i] < above[i]) {
                if (above[i] > best) axis = i, low = false, best = above[i];
            } else if (below[i] > best) {
                axis = i, low = true, best = below[i];
            }
        }
        shift[0] = shift[1] = 0.f;
        shift[axis] = low ? best : -best;
    }
    lp[0] += shift[0], lp[1] += shift[1];
    W2 world2 = to_world(box.m, w2(lp[0], lp[1]));
    if (feature) {  // AABB::project_point_with_feature (point_aabb.rs:83-134, dim2): the manifold generator reports it with the contact
        int z = (shift[0] == 0.f) + (shift[1] == 0.f), moved = shift[0] != 0.f ? 0 : 1;
        uint32_t f = FEAT2_UNKNOWN;
        if (z == 2) {
            for (int i = 0; i < 2 && f == FEAT2_UNKNOWN; ++i) {
                if (lp[i] > he[i] - NCB_EPS)
                    f = FEAT2_FACE | (uint32_t)i;
                else if (lp[i] <= -he[i] + NCB_EPS)
                    f = FEAT2_FACE | (uint32_t)(i + 2);
            }
        } else if (z == 1) {
            f = FEAT2_FACE | (uint32_t)(lp[moved] < (-he[moved] + he[moved]) * 0.5f ? moved + 2 : moved);
        } else {
            f = FEAT2_VERTEX | (lp[0] < 0.f ? 1u : 0u) | (lp[1] < 0.f ? 2u : 0u);
        }
        *feature = f;
    }
    W2 dpt = world2 - center, dir, normal;
    float dist, depth;
    if (unit_get(dpt, NCB_EPS, dir, dist)) {
        depth = inside ? dist + radius : -dist + radius;
        normal = inside ? -dir : dir;
    } else {
        // the centre lies on the boundary: the feature's own normal, in the cuboid's LOCAL frame like the reference (:50)
        int zeros = (shift[0] == 0.f) + (shift[1] == 0.f);
        int moved_axis = shift[0] != 0.f ? 0 : 1;
        W2 fnrm;
        if (zeros == 2) {
            int face = -1;
            for (int i = 0; i < 2 && face < 0; ++i) {
                if (lp[i] > he[i] - NCB_EPS)
                    face = i;
                else if (lp[i] <= -he[i] + NCB_EPS)
                    face = i + 2;
            }
            if (face < 0) return false;  // FeatureId::Unknown
            fnrm = w2(face == 0 ? 1.f : face == 2 ? -1.f : 0.f, face == 1 ? 1.f : face == 3 ? -1.f : 0.f);
        } else if (zeros == 1) {
            bool neg = lp[moved_axis] < (-he[moved_axis] + he[moved_axis]) * 0.5f;
            fnrm = moved_axis == 0 ? w2(neg ? -1.f : 1.f, 0.f) : w2(0.f, neg ? -1.f : 1.f);
        } else {
            fnrm = normalized(w2(lp[0] < 0.f ? -1.f : 1.f, lp[1] < 0.f ? -1.f : 1.f));
        }
        depth = radius;
        normal = -fnrm;
    }
    if (!(depth >= -prediction)) return false;
    h.w1 = center + normal * radius, h.w2 = world2, h.n = normal, h.depth = depth;
    return true;
}

// ConvexPolygon::project_point_with_feature + contact_ball_convex_polyhedron: the ball centre is projected on the polygon with
// GJK against the constant origin (EPA when it lies inside), in the frame translated by -centre
// Segment::project_point_with_feature (query/point/point_segment.rs:14-91, dim2) + Segment::feature_normal (segment.rs:237-284), then
// contact_ball_convex_polyhedron (contact_ball_convex_polyhedron.rs:12-62)
__device__ bool ball_segment(W2 center, float radius, const Operand2& seg, float prediction, Hit2& h, uint32_t* feature = nullptr) {
    W2 a = w2(seg.a, seg.b), b = w2(seg.c, seg.d), ls = to_local(seg.m, center);
    W2 ab = b - a, ap = ls - a;
    float ab_ap = dot(ab, ap), sqnab = nsq(ab);
    W2 world2;
    uint32_t f;
    bool on_edge = false;
    if (ab_ap <= 0.f) {
        f = FEAT2_VERTEX | 0u, world2 = to_world(seg.m, a);
    } else if (ab_ap >= sqnab) {
        f = FEAT2_VERTEX | 1u, world2 = to_world(seg.m, b);
    } else {
        float u = ab_ap / sqnab;
        world2 = to_world(seg.m, a + ab * u);
        on_edge = true;
    }
    const bool inside = relative_eq(world2.x, center.x) && relative_eq(world2.y, center.y);
    if (on_edge) f = perp(center - world2, ab) >= 0.f ? (FEAT2_FACE | 0u) : (FEAT2_FACE | 1u);
    if (feature) *feature = f;
    W2 dpt = world2 - center, dir, normal;
    float dist, depth;
    if (unit_get(dpt, NCB_EPS, dir, dist)) {
        depth = inside ? dist + radius : -dist + radius;
        normal = inside ? -dir : dir;
    } else {
        W2 sd, fn = w2(0.f, 1.f);  // feature_normal: no direction -> the y axis
        if (unit(ab, NCB_EPS, sd)) {
            uint32_t id = f & 0xffffu;
            fn = (f & FEAT2_VERTEX) ? (id == 0 ? sd : -sd) : (id == 0 ? w2(sd.y, -sd.x) : w2(-sd.y, sd.x));
        }
        depth = radius;
        normal = -fn;
    }
    if (!(depth >= -prediction)) return false;
    h.w1 = center + normal * radius, h.w2 = world2, h.n = normal, h.depth = depth;
    return true;
}

__device__ bool ball_polygon(W2 center, float radius, const Operand2& poly, float prediction, float cos_one_degree, Hit2& h, int& epa_status,
                             uint32_t* feature = nullptr) {
    Operand2 g = poly, origin;
    g.m.t = (-center) + poly.m.t;
    origin.kind = D2_ORIGIN, origin.a = origin.b = 0.f, origin.pts = origin.nrm = nullptr, origin.npts = 0;
    origin.m.t = w2(0.f, 0.f), origin.m.re = 1.f, origin.m.im = 0.f;
    W2 d0;
    if (!unit(-g.m.t, NCB_EPS, d0)) d0 = w2(1.f, 0.f);
    Tri2 s;
    for (int i = 0; i < 3; ++i) s.v[i].p = s.v[i].o1 = s.v[i].o2 = w2(0.f, 0.f), s.old_idx[i] = i;
    s.bary[0] = s.bary[1] = s.old_bary[0] = s.old_bary[1] = 0.f;
    s.dim = s.old_dim = 0;
    s.v[0] = minkowski(g, origin, d0);
    W2 p1, p2, n;
    bool inside = gjk2(g, origin, NCB_FMAX, s, p1, p2, n) != G_POINTS;
    W2 world2;
    if (!inside) {
        world2 = p1 + center;
    } else {
        Poly2 e;
        epa_status = epa2(g, origin, s, e, p1, p2, n);
        world2 = epa_status == 1 ? p1 + center : center;
    }
    W2 back = center - world2;
    W2 ldir = unrotate(poly.m, inside ? -back : back), lu;
    int face = -1, vert = -1;
    if (unit(ldir, NCB_EPS, lu)) {  // support_feature_id_toward
        for (uint32_t i = 0; i < poly.npts && face < 0; ++i)
            if (__ldg(poly.nrm + 2 * i) * lu.x + __ldg(poly.nrm + 2 * i + 1) * lu.y >= cos_one_degree) face = (int)i;
        if (face < 0) {
            vert = 0;
            float best = __ldg(poly.pts) * lu.x + __ldg(poly.pts + 1) * lu.y;
            for (uint32_t i = 1; i < poly.npts; ++i) {
                float d = __ldg(poly.pts + 2 * i) * lu.x + __ldg(poly.pts + 2 * i + 1) * lu.y;
                if (d > best) best = d, vert = (int)i;
            }
        }
    }
    if (feature) *feature = face >= 0 ? (FEAT2_FACE | (uint32_t)face) : vert >= 0 ? (FEAT2_VERTEX | (uint32_t)vert) : FEAT2_UNKNOWN;
    W2 dpt = world2 - center, dir, normal;
    float dist, depth;
    if (unit_get(dpt, NCB_EPS, dir, dist)) {
        depth = inside ? dist + radius : -dist + radius;
        normal = inside ? -dir : dir;
    } else {
        if (face < 0 && vert < 0) return false;  // FeatureId::Unknown
        W2 fnrm;
        if (face >= 0) {
            fnrm = w2(__ldg(poly.nrm + 2 * face), __ldg(poly.nrm + 2 * face + 1));
        } else {
            int prev = vert == 0 ? (int)poly.npts - 1 : vert - 1;
            fnrm = normalized(w2(__ldg(poly.nrm + 2 * prev), __ldg(poly.nrm + 2 * prev + 1)) + w2(__ldg(poly.nrm + 2 * vert), __ldg(poly.nrm + 2 * vert + 1)));
        }
        depth = radius;
        normal = -fnrm;
    }
    if (!(depth >= -prediction)) return false;
    h.w1 = center + normal * radius, h.w2 = world2, h.n = normal, h.depth = depth;
    return true;
}

// =============================================================================================================================
// 2-D world update: fat AABBs, the broad phase of the 3-D path on boxes with z = 0 (the pair set is the set of intersecting fat
// boxes whatever the tree), and the contact-manifold generators of the 2-D crate:
//   AABBs        bounding_volume/aabb_ball.rs:8-13, aabb_cuboid.rs:9-14, aabb_convex_polygon.rs + aabb_utils.rs:59-79,
//                pipeline/object/collision_object.rs:89-93 (query limit), dbvt_broad_phase.rs:341 (margin)
//   generators   ball_ball_manifold_generator.rs:29-66, ball_convex_polyhedron_manifold_generator.rs:29-116,
//                convex_polyhedron_convex_polyhedron_manifold_generator.rs:81-165 with shape/convex_polygonal_feature2.rs:95-236,
//                shape/cuboid.rs:186-226,279-350 (dim2), shape/convex_polygon.rs:123-184
//   manifold     query/contact/contact_manifold.rs:165-236 (fresh manifold, DistanceBased(0.02))
// =============================================================================================================================
__device__ __forceinline__ Operand2 load_operand(uint32_t t, float4 p, float4 m, const float* poly, const float* poly_nrm);
struct Edge2 {  // ConvexPolygonalFeature in 2-D: at most two vertices
    W2 v[2];
    uint32_t vid[2], fid;
    W2 normal;
    int nv;
    bool has_normal;
};
__device__ __forceinline__ void edge_clear(Edge2& e) { e.nv = 0, e.has_normal = false, e.fid = FEAT2_UNKNOWN; }
__device__ __forceinline__ void edge_to_world(Edge2& e, const Pose2& m) {
    e.v[0] = to_world(m, e.v[0]), e.v[1] = to_world(m, e.v[1]);
    if (e.has_normal) e.normal = rotate(m, e.normal);
}
__device__ void box_face(const Operand2& g, int i, Edge2& e) {
    edge_clear(e);
    int i1 = i < 2 ? i : i - 2, i2 = (i1 + 1) % 2;
    float sign = i < 2 ? 1.f : -1.f;
    float c[2] = {g.a, g.b};
    c[i1] *= sign;
    c[i2] *= (i1 == 0) ? -sign : sign;
    W2 p1 = w2(c[0], c[1]);
    c[i2] = -c[i2];
    W2 p2 = w2(c[0], c[1]);
    uint32_t id1 = sign < 0.f ? (1u << i1) : 0u, id2 = id1;
    if ((i2 == 0 ? p1.x : p1.y) < 0.f)
        id1 |= 1u << i2;
    else
        id2 |= 1u << i2;
    e.v[0] = p1, e.vid[0] = FEAT2_VERTEX | id1;
    e.v[1] = p2, e.vid[1] = FEAT2_VERTEX | id2;
    e.nv = 2;
    e.normal = i1 == 0 ? w2(sign, 0.f) : w2(0.f, sign), e.has_normal = true;
    e.fid = FEAT2_FACE | (uint32_t)i;
}
__device__ void ngon_face(const Operand2& g, uint32_t ia, Edge2& e) {
    edge_clear(e);
    uint32_t ib = (ia + 1) % g.npts;
    e.v[0] = w2(__ldg(g.pts + 2 * ia), __ldg(g.pts + 2 * ia + 1)), e.vid[0] = FEAT2_VERTEX | ia;
    e.v[1] = w2(__ldg(g.pts + 2 * ib), __ldg(g.pts + 2 * ib + 1)), e.vid[1] = FEAT2_VERTEX | ib;
    e.nv = 2;
    e.normal = w2(__ldg(g.nrm + 2 * ia), __ldg(g.nrm + 2 * ia + 1)), e.has_normal = true;
    e.fid = FEAT2_FACE | ia;
}
// Segment::face (segment.rs:212-235, dim2) for the segment (a, b); a degenerate one is the vertex a
__device__ void segment_face(W2 a, W2 b, uint32_t id, Edge2& e) {
    edge_clear(e);
    W2 ab = b - a, nrm;
    if (unit(w2(ab.y, -ab.x), NCB_EPS, nrm)) {
        e.fid = FEAT2_FACE | id, e.nv = 2, e.has_normal = true;
        if (id == 0)
            e.v[0] = a, e.vid[0] = FEAT2_VERTEX | 0u, e.v[1] = b, e.vid[1] = FEAT2_VERTEX | 1u, e.normal = nrm;
        else
            e.v[0] = b, e.vid[0] = FEAT2_VERTEX | 1u, e.v[1] = a, e.vid[1] = FEAT2_VERTEX | 0u, e.normal = -nrm;
    } else {
        e.v[0] = a, e.vid[0] = FEAT2_VERTEX | 0u, e.nv = 1, e.fid = FEAT2_VERTEX | 0u;
    }
}
__device__ void face_toward(const Operand2& g, W2 dir, Edge2& e) {
    if (g.kind == D2_SEGMENT) {  // segment.rs:286-299 (dim2): the world `dir` against the LOCAL segment direction, as in the reference
        W2 a = w2(g.a, g.b), b = w2(g.c, g.d);
        segment_face(a, b, perp(dir, b - a) >= 0.f ? 0u : 1u, e);
        edge_to_world(e, g.m);
        return;
    }
    W2 ld = unrotate(g.m, dir);
    if (g.kind == D2_CUBOID) {
        int iamax = fabsf(ld.y) > fabsf(ld.x) ? 1 : 0;
        box_face(g, (iamax == 0 ? ld.x : ld.y) > 0.f ? iamax : iamax + 2, e);
    } else {
        uint32_t arg = 0;
        float best = __ldg(g.nrm) * ld.x + __ldg(g.nrm + 1) * ld.y;
        for (uint32_t i = 1; i < g.npts; ++i) {
            float d = __ldg(g.nrm + 2 * i) * ld.x + __ldg(g.nrm + 2 * i + 1) * ld.y;
            if (d > best) best = d, arg = i;
        }
        ngon_face(g, arg, e);
    }
    edge_to_world(e, g.m);
}
__device__ void feature_toward(const Operand2& g, W2 dir, float cang, float sang, Edge2& e) {
    if (g.kind == D2_SEGMENT) {  // segment.rs:315-345 (dim2), sang = sin(angular prediction)
        edge_clear(e);
        W2 a = to_world(g.m, w2(g.a, g.b)), b = to_world(g.m, w2(g.c, g.d)), sd;
        if (unit(b - a, NCB_EPS, sd)) {
            float c = dot(dir, sd);
            if (c > sang)
                e.fid = FEAT2_VERTEX | 1u, e.v[0] = b, e.vid[0] = FEAT2_VERTEX | 1u, e.nv = 1;
            else if (c < -sang)
                e.fid = FEAT2_VERTEX | 0u, e.v[0] = a, e.vid[0] = FEAT2_VERTEX | 0u, e.nv = 1;
            else
                segment_face(a, b, perp(dir, sd) >= 0.f ? 0u : 1u, e);
        }
        return;
    }
    if (g.kind != D2_CUBOID) {  // ConvexPolygon::support_feature_toward is its support face
        face_toward(g, dir, e);
        return;
    }
    W2 ld = unrotate(g.m, dir);
    float l[2] = {ld.x, ld.y}, sp[2] = {g.a, g.b};
    edge_clear(e);
    uint32_t id = 0;
    for (int i = 0; i < 2; ++i) {
        float sign = signbit(l[i]) ? -1.f : 1.f;  // f32::signum (-0.0 -> -1.0); NaN directions do not reach this point
        if (sign * l[i] >= cang) {
            box_face(g, sign > 0.f ? i : i + 2, e);
            edge_to_world(e, g.m);
            return;
        }
        if (sign < 0.f) id |= 1u << i;
        sp[i] *= sign;
    }
    e.v[0] = to_world(g.m, w2(sp[0], sp[1])), e.vid[0] = FEAT2_VERTEX | id, e.nv = 1;
    e.fid = FEAT2_VERTEX | id;
}

#define MAN2_MAX 4
struct Manifold2d {
    Hit2 c[MAN2_MAX];
    uint32_t f1[MAN2_MAX], f2[MAN2_MAX];
    W2 track[MAN2_MAX];
    int n;
    bool overflow;
};
__device__ void man_push(Manifold2d& mf, const Hit2& h, uint32_t f1, uint32_t f2, W2 tracking) {
    int hit = mf.n;
    float lim = 0.02f * 0.02f;
    for (int i = 0; i < mf.n; ++i) {
        float d = nsq(tracking - mf.track[i]);
        if (d < lim) lim = d, hit = i;
    }
    if (hit == mf.n) {
        if (mf.n == MAN2_MAX) {
            mf.overflow = true;
            return;
        }
        mf.n++;
    } else if (!(h.depth > mf.c[hit].depth)) {
        return;  // the contact already there is deeper
    }
    mf.c[hit] = h, mf.f1[hit] = f1, mf.f2[hit] = f2, mf.track[hit] = tracking;
}
// ConvexPolygonalFeature::clip in 2-D: the overlap of the two segments along the direction orthogonal to the normal
__device__ void clip_edges(const Edge2& a, const Edge2& b, W2 normal, float prediction, const Pose2& m1, Manifold2d& mf, int& n_new) {
    if (a.nv <= 1 || b.nv <= 1) return;
    W2 ortho = w2(-normal.y, normal.x);
    W2 a0 = a.v[0], a1 = a.v[1], b0 = b.v[0], b1 = b.v[1];
    float ra0 = dot(a0 - a0, ortho), ra1 = dot(a1 - a0, ortho), rb0 = dot(b0 - a0, ortho), rb1 = dot(b1 - a0, ortho);
    uint32_t fa0 = a.vid[0], fa1 = a.vid[1], fb0 = b.vid[0], fb1 = b.vid[1];
    if (ra1 < ra0) {
        float t = ra0;
        ra0 = ra1, ra1 = t;
        uint32_t u = fa0;
        fa0 = fa1, fa1 = u;
        W2 p = a0;
        a0 = a1, a1 = p;
    }
    if (rb1 < rb0) {
        float t = rb0;
        rb0 = rb1, rb1 = t;
        uint32_t u = fb0;
        fb0 = fb1, fb1 = u;
        W2 p = b0;
        b0 = b1, b1 = p;
    }
    if (rb0 > ra1 || ra0 > rb1) return;
    float la = ra1 - ra0, lb = rb1 - rb0;
    // both candidates are found first and pushed afterwards, in the reference's order
    Hit2 cand[2];
    uint32_t cf1[2], cf2[2];
    for (int side = 0; side < 2; ++side) {
        W2 w1, w2_;
        bool on_a = side == 0 ? rb0 > ra0 : rb1 < ra1;  // the end point of b's range falls inside a's range: project it on a
        if (on_a) {
            float k = ((side == 0 ? rb0 : rb1) - ra0) / la;
            w1 = w2(a0.x * (1.f - k) + a1.x * k, a0.y * (1.f - k) + a1.y * k);
            w2_ = side == 0 ? b0 : b1;
            cf1[side] = a.fid, cf2[side] = side == 0 ? fb0 : fb1;
        } else {
            float k = ((side == 0 ? ra0 : ra1) - rb0) / lb;
            w1 = side == 0 ? a0 : a1;
            w2_ = w2(b0.x * (1.f - k) + b1.x * k, b0.y * (1.f - k) + b1.y * k);
            cf1[side] = side == 0 ? fa0 : fa1, cf2[side] = b.fid;
        }
        cand[side].w1 = w1, cand[side].w2 = w2_, cand[side].n = normal, cand[side].depth = -dot(normal, w2_ - w1);
    }
    for (int side = 0; side < 2; ++side)
        if (-cand[side].depth <= prediction) {
            n_new++;
            if (cf1[side] != FEAT2_UNKNOWN && cf2[side] != FEAT2_UNKNOWN) man_push(mf, cand[side], cf1[side], cf2[side], to_local(m1, cand[side].w1));
        }
}

// bounding_volume::aabb(shape, m) in 2-D
__device__ void aabb_of_shape(const Operand2& g, W2& lo, W2& hi) {
    if (g.kind == D2_PLANE) {  // aabb_plane.rs:13-21: half of f32::MAX either way; the 3-D broad phase keeps such boxes out of the tree
        lo = w2(-NCB_FMAX * 0.5f, -NCB_FMAX * 0.5f), hi = w2(NCB_FMAX * 0.5f, NCB_FMAX * 0.5f);
    } else if (g.kind == D2_BALL) {
        lo = w2(g.m.t.x + (-g.a), g.m.t.y + (-g.a)), hi = w2(g.m.t.x + g.a, g.m.t.y + g.a);
    } else if (g.kind == D2_CUBOID) {
        float are = fabsf(g.m.re), aim = fabsf(g.m.im);
        W2 w = w2(are * g.a + aim * g.b, aim * g.a + are * g.b);
        lo = g.m.t - w, hi = g.m.t + w;
    } else if (g.kind == D2_SEGMENT) {  // support_map_aabb (aabb_utils.rs:9-31): one support point per axis direction
        hi = w2(support(g, w2(1.f, 0.f)).x, support(g, w2(0.f, 1.f)).y);
        lo = w2(support(g, w2(-1.f, 0.f)).x, support(g, w2(0.f, -1.f)).y);
    } else {
        W2 p = to_world(g.m, w2(__ldg(g.pts), __ldg(g.pts + 1)));
        lo = hi = p;
        for (uint32_t k = 1; k < g.npts; ++k) {
            p = to_world(g.m, w2(__ldg(g.pts + 2 * k), __ldg(g.pts + 2 * k + 1)));
            lo = w2(fminf(lo.x, p.x), fminf(lo.y, p.y)), hi = w2(fmaxf(hi.x, p.x), fmaxf(hi.y, p.y));
        }
    }
}
// One pair through its ContactManifoldGenerator into a fresh manifold.  flags: bit 0 = the reference would panic, bit 1 = EPA capacity.
__device__ void manifold_of_pair(const Operand2& g1, const Operand2& g2, float linear, float cang1, float cang2, float sang1, float sang2, float cos_one_degree,
                                 Manifold2d& mf, int& flags) {
    mf.n = 0, mf.overflow = false;
    const uint32_t FACE0 = FEAT2_FACE | 0u;
    Hit2 h;
    if (g1.kind == D2_PLANE && g2.kind == D2_PLANE) {
        // no contact algorithm for two planes: the pair has no interaction edge
    } else if (g1.kind == D2_BALL && g2.kind == D2_BALL) {
        if (ball_ball(g1.m.t, g1.a, g2.m.t, g2.a, linear, h)) man_push(mf, h, FACE0, FACE0, w2(0.f, 0.f));
    } else if (g1.kind == D2_PLANE || g2.kind == D2_PLANE) {
        // PlaneBallManifoldGenerator / PlaneConvexPolyhedronManifoldGenerator (plane_ball_manifold_generator.rs:40-77,
        // plane_convex_polyhedron_manifold_generator.rs:40-85), flip = the plane is the second object
        const bool flip = g1.kind != D2_PLANE;
        const Operand2& pl = flip ? g2 : g1;
        const Operand2& ot = flip ? g1 : g2;
        W2 n = rotate(pl.m, w2(pl.a, pl.b)), center = pl.m.t;
        if (ot.kind == D2_BALL) {
            float dist = dot(ot.m.t - center, n), depth = -dist + ot.a;
            if (depth > -linear) {
                W2 on_plane = ot.m.t + n * (-dist), on_ball = ot.m.t + n * (-ot.a);
                if (!flip)
                    h.w1 = on_plane, h.w2 = on_ball, h.n = n;
                else
                    h.w1 = on_ball, h.w2 = on_plane, h.n = -n;
                h.depth = depth;
                man_push(mf, h, FACE0, FACE0, w2(0.f, 0.f));
            }
        } else {
            Edge2 f;
            face_toward(ot, -n, f);
            for (int i = 0; i < 2; ++i) {  // both slots of the feature's vertex array, like the reference's iteration
                W2 on_shape = f.v[i];
                float dist = dot(on_shape - center, n);
                if (dist <= linear) {
                    W2 on_plane = on_shape + (-n) * dist;
                    W2 track = to_local(ot.m, on_shape);
                    h.depth = -dist;
                    if (!flip) {
                        h.w1 = on_plane, h.w2 = on_shape, h.n = n;
                        man_push(mf, h, FACE0, f.vid[i], track);
                    } else {
                        h.w1 = on_shape, h.w2 = on_plane, h.n = -n;
                        man_push(mf, h, f.vid[i], FACE0, track);
                    }
                }
            }
        }
    } else if (g1.kind == D2_BALL || g2.kind == D2_BALL) {
        const bool flip = g1.kind != D2_BALL;
        const Operand2& ball = flip ? g2 : g1;
        const Operand2& other = flip ? g1 : g2;
        uint32_t f2 = FEAT2_UNKNOWN;
        int q = 1;
        bool ok = other.kind == D2_CUBOID    ? ball_cuboid(ball.m.t, ball.a, other, linear, h, &f2)
                  : other.kind == D2_SEGMENT ? ball_segment(ball.m.t, ball.a, other, linear, h, &f2)
                                             : ball_polygon(ball.m.t, ball.a, other, linear, cos_one_degree, h, q, &f2);
        if (q == -1) flags |= 1;
        if (q == -2) flags |= 2;
        if (ok) {
            if (f2 == FEAT2_UNKNOWN) {
                flags |= 1;  // "Feature id cannot be unknown."
            } else if (!flip) {
                man_push(mf, h, FACE0, f2, w2(0.f, 0.f));
            } else {
                W2 t = h.w1;
                h.w1 = h.w2, h.w2 = t, h.n = -h.n;
                man_push(mf, h, f2, FACE0, w2(0.f, 0.f));
            }
        }
    } else {
        W2 d0;
        if (!unit(g2.m.t - g1.m.t, NCB_EPS, d0)) d0 = w2(1.f, 0.f);
        Tri2 s;
        for (int i = 0; i < 3; ++i) s.v[i].p = s.v[i].o1 = s.v[i].o2 = w2(0.f, 0.f), s.old_idx[i] = i;
        s.bary[0] = s.bary[1] = s.old_bary[0] = s.old_bary[1] = 0.f;
        s.dim = s.old_dim = 0;
        s.v[0] = minkowski(g1, g2, d0);
        W2 p1, p2, n;
        int r = gjk2(g1, g2, linear, s, p1, p2, n);
        bool ok = r == G_POINTS;
        if (r == G_INSIDE) {
            Poly2 e;
            int q = epa2(g1, g2, s, e, p1, p2, n);
            ok = q == 1;
            if (q == -1) flags |= 1;
            if (q == -2) flags |= 2;
        }
        if (ok) {
            h.w1 = p1, h.w2 = p2, h.n = n, h.depth = -dot(n, p2 - p1);
            Edge2 fa, fb;
            if (h.depth > 0.f) {
                face_toward(g1, n, fa);
                face_toward(g2, -n, fb);
            } else {
                feature_toward(g1, n, cang1, sang1, fa);
                feature_toward(g2, -n, cang2, sang2, fb);
            }
            int n_new = 0;
            clip_edges(fa, fb, n, linear, g1.m, mf, n_new);
            if (n_new == 0 && fa.fid != FEAT2_UNKNOWN && fb.fid != FEAT2_UNKNOWN) man_push(mf, h, fa.fid, fb.fid, to_local(g1.m, h.w1));
        }
    }
}
struct Args2 {
    uint32_t n;
    const uint32_t *type1, *type2;
    const float4 *param1, *param2, *pose1, *pose2;
    const float* poly;
    const float* poly_nrm;
    float prediction;
    float cos_one_degree;  // cos(pi / 180) in f32 from the host libm (convex_polygon.rs:187-188)
    uint8_t* found;
    float* out;
    uint32_t* counters;  // [0] reference panics, [1] EPA capacity overflows
};
__device__ __forceinline__ Operand2 load_operand(uint32_t t, float4 p, float4 m, const float* poly, const float* poly_nrm) {
    Operand2 g;
    g.kind = t, g.a = p.x, g.b = p.y, g.c = p.z, g.d = p.w, g.pts = g.nrm = nullptr, g.npts = 0;
    if (t == D2_POLYGON) {
        g.pts = poly + 2 * (size_t)p.x, g.npts = (uint32_t)p.y;
        g.nrm = poly_nrm ? poly_nrm + 2 * (size_t)p.x : nullptr;
    }
    g.m.t = w2(m.x, m.y), g.m.re = m.z, g.m.im = m.w;
    return g;
}
// contact_plane_support_map (query/contact/contact_plane_support_map.rs:8-28)
__device__ bool plane_support(const Operand2& plane, const Operand2& other, float prediction, Hit2& h) {
    W2 n = rotate(plane.m, w2(plane.a, plane.b));
    W2 deepest = other.kind == D2_BALL ? other.m.t + (-n) * other.a : support(other, -n);  // support_point_toward: a ball takes the unit direction as is
    float distance = dot(n, plane.m.t - deepest);
    if (!(distance > -prediction)) return false;
    h.w1 = deepest + n * distance, h.w2 = deepest, h.n = n, h.depth = distance;
    return true;
}

// query::proximity for one pair (query/proximity/proximity_shape_shape.rs:8-33): proximity_ball_ball.rs:8-36,
// proximity_plane_support_map.rs:9-47 (either order), proximity_support_map_support_map.rs:12-75 = the GJK above with exact_dist = false
// (gjk.rs:76-177: its Proximity exits).  Returns NCB_PROXIMITY_INTERSECTING / _WITHIN_MARGIN / _DISJOINT.
__device__ uint8_t proximity_of_pair(const Operand2& g1, const Operand2& g2, float margin) {
    if (g1.kind == D2_BALL && g2.kind == D2_BALL) {
        float d2 = nsq(g2.m.t - g1.m.t), sum = g1.a + g2.a, lim = sum + margin;
        if (d2 <= lim * lim) return d2 <= sum * sum ? NCB_PROXIMITY_INTERSECTING : NCB_PROXIMITY_WITHIN_MARGIN;
        return NCB_PROXIMITY_DISJOINT;
    }
    if (g1.kind == D2_PLANE || g2.kind == D2_PLANE) {
        const Operand2& pl = g1.kind == D2_PLANE ? g1 : g2;
        const Operand2& ot = g1.kind == D2_PLANE ? g2 : g1;
        W2 n = rotate(pl.m, w2(pl.a, pl.b));
        W2 deepest = ot.kind == D2_BALL ? ot.m.t + (-n) * ot.a : support(ot, -n);
        float distance = dot(n, pl.m.t - deepest);
        if (distance >= -margin) return distance >= 0.f ? NCB_PROXIMITY_INTERSECTING : NCB_PROXIMITY_WITHIN_MARGIN;
        return NCB_PROXIMITY_DISJOINT;
    }
    const float rel = sqrtf(TOL10);
    W2 dir, u;
    if (!unit(g2.m.t - g1.m.t, NCB_EPS, dir)) dir = w2(1.f, 0.f);
    Tri2 s;
    for (int i = 0; i < 3; ++i) s.v[i].p = s.v[i].o1 = s.v[i].o2 = w2(0.f, 0.f), s.old_idx[i] = i;
    s.bary[0] = s.bary[1] = s.old_bary[0] = s.old_bary[1] = 0.f;
    s.dim = s.old_dim = 0;
    s.v[0] = minkowski(g1, g2, dir);
    W2 proj = tri_project(s);
    if (!unit(proj, 0.f, u)) return NCB_PROXIMITY_INTERSECTING;
    float upper = NCB_FMAX;
    for (int it = 0;; ++it) {
        float prev_upper = upper, len;
        if (!unit_get(-proj, TOL10, dir, len)) return NCB_PROXIMITY_INTERSECTING;
        upper = len;
        if (upper >= prev_upper) return NCB_PROXIMITY_WITHIN_MARGIN;
        MinkowskiPt c = minkowski(g1, g2, dir);
        float lower = -dot(dir, c.p);
        if (lower > margin) return NCB_PROXIMITY_DISJOINT;
        if ((lower > 0.f && upper <= margin) || upper - lower <= rel * upper || !tri_add(s, c)) return NCB_PROXIMITY_WITHIN_MARGIN;
        proj = tri_project(s);
        if (s.dim == 2) return lower >= TOL10 ? NCB_PROXIMITY_WITHIN_MARGIN : NCB_PROXIMITY_INTERSECTING;
        if (it + 1 == 10000) return NCB_PROXIMITY_DISJOINT;
    }
}

// ---- RayCast for the 2-D shapes, solid = true ------------------------------------------------------------------------------------------
//   query/ray/ray_ball.rs:15-142, ray_cuboid.rs + ray_aabb.rs:52-75,183-300 (the far-side face id carries the reference's `+ 3`),
//   ray_plane.rs:9-79, ray_support_map.rs:15-60,165-189 (ConvexPolygon) -> gjk::cast_ray (gjk.rs:180-365) on the simplex above.
struct RayHit2 {
    bool hit;
    float toi;
    W2 n;
    uint32_t feature;
};
__device__ __forceinline__ RayHit2 ray2_miss() { return RayHit2{false, 0.f, W2{0.f, 0.f}, FEAT2_UNKNOWN}; }

__device__ RayHit2 ray2_ball(W2 center, float radius, W2 o, W2 d, float max_toi) {
    RayHit2 h = ray2_miss();
    W2 dc = o - center;
    float a = nsq(d), b = dot(dc, d), c = nsq(dc) - radius * radius;
    bool inside = false;
    float t = 0.f;
    if (a == 0.f) {
        if (c > 0.f) return h;
        inside = true;
    } else if (c > 0.f && b > 0.f) {
        return h;
    } else {
        float delta = b * b - a * c;
        if (delta < 0.f) return h;
        t = (-b - sqrtf(delta)) / a;
        if (t <= 0.f) inside = true, t = 0.f;
    }
    if (!(t <= max_toi)) return h;
    W2 normal = normalized((o + d * t) - center);
    h.hit = true, h.toi = t, h.n = inside ? -normal : normal, h.feature = FEAT2_FACE;
    return h;
}

__device__ RayHit2 ray2_cuboid(const Operand2& g, W2 o_w, W2 d_w, float max_toi) {
    RayHit2 h = ray2_miss();
    W2 o = to_local(g.m, o_w), d = unrotate(g.m, d_w);
    const float oo[2] = {o.x, o.y}, dd[2] = {d.x, d.y}, he[2] = {g.a, g.b};
    float tmax = NCB_FMAX, tmin = -NCB_FMAX;
    int near_side = 0, far_side = 0;
    bool near_diag = false;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (dd[i] == 0.f) {
            if (oo[i] < -he[i] || oo[i] > he[i]) return h;
        } else {
            float denom = 1.f / dd[i];
            float nr = (-he[i] - oo[i]) * denom, fr = (he[i] - oo[i]) * denom;
            bool flip = false;
            if (nr > fr) {
                float t = nr;
                nr = fr, fr = t, flip = true;
            }
            if (nr > tmin)
                tmin = nr, near_side = flip ? -(i + 1) : (i + 1), near_diag = false;
            else if (nr == tmin)
                near_diag = true;
            if (fr < tmax) tmax = fr, far_side = !flip ? -(i + 1) : (i + 1);
            if (tmax < 0.f || tmin > tmax) return h;
        }
    }
    W2 near_n = w2(0.f, 0.f);
    if (near_diag)
        near_n = -normalized(d);
    else if (near_side == 1 || near_side == -1)
        near_n.x = near_side < 0 ? 1.f : -1.f;
    else if (near_side == 2 || near_side == -2)
        near_n.y = near_side < 0 ? 1.f : -1.f;
    float t;
    W2 n;
    int side;
    if (tmin < 0.f)
        t = 0.f, n = w2(0.f, 0.f), side = far_side;
    else if (tmin <= max_toi)
        t = tmin, n = near_n, side = near_side;
    else
        return h;
    h.hit = true, h.toi = t, h.n = rotate(g.m, n);
    h.feature = FEAT2_FACE | ((uint32_t)(side < 0 ? (-side - 1 + 3) : (side - 1)) & 0x3fffffffu);
    return h;
}

__device__ RayHit2 ray2_plane(const Operand2& g, W2 o_w, W2 d_w, float max_toi) {
    RayHit2 h = ray2_miss();
    W2 o = to_local(g.m, o_w), d = unrotate(g.m, d_w), pn = w2(g.a, g.b);
    float dot_normal_dpos = dot(pn, -o);
    if (dot_normal_dpos > 0.f) {
        h.hit = true, h.toi = 0.f, h.feature = FEAT2_FACE;
        return h;
    }
    float t = dot_normal_dpos / dot(pn, d);
    if (t >= 0.f && t <= max_toi) h.hit = true, h.toi = t, h.n = rotate(g.m, pn), h.feature = FEAT2_FACE;
    return h;
}

__device__ __forceinline__ bool ray2_plane_toi(W2 center, W2 normal, W2 origin, W2 dir, float& t_out) {  // ray_plane.rs:9-42
    float denom = dot(normal, dir);
    if (relative_eq(denom, 0.f)) return false;
    float t = dot(normal, center - origin) / denom;
    if (t >= 0.f) {
        t_out = t;
        return true;
    }
    return false;
}

// RayCast for ConvexPolygon: gjk::cast_ray on the Minkowski difference polygon - ConstantOrigin, in the polygon's frame
__device__ RayHit2 ray2_polygon(const Operand2& g, W2 o_w, W2 d_w, float max_toi) {
    RayHit2 h = ray2_miss();
    W2 ray_origin = to_local(g.m, o_w), ray_dir = unrotate(g.m, d_w);
    Operand2 loc = g, org;  // the polygon at the identity, and special_support_maps::ConstantOrigin
    loc.m.t = w2(0.f, 0.f), loc.m.re = 1.f, loc.m.im = 0.f;
    org = loc, org.kind = D2_ORIGIN;
    const float rel = sqrtf(TOL10);
    float ray_length = sqrtf(nsq(ray_dir));
    if (relative_eq(ray_length, 0.f)) return h;
    float ltoi = 0.f;
    W2 cur_o = ray_origin, cur_d = ray_dir / ray_length;
    W2 ldir = -cur_d, dir;
    Tri2 s;
    for (int i = 0; i < 3; ++i) s.v[i].p = s.v[i].o1 = s.v[i].o2 = w2(0.f, 0.f), s.old_idx[i] = i;
    s.bary[0] = s.bary[1] = s.old_bary[0] = s.old_bary[1] = 0.f;
    s.dim = s.old_dim = 0;
    s.v[0] = minkowski(loc, org, -cur_d);
    s.v[0].p = s.v[0].p + (-cur_o);
    W2 proj = tri_project(s);
    float upper = NCB_FMAX;
    bool last_chance = false;
    for (int it = 0;; ++it) {
        float prev_upper = upper, len;
        if (!unit_get(-proj, TOL10, dir, len)) break;  // Some((ltoi / ray_length, ldir))
        upper = len;
        MinkowskiPt sp;
        if (upper >= prev_upper) {
            last_chance = true;
            sp.p = sp.o1 = proj + cur_o, sp.o2 = w2(0.f, 0.f);
        } else {
            sp = minkowski(loc, org, dir);
        }
        if (last_chance && ltoi > 0.f) break;
        float t;
        if (ray2_plane_toi(sp.p, dir, cur_o, cur_d, t)) {
            if (dot(dir, cur_d) < 0.f && t > 0.f) {
                ldir = dir;
                ltoi += t;
                if (ltoi / ray_length > max_toi) return h;
                W2 shift = cur_d * t;
                cur_o = cur_o + shift;
                upper = NCB_FMAX;
                for (int i = 0; i <= s.dim; ++i) s.v[i].p = s.v[i].p + (-shift);
                last_chance = false;
            }
        } else if (dot(dir, cur_d) > TOL10) {
            return h;
        }
        if (last_chance) return h;
        float lower = -dot(dir, sp.p - cur_o);
        if (upper - lower <= rel * upper) return h;
        sp.p = sp.p + (-cur_o);
        (void)tri_add(s, sp);
        proj = tri_project(s);
        if (s.dim == 2) {
            if (lower >= TOL10) return h;
            break;
        }
        if (it + 1 == 10000) return h;
    }
    h.hit = true, h.toi = ltoi / ray_length, h.n = rotate(g.m, ldir), h.feature = FEAT2_UNKNOWN;
    return h;
}

// RayCast for Segment, dim2 (query/ray/ray_support_map.rs:219-293): line / line parameters (closest_points_line_line.rs:27-70) on the
// segment moved by its pose, the collinear cases; the SCALED normal; max_toi is not applied, as in the reference
__device__ RayHit2 ray2_segment(const Operand2& g, W2 o, W2 d) {
    RayHit2 h = ray2_miss();
    W2 a = to_world(g.m, w2(g.a, g.b)), b = to_world(g.m, w2(g.c, g.d));
    W2 sd = b - a, r = o - a;
    float aa = nsq(d), e = nsq(sd), f = dot(sd, r), s, t;
    bool parallel = false;
    if (aa <= NCB_EPS && e <= NCB_EPS) {
        s = 0.f, t = 0.f;
    } else if (aa <= NCB_EPS) {
        s = 0.f, t = f / e;
    } else {
        float c = dot(d, r);
        if (e <= NCB_EPS) {
            s = -c / aa, t = 0.f;
        } else {
            float bq = dot(d, sd), ae = aa * e, bb = bq * bq, denom = ae - bb;
            parallel = denom <= NCB_EPS || ulps_eq(ae, bb);
            s = !parallel ? (bq * f - c * e) / denom : 0.f;
            t = (bq * s + f) / e;
        }
    }
    W2 nrm = w2(sd.y, -sd.x);
    if (parallel) {
        W2 dpos = a - o;
        if (fabsf(dot(dpos, nrm)) < NCB_EPS) {
            float dist1 = dot(dpos, d), dist2 = dist1 + dot(sd, d);
            if (dist1 >= 0.f && dist2 >= 0.f) {
                h.hit = true, h.n = nrm;
                if (dist1 <= dist2)
                    h.toi = dist1 / nsq(d), h.feature = FEAT2_VERTEX | 0u;
                else
                    h.toi = dist2 / nsq(d), h.feature = FEAT2_VERTEX | 1u;
            } else if (dist1 >= 0.f || dist2 >= 0.f) {
                h.hit = true, h.toi = 0.f, h.n = nrm, h.feature = FEAT2_FACE | 0u;
            }
        }
    } else if (s >= 0.f && t >= 0.f && t <= 1.f) {
        h.hit = true, h.toi = s;
        if (dot(nrm, d) > 0.f)
            h.n = -nrm, h.feature = FEAT2_FACE | 1u;
        else
            h.n = nrm, h.feature = FEAT2_FACE | 0u;
    }
    return h;
}

// RayCast::toi_and_normal_with_ray(m, ray, max_toi, true) of one shape
__device__ RayHit2 shape_ray_cast2(const Operand2& g, W2 o, W2 d, float max_toi) {
    if (g.kind == D2_SEGMENT) return ray2_segment(g, o, d);
    if (g.kind == D2_BALL) return ray2_ball(g.m.t, g.a, o, d, max_toi);
    if (g.kind == D2_CUBOID) return ray2_cuboid(g, o, d, max_toi);
    if (g.kind == D2_POLYGON) return ray2_polygon(g, o, d, max_toi);
    return ray2_plane(g, o, d, max_toi);
}

// PointQuery::contains_point of one shape (point_ball.rs:45-47, point_cuboid.rs:34-38 -> point_aabb.rs:153-156, point_plane.rs:44-48;
// ConvexPolygon: the default, project_point(..).is_inside = gjk::project_origin finds the point inside, point_support_map.rs:14-55)
__device__ bool shape_contains_point2(const Operand2& g, W2 pt) {
    if (g.kind == D2_BALL) return nsq(to_local(g.m, pt)) <= g.a * g.a;
    if (g.kind == D2_CUBOID) {
        W2 l = to_local(g.m, pt);
        return !(l.x < -g.a || l.x > g.a || l.y < -g.b || l.y > g.b);
    }
    if (g.kind == D2_PLANE) return dot(w2(g.a, g.b), to_local(g.m, pt)) <= 0.f;
    if (g.kind == D2_SEGMENT) {  // the trait's default on Segment::project_point: relative_eq!(proj, pt) (point_segment.rs:52-91)
        W2 a = w2(g.a, g.b), ab = w2(g.c, g.d) - a, ap = to_local(g.m, pt) - a;
        float ab_ap = dot(ab, ap), sq = nsq(ab);
        W2 proj = ab_ap <= 0.f ? to_world(g.m, a) : ab_ap >= sq ? to_world(g.m, w2(g.c, g.d)) : to_world(g.m, a + ab * (ab_ap / sq));
        return relative_eq(proj.x, pt.x) && relative_eq(proj.y, pt.y);
    }
    Operand2 s = g, origin;
    s.m.t = (-pt) + g.m.t;  // Translation::from(-point) * m
    origin.kind = D2_ORIGIN, origin.a = origin.b = 0.f, origin.pts = origin.nrm = nullptr, origin.npts = 0;
    origin.m.t = w2(0.f, 0.f), origin.m.re = 1.f, origin.m.im = 0.f;
    W2 d0;
    if (!unit(-s.m.t, NCB_EPS, d0)) d0 = w2(1.f, 0.f);
    Tri2 sx;
    for (int i = 0; i < 3; ++i) sx.v[i].p = sx.v[i].o1 = sx.v[i].o2 = w2(0.f, 0.f), sx.old_idx[i] = i;
    sx.bary[0] = sx.bary[1] = sx.old_bary[0] = sx.old_bary[1] = 0.f;
    sx.dim = sx.old_dim = 0;
    sx.v[0] = minkowski(s, origin, d0);
    W2 p1, p2, n;
    return gjk2(s, origin, NCB_FMAX, sx, p1, p2, n) != G_POINTS;
}

// query::contact for one pair.  flags: bit 0 = the reference would panic, bit 1 = EPA capacity exceeded.
__device__ bool contact_of_pair(const Operand2& g1, const Operand2& g2, float prediction, float cos_one_degree, Hit2& h, int& flags) {
    h.w1 = h.w2 = h.n = w2(0.f, 0.f), h.depth = 0.f;
    bool ok = false;
    if (g1.kind == D2_BALL && g2.kind == D2_BALL) {
        ok = ball_ball(g1.m.t, g1.a, g2.m.t, g2.a, prediction, h);
    } else if (g1.kind == D2_PLANE || g2.kind == D2_PLANE) {  // contact_plane_support_map / contact_support_map_plane (plane x plane is refused on the host)
        const bool flip = g1.kind != D2_PLANE;
        ok = plane_support(flip ? g2 : g1, flip ? g1 : g2, prediction, h);
        if (ok && flip) {
            W2 t = h.w1;
            h.w1 = h.w2, h.w2 = t, h.n = -h.n;
        }
    } else if (g1.kind == D2_BALL || g2.kind == D2_BALL) {  // contact_ball_convex_polyhedron; with the ball second: the same query, flipped
        const bool flip = g1.kind != D2_BALL;
        const Operand2& ball = flip ? g2 : g1;
        const Operand2& other = flip ? g1 : g2;
        int q = 1;
        ok = other.kind == D2_CUBOID    ? ball_cuboid(ball.m.t, ball.a, other, prediction, h)
             : other.kind == D2_SEGMENT ? ball_segment(ball.m.t, ball.a, other, prediction, h)
                                        : ball_polygon(ball.m.t, ball.a, other, prediction, cos_one_degree, h, q);
        if (q == -1) flags |= 1;
        if (q == -2) flags |= 2;
        if (ok && flip) {
            W2 t = h.w1;
            h.w1 = h.w2, h.w2 = t, h.n = -h.n;
        }
    } else {
        W2 d0;
        if (!unit(g2.m.t - g1.m.t, NCB_EPS, d0)) d0 = w2(1.f, 0.f);
        Tri2 s;
        for (int i = 0; i < 3; ++i) s.v[i].p = s.v[i].o1 = s.v[i].o2 = w2(0.f, 0.f), s.old_idx[i] = i;
        s.bary[0] = s.bary[1] = s.old_bary[0] = s.old_bary[1] = 0.f;
        s.dim = s.old_dim = 0;
        s.v[0] = minkowski(g1, g2, d0);
        W2 p1, p2, n;
        int r = gjk2(g1, g2, prediction, s, p1, p2, n);
        ok = r == G_POINTS;
        if (r == G_INSIDE) {
            Poly2 e;
            int q = epa2(g1, g2, s, e, p1, p2, n);
            ok = q == 1;
            if (q == -1) flags |= 1;
            if (q == -2) flags |= 2;
        }
        if (ok) h.w1 = p1, h.w2 = p2, h.n = n, h.depth = -dot(n, p2 - p1);
    }
    return ok;
}

#ifndef NCB_HOST_SHIM  // kernels and the host entry points: CUDA only (tests/host_shim compiles the per-pair functions above for the host)
__global__ void __launch_bounds__(64) k_contact2d(Args2 A) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.n) return;
    Operand2 g1 = load_operand(__ldg(&A.type1[k]), __ldg(&A.param1[k]), __ldg(&A.pose1[k]), A.poly, A.poly_nrm);
    Operand2 g2 = load_operand(__ldg(&A.type2[k]), __ldg(&A.param2[k]), __ldg(&A.pose2[k]), A.poly, A.poly_nrm);
    Hit2 h;
    int flags = 0;
    bool ok = contact_of_pair(g1, g2, A.prediction, A.cos_one_degree, h, flags);
    if (flags & 1) atomicAdd(&A.counters[0], 1u);
    if (flags & 2) atomicAdd(&A.counters[1], 1u);
    A.found[k] = ok ? 1 : 0;
    float* o = A.out + 7 * (size_t)k;
    o[0] = h.w1.x, o[1] = h.w1.y, o[2] = h.w2.x, o[3] = h.w2.y, o[4] = h.n.x, o[5] = h.n.y, o[6] = h.depth;
}
__global__ void __launch_bounds__(128) k_proximity2d(Args2 A, const float* __restrict__ margins, uint8_t* __restrict__ status) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.n) return;
    Operand2 g1 = load_operand(__ldg(&A.type1[k]), __ldg(&A.param1[k]), __ldg(&A.pose1[k]), A.poly, A.poly_nrm);
    Operand2 g2 = load_operand(__ldg(&A.type2[k]), __ldg(&A.param2[k]), __ldg(&A.pose2[k]), A.poly, A.poly_nrm);
    status[k] = proximity_of_pair(g1, g2, __ldg(&margins[k]));
}

// one thread per (shape, ray): rays = origin x y, dir x y, max_toi
__global__ void __launch_bounds__(128) k_ray2d(Args2 A, const float* __restrict__ rays, float* __restrict__ out, uint32_t* __restrict__ feature) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.n) return;
    Operand2 g = load_operand(__ldg(&A.type1[k]), __ldg(&A.param1[k]), __ldg(&A.pose1[k]), A.poly, nullptr);
    const float* q = rays + 5 * (size_t)k;
    RayHit2 h = shape_ray_cast2(g, w2(__ldg(q), __ldg(q + 1)), w2(__ldg(q + 2), __ldg(q + 3)), __ldg(q + 4));
    A.found[k] = h.hit ? 1 : 0;
    out[3 * (size_t)k] = h.toi, out[3 * (size_t)k + 1] = h.n.x, out[3 * (size_t)k + 2] = h.n.y;
    feature[k] = h.hit ? h.feature : FEAT2_UNKNOWN;
}

struct World2Args {
    uint32_t n;
    const float2 *pos, *rot;
    const uint32_t* type;
    const float4* param;
    const float *qlimit, *cos_ang, *sin_ang;  // cos / sin of the angular prediction, from the host libm like the reference's
    const float *poly, *poly_nrm;
    float margin, cos_one_degree;
    float4 *aabb_lo, *aabb_hi;
    // narrow phase
    const uint2* pairs;
    const uint32_t* n_pairs_dev;
    uint32_t cap_pairs, cap_contacts;
    uint32_t* manifold_start;
    uint8_t* manifold_count;
    float* contacts;     // 7 floats per contact
    uint32_t* features;  // 2 words per contact
    uint32_t* counters;  // [0] contacts allocated, [1] reference panics, [2] EPA overflows, [3] manifold overflows
    const uint8_t* qkind;  // nullptr, or per object 1 = GeometricQueryType::Proximity (a sensor)
    uint8_t* prox;         // per pair: NCB_PROXIMITY_* for sensor pairs, NCB_PROXIMITY_NONE otherwise
};
__device__ __forceinline__ Operand2 world_operand(const World2Args& A, uint32_t i) {
    float2 t = __ldg(&A.pos[i]), r = __ldg(&A.rot[i]);
    return load_operand(__ldg(&A.type[i]), __ldg(&A.param[i]), make_float4(t.x, t.y, r.x, r.y), A.poly, A.poly_nrm);
}
__global__ void __launch_bounds__(256) k_aabb2d(World2Args A) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    Operand2 g = world_operand(A, i);
    W2 lo, hi;
    aabb_of_shape(g, lo, hi);
    float ql = __ldg(&A.qlimit[i]), mg = A.margin;
    // the 3-D broad phase reads float4 boxes; the w of the upper corner carries the shape type (ball 0, cuboid 1, polygon = the hull code 2)
    A.aabb_lo[i] = make_float4((lo.x + (-ql)) + (-mg), (lo.y + (-ql)) + (-mg), 0.f, 0.f);
    A.aabb_hi[i] = make_float4((hi.x + ql) + mg, (hi.y + ql) + mg, 0.f, __uint_as_float(g.kind));
}
__global__ void __launch_bounds__(64) k_narrow2d(World2Args A) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t np = min(*A.n_pairs_dev, A.cap_pairs);
    if (p >= np) return;
    uint2 pr = __ldg(&A.pairs[p]);
    Operand2 g1 = world_operand(A, pr.x), g2 = world_operand(A, pr.y);
    float linear = __ldg(&A.qlimit[pr.x]) + __ldg(&A.qlimit[pr.y]);
    Manifold2d mf;
    int flags = 0;
    const bool sensor = A.qkind && (__ldg(&A.qkind[pr.x]) | __ldg(&A.qkind[pr.y]));
    if (sensor) {  // Interaction::Proximity (narrow_phase.rs:138-167): the proximity dispatcher's detector, margin = the two limits added
        mf.n = 0, mf.overflow = false;
        A.prox[p] = (g1.kind == D2_PLANE && g2.kind == D2_PLANE) ? (uint8_t)NCB_PROXIMITY_NONE : proximity_of_pair(g1, g2, linear);
    } else {
        if (A.prox) A.prox[p] = (uint8_t)NCB_PROXIMITY_NONE;
        manifold_of_pair(g1, g2, linear, __ldg(&A.cos_ang[pr.x]), __ldg(&A.cos_ang[pr.y]), __ldg(&A.sin_ang[pr.x]), __ldg(&A.sin_ang[pr.y]),
                         A.cos_one_degree, mf, flags);
    }
    if (flags & 1) atomicAdd(&A.counters[1], 1u);
    if (flags & 2) atomicAdd(&A.counters[2], 1u);
    if (mf.overflow) atomicAdd(&A.counters[3], 1u);
    uint32_t start = mf.n ? atomicAdd(&A.counters[0], (uint32_t)mf.n) : 0u;
    A.manifold_start[p] = start;
    A.manifold_count[p] = (uint8_t)mf.n;
    for (int k = 0; k < mf.n; ++k) {
        uint32_t dst = start + k;
        if (dst >= A.cap_contacts) break;
        float* o = A.contacts + 7 * (size_t)dst;
        const Hit2& c = mf.c[k];
        o[0] = c.w1.x, o[1] = c.w1.y, o[2] = c.w2.x, o[3] = c.w2.y, o[4] = c.n.x, o[5] = c.n.y, o[6] = c.depth;
        A.features[2 * (size_t)dst] = mf.f1[k], A.features[2 * (size_t)dst + 1] = mf.f2[k];
    }
}

// ---- world ray queries: glue::interferences_with_ray / first_interference_with_ray (pipeline/glue/query.rs:13-77,183-224) -------------
// against the world of the last ncb2d_world_update: candidates = objects whose stored (fat) box the ray enters within max_toi
// (AABB::toi_with_ray, DIM = 2), the query's collision groups, then the shape's RayCast.  One thread per ray over the update's LBVH.
struct WorldRay2Args {
    World2Args W;
    const float4 *leaf_lo, *leaf_hi, *nodes;
    uint32_t n_tree;         // leaves in the tree; [n_tree, W.n) are the outliers (planes), tested one by one
    const uint32_t* groups;  // 3 words per object, or nullptr (default groups)
    uint32_t qg[3];
    int use_groups;
    const float* rays;       // origin x y, dir x y, max_toi
    uint32_t n_rays;
    unsigned long long* keys;  // ray << 32 | handle
    float4* vals;              // toi, normal x y, -
    uint32_t* feats;
    uint32_t cap;
    uint32_t* counter;
    uint32_t* trav_overflow;
};
template <bool FIRST>
__device__ __forceinline__ void world_ray_leaf(const WorldRay2Args& A, uint32_t ri, W2 o, W2 d, float max_toi, uint32_t handle, RayHit2& best,
                                               uint32_t& best_h) {
    if (A.use_groups) {  // CollisionGroups::can_interact_with_groups (collision_groups.rs:353-359); objects without groups have the defaults
        uint32_t m1 = 0x3fffffffu, w1 = 0x3fffffffu, b1 = 0u;
        if (A.groups) m1 = __ldg(&A.groups[3 * handle]), w1 = __ldg(&A.groups[3 * handle + 1]), b1 = __ldg(&A.groups[3 * handle + 2]);
        if (!((m1 & A.qg[2]) == 0 && (A.qg[0] & b1) == 0 && (m1 & A.qg[1]) != 0 && (A.qg[0] & w1) != 0)) return;
    }
    RayHit2 h = shape_ray_cast2(world_operand(A.W, handle), o, d, max_toi);
    if (!h.hit) return;
    if (FIRST) {
        if (!best.hit || h.toi < best.toi || (h.toi == best.toi && handle < best_h)) best = h, best_h = handle;
    } else {
        uint32_t k = atomicAdd(A.counter, 1u);
        if (k < A.cap) {
            A.keys[k] = ((unsigned long long)ri << 32) | handle;
            A.vals[k] = make_float4(h.toi, h.n.x, h.n.y, 0.f);
            A.feats[k] = h.feature;
        }
    }
}
__device__ __forceinline__ bool slab2(float4 lo, float4 hi, W2 o, W2 d, W2 inv, float max_toi) {  // ray_aabb.rs:13-50, DIM = 2
    float tmin = 0.f, tmax = max_toi;
    if (d.x == 0.f) {
        if (o.x < lo.x || o.x > hi.x) return false;
    } else {
        float a = (lo.x - o.x) * inv.x, b = (hi.x - o.x) * inv.x;
        tmin = fmaxf(tmin, fminf(a, b)), tmax = fminf(tmax, fmaxf(a, b));
        if (tmin > tmax) return false;
    }
    if (d.y == 0.f) {
        if (o.y < lo.y || o.y > hi.y) return false;
    } else {
        float a = (lo.y - o.y) * inv.y, b = (hi.y - o.y) * inv.y;
        tmin = fmaxf(tmin, fminf(a, b)), tmax = fminf(tmax, fmaxf(a, b));
        if (tmin > tmax) return false;
    }
    return true;
}
template <bool FIRST>
__global__ void __launch_bounds__(128) k_world_ray2d(WorldRay2Args A) {
    uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= A.n_rays) return;
    const float* q = A.rays + 5 * (size_t)ri;
    const W2 o = w2(__ldg(q), __ldg(q + 1)), d = w2(__ldg(q + 2), __ldg(q + 3)), inv = w2(1.f / d.x, 1.f / d.y);
    const float max_toi = __ldg(q + 4);
    RayHit2 best = ray2_miss();
    uint32_t best_h = 0;
    const uint32_t m = A.n_tree;
    if (m >= 2) {
        uint32_t stack[64];
        int sp = 0;
        uint32_t node = 0;
        for (;;) {
            const float4* rec = A.nodes + 4 * (size_t)node;
            float4 Llo = __ldg(rec + 0), Lhi = __ldg(rec + 1), Rlo = __ldg(rec + 2), Rhi = __ldg(rec + 3);
            uint32_t left = __float_as_uint(Llo.w), right = __float_as_uint(Lhi.w);
            bool goL = slab2(Llo, Lhi, o, d, inv, max_toi), goR = slab2(Rlo, Rhi, o, d, inv, max_toi);
            if (goL && (left & 0x80000000u)) {
                world_ray_leaf<FIRST>(A, ri, o, d, max_toi, __float_as_uint(__ldg(&A.leaf_lo[left & 0x7fffffffu].w)), best, best_h);
                goL = false;
            }
            if (goR && (right & 0x80000000u)) {
                world_ray_leaf<FIRST>(A, ri, o, d, max_toi, __float_as_uint(__ldg(&A.leaf_lo[right & 0x7fffffffu].w)), best, best_h);
                goR = false;
            }
            if (goL && goR) {
                if (sp < 64)
                    stack[sp++] = right;
                else
                    atomicAdd(A.trav_overflow, 1u);
                node = left;
            } else if (goL) {
                node = left;
            } else if (goR) {
                node = right;
            } else {
                if (sp == 0) break;
                node = stack[--sp];
            }
        }
    } else if (m == 1) {
        float4 lo = __ldg(&A.leaf_lo[0]), hi = __ldg(&A.leaf_hi[0]);
        if (slab2(lo, hi, o, d, inv, max_toi)) world_ray_leaf<FIRST>(A, ri, o, d, max_toi, __float_as_uint(lo.w), best, best_h);
    }
    for (uint32_t k = m; k < A.W.n; ++k) {
        float4 lo = __ldg(&A.leaf_lo[k]), hi = __ldg(&A.leaf_hi[k]);
        if (slab2(lo, hi, o, d, inv, max_toi)) world_ray_leaf<FIRST>(A, ri, o, d, max_toi, __float_as_uint(lo.w), best, best_h);
    }
    if (FIRST && best.hit) {
        uint32_t k = atomicAdd(A.counter, 1u);
        if (k < A.cap) {
            A.keys[k] = ((unsigned long long)ri << 32) | best_h;
            A.vals[k] = make_float4(best.toi, best.n.x, best.n.y, 0.f);
            A.feats[k] = best.feature;
        }
    }
}
// glue::interferences_with_aabb (KIND 0: mins x y, maxs x y) / interferences_with_point (KIND 2: x y) (pipeline/glue/query.rs:79-181)
template <int KIND>
__device__ __forceinline__ bool box_test2(const float* q, float4 lo, float4 hi) {
    if (KIND == 0) return lo.x <= q[2] && lo.y <= q[3] && hi.x >= q[0] && hi.y >= q[1];  // AABB::intersects
    return !(q[0] < lo.x || q[0] > hi.x || q[1] < lo.y || q[1] > hi.y);                   // AABB::contains_local_point
}
template <int KIND>
__device__ __forceinline__ void world_query_leaf(const WorldRay2Args& A, uint32_t qi, const float* q, uint32_t handle) {
    if (A.use_groups) {
        uint32_t m1 = 0x3fffffffu, w1 = 0x3fffffffu, b1 = 0u;
        if (A.groups) m1 = __ldg(&A.groups[3 * handle]), w1 = __ldg(&A.groups[3 * handle + 1]), b1 = __ldg(&A.groups[3 * handle + 2]);
        if (!((m1 & A.qg[2]) == 0 && (A.qg[0] & b1) == 0 && (m1 & A.qg[1]) != 0 && (A.qg[0] & w1) != 0)) return;
    }
    if (KIND == 2 && !shape_contains_point2(world_operand(A.W, handle), w2(q[0], q[1]))) return;
    uint32_t k = atomicAdd(A.counter, 1u);
    if (k < A.cap) A.keys[k] = ((unsigned long long)qi << 32) | handle;
}
template <int KIND>
__global__ void __launch_bounds__(128) k_world_query2d(WorldRay2Args A) {
    const int W = KIND == 0 ? 4 : 2;
    uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= A.n_rays) return;
    float q[4];
    for (int k = 0; k < W; ++k) q[k] = __ldg(A.rays + (size_t)W * qi + k);
    const uint32_t m = A.n_tree;
    if (m >= 2) {
        uint32_t stack[64];
        int sp = 0;
        uint32_t node = 0;
        for (;;) {
            const float4* rec = A.nodes + 4 * (size_t)node;
            float4 Llo = __ldg(rec + 0), Lhi = __ldg(rec + 1), Rlo = __ldg(rec + 2), Rhi = __ldg(rec + 3);
            uint32_t left = __float_as_uint(Llo.w), right = __float_as_uint(Lhi.w);
            bool goL = box_test2<KIND>(q, Llo, Lhi), goR = box_test2<KIND>(q, Rlo, Rhi);
            if (goL && (left & 0x80000000u)) {
                world_query_leaf<KIND>(A, qi, q, __float_as_uint(__ldg(&A.leaf_lo[left & 0x7fffffffu].w)));
                goL = false;
            }
            if (goR && (right & 0x80000000u)) {
                world_query_leaf<KIND>(A, qi, q, __float_as_uint(__ldg(&A.leaf_lo[right & 0x7fffffffu].w)));
                goR = false;
            }
            if (goL && goR) {
                if (sp < 64)
                    stack[sp++] = right;
                else
                    atomicAdd(A.trav_overflow, 1u);
                node = left;
            } else if (goL) {
                node = left;
            } else if (goR) {
                node = right;
            } else {
                if (sp == 0) break;
                node = stack[--sp];
            }
        }
    } else if (m == 1) {
        float4 lo = __ldg(&A.leaf_lo[0]), hi = __ldg(&A.leaf_hi[0]);
        if (box_test2<KIND>(q, lo, hi)) world_query_leaf<KIND>(A, qi, q, __float_as_uint(lo.w));
    }
    for (uint32_t k = m; k < A.W.n; ++k) {
        float4 lo = __ldg(&A.leaf_lo[k]), hi = __ldg(&A.leaf_hi[k]);
        if (box_test2<KIND>(q, lo, hi)) world_query_leaf<KIND>(A, qi, q, __float_as_uint(lo.w));
    }
}
__global__ void k_iota2d(uint32_t* p, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
__global__ void k_gather_rows2d(const uint32_t* __restrict__ order, const float4* __restrict__ vals, const uint32_t* __restrict__ feats, uint32_t n,
                                float4* __restrict__ vals_out, uint32_t* __restrict__ feats_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vals_out[i] = vals[order[i]], feats_out[i] = feats[order[i]];
}

#endif  // NCB_HOST_SHIM (kernels)

}  // namespace d2
}  // namespace ncb

using namespace ncb;

#ifndef NCB_HOST_SHIM
#define CK2(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            char b__[512];                                                                                \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            ctx->err = b__;                                                                               \
            return NCB_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

extern "C" {

int ncb2d_contact(ncb_ctx* ctx, uint32_t n_pairs, const uint32_t* type1, const float* param1, const float* pose1, const uint32_t* type2,
                  const float* param2, const float* pose2, const float* poly_points, const float* poly_normals, uint32_t n_poly_points,
                  float prediction, uint8_t* found, float* out, uint32_t* ref_panics, uint32_t* epa_overflow) {
    if (!ctx || (n_pairs && (!type1 || !param1 || !pose1 || !type2 || !param2 || !pose2 || !found || !out))) return NCB_ERR_ARG;
    if (ref_panics) *ref_panics = 0;
    if (epa_overflow) *epa_overflow = 0;
    if (n_pairs == 0) return NCB_OK;
    // input validation before device state is touched
    for (uint32_t k = 0; k < n_pairs; ++k) {
        for (int side = 0; side < 2; ++side) {
            uint32_t t = side ? type2[k] : type1[k];
            const float* p = (side ? param2 : param1) + 4 * (size_t)k;
            if (t > 4) {
                ctx->err = "ncb2d_contact: unknown 2-D shape type";
                return NCB_ERR_UNSUPPORTED;
            }
            if (t == 4 && p[0] == p[2] && p[1] == p[3]) {
                ctx->err = "ncb2d_contact: a segment needs two different end points";
                return NCB_ERR_ARG;
            }
            if (t == 2) {
                if (!poly_points || p[1] < 1.f || p[0] < 0.f || (uint64_t)p[0] + (uint64_t)p[1] > n_poly_points) {
                    ctx->err = "ncb2d_contact: polygon point range outside poly_points";
                    return NCB_ERR_ARG;
                }
            }
        }
        if (type1[k] == 3 && type2[k] == 3) {
            ctx->err = "ncb2d_contact: no algorithm for plane x plane (the reference panics)";
            return NCB_ERR_UNSUPPORTED;
        }
        bool b1 = type1[k] == 0, b2 = type2[k] == 0;
        if (((b1 && type2[k] == 2) || (b2 && type1[k] == 2)) && !poly_normals) {
            ctx->err = "ncb2d_contact: a ball x convex polygon pair needs poly_normals (ConvexPolygon::normals)";
            return NCB_ERR_ARG;
        }
    }
    CK2(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    size_t n = n_pairs;
    DevBuf<uint32_t> d_t;   // type1 | type2 | counters
    DevBuf<float4> d_f4;    // param1 | param2 | pose1 | pose2
    DevBuf<float> d_poly, d_nrm, d_out;
    DevBuf<uint8_t> d_found;
    CK2(d_t.reserve(2 * n + 2));
    CK2(d_f4.reserve(4 * n));
    CK2(d_poly.reserve(2 * (size_t)(n_poly_points ? n_poly_points : 1)));
    CK2(d_out.reserve(7 * n));
    CK2(d_found.reserve(n));
    CK2(cudaMemcpyAsync(d_t.p, type1, 4 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_t.p + n, type2, 4 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemsetAsync(d_t.p + 2 * n, 0, 8, s));
    CK2(cudaMemcpyAsync(d_f4.p, param1, 16 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_f4.p + n, param2, 16 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_f4.p + 2 * n, pose1, 16 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_f4.p + 3 * n, pose2, 16 * n, cudaMemcpyHostToDevice, s));
    if (n_poly_points) CK2(cudaMemcpyAsync(d_poly.p, poly_points, 8 * (size_t)n_poly_points, cudaMemcpyHostToDevice, s));
    if (n_poly_points && poly_normals) {
        CK2(d_nrm.reserve(2 * (size_t)n_poly_points));
        CK2(cudaMemcpyAsync(d_nrm.p, poly_normals, 8 * (size_t)n_poly_points, cudaMemcpyHostToDevice, s));
    }
    d2::Args2 A;
    A.n = n_pairs;
    A.type1 = d_t.p, A.type2 = d_t.p + n;
    A.param1 = d_f4.p, A.param2 = d_f4.p + n, A.pose1 = d_f4.p + 2 * n, A.pose2 = d_f4.p + 3 * n;
    A.poly = d_poly.p;
    A.poly_nrm = (n_poly_points && poly_normals) ? d_nrm.p : nullptr;
    A.prediction = prediction;
    A.cos_one_degree = cosf((float)(3.14159265358979323846 / 180.0));
    A.found = d_found.p, A.out = d_out.p;
    A.counters = d_t.p + 2 * n;
    d2::k_contact2d<<<(n_pairs + 63) / 64, 64, 0, s>>>(A);
    CK2(cudaGetLastError());
    uint32_t cnt[2] = {0, 0};
    CK2(cudaMemcpyAsync(found, d_found.p, n, cudaMemcpyDeviceToHost, s));
    CK2(cudaMemcpyAsync(out, d_out.p, 28 * n, cudaMemcpyDeviceToHost, s));
    CK2(cudaMemcpyAsync(cnt, d_t.p + 2 * n, 8, cudaMemcpyDeviceToHost, s));
    CK2(cudaStreamSynchronize(s));
    if (ref_panics) *ref_panics = cnt[0];
    if (epa_overflow) *epa_overflow = cnt[1];
    return NCB_OK;
}

// ncollide2d::query::proximity(m1, g1, m2, g2, margin) for a batch (one margin per pair); out: NCB_PROXIMITY_* per pair.
int ncb2d_proximity(ncb_ctx* ctx, uint32_t n_pairs, const uint32_t* type1, const float* param1, const float* pose1, const uint32_t* type2,
                    const float* param2, const float* pose2, const float* poly_points, uint32_t n_poly_points, const float* margins,
                    uint8_t* out) {
    if (!ctx || (n_pairs && (!type1 || !param1 || !pose1 || !type2 || !param2 || !pose2 || !margins || !out))) return NCB_ERR_ARG;
    if (n_pairs == 0) return NCB_OK;
    for (uint32_t k = 0; k < n_pairs; ++k) {
        for (int side = 0; side < 2; ++side) {
            uint32_t t = side ? type2[k] : type1[k];
            const float* p = (side ? param2 : param1) + 4 * (size_t)k;
            if (t > 4) {
                ctx->err = "ncb2d_proximity: unknown 2-D shape type";
                return NCB_ERR_UNSUPPORTED;
            }
            if (t == 4 && p[0] == p[2] && p[1] == p[3]) {
                ctx->err = "ncb2d_proximity: a segment needs two different end points";
                return NCB_ERR_ARG;
            }
            if (t == 2 && (!poly_points || p[1] < 1.f || p[0] < 0.f || (uint64_t)p[0] + (uint64_t)p[1] > n_poly_points)) {
                ctx->err = "ncb2d_proximity: polygon point range outside poly_points";
                return NCB_ERR_ARG;
            }
        }
        if (type1[k] == 3 && type2[k] == 3) {
            ctx->err = "ncb2d_proximity: no algorithm for plane x plane (the reference panics)";
            return NCB_ERR_UNSUPPORTED;
        }
        if (!(margins[k] >= 0.f)) {
            ctx->err = "ncb2d_proximity: the proximity margin must be positive or zero";
            return NCB_ERR_ARG;
        }
    }
    CK2(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    size_t n = n_pairs;
    DevBuf<uint32_t> d_t;
    DevBuf<float4> d_f4;
    DevBuf<float> d_poly, d_mg;
    DevBuf<uint8_t> d_out;
    CK2(d_t.reserve(2 * n));
    CK2(d_f4.reserve(4 * n));
    CK2(d_poly.reserve(2 * (size_t)(n_poly_points ? n_poly_points : 1)));
    CK2(d_mg.reserve(n));
    CK2(d_out.reserve(n));
    CK2(cudaMemcpyAsync(d_t.p, type1, 4 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_t.p + n, type2, 4 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_f4.p, param1, 16 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_f4.p + n, param2, 16 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_f4.p + 2 * n, pose1, 16 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_f4.p + 3 * n, pose2, 16 * n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_mg.p, margins, 4 * n, cudaMemcpyHostToDevice, s));
    if (n_poly_points) CK2(cudaMemcpyAsync(d_poly.p, poly_points, 8 * (size_t)n_poly_points, cudaMemcpyHostToDevice, s));
    d2::Args2 A;
    memset(&A, 0, sizeof A);
    A.n = n_pairs;
    A.type1 = d_t.p, A.type2 = d_t.p + n;
    A.param1 = d_f4.p, A.param2 = d_f4.p + n, A.pose1 = d_f4.p + 2 * n, A.pose2 = d_f4.p + 3 * n;
    A.poly = d_poly.p;
    d2::k_proximity2d<<<(n_pairs + 127) / 128, 128, 0, s>>>(A, d_mg.p, d_out.p);
    CK2(cudaGetLastError());
    CK2(cudaMemcpyAsync(out, d_out.p, n, cudaMemcpyDeviceToHost, s));
    CK2(cudaStreamSynchronize(s));
    return NCB_OK;
}

// RayCast::toi_and_normal_with_ray(m, ray, max_toi, solid = true) of shape k for ray k, for a batch.
int ncb2d_ray_cast(ncb_ctx* ctx, uint32_t n, const uint32_t* type, const float* param, const float* pose, const float* poly_points,
                   uint32_t n_poly_points, const float* rays, uint8_t* found, float* out, uint32_t* feature) {
    if (!ctx || (n && (!type || !param || !pose || !rays || !found || !out || !feature))) return NCB_ERR_ARG;
    if (n == 0) return NCB_OK;
    for (uint32_t k = 0; k < n; ++k) {
        const float* p = param + 4 * (size_t)k;
        if (type[k] > 4) {
            ctx->err = "ncb2d_ray_cast: unknown 2-D shape type";
            return NCB_ERR_UNSUPPORTED;
        }
        if (type[k] == 4 && p[0] == p[2] && p[1] == p[3]) {
            ctx->err = "ncb2d_ray_cast: a segment needs two different end points";
            return NCB_ERR_ARG;
        }
        if (type[k] == 2 && (!poly_points || p[1] < 1.f || p[0] < 0.f || (uint64_t)p[0] + (uint64_t)p[1] > n_poly_points)) {
            ctx->err = "ncb2d_ray_cast: polygon point range outside poly_points";
            return NCB_ERR_ARG;
        }
    }
    CK2(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    size_t m = n;
    DevBuf<uint32_t> d_t;
    DevBuf<float4> d_f4;
    DevBuf<float> d_poly, d_rays, d_out;
    DevBuf<uint8_t> d_found;
    CK2(d_t.reserve(2 * m));
    CK2(d_f4.reserve(2 * m));
    CK2(d_poly.reserve(2 * (size_t)(n_poly_points ? n_poly_points : 1)));
    CK2(d_rays.reserve(5 * m));
    CK2(d_out.reserve(3 * m));
    CK2(d_found.reserve(m));
    CK2(cudaMemcpyAsync(d_t.p, type, 4 * m, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_f4.p, param, 16 * m, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_f4.p + m, pose, 16 * m, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_rays.p, rays, 20 * m, cudaMemcpyHostToDevice, s));
    if (n_poly_points) CK2(cudaMemcpyAsync(d_poly.p, poly_points, 8 * (size_t)n_poly_points, cudaMemcpyHostToDevice, s));
    d2::Args2 A;
    memset(&A, 0, sizeof A);
    A.n = n;
    A.type1 = d_t.p, A.param1 = d_f4.p, A.pose1 = d_f4.p + m;
    A.poly = d_poly.p;
    A.found = d_found.p;
    d2::k_ray2d<<<(n + 127) / 128, 128, 0, s>>>(A, d_rays.p, d_out.p, d_t.p + m);
    CK2(cudaGetLastError());
    CK2(cudaMemcpyAsync(found, d_found.p, m, cudaMemcpyDeviceToHost, s));
    CK2(cudaMemcpyAsync(out, d_out.p, 12 * m, cudaMemcpyDeviceToHost, s));
    CK2(cudaMemcpyAsync(feature, d_t.p + m, 4 * m, cudaMemcpyDeviceToHost, s));
    CK2(cudaStreamSynchronize(s));
    return NCB_OK;
}

// CollisionWorld::update of ncollide2d for a fresh world of balls, cuboids and convex polygons (pipeline/world.rs:104-119 with the 2-D
// generators above).  Host buffers in and out.  The broad phase is the 3-D path's LBVH on boxes with z = 0: it uses the context's
// broad-phase scratch, so do not interleave it with a stepping world (ncb_sim_*) of the same context.
int ncb2d_world_update(ncb_ctx* ctx, const ncb2d_objects* o, float margin, uint32_t* pairs, uint32_t cap_pairs, uint32_t* manifold_start,
                       uint8_t* manifold_count, float* contacts, uint32_t* features, uint32_t cap_contacts, uint32_t* n_pairs,
                       uint32_t* n_contacts, uint32_t* diag) {
    if (!ctx || !o || !n_pairs || !n_contacts) return NCB_ERR_ARG;
    *n_pairs = *n_contacts = 0;
    if (diag) diag[0] = diag[1] = diag[2] = diag[3] = 0;
    uint32_t n = o->n;
    if (n == 0) return NCB_OK;
    if (!o->pos || !o->rot || !o->shape_type || !o->shape_param || !o->query_limit || !o->ang_pred) return NCB_ERR_ARG;
    bool any_poly = false;
    for (uint32_t i = 0; i < n; ++i) {  // validation before device state is touched
        uint32_t t = o->shape_type[i];
        if (t > 4) {
            ctx->err = "ncb2d_world_update: unknown 2-D shape type";
            return NCB_ERR_UNSUPPORTED;
        }
        if (t == 4 && o->shape_param[4 * (size_t)i] == o->shape_param[4 * (size_t)i + 2] && o->shape_param[4 * (size_t)i + 1] == o->shape_param[4 * (size_t)i + 3]) {
            ctx->err = "ncb2d_world_update: a segment needs two different end points";
            return NCB_ERR_ARG;
        }
        if (o->query_kind && o->query_kind[i] > 1) {
            ctx->err = "ncb2d_world_update: query_kind must be 0 (Contacts) or 1 (Proximity)";
            return NCB_ERR_ARG;
        }
        if (t == 2) {
            const float* p = o->shape_param + 4 * (size_t)i;
            any_poly = true;
            if (!o->poly_points || !o->poly_normals || p[1] < 1.f || p[0] < 0.f || (uint64_t)p[0] + (uint64_t)p[1] > o->n_poly_points) {
                ctx->err = "ncb2d_world_update: polygon point range outside poly_points / normals missing";
                return NCB_ERR_ARG;
            }
        }
    }
    CK2(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    ctx->d2.last_n = 0, ctx->d2.last_pairs = 0;  // set again when this update has succeeded
    // NCB2D_PROFILE=1: wall time per phase (with a stream synchronisation at every boundary) on stderr
    static const bool prof = getenv("NCB2D_PROFILE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!prof) return;
        cudaStreamSynchronize(s);
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[2d] %-12s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    // the device buffers live in the context (no cudaMalloc / cudaFree per update once they have grown to the world's size)
    DevBuf<float2>&d_pos = ctx->d2.pos, &d_rot = ctx->d2.rot;
    DevBuf<uint32_t>&d_type = ctx->d2.type, &d_groups = ctx->d2.groups, &d_start = ctx->d2.start, &d_feat = ctx->d2.feat, &d_cnt = ctx->d2.cnt;
    DevBuf<float4>& d_param = ctx->d2.param;
    DevBuf<float>&d_ql = ctx->d2.ql, &d_cang = ctx->d2.cang, &d_poly = ctx->d2.poly, &d_nrm = ctx->d2.nrm, &d_contacts = ctx->d2.contacts;
    DevBuf<uint8_t>& d_count = ctx->d2.count;
    CK2(d_pos.reserve(n));
    CK2(d_rot.reserve(n));
    CK2(d_type.reserve(n));
    CK2(d_param.reserve(n));
    CK2(d_ql.reserve(n));
    CK2(d_cang.reserve(n));
    CK2(ctx->d2.sang.reserve(n));
    CK2(d_cnt.reserve(4));
    std::vector<float> cang(n);
    std::vector<float> sang(n);
    for (uint32_t i = 0; i < n; ++i) cang[i] = cosf(o->ang_pred[i]), sang[i] = sinf(o->ang_pred[i]);  // ContactPrediction: angle.cos() / .sin() with the host libm
    CK2(cudaMemcpyAsync(d_pos.p, o->pos, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_rot.p, o->rot, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_type.p, o->shape_type, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_param.p, o->shape_param, 16 * (size_t)n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_ql.p, o->query_limit, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(d_cang.p, cang.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    CK2(cudaMemcpyAsync(ctx->d2.sang.p, sang.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    if (o->groups) {
        CK2(d_groups.reserve(3 * (size_t)n));
        CK2(cudaMemcpyAsync(d_groups.p, o->groups, 12 * (size_t)n, cudaMemcpyHostToDevice, s));
    }
    if (any_poly) {
        CK2(d_poly.reserve(2 * (size_t)o->n_poly_points));
        CK2(d_nrm.reserve(2 * (size_t)o->n_poly_points));
        CK2(cudaMemcpyAsync(d_poly.p, o->poly_points, 8 * (size_t)o->n_poly_points, cudaMemcpyHostToDevice, s));
        CK2(cudaMemcpyAsync(d_nrm.p, o->poly_normals, 8 * (size_t)o->n_poly_points, cudaMemcpyHostToDevice, s));
    }
    if (o->query_kind) {
        CK2(ctx->d2.qkind.reserve(n));
        CK2(cudaMemcpyAsync(ctx->d2.qkind.p, o->query_kind, n, cudaMemcpyHostToDevice, s));
    }
    int r = reserve_broad(ctx, n);
    if (r) return r;
    mark("alloc+h2d");
    d2::World2Args A;
    memset(&A, 0, sizeof A);
    A.n = n;
    A.pos = d_pos.p, A.rot = d_rot.p, A.type = d_type.p, A.param = d_param.p, A.qlimit = d_ql.p, A.cos_ang = d_cang.p, A.sin_ang = ctx->d2.sang.p;
    A.poly = d_poly.p, A.poly_nrm = d_nrm.p;
    A.margin = margin;
    A.cos_one_degree = cosf((float)(3.14159265358979323846 / 180.0));
    A.aabb_lo = ctx->aabb_lo.p, A.aabb_hi = ctx->aabb_hi.p;
    size_t capp = cap_pairs ? cap_pairs : 1;
    r = reserve_pairs(ctx, capp);
    if (r) return r;
    r = reset_counters(ctx);
    if (r) return r;
    d2::k_aabb2d<<<(n + 255) / 256, 256, 0, s>>>(A);
    CK2(cudaGetLastError());
    CK2(launch_lbvh_build(ctx, n, nullptr));
    CK2(launch_pair_search(ctx, n, o->groups ? d_groups.p : nullptr, 0, 0xffffffffu, (uint32_t)capp, -1));
    // the 3-D path's counting sort by pair kind (the 2-D type codes are the 3-D ones): warps of the narrow kernel then hold one generator
    CK2(launch_pair_sort(ctx, (uint32_t)capp, nullptr));
    mark("broad");
    // narrow phase over the sorted pairs (count read on the device)
    size_t capc = cap_contacts ? cap_contacts : 1;
    CK2(d_start.reserve(capp));
    CK2(d_count.reserve(capp));
    CK2(d_contacts.reserve(7 * capc));
    CK2(d_feat.reserve(2 * capc));
    CK2(cudaMemsetAsync(d_cnt.p, 0, 16, s));
    A.pairs = ctx->pairs.p;
    A.n_pairs_dev = &ctx->counters.p->n_pairs;
    A.cap_pairs = (uint32_t)capp, A.cap_contacts = (uint32_t)capc;
    A.manifold_start = d_start.p, A.manifold_count = d_count.p, A.contacts = d_contacts.p, A.features = d_feat.p;
    A.counters = d_cnt.p;
    CK2(ctx->d2.prox.reserve(capp));
    A.qkind = o->query_kind ? ctx->d2.qkind.p : nullptr, A.prox = ctx->d2.prox.p;
    d2::k_narrow2d<<<(uint32_t)((capp + 63) / 64), 64, 0, s>>>(A);
    CK2(cudaGetLastError());
    mark("narrow");
    r = read_counters(ctx);
    if (r) return r;
    uint32_t cnt[4];
    CK2(cudaMemcpyAsync(cnt, d_cnt.p, 16, cudaMemcpyDeviceToHost, s));
    CK2(cudaStreamSynchronize(s));
    uint32_t np = ctx->last_counters.n_pairs, nc = cnt[0];
    *n_pairs = np, *n_contacts = nc;
    ctx->d2.last_pairs = np < capp ? np : (uint32_t)capp;
    ctx->d2.last_n = n, ctx->d2.last_groups = o->groups != nullptr;
    if (diag) diag[0] = cnt[1], diag[1] = cnt[2], diag[2] = cnt[3], diag[3] = ctx->last_counters.stack_overflow;
    uint32_t wp = np < cap_pairs ? np : cap_pairs, wc = nc < cap_contacts ? nc : cap_contacts;
    if (pairs && wp) CK2(cudaMemcpyAsync(pairs, ctx->pairs.p, 8 * (size_t)wp, cudaMemcpyDeviceToHost, s));
    if (manifold_start && wp) CK2(cudaMemcpyAsync(manifold_start, d_start.p, 4 * (size_t)wp, cudaMemcpyDeviceToHost, s));
    if (manifold_count && wp) CK2(cudaMemcpyAsync(manifold_count, d_count.p, wp, cudaMemcpyDeviceToHost, s));
    if (contacts && wc) CK2(cudaMemcpyAsync(contacts, d_contacts.p, 28 * (size_t)wc, cudaMemcpyDeviceToHost, s));
    if (features && wc) CK2(cudaMemcpyAsync(features, d_feat.p, 8 * (size_t)wc, cudaMemcpyDeviceToHost, s));
    CK2(cudaStreamSynchronize(s));
    mark("d2h");
    return (np > cap_pairs || nc > cap_contacts) ? 1 : NCB_OK;
}

// glue::interferences_with_ray (first_only = 0) / first_interference_with_ray (first_only = 1) against the world of the last
// ncb2d_world_update.  Rows sorted by (ray, handle); returns 1 when they were truncated at cap.
int ncb2d_world_ray_cast(ncb_ctx* ctx, uint32_t n_rays, const float* rays, const uint32_t* groups, int first_only, uint32_t* idx, float* val,
                         uint32_t* feat, uint32_t cap, uint32_t* n_out) {
    if (!ctx || !n_out || (n_rays && !rays) || (cap && (!idx || !val || !feat))) return NCB_ERR_ARG;
    *n_out = 0;
    uint32_t n = ctx->d2.last_n;
    if (n == 0) {
        ctx->err = "ncb2d_world_ray_cast: no 2-D world on the device (call ncb2d_world_update first)";
        return NCB_ERR_ARG;
    }
    if (n_rays == 0) return NCB_OK;
    CK2(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    auto& D = ctx->d2;
    size_t capd = cap ? cap : 1;
    CK2(D.q_rays.reserve(5 * (size_t)n_rays));
    CK2(D.q_keys.reserve(capd));
    CK2(D.q_keys2.reserve(capd));
    CK2(D.q_vals.reserve(capd));
    CK2(D.q_vals2.reserve(capd));
    CK2(D.q_feat.reserve(capd));
    CK2(D.q_feat2.reserve(capd));
    CK2(D.q_order.reserve(capd));
    CK2(D.q_order2.reserve(capd));
    CK2(D.q_cnt.reserve(1));
    CK2(cudaMemcpyAsync(D.q_rays.p, rays, 20 * (size_t)n_rays, cudaMemcpyHostToDevice, s));
    CK2(cudaMemsetAsync(D.q_cnt.p, 0, 4, s));
    d2::WorldRay2Args A;
    memset(&A, 0, sizeof A);
    A.W.n = n;
    A.W.pos = D.pos.p, A.W.rot = D.rot.p, A.W.type = D.type.p, A.W.param = D.param.p, A.W.poly = D.poly.p;
    A.leaf_lo = ctx->leaf_lo.p, A.leaf_hi = ctx->leaf_hi.p, A.nodes = ctx->nodes.p;
    A.n_tree = n - ctx->last_counters.n_outliers;
    A.groups = D.last_groups ? D.groups.p : nullptr;
    A.use_groups = groups != nullptr;
    if (groups) A.qg[0] = groups[0], A.qg[1] = groups[1], A.qg[2] = groups[2];
    A.rays = D.q_rays.p, A.n_rays = n_rays;
    A.keys = D.q_keys.p, A.vals = D.q_vals.p, A.feats = D.q_feat.p, A.cap = cap, A.counter = D.q_cnt.p;
    A.trav_overflow = trav_overflow_counter(ctx);
    if (first_only)
        d2::k_world_ray2d<true><<<(n_rays + 127) / 128, 128, 0, s>>>(A);
    else
        d2::k_world_ray2d<false><<<(n_rays + 127) / 128, 128, 0, s>>>(A);
    CK2(cudaGetLastError());
    uint32_t found = 0;
    CK2(cudaMemcpyAsync(&found, D.q_cnt.p, 4, cudaMemcpyDeviceToHost, s));
    CK2(cudaStreamSynchronize(s));
    *n_out = found;
    uint32_t w = found < cap ? found : cap;
    if (w == 0) return found > cap ? 1 : NCB_OK;
    // rows in (ray, handle) order: 64-bit radix sort of the keys with their positions, then one gather
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)w, 0, 64);
    CK2(D.q_tmp.reserve(tmp + 256));
    d2::k_iota2d<<<(w + 255) / 256, 256, 0, s>>>(D.q_order.p, w);
    tmp = D.q_tmp.cap;
    CK2(cub::DeviceRadixSort::SortPairs(D.q_tmp.p, tmp, D.q_keys.p, D.q_keys2.p, D.q_order.p, D.q_order2.p, (int)w, 0, 64, s));
    d2::k_gather_rows2d<<<(w + 255) / 256, 256, 0, s>>>(D.q_order2.p, D.q_vals.p, D.q_feat.p, w, D.q_vals2.p, D.q_feat2.p);
    CK2(cudaGetLastError());
    std::vector<unsigned long long> hk(w);
    std::vector<float4> hv(w);
    CK2(cudaMemcpyAsync(hk.data(), D.q_keys2.p, 8 * (size_t)w, cudaMemcpyDeviceToHost, s));
    CK2(cudaMemcpyAsync(hv.data(), D.q_vals2.p, 16 * (size_t)w, cudaMemcpyDeviceToHost, s));
    CK2(cudaMemcpyAsync(feat, D.q_feat2.p, 4 * (size_t)w, cudaMemcpyDeviceToHost, s));
    CK2(cudaStreamSynchronize(s));
    for (uint32_t k = 0; k < w; ++k) {
        idx[2 * (size_t)k] = (uint32_t)(hk[k] >> 32), idx[2 * (size_t)k + 1] = (uint32_t)hk[k];
        val[3 * (size_t)k] = hv[k].x, val[3 * (size_t)k + 1] = hv[k].y, val[3 * (size_t)k + 2] = hv[k].z;
    }
    return found > cap ? 1 : NCB_OK;
}

// glue::interferences_with_aabb (kind 0) / interferences_with_point (kind 2) against the world of the last ncb2d_world_update.
int ncb2d_world_query(ncb_ctx* ctx, int kind, uint32_t n_queries, const float* queries, const uint32_t* groups, uint32_t* idx, uint32_t cap,
                      uint32_t* n_out) {
    if (!ctx || !n_out || (kind != 0 && kind != 2) || (n_queries && !queries) || (cap && !idx)) return NCB_ERR_ARG;
    *n_out = 0;
    uint32_t n = ctx->d2.last_n;
    if (n == 0) {
        ctx->err = "ncb2d_world_query: no 2-D world on the device (call ncb2d_world_update first)";
        return NCB_ERR_ARG;
    }
    if (n_queries == 0) return NCB_OK;
    CK2(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    auto& D = ctx->d2;
    const size_t W = kind == 0 ? 4 : 2, capd = cap ? cap : 1;
    CK2(D.q_rays.reserve(W * (size_t)n_queries));
    CK2(D.q_keys.reserve(capd));
    CK2(D.q_keys2.reserve(capd));
    CK2(D.q_cnt.reserve(1));
    CK2(cudaMemcpyAsync(D.q_rays.p, queries, 4 * W * (size_t)n_queries, cudaMemcpyHostToDevice, s));
    CK2(cudaMemsetAsync(D.q_cnt.p, 0, 4, s));
    d2::WorldRay2Args A;
    memset(&A, 0, sizeof A);
    A.W.n = n;
    A.W.pos = D.pos.p, A.W.rot = D.rot.p, A.W.type = D.type.p, A.W.param = D.param.p, A.W.poly = D.poly.p;
    A.leaf_lo = ctx->leaf_lo.p, A.leaf_hi = ctx->leaf_hi.p, A.nodes = ctx->nodes.p;
    A.n_tree = n - ctx->last_counters.n_outliers;
    A.groups = D.last_groups ? D.groups.p : nullptr;
    A.use_groups = groups != nullptr;
    if (groups) A.qg[0] = groups[0], A.qg[1] = groups[1], A.qg[2] = groups[2];
    A.rays = D.q_rays.p, A.n_rays = n_queries;
    A.keys = D.q_keys.p, A.cap = cap, A.counter = D.q_cnt.p;
    A.trav_overflow = trav_overflow_counter(ctx);
    if (kind == 0)
        d2::k_world_query2d<0><<<(n_queries + 127) / 128, 128, 0, s>>>(A);
    else
        d2::k_world_query2d<2><<<(n_queries + 127) / 128, 128, 0, s>>>(A);
    CK2(cudaGetLastError());
    uint32_t found = 0;
    CK2(cudaMemcpyAsync(&found, D.q_cnt.p, 4, cudaMemcpyDeviceToHost, s));
    CK2(cudaStreamSynchronize(s));
    *n_out = found;
    uint32_t w = found < cap ? found : cap;
    if (w == 0) return found > cap ? 1 : NCB_OK;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)w, 0, 64);
    CK2(D.q_tmp.reserve(tmp + 256));
    tmp = D.q_tmp.cap;
    CK2(cub::DeviceRadixSort::SortKeys(D.q_tmp.p, tmp, D.q_keys.p, D.q_keys2.p, (int)w, 0, 64, s));
    std::vector<unsigned long long> hk(w);
    CK2(cudaMemcpyAsync(hk.data(), D.q_keys2.p, 8 * (size_t)w, cudaMemcpyDeviceToHost, s));
    CK2(cudaStreamSynchronize(s));
    for (uint32_t k = 0; k < w; ++k) idx[2 * (size_t)k] = (uint32_t)(hk[k] >> 32), idx[2 * (size_t)k + 1] = (uint32_t)hk[k];
    return found > cap ? 1 : NCB_OK;
}

int ncb2d_world_fetch_proximity(ncb_ctx* ctx, uint8_t* prox, uint32_t cap_pairs) {
    if (!ctx || (cap_pairs && !prox)) return NCB_ERR_ARG;
    CK2(cudaSetDevice(ctx->device));
    uint32_t np = ctx->d2.last_pairs, w = np < cap_pairs ? np : cap_pairs;
    if (w) {
        CK2(cudaMemcpyAsync(prox, ctx->d2.prox.p, w, cudaMemcpyDeviceToHost, ctx->stream));
        CK2(cudaStreamSynchronize(ctx->stream));
    }
    return np > cap_pairs ? 1 : NCB_OK;
}

}  // extern "C"
#endif  // NCB_HOST_SHIM (host entry points)

