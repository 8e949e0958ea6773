// ORACLE — TEST INFRASTRUCTURE ONLY (see na.hpp header).
// Point projection on segment / triangle / tetrahedron and the Voronoi simplex used by GJK, restated from
//   query/point/point_segment.rs:52-91, query/point/point_triangle.rs:61-309,
//   query/point/point_tetrahedron.rs:35-353, query/algorithms/voronoi_simplex3.rs:12-352,
//   query/algorithms/cso_point.rs:14-85.
// All projections here are of a point `p` with the identity isometry (the only way the path calls them).
#pragma once
#include "na.hpp"

namespace orc {

struct CSOPoint {
    V3 point, orig1, orig2;
};
static inline CSOPoint cso_new(V3 o1, V3 o2) { return {o1 - o2, o1, o2}; }
static inline CSOPoint cso_origin() { return cso_new(v3(0, 0, 0), v3(0, 0, 0)); }

enum LocKind { ON_VERTEX, ON_EDGE, ON_FACE, ON_SOLID };
struct Location {
    LocKind kind;
    int id;         // vertex / edge / face index (for triangle OnFace: the face side)
    real bc[3];     // barycentric coordinates (2 for edges, 3 for faces)
};

// point_segment.rs:52-91
static inline V3 project_on_segment(V3 a, V3 b, V3 p, Location* loc) {
    V3 ab = b - a, ap = p - a;
    real ab_ap = dot(ab, ap), sqnab = norm_squared(ab);
    if (ab_ap <= real(0)) {
        *loc = {ON_VERTEX, 0, {0, 0, 0}};
        return a;
    } else if (ab_ap >= sqnab) {
        *loc = {ON_VERTEX, 1, {0, 0, 0}};
        return b;
    }
    real u = ab_ap / sqnab;
    *loc = {ON_EDGE, 0, {real(1) - u, u, 0}};
    return a + ab * u;
}

// point_triangle.rs:61-309 (dim3, `solid` as passed by the caller)
static inline V3 project_on_triangle(V3 a, V3 b, V3 c, V3 p, bool solid, Location* loc) {
    const real _1 = 1, _0 = 0;
    V3 ab = b - a, ac = c - a, ap = p - a;
    real ab_ap = dot(ab, ap), ac_ap = dot(ac, ap);
    if (ab_ap <= _0 && ac_ap <= _0) {
        *loc = {ON_VERTEX, 0, {0, 0, 0}};
        return a;
    }
    V3 bp = p - b;
    real ab_bp = dot(ab, bp), ac_bp = dot(ac, bp);
    if (ab_bp >= _0 && ac_bp <= ab_bp) {
        *loc = {ON_VERTEX, 1, {0, 0, 0}};
        return b;
    }
    V3 cp = p - c;
    real ab_cp = dot(ab, cp), ac_cp = dot(ac, cp);
    if (ac_cp >= _0 && ab_cp <= ac_cp) {
        *loc = {ON_VERTEX, 2, {0, 0, 0}};
        return c;
    }
    V3 bc = c - b;
    // stable_check_edges_voronoi (dim3, without improved_fixed_point_support)
    V3 n = cross(ab, ac);
    real vc = dot(n, cross(ab, ap));
    if (vc < _0 && ab_ap >= _0 && ab_bp <= _0) {
        real v = ab_ap / norm_squared(ab);
        *loc = {ON_EDGE, 0, {_1 - v, v, 0}};
        return a + ab * v;
    }
    real vb = -dot(n, cross(ac, cp));
    if (vb < _0 && ac_ap >= _0 && ac_cp <= _0) {
        real w = ac_ap / norm_squared(ac);
        *loc = {ON_EDGE, 2, {_1 - w, w, 0}};
        return a + ac * w;
    }
    real va = dot(n, cross(bc, bp));
    if (va < _0 && ac_bp - ab_bp >= _0 && ab_cp - ac_cp >= _0) {
        real w = dot(bc, bp) / norm_squared(bc);
        *loc = {ON_EDGE, 1, {_1 - w, w, 0}};
        return b + bc * w;
    }
    int clockwise = dot(n, ap) >= _0 ? 0 : 1;
    if (va + vb + vc != _0) {
        real denom = _1 / (va + vb + vc);
        real v = vb * denom, w = vc * denom;
        *loc = {ON_FACE, clockwise, {_1 - v - w, v, w}};
        return a + ab * v + ac * w;
    }
    if (solid) {
        *loc = {ON_SOLID, 0, {0, 0, 0}};
        return p;
    }
    // non-solid fallback: project on the closest edge (point_triangle.rs:262-307)
    real v = ab_ap / (ab_ap - ab_bp);
    real w = ac_ap / (ac_ap - ac_cp);
    real u = (ac_bp - ab_bp) / (ac_bp - ab_bp + ab_cp - ac_cp);
    real d_ab = norm_squared(ap) - (norm_squared(ab) * v * v);
    real d_ac = norm_squared(ap) - (norm_squared(ac) * u * u);
    real d_bc = norm_squared(bp) - (norm_squared(bc) * w * w);
    if (d_ab < d_ac) {
        if (d_ab < d_bc) {
            *loc = {ON_EDGE, 0, {_1 - v, v, 0}};
            return a + ab * v;
        }
        *loc = {ON_EDGE, 1, {_1 - u, u, 0}};
        return b + bc * u;
    }
    if (d_ac < d_bc) {
        *loc = {ON_EDGE, 2, {_1 - w, w, 0}};
        return a + ac * w;
    }
    *loc = {ON_EDGE, 1, {_1 - u, u, 0}};
    return b + bc * u;
}

// point_tetrahedron.rs:35-353 (solid = true; the non-solid branch is unimplemented!() in the reference)
static inline bool tetra_check_edge(int i, V3 a, V3 nabc, V3 nabd, V3 ap, V3 ab, real ap_ab, real bp_ab, real* dabc, real* dabd,
                                    V3* proj, Location* loc) {
    const real _0 = 0, _1 = 1;
    real ab_ab = ap_ab - bp_ab;
    V3 ap_x_ab = cross(ap, ab);
    *dabc = dot(ap_x_ab, nabc);
    *dabd = dot(ap_x_ab, nabd);
    if (ab_ab != _0 && *dabc >= _0 && *dabd >= _0 && ap_ab >= _0 && ap_ab <= ab_ab) {
        real u = ap_ab / ab_ab;
        *loc = {ON_EDGE, i, {_1 - u, u, 0}};
        *proj = a + ab * u;
        return true;
    }
    return false;
}
static inline bool tetra_check_face(int i, V3 a, V3 b, V3 c, V3 ap, V3 bp, V3 cp, V3 ab, V3 ac, V3 ad, real dabc, real dbca,
                                    real dacb, V3* proj, Location* loc) {
    const real _0 = 0, _1 = 1;
    if (dabc < _0 && dbca < _0 && dacb < _0) {
        V3 n = cross(ab, ac);
        if (dot(n, ad) * dot(n, ap) < _0) {
            V3 normal;
            if (!try_normalize(n, EPS, &normal)) return false;  // `?` inside check_face
            real vc = dot(normal, cross(ap, bp));
            real va = dot(normal, cross(bp, cp));
            real vb = dot(normal, cross(cp, ap));
            real denom = va + vb + vc;
            real inv_denom = _1 / denom;
            real b0 = va * inv_denom, b1 = vb * inv_denom, b2 = vc * inv_denom;
            *loc = {ON_FACE, i, {b0, b1, b2}};
            *proj = a * b0 + b * b1 + c * b2;
            return true;
        }
    }
    return false;
}
static inline V3 project_on_tetrahedron(V3 a, V3 b, V3 c, V3 d, V3 p, Location* loc) {
    const real _0 = 0;
    V3 ab = b - a, ac = c - a, ad = d - a, ap = p - a;
    real ap_ab = dot(ap, ab), ap_ac = dot(ap, ac), ap_ad = dot(ap, ad);
    if (ap_ab <= _0 && ap_ac <= _0 && ap_ad <= _0) {
        *loc = {ON_VERTEX, 0, {0, 0, 0}};
        return a;
    }
    V3 bc = c - b, bd = d - b, bp = p - b;
    real bp_bc = dot(bp, bc), bp_bd = dot(bp, bd), bp_ab = dot(bp, ab);
    if (bp_bc <= _0 && bp_bd <= _0 && bp_ab >= _0) {
        *loc = {ON_VERTEX, 1, {0, 0, 0}};
        return b;
    }
    V3 cd = d - c, cp = p - c;
    real cp_ac = dot(cp, ac), cp_bc = dot(cp, bc), cp_cd = dot(cp, cd);
    if (cp_cd <= _0 && cp_bc >= _0 && cp_ac >= _0) {
        *loc = {ON_VERTEX, 2, {0, 0, 0}};
        return c;
    }
    V3 dp = p - d;
    real dp_cd = dot(dp, cd), dp_bd = dot(dp, bd), dp_ad = dot(dp, ad);
    if (dp_ad >= _0 && dp_bd >= _0 && dp_cd >= _0) {
        *loc = {ON_VERTEX, 3, {0, 0, 0}};
        return d;
    }
    V3 proj;
    V3 nabc = cross(ab, ac), nabd = cross(ab, ad);
    real dabc, dabd;
    if (tetra_check_edge(0, a, nabc, nabd, ap, ab, ap_ab, bp_ab, &dabc, &dabd, &proj, loc)) return proj;
    V3 nacd = cross(ac, ad);
    real dacd, dacb;
    if (tetra_check_edge(1, a, nacd, -nabc, ap, ac, ap_ac, cp_ac, &dacd, &dacb, &proj, loc)) return proj;
    real dadb, dadc;
    if (tetra_check_edge(2, a, -nabd, -nacd, ap, ad, ap_ad, dp_ad, &dadb, &dadc, &proj, loc)) return proj;
    V3 nbcd = cross(bc, bd);
    real dbca, dbcd;
    if (tetra_check_edge(3, b, nabc, nbcd, bp, bc, bp_bc, cp_bc, &dbca, &dbcd, &proj, loc)) return proj;
    real dbdc, dbda;
    if (tetra_check_edge(4, b, -nbcd, nabd, bp, bd, bp_bd, dp_bd, &dbdc, &dbda, &proj, loc)) return proj;
    real dcda, dcdb;
    if (tetra_check_edge(5, c, nacd, nbcd, cp, cd, cp_cd, dp_cd, &dcda, &dcdb, &proj, loc)) return proj;

    if (tetra_check_face(0, a, b, c, ap, bp, cp, ab, ac, ad, dabc, dbca, dacb, &proj, loc)) return proj;
    if (tetra_check_face(1, a, b, d, ap, bp, dp, ab, ad, ac, dadb, dabd, dbda, &proj, loc)) return proj;
    if (tetra_check_face(2, a, c, d, ap, cp, dp, ac, ad, ab, dacd, dcda, dadc, &proj, loc)) return proj;
    if (tetra_check_face(3, b, c, d, bp, cp, dp, bc, bd, -ab, dbcd, dcdb, dbdc, &proj, loc)) return proj;
    *loc = {ON_SOLID, 0, {0, 0, 0}};
    return p;
}

// voronoi_simplex3.rs
struct VoronoiSimplex {
    int prev_vertices[4] = {0, 1, 2, 3};
    real prev_proj[3] = {0, 0, 0};
    int prev_dim = 0;
    CSOPoint vertices[4];
    real proj[3] = {0, 0, 0};
    int dim = 0;

    VoronoiSimplex() {
        for (auto& v : vertices) v = cso_origin();
    }
    void swap(int i1, int i2) {
        std::swap(vertices[i1], vertices[i2]);
        std::swap(prev_vertices[i1], prev_vertices[i2]);
    }
    void reset(const CSOPoint& pt) {
        dim = 0;
        prev_dim = 0;
        vertices[0] = pt;
    }
    bool add_point(const CSOPoint& pt, real eps_tol) {  // :49-85
        prev_dim = dim;
        for (int i = 0; i < 3; ++i) prev_proj[i] = proj[i];
        for (int i = 0; i < 4; ++i) prev_vertices[i] = i;
        if (dim == 0) {
            if (norm_squared(vertices[0].point - pt.point) < eps_tol) return false;
        } else if (dim == 1) {
            V3 ab = vertices[1].point - vertices[0].point, ac = pt.point - vertices[0].point;
            if (norm_squared(cross(ab, ac)) < eps_tol) return false;
        } else {
            V3 ab = vertices[1].point - vertices[0].point, ac = vertices[2].point - vertices[0].point;
            V3 ap = pt.point - vertices[0].point;
            V3 n = normalize(cross(ab, ac));
            if (std::fabs(dot(n, ap)) < eps_tol) return false;
        }
        dim += 1;
        vertices[dim] = pt;
        return true;
    }
    const CSOPoint& prev_point(int i) const { return vertices[prev_vertices[i]]; }

    V3 project_origin_and_reduce() {  // :115-280
        const V3 O = v3(0, 0, 0);
        Location loc;
        if (dim == 0) {
            proj[0] = 1;
            return vertices[0].point;
        } else if (dim == 1) {
            V3 p = project_on_segment(vertices[0].point, vertices[1].point, O, &loc);
            if (loc.kind == ON_VERTEX && loc.id == 0) {
                proj[0] = 1;
                dim = 0;
            } else if (loc.kind == ON_VERTEX) {
                swap(0, 1);
                proj[0] = 1;
                dim = 0;
            } else {
                proj[0] = loc.bc[0];
                proj[1] = loc.bc[1];
            }
            return p;
        } else if (dim == 2) {
            V3 p = project_on_triangle(vertices[0].point, vertices[1].point, vertices[2].point, O, true, &loc);
            if (loc.kind == ON_VERTEX) {
                swap(0, loc.id);
                proj[0] = 1;
                dim = 0;
            } else if (loc.kind == ON_EDGE && loc.id == 0) {
                proj[0] = loc.bc[0];
                proj[1] = loc.bc[1];
                dim = 1;
            } else if (loc.kind == ON_EDGE && loc.id == 1) {
                swap(0, 2);
                proj[0] = loc.bc[1];
                proj[1] = loc.bc[0];
                dim = 1;
            } else if (loc.kind == ON_EDGE && loc.id == 2) {
                swap(1, 2);
                proj[0] = loc.bc[0];
                proj[1] = loc.bc[1];
                dim = 1;
            } else if (loc.kind == ON_FACE) {
                proj[0] = loc.bc[0], proj[1] = loc.bc[1], proj[2] = loc.bc[2];
            }
            return p;
        } else {
            V3 p = project_on_tetrahedron(vertices[0].point, vertices[1].point, vertices[2].point, vertices[3].point, O, &loc);
            if (loc.kind == ON_VERTEX) {
                swap(0, loc.id);
                proj[0] = 1;
                dim = 0;
            } else if (loc.kind == ON_EDGE) {
                switch (loc.id) {
                    case 0: break;
                    case 1: swap(1, 2); break;
                    case 2: swap(1, 3); break;
                    case 3: swap(0, 2); break;
                    case 4: swap(0, 3); break;
                    default: swap(0, 2); swap(1, 3); break;
                }
                if (loc.id == 3 || loc.id == 4) {
                    proj[0] = loc.bc[1];
                    proj[1] = loc.bc[0];
                } else {
                    proj[0] = loc.bc[0];
                    proj[1] = loc.bc[1];
                }
                dim = 1;
            } else if (loc.kind == ON_FACE) {
                switch (loc.id) {
                    case 0:
                        proj[0] = loc.bc[0], proj[1] = loc.bc[1], proj[2] = loc.bc[2];
                        break;
                    case 1:
                        vertices[2] = vertices[3];
                        proj[0] = loc.bc[0], proj[1] = loc.bc[1], proj[2] = loc.bc[2];
                        break;
                    case 2:
                        vertices[1] = vertices[3];
                        proj[0] = loc.bc[0], proj[1] = loc.bc[2], proj[2] = loc.bc[1];
                        break;
                    default:
                        vertices[0] = vertices[3];
                        proj[0] = loc.bc[2], proj[1] = loc.bc[0], proj[2] = loc.bc[1];
                        break;
                }
                dim = 2;
            }
            return p;
        }
    }
};

}  // namespace orc
