// Device GJK for one pair per thread (the EPA that takes over when the origin is inside the CSO is in epa.cuh, included at
// the end of this file): fixed capacity, no recursion, no heap allocation.
//
// Replaces (reference, file:line): query/algorithms/gjk.rs:76-177,367-388; voronoi_simplex3.rs:49-280;
// cso_point.rs:70-85; query/point/point_segment.rs:52-91, point_triangle.rs:61-309, point_tetrahedron.rs:35-353;
// query/algorithms/epa3.rs:219-454 (std BinaryHeap + Vec + recursive silhouette flood -> fixed arrays + explicit
// stack); query/contact/contact_support_map_support_map.rs:38-79; shape/support_map.rs:26-29; shape/cuboid.rs:137-145;
// utils/point_cloud_support_point.rs:6-24.
//
// Contact parity needs the SAME iteration path as the reference (SURVEY.md §7 hard part 3b): same start direction,
// same support tie-breaks, same simplex permutations, same exit tests, no FMA contraction.
#pragma once
#include "ncb_internal.h"
#include "vec.cuh"

namespace ncb {

struct HullView {
    uint32_t nv, nf;
    const float* pts;
    const float* fnormal;
    const uint32_t *face_first, *face_num, *vaf, *eaf;
    const uint32_t *vfirst, *vnum, *fav, *eav;
    const uint32_t *edge_vertices, *edge_faces;
    const float* edge_dir;
    NCB_HD V3 pt(uint32_t i) const { return v3(__ldg(pts + 3 * i), __ldg(pts + 3 * i + 1), __ldg(pts + 3 * i + 2)); }
    NCB_HD V3 fn(uint32_t i) const { return v3(__ldg(fnormal + 3 * i), __ldg(fnormal + 3 * i + 1), __ldg(fnormal + 3 * i + 2)); }
    NCB_HD V3 edir(uint32_t i) const { return v3(__ldg(edge_dir + 3 * i), __ldg(edge_dir + 3 * i + 1), __ldg(edge_dir + 3 * i + 2)); }
};

NCB_HD HullView hull_view(const DevHulls& L, uint32_t h) {
    HullView H;
    uint32_t v0 = __ldg(L.vert_off + h), f0 = __ldg(L.face_off + h), e0 = __ldg(L.edge_off + h);
    uint32_t fa0 = __ldg(L.fadj_off + h), va0 = __ldg(L.vadj_off + h);
    H.nv = __ldg(L.vert_off + h + 1) - v0;
    H.nf = __ldg(L.face_off + h + 1) - f0;
    H.pts = L.points + 3 * (size_t)v0;
    H.fnormal = L.face_normal + 3 * (size_t)f0;
    H.face_first = L.face_first + f0;
    H.face_num = L.face_num + f0;
    H.vaf = L.vaf + fa0;
    H.eaf = L.eaf + fa0;
    H.vfirst = L.vert_first_adj + v0;
    H.vnum = L.vert_num_adj + v0;
    H.fav = L.fav + va0;
    H.eav = L.eav + va0;
    H.edge_vertices = L.edge_vertices + 2 * (size_t)e0;
    H.edge_faces = L.edge_faces + 2 * (size_t)e0;
    H.edge_dir = L.edge_dir + 3 * (size_t)e0;
    return H;
}

// A support-mapped operand.  kind: 0 cuboid, 1 hull, 2 constant origin.
struct Support {
    int kind;
    V3 he;
    HullView hull;
    NCB_HD uint32_t nv() const { return hull.nv; }
    NCB_HD V3 pt(uint32_t i) const { return hull.pt(i); }
};
// The same operand with only what a support evaluation reads (kind, half extents | vertex array): the EPA kernels keep two of
// these per lane in registers across expansion steps instead of the 13-pointer HullView.
struct SupportS {
    int kind;
    V3 he;
    uint32_t nverts;
#ifndef NCB_HOST_SHIM
    const float4* pts4;  // vertices padded to 16 B (DevHulls::points4): one load per vertex
    NCB_HD V3 pt(uint32_t i) const {
        float4 q = __ldg(pts4 + i);
        return v3(q.x, q.y, q.z);
    }
#else  // the host shim reads the packed array of the hull view
    const float* pts;
    NCB_HD V3 pt(uint32_t i) const { return v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]); }
#endif
    NCB_HD uint32_t nv() const { return nverts; }
};
#ifdef NCB_HOST_SHIM
NCB_HD SupportS slim_support(const Support& g) {
    SupportS r;
    r.kind = g.kind, r.he = g.he, r.nverts = g.hull.nv, r.pts = g.hull.pts;
    return r;
}
#endif

template <class G>
NCB_HD V3 local_support_point(const G& g, V3 dir) {
    if (g.kind == 0) return v3(copysignf(g.he.x, dir.x), copysignf(g.he.y, dir.y), copysignf(g.he.z, dir.z));
    uint32_t best = 0;
    float best_dot = dot(g.pt(0), dir);
    for (uint32_t i = 1; i < g.nv(); ++i) {
        float d = dot(g.pt(i), dir);
        if (d > best_dot) {
            best_dot = d;
            best = i;
        }
    }
    return g.pt(best);
}
template <class G>
NCB_HD V3 support_point(const G& g, const Iso& m, V3 dir) {
    if (g.kind == 2) return v3(0.f, 0.f, 0.f);
    V3 ld = iso_inv_vec(m, dir);
    return iso_mul_point(m, local_support_point(g, ld));
}

struct CSOPoint {
    V3 point, orig1, orig2;
};
// CSOPoint::from_shapes (cso_point.rs:70-85).  The two support evaluations are independent, so they are issued in a
// canonical order (the O(1) operand first, the vertex-scanning hull second) whatever the pair's orientation is: lanes
// of a warp holding (cuboid, hull) and (hull, cuboid) pairs then run the same code at the same time.  Values are
// unchanged: each operand still sees its own isometry and direction.
template <class G>
NCB_HD CSOPoint cso_from_shapes(const Iso& m1, const G& g1, const Iso& m2, const G& g2, V3 dir) {
    CSOPoint c;
    bool swap = g1.kind == 1 && g2.kind != 1;
    const G& ga = swap ? g2 : g1;
    const G& gb = swap ? g1 : g2;
    const Iso& ia = swap ? m2 : m1;
    const Iso& ib = swap ? m1 : m2;
    V3 da = swap ? -dir : dir;
    V3 sa = support_point(ga, ia, da);
    V3 sb = support_point(gb, ib, -da);
    c.orig1 = swap ? sb : sa;
    c.orig2 = swap ? sa : sb;
    c.point = c.orig1 - c.orig2;
    return c;
}

// ---- Voronoi-region projections of the ORIGIN's generalisation `p` (identity isometry) ---------------------
enum { LOC_VERTEX = 0, LOC_EDGE = 1, LOC_FACE = 2, LOC_SOLID = 3 };
struct Loc {
    int kind, id;
    float b0, b1, b2;
};
NCB_HD Loc mkloc(int kind, int id, float b0 = 0.f, float b1 = 0.f, float b2 = 0.f) { return Loc{kind, id, b0, b1, b2}; }

NCB_HD V3 proj_segment(V3 a, V3 b, V3 p, Loc& loc) {
    V3 ab = b - a, ap = p - a;
    float ab_ap = dot(ab, ap), sqnab = norm_squared(ab);
    if (ab_ap <= 0.f) {
        loc = mkloc(LOC_VERTEX, 0);
        return a;
    }
    if (ab_ap >= sqnab) {
        loc = mkloc(LOC_VERTEX, 1);
        return b;
    }
    float u = ab_ap / sqnab;
    loc = mkloc(LOC_EDGE, 0, 1.f - u, u);
    return a + ab * u;
}

// solid = true variant only (the one the simplex and EPA use).  POINT = false: only the location (region + barycentric
// coordinates) is wanted; the decisions and the coordinates come from the same expressions either way.
// Code shape: the three edge regions share ONE tail (the quotient, the location record, the point) — a region test only selects its
// operands.  The kernels that run this are instruction-fetch bound (profiles/r2_ncu_gjk_fetch_bound.txt): one copy of each tail keeps
// the code short, and lanes of a warp that land in different edge regions run the tail together.  Values are those of the
// straight-line version, expression for expression.
template <bool POINT>
NCB_HD V3 proj_triangle_core(V3 a, V3 b, V3 c, V3 p, Loc& loc) {
    V3 ab = b - a, ac = c - a, ap = p - a;
    float ab_ap = dot(ab, ap), ac_ap = dot(ac, ap);
    if (ab_ap <= 0.f && ac_ap <= 0.f) {
        loc = mkloc(LOC_VERTEX, 0);
        return a;
    }
    V3 bp = p - b;
    float ab_bp = dot(ab, bp), ac_bp = dot(ac, bp);
    if (ab_bp >= 0.f && ac_bp <= ab_bp) {
        loc = mkloc(LOC_VERTEX, 1);
        return b;
    }
    V3 cp = p - c;
    float ab_cp = dot(ab, cp), ac_cp = dot(ac, cp);
    if (ac_cp >= 0.f && ab_cp <= ac_cp) {
        loc = mkloc(LOC_VERTEX, 2);
        return c;
    }
    V3 bc = c - b;
    V3 n = cross(ab, ac);
    float vc = dot(n, cross(ab, ap));
    int edge = -1;  // the edge region found, with the operands of its tail: coordinate = num / norm_squared(dir), point = from + dir * coordinate
    float num = 0.f;
    V3 from = a, dir = ab;
    float vb = 0.f, va = 0.f;
    if (vc < 0.f && ab_ap >= 0.f && ab_bp <= 0.f) {
        edge = 0, num = ab_ap;
    } else {
        vb = -dot(n, cross(ac, cp));
        if (vb < 0.f && ac_ap >= 0.f && ac_cp <= 0.f) {
            edge = 2, num = ac_ap, dir = ac;
        } else {
            va = dot(n, cross(bc, bp));
            if (va < 0.f && ac_bp - ab_bp >= 0.f && ab_cp - ac_cp >= 0.f) edge = 1, num = dot(bc, bp), from = b, dir = bc;
        }
    }
    if (edge >= 0) {
        float w = num / norm_squared(dir);
        loc = mkloc(LOC_EDGE, edge, 1.f - w, w);
        return POINT ? from + dir * w : a;
    }
    int clockwise = dot(n, ap) >= 0.f ? 0 : 1;
    if (va + vb + vc != 0.f) {
        float denom = 1.f / (va + vb + vc);
        float v = vb * denom, w = vc * denom;
        loc = mkloc(LOC_FACE, clockwise, 1.f - v - w, v, w);
        return POINT ? a + ab * v + ac * w : a;
    }
    loc = mkloc(LOC_SOLID, 0);
    return p;
}
static __device__ __noinline__ V3 proj_triangle(V3 a, V3 b, V3 c, V3 p, Loc& loc) { return proj_triangle_core<true>(a, b, c, p, loc); }

// The location-only form the EPA kernels inline (two sites): straight-line, every region returns at once (measured: the shared-tail
// form above costs the first EPA tier 4 %; its selects buy nothing when no point is computed).
template <>
NCB_HD V3 proj_triangle_core<false>(V3 a, V3 b, V3 c, V3 p, Loc& loc) {
    V3 ab = b - a, ac = c - a, ap = p - a;
    float ab_ap = dot(ab, ap), ac_ap = dot(ac, ap);
    if (ab_ap <= 0.f && ac_ap <= 0.f) {
        loc = mkloc(LOC_VERTEX, 0);
        return a;
    }
    V3 bp = p - b;
    float ab_bp = dot(ab, bp), ac_bp = dot(ac, bp);
    if (ab_bp >= 0.f && ac_bp <= ab_bp) {
        loc = mkloc(LOC_VERTEX, 1);
        return b;
    }
    V3 cp = p - c;
    float ab_cp = dot(ab, cp), ac_cp = dot(ac, cp);
    if (ac_cp >= 0.f && ab_cp <= ac_cp) {
        loc = mkloc(LOC_VERTEX, 2);
        return c;
    }
    V3 bc = c - b;
    V3 n = cross(ab, ac);
    float vc = dot(n, cross(ab, ap));
    if (vc < 0.f && ab_ap >= 0.f && ab_bp <= 0.f) {
        float v = ab_ap / norm_squared(ab);
        loc = mkloc(LOC_EDGE, 0, 1.f - v, v);
        return a;
    }
    float vb = -dot(n, cross(ac, cp));
    if (vb < 0.f && ac_ap >= 0.f && ac_cp <= 0.f) {
        float w = ac_ap / norm_squared(ac);
        loc = mkloc(LOC_EDGE, 2, 1.f - w, w);
        return a;
    }
    float va = dot(n, cross(bc, bp));
    if (va < 0.f && ac_bp - ab_bp >= 0.f && ab_cp - ac_cp >= 0.f) {
        float w = dot(bc, bp) / norm_squared(bc);
        loc = mkloc(LOC_EDGE, 1, 1.f - w, w);
        return a;
    }
    int clockwise = dot(n, ap) >= 0.f ? 0 : 1;
    if (va + vb + vc != 0.f) {
        float denom = 1.f / (va + vb + vc);
        float v = vb * denom, w = vc * denom;
        loc = mkloc(LOC_FACE, clockwise, 1.f - v - w, v, w);
        return a;
    }
    loc = mkloc(LOC_SOLID, 0);
    return p;
}

// Tetrahedron::project_point_with_location's edge and face tests (point_tetrahedron.rs): the test of a region; its tail (quotients,
// location, point) exists once in proj_tetrahedron.
NCB_HD bool tetra_edge_test(V3 nabc, V3 nabd, V3 ap, V3 ab, float ap_ab, float bp_ab, float& dabc, float& dabd) {
    float ab_ab = ap_ab - bp_ab;
    V3 ap_x_ab = cross(ap, ab);
    dabc = dot(ap_x_ab, nabc);
    dabd = dot(ap_x_ab, nabd);
    return ab_ab != 0.f && dabc >= 0.f && dabd >= 0.f && ap_ab >= 0.f && ap_ab <= ab_ab;
}
NCB_HD bool tetra_face_test(V3 ap, V3 ab, V3 ac, V3 ad, float dabc, float dbca, float dacb, V3& n) {
    if (dabc < 0.f && dbca < 0.f && dacb < 0.f) {
        n = cross(ab, ac);
        return dot(n, ad) * dot(n, ap) < 0.f;
    }
    return false;
}
static __device__ __noinline__ V3 proj_tetrahedron(V3 a, V3 b, V3 c, V3 d, V3 p, Loc& loc) {
    V3 ab = b - a, ac = c - a, ad = d - a, ap = p - a;
    float ap_ab = dot(ap, ab), ap_ac = dot(ap, ac), ap_ad = dot(ap, ad);
    if (ap_ab <= 0.f && ap_ac <= 0.f && ap_ad <= 0.f) {
        loc = mkloc(LOC_VERTEX, 0);
        return a;
    }
    V3 bc = c - b, bd = d - b, bp = p - b;
    float bp_bc = dot(bp, bc), bp_bd = dot(bp, bd), bp_ab = dot(bp, ab);
    if (bp_bc <= 0.f && bp_bd <= 0.f && bp_ab >= 0.f) {
        loc = mkloc(LOC_VERTEX, 1);
        return b;
    }
    V3 cd = d - c, cp = p - c;
    float cp_ac = dot(cp, ac), cp_bc = dot(cp, bc), cp_cd = dot(cp, cd);
    if (cp_cd <= 0.f && cp_bc >= 0.f && cp_ac >= 0.f) {
        loc = mkloc(LOC_VERTEX, 2);
        return c;
    }
    V3 dp = p - d;
    float dp_cd = dot(dp, cd), dp_bd = dot(dp, bd), dp_ad = dot(dp, ad);
    if (dp_ad >= 0.f && dp_bd >= 0.f && dp_cd >= 0.f) {
        loc = mkloc(LOC_VERTEX, 3);
        return d;
    }
    // the six edge regions, in the reference's order; the first that holds gives the operands of the shared tail
    V3 nabc = cross(ab, ac), nabd = cross(ab, ad), nacd, nbcd;
    float dabc, dabd, dacd, dacb, dadb, dadc, dbca, dbcd, dbdc, dbda, dcda, dcdb;
    int edge = -1;
    V3 from = a, dir = ab;
    float num = ap_ab, den = ap_ab - bp_ab;
    if (tetra_edge_test(nabc, nabd, ap, ab, ap_ab, bp_ab, dabc, dabd)) {
        edge = 0;
    } else {
        nacd = cross(ac, ad);
        if (tetra_edge_test(nacd, -nabc, ap, ac, ap_ac, cp_ac, dacd, dacb)) {
            edge = 1, dir = ac, num = ap_ac, den = ap_ac - cp_ac;
        } else if (tetra_edge_test(-nabd, -nacd, ap, ad, ap_ad, dp_ad, dadb, dadc)) {
            edge = 2, dir = ad, num = ap_ad, den = ap_ad - dp_ad;
        } else {
            nbcd = cross(bc, bd);
            if (tetra_edge_test(nabc, nbcd, bp, bc, bp_bc, cp_bc, dbca, dbcd))
                edge = 3, from = b, dir = bc, num = bp_bc, den = bp_bc - cp_bc;
            else if (tetra_edge_test(-nbcd, nabd, bp, bd, bp_bd, dp_bd, dbdc, dbda))
                edge = 4, from = b, dir = bd, num = bp_bd, den = bp_bd - dp_bd;
            else if (tetra_edge_test(nacd, nbcd, cp, cd, cp_cd, dp_cd, dcda, dcdb))
                edge = 5, from = c, dir = cd, num = cp_cd, den = cp_cd - dp_cd;
        }
    }
    if (edge >= 0) {
        float u = num / den;
        loc = mkloc(LOC_EDGE, edge, 1.f - u, u);
        return from + dir * u;
    }
    // the four face regions: a face whose normal is too short to normalise is skipped, like the reference's `return false`
    for (int first = 0; first < 4;) {
        int face = -1;
        V3 n, fa = a, fb = b, fc = c, fap = ap, fbp = bp, fcp = cp;
        if (first <= 0 && tetra_face_test(ap, ab, ac, ad, dabc, dbca, dacb, n)) {
            face = 0;
        } else if (first <= 1 && tetra_face_test(ap, ab, ad, ac, dadb, dabd, dbda, n)) {
            face = 1, fc = d, fcp = dp;
        } else if (first <= 2 && tetra_face_test(ap, ac, ad, ab, dacd, dcda, dadc, n)) {
            face = 2, fb = c, fc = d, fbp = cp, fcp = dp;
        } else if (first <= 3 && tetra_face_test(bp, bc, bd, -ab, dbcd, dcdb, dbdc, n)) {
            face = 3, fa = b, fb = c, fc = d, fap = bp, fbp = cp, fcp = dp;
        }
        if (face < 0) break;
        V3 normal;
        if (try_normalize(n, NCB_EPS, normal)) {
            float vc = dot(normal, cross(fap, fbp));
            float va = dot(normal, cross(fbp, fcp));
            float vb = dot(normal, cross(fcp, fap));
            float denom = va + vb + vc;
            float inv_denom = 1.f / denom;
            float b0 = va * inv_denom, b1 = vb * inv_denom, b2 = vc * inv_denom;
            loc = mkloc(LOC_FACE, face, b0, b1, b2);
            return fa * b0 + fb * b1 + fc * b2;
        }
        first = face + 1;
    }
    loc = mkloc(LOC_SOLID, 0);
    return p;
}

// ---- VoronoiSimplex -----------------------------------------------------------------------------------------
struct Simplex {
    CSOPoint v[4];
    float proj[3], prev_proj[3];
    int prev_vertices[4];
    int dim, prev_dim;
};

NCB_HD void simplex_init(Simplex& s, const CSOPoint& p) {
    // VoronoiSimplex::new() + reset(p)
    CSOPoint o;
    o.point = o.orig1 = o.orig2 = v3(0.f, 0.f, 0.f);
    s.v[1] = s.v[2] = s.v[3] = o;
    s.v[0] = p;
    for (int i = 0; i < 3; ++i) s.proj[i] = s.prev_proj[i] = 0.f;
    for (int i = 0; i < 4; ++i) s.prev_vertices[i] = i;
    s.dim = 0;
    s.prev_dim = 0;
}
NCB_HD void simplex_swap(Simplex& s, int a, int b) {
    CSOPoint t = s.v[a];
    s.v[a] = s.v[b];
    s.v[b] = t;
    int u = s.prev_vertices[a];
    s.prev_vertices[a] = s.prev_vertices[b];
    s.prev_vertices[b] = u;
}
NCB_HD bool simplex_add_point(Simplex& s, const CSOPoint& pt) {
    const float eps_tol = NCB_EPS * 10.0f;
    s.prev_dim = s.dim;
    for (int i = 0; i < 3; ++i) s.prev_proj[i] = s.proj[i];
    for (int i = 0; i < 4; ++i) s.prev_vertices[i] = i;
    if (s.dim == 0) {
        if (norm_squared(s.v[0].point - pt.point) < eps_tol) return false;
    } else if (s.dim == 1) {
        V3 ab = s.v[1].point - s.v[0].point, ac = pt.point - s.v[0].point;
        if (norm_squared(cross(ab, ac)) < eps_tol) return false;
    } else {
        V3 ab = s.v[1].point - s.v[0].point, ac = s.v[2].point - s.v[0].point, ap = pt.point - s.v[0].point;
        V3 n = normalize(cross(ab, ac));
        if (fabsf(dot(n, ap)) < eps_tol) return false;
    }
    s.dim += 1;
    s.v[s.dim] = pt;
    return true;
}
static __device__ __noinline__ V3 simplex_project_origin_and_reduce(Simplex& s) {
    const V3 O = v3(0.f, 0.f, 0.f);
    Loc loc;
    if (s.dim == 0) {
        s.proj[0] = 1.f;
        return s.v[0].point;
    }
    if (s.dim == 1) {
        V3 p = proj_segment(s.v[0].point, s.v[1].point, O, loc);
        if (loc.kind == LOC_VERTEX) {
            if (loc.id == 1) simplex_swap(s, 0, 1);
            s.proj[0] = 1.f;
            s.dim = 0;
        } else {
            s.proj[0] = loc.b0;
            s.proj[1] = loc.b1;
        }
        return p;
    }
    if (s.dim == 2) {
        V3 p = proj_triangle(s.v[0].point, s.v[1].point, s.v[2].point, O, loc);
        if (loc.kind == LOC_VERTEX) {
            simplex_swap(s, 0, loc.id);
            s.proj[0] = 1.f;
            s.dim = 0;
        } else if (loc.kind == LOC_EDGE) {
            if (loc.id == 0) {
                s.proj[0] = loc.b0, s.proj[1] = loc.b1;
            } else if (loc.id == 1) {
                simplex_swap(s, 0, 2);
                s.proj[0] = loc.b1, s.proj[1] = loc.b0;
            } else {
                simplex_swap(s, 1, 2);
                s.proj[0] = loc.b0, s.proj[1] = loc.b1;
            }
            s.dim = 1;
        } else if (loc.kind == LOC_FACE) {
            s.proj[0] = loc.b0, s.proj[1] = loc.b1, s.proj[2] = loc.b2;
        }
        return p;
    }
    V3 p = proj_tetrahedron(s.v[0].point, s.v[1].point, s.v[2].point, s.v[3].point, O, loc);
    if (loc.kind == LOC_VERTEX) {
        simplex_swap(s, 0, loc.id);
        s.proj[0] = 1.f;
        s.dim = 0;
    } else if (loc.kind == LOC_EDGE) {
        switch (loc.id) {
            case 0: break;
            case 1: simplex_swap(s, 1, 2); break;
            case 2: simplex_swap(s, 1, 3); break;
            case 3: simplex_swap(s, 0, 2); break;
            case 4: simplex_swap(s, 0, 3); break;
            default:
                simplex_swap(s, 0, 2);
                simplex_swap(s, 1, 3);
                break;
        }
        if (loc.id == 3 || loc.id == 4) {
            s.proj[0] = loc.b1, s.proj[1] = loc.b0;
        } else {
            s.proj[0] = loc.b0, s.proj[1] = loc.b1;
        }
        s.dim = 1;
    } else if (loc.kind == LOC_FACE) {
        if (loc.id == 0) {
            s.proj[0] = loc.b0, s.proj[1] = loc.b1, s.proj[2] = loc.b2;
        } else if (loc.id == 1) {
            s.v[2] = s.v[3];
            s.proj[0] = loc.b0, s.proj[1] = loc.b1, s.proj[2] = loc.b2;
        } else if (loc.id == 2) {
            s.v[1] = s.v[3];
            s.proj[0] = loc.b0, s.proj[1] = loc.b2, s.proj[2] = loc.b1;
        } else {
            s.v[0] = s.v[3];
            s.proj[0] = loc.b2, s.proj[1] = loc.b0, s.proj[2] = loc.b1;
        }
        s.dim = 2;
    }
    return p;
}

// gjk.rs:367-388
NCB_HD void gjk_result(const Simplex& s, bool prev, V3& p1, V3& p2) {
    V3 r0 = v3(0.f, 0.f, 0.f), r1 = v3(0.f, 0.f, 0.f);
    if (prev) {
        for (int i = 0; i < s.prev_dim + 1; ++i) {
            float coord = s.prev_proj[i];
            const CSOPoint& pt = s.v[s.prev_vertices[i]];
            r0 = r0 + pt.orig1 * coord;
            r1 = r1 + pt.orig2 * coord;
        }
    } else {
        for (int i = 0; i < s.dim + 1; ++i) {
            float coord = s.proj[i];
            r0 = r0 + s.v[i].orig1 * coord;
            r1 = r1 + s.v[i].orig2 * coord;
        }
    }
    p1 = r0;
    p2 = r1;
}

enum { GJK_INTERSECTION = 0, GJK_CLOSEST_POINTS = 1, GJK_NO_INTERSECTION = 3 };
#ifdef NCB_EPA_STATS  // host-side work statistics only (tests/host_shim/gjk_host.cpp, scripts/epa_work_stats.py)
static int g_gjk_support_evals = 0;
#define NCB_GJK_COUNT() (++g_gjk_support_evals)
#else
#define NCB_GJK_COUNT()
#endif

// gjk::closest_points with exact_dist = true, preceded by the caller's `simplex.reset(CSOPoint::from_shapes(.., init_dir))`
// (contact_support_map_support_map.rs:52-63): the first support evaluation shares the loop's code (one copy of the two support
// maps in the instruction stream instead of two) and the three ClosestPoints exits share one witness computation.
// G: Support, or SupportS (kind, half extents | vertex array: all a support evaluation reads; k_cc_gjk keeps its operands that slim).
template <class G>
static __device__ __noinline__ int gjk_closest_points(const Iso& m1, const G& g1, const Iso& m2, const G& g2, float max_dist,
                                                      V3 init_dir, Simplex& s, V3& p1, V3& p2, V3& out_dir) {
    const float eps_tol = NCB_EPS * 10.0f;
    const float eps_rel = sqrtf(eps_tol);
    V3 proj = v3(0.f, 0.f, 0.f), old_dir = v3(0.f, 0.f, 0.f);
    float max_bound = NCB_FMAX;
    V3 dir = init_dir;
    int niter = 0;
    bool first = true;
    bool res_prev = false;
    for (;;) {
        if (!first) {
            float old_max_bound = max_bound;
            float dist;
            if (!unit_try_new_and_get(-proj, eps_tol, dir, dist)) return GJK_INTERSECTION;
            max_bound = dist;
            if (max_bound >= old_max_bound) {
                res_prev = true;
                out_dir = old_dir;
                break;
            }
        }
        CSOPoint cso = cso_from_shapes(m1, g1, m2, g2, dir);
        NCB_GJK_COUNT();
        if (first) {
            simplex_init(s, cso);
            proj = simplex_project_origin_and_reduce(s);
            V3 pd;
            if (!unit_try_new(proj, 0.f, pd)) return GJK_INTERSECTION;
            old_dir = -pd;
            first = false;
            continue;
        }
        float min_bound = -dot(dir, cso.point);
        if (min_bound > max_dist) {
            out_dir = dir;
            return GJK_NO_INTERSECTION;
        } else if (max_bound - min_bound <= eps_rel * max_bound) {
            out_dir = dir;
            break;
        }
        if (!simplex_add_point(s, cso)) {
            out_dir = dir;
            break;
        }
        old_dir = dir;
        proj = simplex_project_origin_and_reduce(s);
        if (s.dim == 3) {
            if (min_bound >= eps_tol) {
                res_prev = true;
                out_dir = old_dir;
                break;
            }
            return GJK_INTERSECTION;
        }
        niter += 1;
        if (niter == 10000) {
            out_dir = v3(1.f, 0.f, 0.f);
            return GJK_NO_INTERSECTION;
        }
    }
    gjk_result(s, res_prev, p1, p2);
    return GJK_CLOSEST_POINTS;
}

}  // namespace ncb

#include "epa.cuh"
