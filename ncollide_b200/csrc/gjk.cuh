// Device GJK / EPA for one pair per thread, fixed capacity, no recursion, no heap allocation.
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

#define EPA_MAX_VERTS 48
#define EPA_MAX_FACES 192
#define EPA_MAX_HEAP 160
#define EPA_MAX_STACK 128

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
};

NCB_HD V3 local_support_point(const Support& g, V3 dir) {
    if (g.kind == 0) return v3(copysignf(g.he.x, dir.x), copysignf(g.he.y, dir.y), copysignf(g.he.z, dir.z));
    uint32_t best = 0;
    float best_dot = dot(g.hull.pt(0), dir);
    for (uint32_t i = 1; i < g.hull.nv; ++i) {
        float d = dot(g.hull.pt(i), dir);
        if (d > best_dot) {
            best_dot = d;
            best = i;
        }
    }
    return g.hull.pt(best);
}
NCB_HD V3 support_point(const Support& g, const Iso& m, V3 dir) {
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
NCB_HD CSOPoint cso_from_shapes(const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, V3 dir) {
    CSOPoint c;
    bool swap = g1.kind == 1 && g2.kind != 1;
    const Support& ga = swap ? g2 : g1;
    const Support& gb = swap ? g1 : g2;
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

// solid = true variant only (the one the simplex and EPA use)
static __device__ __noinline__ V3 proj_triangle(V3 a, V3 b, V3 c, V3 p, Loc& loc) {
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
        return a + ab * v;
    }
    float vb = -dot(n, cross(ac, cp));
    if (vb < 0.f && ac_ap >= 0.f && ac_cp <= 0.f) {
        float w = ac_ap / norm_squared(ac);
        loc = mkloc(LOC_EDGE, 2, 1.f - w, w);
        return a + ac * w;
    }
    float va = dot(n, cross(bc, bp));
    if (va < 0.f && ac_bp - ab_bp >= 0.f && ab_cp - ac_cp >= 0.f) {
        float w = dot(bc, bp) / norm_squared(bc);
        loc = mkloc(LOC_EDGE, 1, 1.f - w, w);
        return b + bc * w;
    }
    int clockwise = dot(n, ap) >= 0.f ? 0 : 1;
    if (va + vb + vc != 0.f) {
        float denom = 1.f / (va + vb + vc);
        float v = vb * denom, w = vc * denom;
        loc = mkloc(LOC_FACE, clockwise, 1.f - v - w, v, w);
        return a + ab * v + ac * w;
    }
    loc = mkloc(LOC_SOLID, 0);
    return p;
}

NCB_HD bool tetra_edge(int i, V3 a, V3 nabc, V3 nabd, V3 ap, V3 ab, float ap_ab, float bp_ab, float& dabc, float& dabd, V3& proj,
                       Loc& loc) {
    float ab_ab = ap_ab - bp_ab;
    V3 ap_x_ab = cross(ap, ab);
    dabc = dot(ap_x_ab, nabc);
    dabd = dot(ap_x_ab, nabd);
    if (ab_ab != 0.f && dabc >= 0.f && dabd >= 0.f && ap_ab >= 0.f && ap_ab <= ab_ab) {
        float u = ap_ab / ab_ab;
        loc = mkloc(LOC_EDGE, i, 1.f - u, u);
        proj = a + ab * u;
        return true;
    }
    return false;
}
NCB_HD bool tetra_face(int i, V3 a, V3 b, V3 c, V3 ap, V3 bp, V3 cp, V3 ab, V3 ac, V3 ad, float dabc, float dbca, float dacb, V3& proj,
                       Loc& loc) {
    if (dabc < 0.f && dbca < 0.f && dacb < 0.f) {
        V3 n = cross(ab, ac);
        if (dot(n, ad) * dot(n, ap) < 0.f) {
            V3 normal;
            if (!try_normalize(n, NCB_EPS, normal)) return false;
            float vc = dot(normal, cross(ap, bp));
            float va = dot(normal, cross(bp, cp));
            float vb = dot(normal, cross(cp, ap));
            float denom = va + vb + vc;
            float inv_denom = 1.f / denom;
            float b0 = va * inv_denom, b1 = vb * inv_denom, b2 = vc * inv_denom;
            loc = mkloc(LOC_FACE, i, b0, b1, b2);
            proj = a * b0 + b * b1 + c * b2;
            return true;
        }
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
    V3 proj;
    V3 nabc = cross(ab, ac), nabd = cross(ab, ad);
    float dabc, dabd;
    if (tetra_edge(0, a, nabc, nabd, ap, ab, ap_ab, bp_ab, dabc, dabd, proj, loc)) return proj;
    V3 nacd = cross(ac, ad);
    float dacd, dacb;
    if (tetra_edge(1, a, nacd, -nabc, ap, ac, ap_ac, cp_ac, dacd, dacb, proj, loc)) return proj;
    float dadb, dadc;
    if (tetra_edge(2, a, -nabd, -nacd, ap, ad, ap_ad, dp_ad, dadb, dadc, proj, loc)) return proj;
    V3 nbcd = cross(bc, bd);
    float dbca, dbcd;
    if (tetra_edge(3, b, nabc, nbcd, bp, bc, bp_bc, cp_bc, dbca, dbcd, proj, loc)) return proj;
    float dbdc, dbda;
    if (tetra_edge(4, b, -nbcd, nabd, bp, bd, bp_bd, dp_bd, dbdc, dbda, proj, loc)) return proj;
    float dcda, dcdb;
    if (tetra_edge(5, c, nacd, nbcd, cp, cd, cp_cd, dp_cd, dcda, dcdb, proj, loc)) return proj;
    if (tetra_face(0, a, b, c, ap, bp, cp, ab, ac, ad, dabc, dbca, dacb, proj, loc)) return proj;
    if (tetra_face(1, a, b, d, ap, bp, dp, ab, ad, ac, dadb, dabd, dbda, proj, loc)) return proj;
    if (tetra_face(2, a, c, d, ap, cp, dp, ac, ad, ab, dacd, dcda, dadc, proj, loc)) return proj;
    if (tetra_face(3, b, c, d, bp, cp, dp, bc, bd, -ab, dbcd, dcdb, dbdc, proj, loc)) return proj;
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

// gjk::closest_points with exact_dist = true, preceded by the caller's `simplex.reset(CSOPoint::from_shapes(.., init_dir))`
// (contact_support_map_support_map.rs:52-63): the first support evaluation shares the loop's code (one copy of the two support
// maps in the instruction stream instead of two) and the three ClosestPoints exits share one witness computation.
static __device__ __noinline__ int gjk_closest_points(const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, float max_dist,
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

// ---- EPA ----------------------------------------------------------------------------------------------------
// Per-thread polytope in local memory, split hot / cold so that the expansion loop touches as few bytes as possible:
//   hot : vertex CSO points, packed face topology (3 vertex ids + deleted flag | 3 neighbour ids), face normals, heap
//   cold: the original support points (orig1 / orig2) of each vertex, read once for the result
// Barycentric coordinates of a face are NOT stored: they are recomputed (same inputs, same arithmetic, same bits) for
// the one face the result is read from.
struct EpaHeapItem {
    uint32_t id;
    float neg_dist;
};
struct EpaState {
    V3 vpoint[EPA_MAX_VERTS];
    uint32_t ftopo[EPA_MAX_FACES][2];  // [0] = pts0 | pts1 << 8 | pts2 << 16 | deleted << 24 ; [1] = adj0 | adj1 << 8 | adj2 << 16
    V3 fnormal[EPA_MAX_FACES];
    float hdist[EPA_MAX_HEAP];
    uint8_t hid[EPA_MAX_HEAP];
    uint8_t sil_face[EPA_MAX_STACK], sil_opp[EPA_MAX_STACK];
    uint8_t stk_face[EPA_MAX_STACK], stk_opp[EPA_MAX_STACK];
    float hpend[EPA_MAX_STACK];  // -dist of the faces created in this turn, pushed on the heap after the loop
    V3 vorig1[EPA_MAX_VERTS], vorig2[EPA_MAX_VERTS];
    int nverts, nfaces, nheap, nsil, niter;
    float max_dist;
    EpaHeapItem best_face_id;
    bool overflow, panicked;
};
static_assert(EPA_MAX_FACES <= 255 && EPA_MAX_VERTS <= 255, "ids are packed in 8 bits");

NCB_HD uint32_t f_pt(const EpaState& e, uint32_t f, uint32_t k) { return (e.ftopo[f][0] >> (8 * k)) & 0xffu; }
NCB_HD uint32_t f_adj(const EpaState& e, uint32_t f, uint32_t k) { return (e.ftopo[f][1] >> (8 * k)) & 0xffu; }
NCB_HD bool f_deleted(const EpaState& e, uint32_t f) { return (e.ftopo[f][0] >> 24) != 0; }
NCB_HD void f_set_deleted(EpaState& e, uint32_t f) { e.ftopo[f][0] |= 0x01000000u; }
NCB_HD void f_set_adj(EpaState& e, uint32_t f, uint32_t k, uint32_t v) {
    e.ftopo[f][1] = (e.ftopo[f][1] & ~(0xffu << (8 * k))) | (v << (8 * k));
}
NCB_HD void epa_push_vertex(EpaState& e, const CSOPoint& c) {
    e.vpoint[e.nverts] = c.point;
    e.vorig1[e.nverts] = c.orig1;
    e.vorig2[e.nverts] = c.orig2;
    e.nverts++;
}

// Rust std BinaryHeap<FaceId>: `<=` comes from partial_cmp on neg_dist.
NCB_HD void heap_sift_up(EpaState& e, int start, int pos) {
    float ed = e.hdist[pos];
    uint8_t ei = e.hid[pos];
    while (pos > start) {
        int parent = (pos - 1) / 2;
        if (ed <= e.hdist[parent]) break;
        e.hdist[pos] = e.hdist[parent];
        e.hid[pos] = e.hid[parent];
        pos = parent;
    }
    e.hdist[pos] = ed;
    e.hid[pos] = ei;
}
NCB_HD void heap_push(EpaState& e, uint32_t id, float nd) {
    if (e.nheap >= EPA_MAX_HEAP) {
        e.overflow = true;
        return;
    }
    e.hid[e.nheap] = (uint8_t)id;
    e.hdist[e.nheap] = nd;
    e.nheap++;
    heap_sift_up(e, 0, e.nheap - 1);
}
NCB_HD bool heap_pop(EpaState& e, EpaHeapItem& out) {
    if (e.nheap == 0) return false;
    --e.nheap;
    float item_d = e.hdist[e.nheap];
    uint8_t item_i = e.hid[e.nheap];
    if (e.nheap > 0) {
        float td = e.hdist[0];
        uint8_t ti = e.hid[0];
        // swap(item, data[0]); sift_down_to_bottom(0)
        int end = e.nheap, pos = 0, child = 1;
        float ed = item_d;
        uint8_t ei = item_i;
        item_d = td;
        item_i = ti;
        while (end >= 2 && child <= end - 2) {
            if (e.hdist[child] <= e.hdist[child + 1]) child += 1;
            e.hdist[pos] = e.hdist[child];
            e.hid[pos] = e.hid[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            e.hdist[pos] = e.hdist[child];
            e.hid[pos] = e.hid[child];
            pos = child;
        }
        e.hdist[pos] = ed;
        e.hid[pos] = ei;
        heap_sift_up(e, 0, pos);
    }
    out.id = item_i;
    out.neg_dist = item_d;
    return true;
}

// Face::new (epa3.rs:93-114): normal + "projection of the origin lies inside the face".  false on overflow.
static __device__ __noinline__ bool epa_face_new(EpaState& e, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t a0, uint32_t a1, uint32_t a2,
                                          bool& proj_inside) {
    if (e.nfaces >= EPA_MAX_FACES) {
        e.overflow = true;
        return false;
    }
    V3 A = e.vpoint[p0], B = e.vpoint[p1], C = e.vpoint[p2];
    Loc loc;
    proj_triangle(A, B, C, v3(0.f, 0.f, 0.f), loc);
    int f = e.nfaces++;
    e.ftopo[f][0] = p0 | (p1 << 8) | (p2 << 16);
    e.ftopo[f][1] = a0 | (a1 << 8) | (a2 << 16);
    V3 n;
    if (!unit_try_new(cross(B - A, C - A), NCB_EPS, n)) n = v3(0.f, 0.f, 0.f);  // utils::ccw_face_normal
    e.fnormal[f] = n;
    proj_inside = loc.kind == LOC_FACE;
    return true;
}
// Face::closest_points (epa3.rs:116-126) with the barycentric coordinates recomputed as Face::new computed them.
static __device__ __noinline__ void epa_face_closest_points(const EpaState& e, uint32_t f, V3& p1, V3& p2) {
    uint32_t i0 = f_pt(e, f, 0), i1 = f_pt(e, f, 1), i2 = f_pt(e, f, 2);
    Loc loc;
    proj_triangle(e.vpoint[i0], e.vpoint[i1], e.vpoint[i2], v3(0.f, 0.f, 0.f), loc);
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (loc.kind == LOC_FACE) b0 = loc.b0, b1 = loc.b1, b2 = loc.b2;
    p1 = e.vorig1[i0] * b0 + e.vorig1[i1] * b1 + e.vorig1[i2] * b2;
    p2 = e.vorig2[i0] * b0 + e.vorig2[i1] * b1 + e.vorig2[i2] * b2;
}
NCB_HD uint32_t epa_next_ccw(EpaState& e, uint32_t f, uint32_t id) {
    uint32_t t = e.ftopo[f][0];
    if ((t & 0xffu) == id) return 1;
    if (((t >> 8) & 0xffu) == id) return 2;
    if (((t >> 16) & 0xffu) != id) e.panicked = true;  // assert_eq! in the reference
    return 0;
}
NCB_HD bool epa_can_be_seen_by(const EpaState& e, uint32_t f, uint32_t point, uint32_t opp) {
    V3 p0 = e.vpoint[f_pt(e, f, opp)];
    V3 pt = e.vpoint[point];
    if (dot(pt - p0, e.fnormal[f]) >= -(NCB_EPS * 10.0f)) return true;
    V3 p1 = e.vpoint[f_pt(e, f, (opp + 1) % 3)], p2 = e.vpoint[f_pt(e, f, (opp + 2) % 3)];
    // utils::is_affinely_dependent_triangle(p1, p2, pt)
    V3 p1p2 = p2 - p1, p1p3 = pt - p1;
    float eps_tol = NCB_EPS * 100.0f;
    return relative_eq(norm_squared(cross(p1p2, p1p3)), 0.f, eps_tol * eps_tol);
}
// compute_silhouette (epa3.rs:432-454): the recursion becomes a LIFO of (face, opp) visits in the same order.
static __device__ __noinline__ void epa_compute_silhouette3(EpaState& e, uint32_t point, uint32_t id0, uint32_t opp0, uint32_t id1, uint32_t opp1,
                                                     uint32_t id2, uint32_t opp2) {
    int sp = 0;
    e.stk_face[sp] = (uint8_t)id2, e.stk_opp[sp] = (uint8_t)opp2, sp++;
    e.stk_face[sp] = (uint8_t)id1, e.stk_opp[sp] = (uint8_t)opp1, sp++;
    e.stk_face[sp] = (uint8_t)id0, e.stk_opp[sp] = (uint8_t)opp0, sp++;
    while (sp > 0) {
        sp--;
        uint32_t id = e.stk_face[sp], opp = e.stk_opp[sp];
        if (f_deleted(e, id)) continue;
        if (!epa_can_be_seen_by(e, id, point, opp)) {
            if (e.nsil >= EPA_MAX_STACK) {
                e.overflow = true;
                return;
            }
            e.sil_face[e.nsil] = (uint8_t)id, e.sil_opp[e.nsil] = (uint8_t)opp, e.nsil++;
        } else {
            f_set_deleted(e, id);
            uint32_t adj_pt_id1 = (opp + 2) % 3, adj_pt_id2 = opp;
            uint32_t adj1 = f_adj(e, id, adj_pt_id1), adj2 = f_adj(e, id, adj_pt_id2);
            uint32_t o1 = epa_next_ccw(e, adj1, f_pt(e, id, adj_pt_id1));
            uint32_t o2 = epa_next_ccw(e, adj2, f_pt(e, id, adj_pt_id2));
            if (e.panicked) return;
            if (sp + 2 > EPA_MAX_STACK) {
                e.overflow = true;
                return;
            }
            // visit adj1 first, then adj2
            e.stk_face[sp] = (uint8_t)adj2, e.stk_opp[sp] = (uint8_t)o2, sp++;
            e.stk_face[sp] = (uint8_t)adj1, e.stk_opp[sp] = (uint8_t)o1, sp++;
        }
    }
}

enum { EPA_CONTINUE = 0, EPA_DONE_OK = 1, EPA_DONE_FAIL = 2 };

#define NCB_EPA_PUSH(ID, ND)                              \
    {                                                     \
        float nd__ = (ND);                                \
        if (nd__ > NCB_EPS * 10.0f) return EPA_DONE_FAIL; \
        heap_push(e, (ID), nd__);                         \
    }

// EPA::closest_points, part 1 (epa3.rs:219-328): initial polytope from the GJK simplex.
static __device__ __noinline__ int epa_init(EpaState& e, const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, int sdim,
                                     const CSOPoint* sv, V3& out1, V3& out2, V3& out_n) {
    e.nverts = e.nfaces = e.nheap = e.nsil = 0;
    e.niter = 0;
    e.overflow = false;
    e.panicked = false;
    for (int i = 0; i < sdim + 1; ++i) epa_push_vertex(e, sv[i]);
    if (sdim == 0) {
        out1 = v3(0.f, 0.f, 0.f);
        out2 = v3(0.f, 0.f, 0.f);
        out_n = v3(0.f, 1.f, 0.f);
        return EPA_DONE_OK;
    } else if (sdim == 3) {
        V3 dp1 = e.vpoint[1] - e.vpoint[0];
        V3 dp2 = e.vpoint[2] - e.vpoint[0];
        V3 dp3 = e.vpoint[3] - e.vpoint[0];
        if (dot(cross(dp1, dp2), dp3) > 0.f) {
            V3 t = e.vpoint[1];
            e.vpoint[1] = e.vpoint[2];
            e.vpoint[2] = t;
            t = e.vorig1[1], e.vorig1[1] = e.vorig1[2], e.vorig1[2] = t;
            t = e.vorig2[1], e.vorig2[1] = e.vorig2[2], e.vorig2[2] = t;
        }
        bool in1, in2, in3, in4;
        epa_face_new(e, 0, 1, 2, 3, 1, 2, in1);
        epa_face_new(e, 1, 3, 2, 3, 2, 0, in2);
        epa_face_new(e, 0, 2, 3, 0, 1, 3, in3);
        epa_face_new(e, 0, 3, 1, 2, 1, 0, in4);
        if (in1) NCB_EPA_PUSH(0, -dot(e.fnormal[0], e.vpoint[0]));
        if (in2) NCB_EPA_PUSH(1, -dot(e.fnormal[1], e.vpoint[1]));
        if (in3) NCB_EPA_PUSH(2, -dot(e.fnormal[2], e.vpoint[2]));
        if (in4) NCB_EPA_PUSH(3, -dot(e.fnormal[3], e.vpoint[3]));
    } else {
        if (sdim == 1) {
            V3 dpt = e.vpoint[1] - e.vpoint[0];
            V3 first, second;
            orthonormal_basis(dpt, first, second);
            epa_push_vertex(e, cso_from_shapes(m1, g1, m2, g2, first));
        }
        bool in;
        epa_face_new(e, 0, 1, 2, 1, 1, 1, in);
        epa_face_new(e, 0, 2, 1, 0, 0, 0, in);
        NCB_EPA_PUSH(0, 0.f);
        NCB_EPA_PUSH(1, 0.f);
    }
    e.max_dist = NCB_FMAX;
    if (e.nheap == 0) {  // heap.peek().unwrap() panics in the reference
        e.panicked = true;
        return EPA_DONE_FAIL;
    }
    e.best_face_id.id = e.hid[0];
    e.best_face_id.neg_dist = e.hdist[0];
    return EPA_CONTINUE;
}

// EPA::closest_points, part 2: ONE turn of `while let Some(face_id) = self.heap.pop()` (epa3.rs:330-425).
static __device__ __noinline__ int epa_step(EpaState& e, const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, V3& out1, V3& out2,
                                     V3& out_n) {
    const float eps_tol = NCB_EPS * 100.0f;
    EpaHeapItem face_id;
    // `if face.deleted { continue; }` (epa3.rs:334-336): stale heap entries are skipped inside the same turn
    do {
        if (!heap_pop(e, face_id)) {  // heap exhausted: the best face so far (epa3.rs:427-429)
            epa_face_closest_points(e, e.best_face_id.id, out1, out2);
            out_n = e.fnormal[e.best_face_id.id];
            return EPA_DONE_OK;
        }
    } while (f_deleted(e, face_id.id));
    uint32_t fid = face_id.id;
    // snapshot of the popped face (the reference clones it before the polytope is edited)
    uint32_t fp0 = f_pt(e, fid, 0), fp1 = f_pt(e, fid, 1), fp2 = f_pt(e, fid, 2);
    uint32_t fa0 = f_adj(e, fid, 0), fa1 = f_adj(e, fid, 1), fa2 = f_adj(e, fid, 2);
    V3 fnorm = e.fnormal[fid];
    if (e.nverts >= EPA_MAX_VERTS) {
        e.overflow = true;
        return EPA_DONE_FAIL;
    }
    CSOPoint cso = cso_from_shapes(m1, g1, m2, g2, fnorm);
    uint32_t support_point_id = (uint32_t)e.nverts;
    epa_push_vertex(e, cso);
    float candidate_max_dist = dot(cso.point, fnorm);
    if (candidate_max_dist < e.max_dist) {
        e.best_face_id = face_id;
        e.max_dist = candidate_max_dist;
    }
    float curr_dist = -face_id.neg_dist;
    if (e.max_dist - curr_dist < eps_tol) {
        epa_face_closest_points(e, e.best_face_id.id, out1, out2);
        out_n = e.fnormal[e.best_face_id.id];
        return EPA_DONE_OK;
    }
    f_set_deleted(e, fid);
    uint32_t o1 = epa_next_ccw(e, fa0, fp0);
    uint32_t o2 = epa_next_ccw(e, fa1, fp1);
    uint32_t o3 = epa_next_ccw(e, fa2, fp2);
    if (e.panicked) return EPA_DONE_FAIL;
    // compute_silhouette x3 (epa3.rs:364-366) as ONE LIFO walk: the three roots are stacked in reverse order, so the
    // flood from adj[0] completes before adj[1] is looked at, exactly like the three sequential recursive calls
    epa_compute_silhouette3(e, support_point_id, fa0, o1, fa1, o2, fa2, o3);
    if (e.panicked || e.overflow) return EPA_DONE_FAIL;
    uint32_t first_new_face_id = (uint32_t)e.nfaces;
    if (e.nsil == 0) return EPA_DONE_FAIL;
    int npend = 0;
    for (int k = 0; k < e.nsil; ++k) {
        uint32_t efid = e.sil_face[k], eopp = e.sil_opp[k];
        if (!f_deleted(e, efid)) {
            uint32_t new_face_id = (uint32_t)e.nfaces;
            uint32_t pt_id1 = f_pt(e, efid, (eopp + 2) % 3);
            uint32_t pt_id2 = f_pt(e, efid, (eopp + 1) % 3);
            bool inside;
            // adj = [edge.face_id, new_face_id + 1, new_face_id - 1] (the last two are patched below for the ends)
            if (!epa_face_new(e, pt_id1, pt_id2, support_point_id, efid, (new_face_id + 1) & 0xffu, (new_face_id - 1) & 0xffu, inside))
                return EPA_DONE_FAIL;
            f_set_adj(e, efid, (eopp + 1) % 3, new_face_id);
            if (inside) {
                V3 pt = e.vpoint[f_pt(e, new_face_id, 0)];
                float dist = dot(e.fnormal[new_face_id], pt);
                if (dist < curr_dist) {
                    // the popped face as it was when cloned (epa3.rs:393-398)
                    // its topology words are unchanged except the deleted flag, which closest_points does not read
                    epa_face_closest_points(e, fid, out1, out2);
                    out_n = fnorm;
                    return EPA_DONE_OK;
                }
                // FaceId::new(new_face_id, -dist)? then heap.push: the validity test stays here, in order; the sift
                // itself is deferred to one converged loop below (pushes commute with nothing else in this loop)
                if (-dist > NCB_EPS * 10.0f) return EPA_DONE_FAIL;
                e.stk_face[npend] = (uint8_t)new_face_id;  // the DFS stack is free at this point: reuse it
                e.hpend[npend] = -dist;
                npend++;
            }
        }
    }
    for (int k = 0; k < npend; ++k) heap_push(e, e.stk_face[k], e.hpend[k]);
    if (e.overflow) return EPA_DONE_FAIL;
    if (first_new_face_id == (uint32_t)e.nfaces) return EPA_DONE_FAIL;
    f_set_adj(e, first_new_face_id, 2, (uint32_t)(e.nfaces - 1));
    f_set_adj(e, (uint32_t)(e.nfaces - 1), 1, first_new_face_id);
    e.nsil = 0;
    e.niter += 1;
    if (e.niter > 10000) return EPA_DONE_FAIL;
    return EPA_CONTINUE;
}
#undef NCB_EPA_PUSH

// EPA::closest_points (epa3.rs:219-430).  false = None (also on capacity overflow, flagged in e.overflow).
static __device__ __noinline__ bool epa_closest_points(EpaState& e, const Iso& m1, const Support& g1, const Iso& m2, const Support& g2,
                                                int sdim, const CSOPoint* sv, V3& out1, V3& out2, V3& out_n) {
    int st = epa_init(e, m1, g1, m2, g2, sdim, sv, out1, out2, out_n);
    while (st == EPA_CONTINUE) st = epa_step(e, m1, g1, m2, g2, out1, out2, out_n);
    return st == EPA_DONE_OK;
}

// contact_support_map_support_map_with_params (init_dir = None: fresh generator).
// Returns GJK_CLOSEST_POINTS / GJK_NO_INTERSECTION.
static __device__ __noinline__ int contact_sm_sm(EpaState& e, const Iso& m1, const Support& g1, const Iso& m2, const Support& g2,
                                          float prediction, V3& p1, V3& p2, V3& dir_out, uint32_t* epa_overflow, uint32_t* ref_panics) {
    V3 dir;
    if (!unit_try_new(m2.t - m1.t, NCB_EPS, dir)) dir = v3(1.f, 0.f, 0.f);
    Simplex s;
    int r = gjk_closest_points(m1, g1, m2, g2, prediction, dir, s, p1, p2, dir_out);
    if (r != GJK_INTERSECTION) return r;
    if (epa_closest_points(e, m1, g1, m2, g2, s.dim, s.v, p1, p2, dir_out)) return GJK_CLOSEST_POINTS;
    if (e.overflow) atomicAdd(epa_overflow, 1u);
    if (e.panicked) atomicAdd(ref_panics, 1u);
    dir_out = v3(1.f, 0.f, 0.f);
    return GJK_NO_INTERSECTION;
}

}  // namespace ncb
