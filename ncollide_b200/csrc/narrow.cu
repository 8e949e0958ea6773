// Narrow phase on the device: one pair per thread, one persistent grid-stride kernel per shape-type key segment
// (pairs were counting-sorted by key in broad.cu, so a warp runs a single algorithm).
//
// Replaces (reference, file:line):
//   NarrowPhase::update / update_contact      pipeline/narrow_phase/narrow_phase.rs:56-104,168-197
//   DefaultContactDispatcher                  contact_generator/default_contact_dispatcher.rs:27-97
//   BallBallManifoldGenerator                 contact_generator/ball_ball_manifold_generator.rs:28-67, query/contact/contact_ball_ball.rs:8-38
//   PlaneBallManifoldGenerator                contact_generator/plane_ball_manifold_generator.rs:30-83
//   PlaneConvexPolyhedronManifoldGenerator    contact_generator/plane_convex_polyhedron_manifold_generator.rs:29-81
//   BallConvexPolyhedronManifoldGenerator     contact_generator/ball_convex_polyhedron_manifold_generator.rs:28-122,
//                                             query/point/point_aabb.rs:9-135, query/point/point_support_map.rs:15-53,90-117
//   ConvexPolyhedronConvexPolyhedronManifoldGenerator  contact_generator/convex_polyhedron_convex_polyhedron_manifold_generator.rs:83-167
//   Cuboid / ConvexHull features              shape/cuboid.rs:185-405,502-563, shape/convex.rs:388-540
//   ConvexPolygonalFeature::clip / add_contact_to_manifold   shape/convex_polygonal_feature3.rs:217-401
//   ContactManifold::push (DistanceBased 0.02) query/contact/contact_manifold.rs:165-236
#ifndef NCB_HOST_SHIM  // tests/host_shim compiles the per-pair functions of this file for the host (test infrastructure)
#include <cooperative_groups.h>
#endif
#include <cmath>
#include <cstdlib>
#include "gjk.cuh"
#include "ncb_internal.h"
#include "shapes.cuh"
#include "vec.cuh"

namespace ncb {

#define FEAT_MAX 16   // vertices per polygonal feature
#define MANIFOLD_MAX 32

// ---- ConvexPolygonalFeature (vertices, ids, optional normal); edge normals are never read on this path -------
struct Feature {
    V3 v[FEAT_MAX];
    uint32_t vid[FEAT_MAX], eid[FEAT_MAX];
    int nv, ne;
    bool has_normal;
    V3 normal;
    uint32_t feature_id;
};
NCB_HD void feat_clear(Feature& f) {
    f.nv = 0;
    f.ne = 0;
    f.has_normal = false;
    f.feature_id = FID_UNKNOWN;
}
NCB_HD void feat_push(Feature& f, V3 p, uint32_t id) {
    if (f.nv < FEAT_MAX) {
        f.v[f.nv] = p;
        f.vid[f.nv] = id;
        f.nv++;
    }
}
NCB_HD void feat_push_edge(Feature& f, uint32_t id) {
    if (f.ne < FEAT_MAX) f.eid[f.ne++] = id;
}
NCB_HD void feat_transform(Feature& f, const Iso& m) {
    for (int i = 0; i < f.nv; ++i) f.v[i] = iso_mul_point(m, f.v[i]);
    if (f.has_normal) f.normal = iso_mul_vec(m, f.normal);
}
NCB_HD int feat_nedges(const Feature& f) { return f.nv == 1 ? 0 : (f.nv == 2 ? 1 : f.nv); }

// Cuboid::face (cuboid.rs:185-277, dim3)
__device__ __noinline__ void cuboid_face(V3 he, uint32_t i, Feature& out) {
    feat_clear(out);
    uint32_t i1;
    float sign;
    if (i < 3) {
        i1 = i;
        sign = 1.f;
    } else {
        i1 = i - 3;
        sign = -1.f;
    }
    uint32_t i2 = (i1 + 1) % 3, i3 = (i1 + 2) % 3;
    uint32_t edge_i2 = sign > 0.f ? i2 : i3, edge_i3 = sign > 0.f ? i3 : i2;
    uint32_t mask_i2 = ~(1u << edge_i2), mask_i3 = ~(1u << edge_i3);
    V3 vertex = he;
    vset(vertex, i1, vget(vertex, i1) * sign);
    uint32_t sbit = sign < 0.f ? 1 : 0, msbit = sign < 0.f ? 0 : 1;
    uint32_t vertex_id = sbit << i1;
    feat_push(out, vertex, FID(NCB_FEATURE_VERTEX, vertex_id));
    feat_push_edge(out, FID(NCB_FEATURE_EDGE, edge_i2 | ((vertex_id & mask_i2) << 2)));
    vset(vertex, i2, -sign * vget(he, i2));
    vset(vertex, i3, sign * vget(he, i3));
    vertex_id |= (msbit << i2) | (sbit << i3);
    feat_push(out, vertex, FID(NCB_FEATURE_VERTEX, vertex_id));
    feat_push_edge(out, FID(NCB_FEATURE_EDGE, edge_i3 | ((vertex_id & mask_i3) << 2)));
    vset(vertex, i2, -vget(he, i2));
    vset(vertex, i3, -vget(he, i3));
    vertex_id |= (1u << i2) | (1u << i3);
    feat_push(out, vertex, FID(NCB_FEATURE_VERTEX, vertex_id));
    feat_push_edge(out, FID(NCB_FEATURE_EDGE, edge_i2 | ((vertex_id & mask_i2) << 2)));
    vset(vertex, i2, sign * vget(he, i2));
    vset(vertex, i3, -sign * vget(he, i3));
    vertex_id = (sbit << i1) | (sbit << i2) | (msbit << i3);
    feat_push(out, vertex, FID(NCB_FEATURE_VERTEX, vertex_id));
    feat_push_edge(out, FID(NCB_FEATURE_EDGE, edge_i3 | ((vertex_id & mask_i3) << 2)));
    V3 normal = v3(0.f, 0.f, 0.f);
    vset(normal, i1, sign);
    out.normal = normal;
    out.has_normal = true;
    out.feature_id = sign > 0.f ? FID(NCB_FEATURE_FACE, i1) : FID(NCB_FEATURE_FACE, i1 + 3);
}
// Cuboid::support_face_toward (cuboid.rs:279-307)
NCB_HD void cuboid_support_face_toward(V3 he, const Iso& m, V3 dir, Feature& out) {
    V3 ld = iso_inv_vec(m, dir);
    uint32_t iamax = 0;
    float amax = fabsf(ld.x);
    if (fabsf(ld.y) > amax) {
        amax = fabsf(ld.y);
        iamax = 1;
    }
    if (fabsf(ld.z) > amax) {
        amax = fabsf(ld.z);
        iamax = 2;
    }
    cuboid_face(he, vget(ld, iamax) > 0.f ? iamax : iamax + 3, out);
    feat_transform(out, m);
}
// Cuboid::support_feature_toward (cuboid.rs:309-405, dim3)
__device__ __noinline__ void cuboid_support_feature_toward(V3 he, const Iso& m, V3 dir, float2 ang_cs, Feature& out) {
    V3 ld = iso_inv_vec(m, dir);
    float cang = ang_cs.x, sang = ang_cs.y;
    V3 sp = he;
    feat_clear(out);
    uint32_t sp_id = 0;
    for (uint32_t i1 = 0; i1 < 3; ++i1) {
        float c = vget(ld, i1);
        float sign = signumf(c);
        if (sign * c >= cang) {
            cuboid_face(he, sign > 0.f ? i1 : i1 + 3, out);
            feat_transform(out, m);
            return;
        } else if (sign < 0.f) {
            vset(sp, i1, vget(sp, i1) * sign);
            sp_id |= 1u << i1;
        }
    }
    for (uint32_t i = 0; i < 3; ++i) {
        float c = vget(ld, i);
        float sign = signumf(c);
        if (sign * c <= sang) {
            vset(sp, i, -vget(he, i));
            V3 p1 = sp;
            vset(sp, i, vget(he, i));
            V3 p2 = sp;
            uint32_t p2_id = sp_id & ~(1u << i);
            feat_push(out, iso_mul_point(m, p1), FID(NCB_FEATURE_VERTEX, sp_id | (1u << i)));
            feat_push(out, iso_mul_point(m, p2), FID(NCB_FEATURE_VERTEX, p2_id));
            uint32_t edge_id = FID(NCB_FEATURE_EDGE, i | (p2_id << 2));
            feat_push_edge(out, edge_id);
            out.feature_id = edge_id;
            return;
        }
    }
    feat_push(out, iso_mul_point(m, sp), FID(NCB_FEATURE_VERTEX, sp_id));
    out.feature_id = FID(NCB_FEATURE_VERTEX, sp_id);
}
// Cuboid::feature_normal (cuboid.rs:502-563)
NCB_HD V3 cuboid_feature_normal(uint32_t f) {
    uint32_t id = FID_ID(f);
    V3 dir = v3(0.f, 0.f, 0.f);
    uint32_t kind = FID_KIND(f);
    if (kind == NCB_FEATURE_FACE) {
        if (id < 3)
            vset(dir, id, 1.f);
        else
            vset(dir, id - 3, -1.f);
        return dir;
    }
    if (kind == NCB_FEATURE_EDGE) {
        uint32_t edge = id & 3u, face1 = (edge + 1) % 3, face2 = (edge + 2) % 3, signs = id >> 2;
        vset(dir, face1, (signs & (1u << face1)) ? -1.f : 1.f);
        vset(dir, face2, (signs & (1u << face2)) ? -1.f : 1.f);
        return normalize(dir);
    }
    for (uint32_t i = 0; i < 3; ++i) vset(dir, i, (id & (1u << i)) ? -1.f : 1.f);
    return normalize(dir);
}

// ConvexHull::face (convex.rs:443-461)
NCB_HD void hull_face(const HullView& H, uint32_t id, Feature& out) {
    feat_clear(out);
    uint32_t first = __ldg(H.face_first + id), last = first + __ldg(H.face_num + id);
    for (uint32_t i = first; i < last; ++i) {
        uint32_t vid = __ldg(H.vaf + i), eid = __ldg(H.eaf + i);
        feat_push(out, H.pt(vid), FID(NCB_FEATURE_VERTEX, vid));
        feat_push_edge(out, FID(NCB_FEATURE_EDGE, eid));
    }
    out.normal = H.fn(id);
    out.has_normal = true;
    out.feature_id = FID(NCB_FEATURE_FACE, id);
}
// ConvexHull::support_face_toward (convex.rs:487-509)
__device__ __noinline__ void hull_support_face_toward(const HullView& H, const Iso& m, V3 dir, Feature& out) {
    V3 ls_dir = iso_inv_vec(m, dir);
    uint32_t best = 0;
    float max_dot = dot(H.fn(0), ls_dir);
    for (uint32_t i = 1; i < H.nf; ++i) {
        float d = dot(H.fn(i), ls_dir);
        if (d > max_dot) {
            max_dot = d;
            best = i;
        }
    }
    hull_face(H, best, out);
    feat_transform(out, m);
}
// ConvexHull::support_feature_id_toward_eps (convex.rs:388-415)
__device__ __noinline__ uint32_t hull_support_feature_id_toward_eps(const HullView& H, V3 local_dir, float2 eps_cs) {
    float seps = eps_cs.y, ceps = eps_cs.x;
    uint32_t sp = 0;
    float best_dot = dot(H.pt(0), local_dir);
    for (uint32_t i = 1; i < H.nv; ++i) {
        float d = dot(H.pt(i), local_dir);
        if (d > best_dot) {
            best_dot = d;
            sp = i;
        }
    }
    uint32_t first = __ldg(H.vfirst + sp), num = __ldg(H.vnum + sp);
    for (uint32_t i = 0; i < num; ++i) {
        uint32_t face_id = __ldg(H.fav + first + i);
        if (dot(H.fn(face_id), local_dir) >= ceps) return FID(NCB_FEATURE_FACE, face_id);
    }
    for (uint32_t i = 0; i < num; ++i) {
        uint32_t edge_id = __ldg(H.eav + first + i);
        if (fabsf(dot(H.edir(edge_id), local_dir)) <= seps) return FID(NCB_FEATURE_EDGE, edge_id);
    }
    return FID(NCB_FEATURE_VERTEX, sp);
}
// ConvexHull::support_feature_toward (convex.rs:511-540)
NCB_HD void hull_support_feature_toward(const HullView& H, const Iso& m, V3 dir, float2 angle, Feature& out) {
    feat_clear(out);
    V3 local_dir = iso_inv_vec(m, dir);
    uint32_t f = hull_support_feature_id_toward_eps(H, local_dir, angle);
    uint32_t kind = FID_KIND(f);
    if (kind == NCB_FEATURE_VERTEX) {
        feat_push(out, H.pt(FID_ID(f)), f);
        out.feature_id = f;
    } else if (kind == NCB_FEATURE_EDGE) {
        uint32_t e = FID_ID(f), v1 = __ldg(H.edge_vertices + 2 * e), v2 = __ldg(H.edge_vertices + 2 * e + 1);
        feat_push(out, H.pt(v1), FID(NCB_FEATURE_VERTEX, v1));
        feat_push(out, H.pt(v2), FID(NCB_FEATURE_VERTEX, v2));
        out.feature_id = f;
        feat_push_edge(out, f);
    } else {
        hull_face(H, FID_ID(f), out);
    }
    feat_transform(out, m);
}
// ConvexHull::feature_normal (convex.rs:463-485)
NCB_HD V3 hull_feature_normal(const HullView& H, uint32_t f) {
    uint32_t id = FID_ID(f), kind = FID_KIND(f);
    if (kind == NCB_FEATURE_FACE) return H.fn(id);
    if (kind == NCB_FEATURE_EDGE) return normalize(H.fn(__ldg(H.edge_faces + 2 * id)) + H.fn(__ldg(H.edge_faces + 2 * id + 1)));
    uint32_t first = __ldg(H.vfirst + id), last = first + __ldg(H.vnum + id);
    V3 n = v3(0.f, 0.f, 0.f);
    for (uint32_t i = first; i < last; ++i) n = n + H.fn(__ldg(H.fav + i));
    return normalize(n);
}

NCB_HD void support_face_toward(const Shape& s, const Iso& m, V3 dir, Feature& out) {
    if (s.type == NCB_SHAPE_CUBOID)
        cuboid_support_face_toward(s.he, m, dir, out);
    else
        hull_support_face_toward(s.hull, m, dir, out);
}
NCB_HD void support_feature_toward(const Shape& s, const Iso& m, V3 dir, float2 angle, Feature& out) {
    if (s.type == NCB_SHAPE_CUBOID)
        cuboid_support_feature_toward(s.he, m, dir, angle, out);
    else
        hull_support_feature_toward(s.hull, m, dir, angle, out);
}

// ---- ContactManifold (fresh manifold, DistanceBased(0.02)) --------------------------------------------------
struct ManifoldContact {
    V3 w1, w2, n;
    float depth;
    uint32_t f1, f2;
};
// ContactKinematic of a contact (query/contact/contact_kinematic.rs:57-66) besides its two feature ids: the tracked local points,
// the NeighborhoodGeometry of each side (g: 0 Point, 1 Line(dir), 2 Plane(dir)) and the dilations.  Only produced when the caller
// asked for it (ncb_set_kinematics): a fresh manifold then carries a side array, one entry per contact slot.
enum { G_POINT = 0, G_LINE = 1, G_PLANE = 2 };
struct Kin {
    V3 local1, local2, dir1, dir2;
    float dil1, dil2;
    uint32_t g1, g2;
};
NCB_HD Kin kin_zero() {
    Kin k;
    k.local1 = k.local2 = k.dir1 = k.dir2 = v3(0.f, 0.f, 0.f);
    k.dil1 = k.dil2 = 0.f;
    k.g1 = k.g2 = G_POINT;
    return k;
}
// P = false: a fresh manifold (one-shot update).  P = true: the persistent manifold of a stepping world — the reference's
// Slab<(TrackedContact, usize)> + DistanceBased cache with persistence 1 (contact_manifold.rs:14-236): entries are kept in
// CACHE order; `slot` is the slab key (free slots are reused LIFO), `live` = "remaining == persistence", `id` = insertion
// counter of this manifold (stable exactly as long as the reference's ContactId is).
template <bool P>
struct ManifoldT {
    ManifoldContact c[MANIFOLD_MAX];
    V3 track[MANIFOLD_MAX];
    int n;
    int deepest;  // >= 0; set to -1 when a contact had to be dropped (capacity), reported through the overflow counter
    Kin* kin = nullptr;  // nullptr, or MANIFOLD_MAX entries aligned with c[] (kinematics wanted)
    NCB_HD bool wants_kin() const { return kin != nullptr; }
};
template <>
struct ManifoldT<true> {
    ManifoldContact c[PM_CAP];
    V3 track[PM_CAP];
    uint8_t slot[PM_CAP], live[PM_CAP];
    uint32_t id[PM_CAP];
    uint8_t free_stack[PM_CAP];
    int n, deepest;
    int nfree, slab_len, had_live;
    uint32_t next_id;
    bool overflow;
    // the stepping world does not keep kinematics (they would have to live in every cache entry): documented in ncb200.h
    NCB_HD bool wants_kin() const { return false; }
};
typedef ManifoldT<false> Manifold;
typedef ManifoldT<true> PManifold;

template <bool P>
NCB_HD void manifold_push(ManifoldT<P>& mf, V3 w1, V3 w2, V3 n, float depth, uint32_t f1, uint32_t f2, V3 tracking_pt,
                          const Kin* kin = nullptr) {
    const float threshold = 0.02f;
    int closest = mf.n;
    float closest_dist = threshold * threshold;
    for (int i = 0; i < mf.n; ++i) {
        float d = norm_squared(tracking_pt - mf.track[i]);
        if (d < closest_dist) {
            closest_dist = d;
            closest = i;
        }
    }
    if (closest == mf.n) {
        if constexpr (P) {
            if (mf.n < PM_CAP) {
                ManifoldContact& c = mf.c[mf.n];
                c.w1 = w1, c.w2 = w2, c.n = n, c.depth = depth, c.f1 = f1, c.f2 = f2;
                mf.track[mf.n] = tracking_pt;
                mf.slot[mf.n] = mf.nfree > 0 ? mf.free_stack[--mf.nfree] : (uint8_t)mf.slab_len++;  // Slab::insert
                mf.live[mf.n] = 1;
                mf.id[mf.n] = ++mf.next_id;
                mf.n++;
            } else {
                mf.overflow = true;
            }
        } else {
            if (mf.n < MANIFOLD_MAX) {
                ManifoldContact& c = mf.c[mf.n];
                c.w1 = w1, c.w2 = w2, c.n = n, c.depth = depth, c.f1 = f1, c.f2 = f2;
                mf.track[mf.n] = tracking_pt;
                if (mf.kin && kin) mf.kin[mf.n] = *kin;
                mf.n++;
            } else {
                mf.deepest = -1;
            }
        }
    } else {
        ManifoldContact& c = mf.c[closest];
        if constexpr (P) {
            if (mf.live[closest]) {
                if (depth <= c.depth) return;  // the contact already in the cache is deeper
            } else {
                mf.live[closest] = 1;  // a contact of the previous step is matched: it keeps its slot and its id
            }
        } else {
            if (depth <= c.depth) return;
            if (mf.kin && kin) mf.kin[closest] = *kin;  // c.0.kinematic = kinematic (contact_manifold.rs:229)
        }
        c.w1 = w1, c.w2 = w2, c.n = n, c.depth = depth, c.f1 = f1, c.f2 = f2;
        mf.track[closest] = tracking_pt;
    }
}

// ---- generators -----------------------------------------------------------------------------------------------
template <bool P>
NCB_HD void gen_ball_ball(const Iso& ma, float r1, const Iso& mb, float r2, float prediction, ManifoldT<P>& mf) {
    V3 c1 = ma.t, c2 = mb.t;
    V3 delta = c2 - c1;
    float d2 = norm_squared(delta);
    float sum_radius = r1 + r2;
    float sre = sum_radius + prediction;
    if (d2 < sre * sre) {
        V3 normal = d2 != 0.f ? normalize(delta) : v3(1.f, 0.f, 0.f);
        Kin k = kin_zero();  // both sides (Face(0), origin, Point), dilated by the radii (ball_ball_manifold_generator.rs:46-58)
        k.dil1 = r1, k.dil2 = r2;
        manifold_push(mf, c1 + normal * r1, c2 + normal * (-r2), normal, sum_radius - sqrtf(d2), FACE0, FACE0, v3(0.f, 0.f, 0.f), &k);
    }
}
template <bool P>
NCB_HD void gen_plane_ball(const Iso& m1, V3 plane_n, const Iso& m2, float radius, float prediction, bool flip, ManifoldT<P>& mf) {
    V3 n = iso_mul_vec(m1, plane_n);
    V3 pc = m1.t, bc = m2.t;
    float dist = dot(bc - pc, n);
    float depth = -dist + radius;
    if (depth > -prediction) {
        V3 world1 = bc + n * (-dist);
        V3 world2 = bc + n * (-radius);
        Kin k = kin_zero();  // plane side (Face(0), local1, Plane(plane.normal)), ball side (Face(0), origin, Point) + dilation (:53-75)
        if (mf.wants_kin()) {
            V3 local1 = iso_inv_point(m1, world1);
            if (!flip)
                k.local1 = local1, k.g1 = G_PLANE, k.dir1 = plane_n, k.dil2 = radius;
            else
                k.local2 = local1, k.g2 = G_PLANE, k.dir2 = plane_n, k.dil1 = radius;
        }
        if (!flip)
            manifold_push(mf, world1, world2, n, depth, FACE0, FACE0, v3(0.f, 0.f, 0.f), &k);
        else
            manifold_push(mf, world2, world1, -n, depth, FACE0, FACE0, v3(0.f, 0.f, 0.f), &k);
    }
}
template <bool P>
__device__ __noinline__ void gen_plane_convex(const Iso& m1, V3 plane_n, const Iso& m2, const Shape& cp, float prediction, bool flip,
                                              ManifoldT<P>& mf, Feature& feat) {
    V3 n = iso_mul_vec(m1, plane_n);
    V3 pc = m1.t;
    support_face_toward(cp, m2, -n, feat);
    for (int i = 0; i < feat.nv; ++i) {
        V3 world2 = feat.v[i];
        float dist = dot(world2 - pc, n);
        if (dist <= prediction) {
            V3 world1 = world2 + (-n * dist);
            V3 local2 = iso_inv_point(m2, world2);
            uint32_t f2 = feat.vid[i];
            Kin k = kin_zero();  // plane side (Face(0), local1, Plane(plane.normal)), polyhedron side (vertex id, local2, Point) (:57-72)
            if (mf.wants_kin()) {
                V3 local1 = iso_inv_point(m1, world1);
                if (!flip)
                    k.local1 = local1, k.g1 = G_PLANE, k.dir1 = plane_n, k.local2 = local2;
                else
                    k.local1 = local2, k.local2 = local1, k.g2 = G_PLANE, k.dir2 = plane_n;
            }
            if (!flip)
                manifold_push(mf, world1, world2, n, -dist, FACE0, f2, local2, &k);
            else
                manifold_push(mf, world2, world1, -n, -dist, f2, FACE0, local2, &k);
        }
    }
}

// AABB::project_point_with_feature through Cuboid (point_cuboid.rs:17-27, point_aabb.rs:9-135)
__device__ __noinline__ void cuboid_project_point_with_feature(V3 he, const Iso& m, V3 pt, bool& inside_out, V3& proj_out,
                                                               uint32_t& feature) {
    V3 mins = v3(0.f, 0.f, 0.f) + (-he), maxs = v3(0.f, 0.f, 0.f) + he;
    V3 ls_pt = iso_inv_point(m, pt);
    V3 mins_pt = mins - ls_pt, pt_maxs = ls_pt - maxs;
    V3 zero = v3(0.f, 0.f, 0.f);
    V3 shift = vmax(mins_pt, zero) - vmax(pt_maxs, zero);
    bool inside = shift.x == 0.f && shift.y == 0.f && shift.z == 0.f;
    V3 ls_proj;
    if (!inside) {
        ls_proj = ls_pt + shift;
    } else {
        float best = -NCB_FMAX;
        bool is_mins = false;
        int best_id = 0;
        for (int i = 0; i < 3; ++i) {
            float a = vget(mins_pt, i), b = vget(pt_maxs, i);
            if (a < b) {
                if (b > best) {
                    best_id = i;
                    is_mins = false;
                    best = b;
                }
            } else if (a > best) {
                best_id = i;
                is_mins = true;
                best = a;
            }
        }
        shift = v3(0.f, 0.f, 0.f);
        vset(shift, best_id, is_mins ? best : -best);
        ls_proj = ls_pt + shift;
    }
    inside_out = inside;
    proj_out = iso_mul_point(m, ls_proj);
    int nzero = 0, last_zero = 0, last_not_zero = 0;
    for (int i = 0; i < 3; ++i) {
        if (vget(shift, i) == 0.f) {
            nzero++;
            last_zero = i;
        } else
            last_not_zero = i;
    }
    V3 center = (mins + maxs) * 0.5f;
    if (nzero == 3) {
        for (int i = 0; i < 3; ++i) {
            if (vget(ls_proj, i) > vget(maxs, i) - NCB_EPS) {
                feature = FID(NCB_FEATURE_FACE, i);
                return;
            }
            if (vget(ls_proj, i) <= vget(mins, i) + NCB_EPS) {
                feature = FID(NCB_FEATURE_FACE, i + 3);
                return;
            }
        }
        feature = FID_UNKNOWN;
    } else if (nzero == 2) {
        feature = vget(ls_proj, last_not_zero) < vget(center, last_not_zero) ? FID(NCB_FEATURE_FACE, last_not_zero + 3)
                                                                               : FID(NCB_FEATURE_FACE, last_not_zero);
    } else {
        uint32_t id = 0;
        for (int i = 0; i < 3; ++i)
            if (vget(ls_proj, i) < vget(center, i)) id |= 1u << i;
        feature = nzero == 0 ? FID(NCB_FEATURE_VERTEX, id) : FID(NCB_FEATURE_EDGE, (id << 2) | (uint32_t)last_zero);
    }
}

// the feature of the projection (point_support_map.rs:104-116)
NCB_HD uint32_t hull_project_feature(const HullView& H, const Iso& m_in, V3 point, bool inside, V3 proj, float2 one_degree_cs) {
    V3 dpt = point - proj;
    V3 local_dir = inside ? iso_inv_vec(m_in, -dpt) : iso_inv_vec(m_in, dpt);
    V3 u;
    if (unit_try_new(local_dir, NCB_EPS, u)) return hull_support_feature_id_toward_eps(H, u, one_degree_cs);
    return FID_UNKNOWN;
}

// (m1, ball) (m2, convex polyhedron)
// ConvexPolyhedron::edge(id) in local coordinates (cuboid.rs:163-183, convex.rs:430-441)
NCB_HD void shape_edge(const Shape& cp, uint32_t f, V3& p1, V3& p2) {
    uint32_t eid = FID_ID(f);
    if (cp.type == NCB_SHAPE_CUBOID) {
        uint32_t edge_i = eid & 3u, vertex_i = eid >> 2;
        V3 res = cp.he;
        for (uint32_t i = 0; i < 3; ++i)
            if (i != edge_i && (vertex_i & (1u << i))) vset(res, (int)i, -vget(res, (int)i));
        p1 = res;
        vset(res, (int)edge_i, -vget(res, (int)edge_i));
        p2 = res;
    } else {
        p1 = cp.hull.pt(__ldg(cp.hull.edge_vertices + 2 * eid)), p2 = cp.hull.pt(__ldg(cp.hull.edge_vertices + 2 * eid + 1));
    }
}
// kinematic of a ball x polyhedron contact (ball_convex_polyhedron_manifold_generator.rs:76-114): the ball side is (Face(0), origin,
// Point) dilated by the radius; the polyhedron side is local2 with the geometry of feature f2 (edge end points e0, e1 in local space)
NCB_HD Kin kin_ball_polyhedron(float radius, const Iso& mcp, V3 world2, V3 normal, uint32_t f2, V3 e0, V3 e1, bool flip) {
    Kin k = kin_zero();
    V3 local2 = iso_inv_point(mcp, world2);
    uint32_t g2 = G_POINT;
    V3 d2 = v3(0.f, 0.f, 0.f);
    if (FID_KIND(f2) == NCB_FEATURE_FACE)
        g2 = G_PLANE, d2 = iso_inv_vec(mcp, -normal);
    else if (FID_KIND(f2) == NCB_FEATURE_EDGE)
        g2 = G_LINE, d2 = normalize(e1 - e0);
    if (!flip)
        k.dil1 = radius, k.local2 = local2, k.g2 = g2, k.dir2 = d2;
    else
        k.dil2 = radius, k.local1 = local2, k.g1 = g2, k.dir1 = d2;
    return k;
}
template <bool P>
NCB_HD void gen_ball_convex_finish(V3 ball_center, float radius, const Shape& cp, const Iso& mcp, bool inside, V3 world2, uint32_t f2,
                                   float prediction, bool flip, ManifoldT<P>& mf) {
    V3 dpt = world2 - ball_center;
    float depth, dist;
    V3 normal, dir;
    if (unit_try_new_and_get(dpt, NCB_EPS, dir, dist)) {
        if (inside) {
            depth = dist + radius;
            normal = -dir;
        } else {
            depth = -dist + radius;
            normal = dir;
        }
    } else {
        if (f2 == FID_UNKNOWN) return;
        depth = radius;
        normal = -(cp.type == NCB_SHAPE_CUBOID ? cuboid_feature_normal(f2) : hull_feature_normal(cp.hull, f2));
    }
    if (depth >= -prediction) {
        V3 world1 = ball_center + normal * radius;
        Kin k = kin_zero();
        if (mf.wants_kin()) {
            V3 e0 = v3(0.f, 0.f, 0.f), e1 = e0;
            if (FID_KIND(f2) == NCB_FEATURE_EDGE) shape_edge(cp, f2, e0, e1);
            k = kin_ball_polyhedron(radius, mcp, world2, normal, f2, e0, e1, flip);
        }
        if (!flip)
            manifold_push(mf, world1, world2, normal, depth, FACE0, f2, v3(0.f, 0.f, 0.f), &k);
        else
            manifold_push(mf, world2, world1, -normal, depth, f2, FACE0, v3(0.f, 0.f, 0.f), &k);
    }
}

// ---- polygon clipping -----------------------------------------------------------------------------------------
NCB_HD bool point_in_poly2d(V2 pt, const V2* poly, int n) {
    if (n == 0) return false;
    float sign = 0.f;
    for (int i1 = 0; i1 < n; ++i1) {
        int i2 = i1 + 1 == n ? 0 : i1 + 1;  // (i1 + 1) % n without the integer division
        V2 seg_dir = V2{poly[i2].x - poly[i1].x, poly[i2].y - poly[i1].y};
        V2 dpt = V2{pt.x - poly[i1].x, pt.y - poly[i1].y};
        float perp = dpt.x * seg_dir.y - dpt.y * seg_dir.x;
        if (sign == 0.f)
            sign = perp;
        else if (sign * perp < 0.f)
            return false;
    }
    return true;
}
NCB_HD bool line_toi_with_plane(V3 plane_center, V3 plane_normal, V3 line_origin, V3 line_dir, float& toi) {
    V3 dpos = plane_center - line_origin;
    float denom = dot(plane_normal, line_dir);
    if (relative_eq(denom, 0.f)) return false;
    toi = dot(plane_normal, dpos) / denom;
    return true;
}
NCB_HD float dot2(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
// closest_points_segment_segment_with_locations_nD, D = 2; true iff both locations are OnEdge
NCB_HD bool seg_seg_2d(V2 a1, V2 b1, V2 a2, V2 b2, float& s_out, float& t_out) {
    const float eps = NCB_EPS;
    V2 d1 = V2{b1.x - a1.x, b1.y - a1.y}, d2 = V2{b2.x - a2.x, b2.y - a2.y}, r = V2{a1.x - a2.x, a1.y - a2.y};
    float a = dot2(d1, d1), e = dot2(d2, d2), f = dot2(d2, r);
    float s, t;
    if (a <= eps && e <= eps) {
        s = 0.f;
        t = 0.f;
    } else if (a <= eps) {
        s = 0.f;
        t = clampf(f / e, 0.f, 1.f);
    } else {
        float c = dot2(d1, r);
        if (e <= eps) {
            t = 0.f;
            s = clampf(-c / a, 0.f, 1.f);
        } else {
            float b = dot2(d1, d2);
            float ae = a * e, bb = b * b, denom = ae - bb;
            bool parallel = denom <= eps || ulps_eq(ae, bb);
            s = !parallel ? clampf((b * f - c * e) / denom, 0.f, 1.f) : 0.f;
            t = (b * s + f) / e;
            if (t < 0.f) {
                t = 0.f;
                s = clampf(-c / a, 0.f, 1.f);
            } else if (t > 1.f) {
                t = 1.f;
                s = clampf((b - c) / a, 0.f, 1.f);
            }
        }
    }
    s_out = s;
    t_out = t;
    return s != 0.f && s != 1.f && t != 0.f && t != 1.f;
}

// add_contact_to_manifold's early returns (convex_polygonal_feature3.rs:355-393)
NCB_HD bool feature_ok_for_manifold(const Feature& ft, uint32_t f) {
    uint32_t kind = FID_KIND(f);
    if (kind == NCB_FEATURE_FACE || kind == NCB_FEATURE_VERTEX) return true;
    if (kind == NCB_FEATURE_EDGE) {
        // the first i1 < min(nv, ne) with eid[i1] == f; scanned from the back without a data-dependent exit, so the loads of the
        // (local-memory) id array do not wait for one another
        int lim = ft.nv < ft.ne ? ft.nv : ft.ne, i1 = -1;
#pragma unroll 4
        for (int i = lim - 1; i >= 0; --i) i1 = ft.eid[i] == f ? i : i1;
        if (i1 < 0) return false;
        int i2 = i1 + 1 == ft.nv ? 0 : i1 + 1;
        V3 d;
        return unit_try_new(ft.v[i2] - ft.v[i1], NCB_EPS, d);
    }
    return false;
}

// NeighborhoodGeometry of feature f of the polygonal feature ft in the local frame of m (add_contact_to_manifold,
// convex_polygonal_feature3.rs:356-397); f passed feature_ok_for_manifold
NCB_HD void feature_geometry(const Feature& ft, uint32_t f, const Iso& m, uint32_t& g, V3& dir) {
    g = G_POINT, dir = v3(0.f, 0.f, 0.f);
    uint32_t kind = FID_KIND(f);
    if (kind == NCB_FEATURE_FACE) {
        g = G_PLANE, dir = iso_inv_vec(m, ft.normal);
    } else if (kind == NCB_FEATURE_EDGE) {
        for (int i1 = 0; i1 < ft.nv; ++i1) {
            if (i1 < ft.ne && ft.eid[i1] == f) {
                int i2 = i1 + 1 == ft.nv ? 0 : i1 + 1;
                V3 d = v3(0.f, 0.f, 0.f);
                unit_try_new(ft.v[i2] - ft.v[i1], NCB_EPS, d);
                g = G_LINE, dir = iso_inv_vec(m, d);
                return;
            }
        }
    }
}
NCB_HD Kin kin_from_features(const Feature& m1, const Feature& m2, const Iso& ma, const Iso& mb, V3 w1, V3 w2, uint32_t f1, uint32_t f2, V3 local1) {
    Kin k = kin_zero();
    k.local1 = local1;
    k.local2 = iso_inv_point(mb, w2);
    feature_geometry(m1, f1, ma, k.g1, k.dir1);
    feature_geometry(m2, f2, mb, k.g2, k.dir2);
    return k;
}

// Candidates found by clip() are buffered and handed to the manifold afterwards, like the reference's `new_contacts`
// Vec (convex_polyhedron_convex_polyhedron_manifold_generator.rs:147-161).  Besides mirroring the reference, this keeps
// the (rare, lane-dependent) manifold work out of clip's nested loops so that lanes stay converged in both parts.
#define CLIP_CAND_MAX 24
struct ClipCand {
    V3 w1, w2;
    uint32_t f1, f2;
};
template <bool P>
struct ClipCtxT {
    const Iso *ma, *mb;
    ManifoldT<P>* mf;
    const Feature *m1, *m2;
    V3 normal;
    int n_new;   // candidates within the prediction distance so far (buffered + already flushed)
    int n_buf;
    ClipCand buf[CLIP_CAND_MAX];
};
template <bool P>
NCB_HD void clip_flush(ClipCtxT<P>& cc) {
    for (int k = 0; k < cc.n_buf; ++k) {
        const ClipCand& c = cc.buf[k];
        if (!feature_ok_for_manifold(*cc.m1, c.f1)) continue;
        if (!feature_ok_for_manifold(*cc.m2, c.f2)) continue;
        float depth = -dot(cc.normal, c.w2 - c.w1);  // Contact::new_wo_depth
        V3 local1 = iso_inv_point(*cc.ma, c.w1);
        Kin kin = kin_zero();
        if (cc.mf->wants_kin()) kin = kin_from_features(*cc.m1, *cc.m2, *cc.ma, *cc.mb, c.w1, c.w2, c.f1, c.f2, local1);
        manifold_push(*cc.mf, c.w1, c.w2, cc.normal, depth, c.f1, c.f2, local1, &kin);
    }
    cc.n_buf = 0;
}
template <bool P>
NCB_HD void clip_emit(ClipCtxT<P>& cc, V3 w1, V3 w2, V3 normal, float prediction, uint32_t f1, uint32_t f2) {
    float depth = -dot(normal, w2 - w1);  // Contact::new_wo_depth
    if (-depth <= prediction) {
        cc.n_new++;
        if (cc.n_buf == CLIP_CAND_MAX) clip_flush(cc);  // order-preserving spill, practically never taken
        ClipCand& c = cc.buf[cc.n_buf++];
        c.w1 = w1, c.w2 = w2, c.f1 = f1, c.f2 = f2;
    }
}

// ConvexPolygonalFeature::clip (convex_polygonal_feature3.rs:217-338); candidates go straight into the manifold
// in the reference's order (the reference buffers them in a Vec and pushes them afterwards: same result).
template <bool P>
__device__ __noinline__ void clip(const Feature& self, const Feature& other, V3 normal, float prediction, ClipCtxT<P>& cc) {
    if (self.nv <= 2 && other.nv <= 2) return;
    V3 b0, b1;
    orthonormal_basis(normal, b0, b1);
    V3 ref_pt = self.v[0];
    V2 poly1[FEAT_MAX], poly2[FEAT_MAX];
    for (int i = 0; i < self.nv; ++i) {
        V3 dpt = self.v[i] - ref_pt;
        poly1[i] = V2{dot(b0, dpt), dot(b1, dpt)};
    }
    for (int i = 0; i < other.nv; ++i) {
        V3 dpt = other.v[i] - ref_pt;
        poly2[i] = V2{dot(b0, dpt), dot(b1, dpt)};
    }
    if (other.nv > 2) {
        for (int i = 0; i < self.nv; ++i) {
            V2 pt = poly1[i];
            if (point_in_poly2d(pt, poly2, other.nv)) {
                V3 origin = ref_pt + b0 * pt.x + b1 * pt.y;
                float toi2;
                if (line_toi_with_plane(other.v[0], other.normal, origin, normal, toi2)) {
                    V3 world2 = origin + normal * toi2;
                    clip_emit(cc, self.v[i], world2, normal, prediction, self.vid[i], other.feature_id);
                }
            }
        }
    }
    if (self.nv > 2) {
        for (int i = 0; i < other.nv; ++i) {
            V2 pt = poly2[i];
            if (point_in_poly2d(pt, poly1, self.nv)) {
                V3 origin = ref_pt + b0 * pt.x + b1 * pt.y;
                float toi1;
                if (line_toi_with_plane(self.v[0], self.normal, origin, normal, toi1)) {
                    V3 world1 = origin + normal * toi1;
                    clip_emit(cc, world1, other.v[i], normal, prediction, self.feature_id, other.vid[i]);
                }
            }
        }
    }
    int nedges1 = feat_nedges(self), nedges2 = feat_nedges(other);
    for (int i1 = 0; i1 < nedges1; ++i1) {
        int j1 = i1 + 1 == self.nv ? 0 : i1 + 1;
        for (int i2 = 0; i2 < nedges2; ++i2) {
            int j2 = i2 + 1 == other.nv ? 0 : i2 + 1;
            float s, t;
            if (seg_seg_2d(poly1[i1], poly1[j1], poly2[i2], poly2[j2], s, t)) {
                V3 world1 = self.v[i1] * (1.f - s) + self.v[j1] * s;
                V3 world2 = other.v[i2] * (1.f - t) + other.v[j2] * t;
                clip_emit(cc, world1, world2, normal, prediction, self.eid[i1], other.eid[i2]);
            }
        }
    }
}

// ConvexPolyhedronConvexPolyhedronManifoldGenerator::generate_contacts after the GJK/EPA result is known
// (convex_polyhedron_convex_polyhedron_manifold_generator.rs:112-161).
template <bool P>
__device__ __noinline__ void convex_convex_manifold(const Iso& ma, const Shape& a, const Iso& mb, const Shape& b, float linear, float2 ang1,
                                                    float2 ang2, V3 p1, V3 p2, V3 dir, ManifoldT<P>& mf, Feature& m1, Feature& m2) {
    float depth = -dot(dir, p2 - p1);
    {
        // the two feature extractions are independent: issue the cuboid one first and the hull one second whatever the
        // pair's orientation is, so that mixed (cuboid, hull) / (hull, cuboid) lanes stay converged
        bool swap = a.type == NCB_SHAPE_CONVEX_HULL && b.type != NCB_SHAPE_CONVEX_HULL;
        const Shape& sa = swap ? b : a;
        const Shape& sb = swap ? a : b;
        const Iso& ia = swap ? mb : ma;
        const Iso& ib = swap ? ma : mb;
        V3 da = swap ? -dir : dir;
        float2 anga = swap ? ang2 : ang1, angb = swap ? ang1 : ang2;
        Feature& fa = swap ? m2 : m1;
        Feature& fb = swap ? m1 : m2;
        if (depth > 0.f) {
            support_face_toward(sa, ia, da, fa);
            support_face_toward(sb, ib, -da, fb);
        } else {
            support_feature_toward(sa, ia, da, anga, fa);
            support_feature_toward(sb, ib, -da, angb, fb);
        }
    }
    ClipCtxT<P> cc;
    cc.ma = &ma;
    cc.mb = &mb;
    cc.mf = &mf;
    cc.m1 = &m1;
    cc.m2 = &m2;
    cc.n_new = 0;
    cc.n_buf = 0;
    cc.normal = dir;
    clip(m1, m2, dir, linear, cc);
    clip_flush(cc);
    if (cc.n_new == 0) {
        if (feature_ok_for_manifold(m1, m1.feature_id) && feature_ok_for_manifold(m2, m2.feature_id)) {
            V3 local1 = iso_inv_point(ma, p1);
            Kin k = kin_zero();
            if (mf.wants_kin()) k = kin_from_features(m1, m2, ma, mb, p1, p2, m1.feature_id, m2.feature_id, local1);
            manifold_push(mf, p1, p2, dir, depth, m1.feature_id, m2.feature_id, local1, &k);
        }
    }
}

#ifndef NCB_HOST_SHIM  // warp-level code: CUDA only
// ---- result write-out: warp-aggregated allocation of contact slots --------------------------------------------
NCB_HD void write_manifold(const Manifold& mf, bool valid, uint32_t pair_slot, uint32_t out_index, ncb_contact* __restrict__ contacts,
                           uint32_t cap_contacts, uint32_t* __restrict__ manifold_start, uint8_t* __restrict__ manifold_count,
                           DevCounters* cnt, ncb_kinematic* __restrict__ kin_out = nullptr) {
    // all 32 lanes call this (valid = false for idle lanes)
    if (valid && mf.deepest < 0) atomicAdd(&cnt->epa_overflow, 1u);  // more than MANIFOLD_MAX distinct contacts: never silent
    uint32_t n = valid ? (uint32_t)mf.n : 0u;
    uint32_t incl = n;
    int lane = threadIdx.x & 31;
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t nz = __popc(__ballot_sync(0xffffffffu, n != 0));
    uint32_t base = 0;
    if (lane == 31 && total) {
        base = atomicAdd(&cnt->n_contacts, total);
        atomicAdd(&cnt->n_contact_pairs, nz);
    }
    base = __shfl_sync(0xffffffffu, base, 31);
    if (!valid) return;
    uint32_t start = base + incl - n;
    manifold_start[out_index] = start;
    manifold_count[out_index] = (uint8_t)n;
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t dst = start + k;
        if (dst >= cap_contacts) break;
        const ManifoldContact& c = mf.c[k];
        ncb_contact o;
        o.world1[0] = c.w1.x, o.world1[1] = c.w1.y, o.world1[2] = c.w1.z;
        o.world2[0] = c.w2.x, o.world2[1] = c.w2.y, o.world2[2] = c.w2.z;
        o.normal[0] = c.n.x, o.normal[1] = c.n.y, o.normal[2] = c.n.z;
        o.depth = c.depth;
        o.f1 = c.f1, o.f2 = c.f2;
        o.pair = out_index;
        contacts[dst] = o;
        if (kin_out && mf.kin) {
            const Kin& q = mf.kin[k];
            ncb_kinematic w;
            w.local1[0] = q.local1.x, w.local1[1] = q.local1.y, w.local1[2] = q.local1.z;
            w.local2[0] = q.local2.x, w.local2[1] = q.local2.y, w.local2[2] = q.local2.z;
            w.dir1[0] = q.dir1.x, w.dir1[1] = q.dir1.y, w.dir1[2] = q.dir1.z;
            w.dir2[0] = q.dir2.x, w.dir2[1] = q.dir2.y, w.dir2[2] = q.dir2.z;
            w.dilation1 = q.dil1, w.dilation2 = q.dil2;
            w.geometry1 = q.g1, w.geometry2 = q.g2;
            kin_out[dst] = w;
        }
    }
}

#endif  // NCB_HOST_SHIM (write_manifold)

// ---- persistent manifold / generator state of a stepping world (indexed by state slot = pair_index[p]) ----------
// header: 8 words  [0] n | nfree << 8 | slab_len << 16   [1] next_id   [2..7] free stack (one byte per entry)
// entry : 4 float4 [w1, depth] [w2, f1] [normal, f2] [tracking point, slot | live << 8 | id << 9], in cache order
__device__ __forceinline__ void pm_load_and_age(const PersistArgs& ps, uint32_t slot, PManifold& mf) {
    const uint32_t* hdr = ps.pm_hdr + (size_t)slot * PM_HDR_WORDS;
    uint32_t w0 = hdr[0];
    int n0 = (int)(w0 & 0xffu);
    mf.nfree = (int)((w0 >> 8) & 0xffu);
    mf.slab_len = (int)((w0 >> 16) & 0xffu);
    mf.next_id = hdr[1];
    for (int k = 0; k < PM_CAP / 4; ++k) {
        uint32_t w = hdr[2 + k];
        mf.free_stack[4 * k] = (uint8_t)w, mf.free_stack[4 * k + 1] = (uint8_t)(w >> 8);
        mf.free_stack[4 * k + 2] = (uint8_t)(w >> 16), mf.free_stack[4 * k + 3] = (uint8_t)(w >> 24);
    }
    mf.n = 0, mf.deepest = 0, mf.had_live = 0, mf.overflow = false;
    // ContactManifold::save_cache_and_clear (contact_manifold.rs:134-156): contacts of the previous update stay in the
    // cache as match candidates (live -> stale); contacts that were already stale leave the cache and the slab
    uint32_t stale_slots = 0;
    const float4* e = ps.pm_entry + (size_t)slot * PM_CAP * PM_ENTRY_F4;
    for (int i = 0; i < n0; ++i) {
        float4 d = e[4 * i + 3];
        uint32_t meta = __float_as_uint(d.w);
        if ((meta >> 8) & 1u) {
            float4 a = e[4 * i], b = e[4 * i + 1], c = e[4 * i + 2];
            ManifoldContact& mc = mf.c[mf.n];
            mc.w1 = v3(a.x, a.y, a.z), mc.depth = a.w;
            mc.w2 = v3(b.x, b.y, b.z), mc.f1 = __float_as_uint(b.w);
            mc.n = v3(c.x, c.y, c.z), mc.f2 = __float_as_uint(c.w);
            mf.track[mf.n] = v3(d.x, d.y, d.z);
            mf.slot[mf.n] = (uint8_t)meta;
            mf.live[mf.n] = 0;
            mf.id[mf.n] = meta >> 9;
            mf.n++;
            mf.had_live++;
        } else {
            stale_slots |= 1u << (meta & 0xffu);
        }
    }
    while (stale_slots) {  // Slab::retain visits the keys in increasing order; every removal pushes its key on the free list
        int sl = __ffs(stale_slots) - 1;
        stale_slots &= stale_slots - 1;
        mf.free_stack[mf.nfree++] = (uint8_t)sl;
    }
}
__device__ __forceinline__ void pm_store(const PersistArgs& ps, uint32_t slot, const PManifold& mf, uint32_t h1, uint32_t h2) {
    uint32_t* hdr = ps.pm_hdr + (size_t)slot * PM_HDR_WORDS;
    hdr[0] = (uint32_t)mf.n | ((uint32_t)mf.nfree << 8) | ((uint32_t)mf.slab_len << 16);
    hdr[1] = mf.next_id;
    for (int k = 0; k < PM_CAP / 4; ++k)
        hdr[2 + k] = (uint32_t)mf.free_stack[4 * k] | ((uint32_t)mf.free_stack[4 * k + 1] << 8) | ((uint32_t)mf.free_stack[4 * k + 2] << 16) |
                     ((uint32_t)mf.free_stack[4 * k + 3] << 24);
    float4* e = ps.pm_entry + (size_t)slot * PM_CAP * PM_ENTRY_F4;
    int has = 0;
    for (int i = 0; i < mf.n; ++i) {
        const ManifoldContact& mc = mf.c[i];
        e[4 * i] = make_float4(mc.w1.x, mc.w1.y, mc.w1.z, mc.depth);
        e[4 * i + 1] = make_float4(mc.w2.x, mc.w2.y, mc.w2.z, __uint_as_float(mc.f1));
        e[4 * i + 2] = make_float4(mc.n.x, mc.n.y, mc.n.z, __uint_as_float(mc.f2));
        uint32_t meta = (uint32_t)mf.slot[i] | ((uint32_t)mf.live[i] << 8) | (mf.id[i] << 9);
        e[4 * i + 3] = make_float4(mf.track[i].x, mf.track[i].y, mf.track[i].z, __uint_as_float(meta));
        has += mf.live[i];
    }
    if (mf.overflow) atomicAdd(ps.pm_overflow, 1u);
    // NarrowPhase::update_contact (narrow_phase.rs:92-103): ContactEvent::Started / Stopped
    if ((mf.had_live > 0) != (has > 0)) {
        uint32_t k = atomicAdd(ps.n_events, 1u);
        if (k < ps.cap_events) ps.events[k] = ((has > 0) ? (1ull << 63) : 0ull) | ((unsigned long long)h1 << 32) | h2;
    }
}
// An updated pair whose generator pushes nothing (separated shapes) still went through save_cache_and_clear.
__device__ __noinline__ void pm_age_only(const PersistArgs& ps, uint32_t slot, uint32_t h1, uint32_t h2) {
    if ((ps.pm_hdr[(size_t)slot * PM_HDR_WORDS] & 0xffu) == 0) return;
    PManifold mf;
    pm_load_and_age(ps, slot, mf);
    pm_store(ps, slot, mf, h1, h2);
}

}  // namespace ncb
#include "capsule.cuh"
namespace ncb {

#ifndef NCB_HOST_SHIM  // kernels, work queues and launchers: CUDA only
struct NarrowArgs {
    PersistArgs ps;
    DevObjects o;
    DevHulls H;
    const uint2* pairs;
    const uint32_t* pair_index;  // optional: original index of each sorted pair
    ncb_contact* contacts;
    ncb_kinematic* kin_out;  // nullptr: kinematics not wanted (ncb_set_kinematics)
    uint32_t cap_contacts;
    uint32_t* manifold_start;
    uint8_t* manifold_count;
    DevCounters* cnt;
    uint32_t cap_pairs;
    float2 one_degree_cs;  // cos / sin of (pi / 180) as f32, from the host libm (convex.rs:543)
    uint32_t* epa_queue;   // EPA_REC_WORDS per record
    uint32_t* cp_queue;    // CP_REC_WORDS per record
    int epa_refill_min;    // idle lanes needed before a warp of k_cc_epa_s refills (batched initialisation)
    int man_part, man_parts;  // k_cc_manifold walks part man_part of man_parts equal parts of the manifold queue
    uint32_t* epa_long;    // [0, cap_pairs): EPA-queue indices tier 1 deferred to tier 2; [cap_pairs, 2 cap_pairs): tier 2 to the last resort
};

// ---- convex x convex in three compacted phases -------------------------------------------------------------------
// One thread per pair through GJK + EPA + clipping makes a warp wait for its slowest lane (penetrating pairs cost
// 10-100x a separated pair: measured 4.8 of 32 lanes active).  Instead:
//   k_cc_gjk      all pairs of the key segment: GJK only; separated pairs finish here, penetrating pairs append their
//                 simplex to the EPA queue, pairs with closest points append a record to the manifold queue;
//   k_cc_epa      EPA over the compacted EPA queue, appends to the manifold queue;
//   k_cc_manifold support features + clipping + manifold over the compacted manifold queue.
// Queues live at [key_start[key], cursor[key]) of two arrays sized like the pair array; cursors are device counters.
#define EPA_REC_WORDS 26
#define CP_REC_WORDS 10

NCB_HD uint32_t queue_append(uint32_t* cursor, bool want) {
    // warp-aggregated slot allocation; all 32 lanes call it
    unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return 0;
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
}
NCB_HD void cp_store(uint32_t* q, uint32_t slot, uint32_t p, V3 p1, V3 p2, V3 dir) {
    float* r = reinterpret_cast<float*>(q + (size_t)slot * CP_REC_WORDS);
    q[(size_t)slot * CP_REC_WORDS] = p;
    r[1] = p1.x, r[2] = p1.y, r[3] = p1.z, r[4] = p2.x, r[5] = p2.y, r[6] = p2.z, r[7] = dir.x, r[8] = dir.y, r[9] = dir.z;
}

// The three convex-convex keys are adjacent in the key order, so their pairs form ONE contiguous range of the sorted
// pair array and share one EPA queue and one manifold queue (cursor slot CCQ): a single launch per phase, one tail.
#define CCQ K_CUBOID_CUBOID
#ifndef NCB_GJK_MINBLOCKS
#define NCB_GJK_MINBLOCKS 6
#endif
template <bool PS>
__global__ void __launch_bounds__(128, NCB_GJK_MINBLOCKS) k_cc_gjk(NarrowArgs A) {
    const int KEY = CCQ;
    uint32_t seg_begin = A.cnt->key_start[K_CUBOID_CUBOID];
    uint32_t seg_end = A.cnt->key_start[K_HULL_HULL] + A.cnt->key_hist[K_HULL_HULL];
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = seg_begin + blockIdx.x * blockDim.x; base < seg_end; base += stride) {
        uint32_t p = base + threadIdx.x;
        bool valid = p < seg_end;
        int r = GJK_NO_INTERSECTION;
        V3 p1, p2, dir;
        Simplex s;
        if (valid) {
            uint2 pr = __ldg(&A.pairs[p]);
            uint32_t i1 = pr.x, i2 = pr.y;
            uint32_t t1 = __ldg(&A.o.type[i1]), t2 = __ldg(&A.o.type[i2]);
            Iso ma = load_iso(A.o, i1), mb = load_iso(A.o, i2);
            float linear = __ldg(&A.o.qlimit[i1]) + __ldg(&A.o.qlimit[i2]);
            // GJK only evaluates support points: slim operands (kind, half extents | vertex array) instead of the full hull views
            SupportS ga = load_slim_support(A.o, A.H, i1, t1), gb = load_slim_support(A.o, A.H, i2, t2);
            // contact_support_map_support_map_with_params, init_dir = None (fresh generator)
            V3 d0;
            uint32_t out_index = 0;
            bool warm = false;
            if constexpr (PS) {  // the generator's last_gjk_dir (convex_polyhedron_convex_polyhedron_manifold_generator.rs:98)
                out_index = A.pair_index ? __ldg(&A.pair_index[p]) : p;
                float4 pd = A.ps.dir[out_index];
                if (pd.w != 0.f) d0 = v3(pd.x, pd.y, pd.z), warm = true;
            }
            if (!warm && !unit_try_new(mb.t - ma.t, NCB_EPS, d0)) d0 = v3(1.f, 0.f, 0.f);
            r = gjk_closest_points(ma, ga, mb, gb, linear, d0, s, p1, p2, dir);
            if constexpr (PS) {
                if (r != GJK_INTERSECTION) A.ps.dir[out_index] = make_float4(dir.x, dir.y, dir.z, 1.f);  // :106 / :139
                if (r == GJK_NO_INTERSECTION) pm_age_only(A.ps, out_index, i1, i2);
            } else if (r == GJK_NO_INTERSECTION) {
                out_index = A.pair_index ? __ldg(&A.pair_index[p]) : p;
                A.manifold_start[out_index] = 0;
                A.manifold_count[out_index] = 0;
            }
        }
        uint32_t slot = queue_append(&A.cnt->cp_cursor[KEY], valid && r == GJK_CLOSEST_POINTS);
        if (valid && r == GJK_CLOSEST_POINTS) cp_store(A.cp_queue, slot, p, p1, p2, dir);
        slot = queue_append(&A.cnt->epa_cursor[KEY], valid && r == GJK_INTERSECTION);
        if (valid && r == GJK_INTERSECTION) {
            uint32_t* q = A.epa_queue + (size_t)slot * EPA_REC_WORDS;
            float* f = reinterpret_cast<float*>(q);
            q[0] = p;
            q[1] = (uint32_t)s.dim;
            for (int i = 0; i < 4; ++i) {
                f[2 + 6 * i + 0] = s.v[i].orig1.x, f[2 + 6 * i + 1] = s.v[i].orig1.y, f[2 + 6 * i + 2] = s.v[i].orig1.z;
                f[2 + 6 * i + 3] = s.v[i].orig2.x, f[2 + 6 * i + 4] = s.v[i].orig2.y, f[2 + 6 * i + 5] = s.v[i].orig2.z;
            }
        }
    }
}

// EPA over the compacted queue.  Lanes are independent workers: an idle lane fetches the next queue entry and builds
// its initial polytope; a busy lane executes ONE expansion step per turn of the outer loop.  All busy lanes therefore
// run the same code (one step) regardless of how many steps their pair needs; refills are batched (>= refill_min idle
// lanes) so that the initialisation path is not paid on every turn.
//
// Three tiers, one after the other on the stream, each restarting what the previous one could not hold (a restart repeats the same
// arithmetic, so the result does not depend on the tier):
//   k_cc_epa_tier<PS, 1>  the whole EPA queue; polytope in SHARED memory (EpaTier1: 16 / 48 / 24 in 133 lane-strided words, face
//                         normals recomputed), 6 CTAs of 64 threads per SM, operands slim (kind, half extents | vertex array) in
//                         registers.  Round 1's kernel kept a 7.4 KB polytope per thread in local memory: 32 warps x 55 KB of touched
//                         lines per SM overflowed L1 and, over 148 SMs, L2, and ncu counted 3.0 GB of DRAM traffic for 82 MB of
//                         algorithmic bytes; here the expansion loop touches no global or local memory except the hull vertices
//                         (L1-resident library) and 24 B of cold support points per new vertex (0.13 GB of DRAM traffic).
//   k_cc_epa_tier<PS, 2>  the 1 % that outgrew tier 1 (queue A.epa_long): EpaTier2 (32 / 160 / 96) in shared memory, CTAs of one
//                         warp, the pairs spread over all warps (a few lanes each): what is left is the latency of the longest runs.
//   k_cc_epa_big          last resort on the local-memory store (48 / 192 / 160): beyond tier 2, segment simplices.  Empty on every
//                         scene measured so far (it then returns at once).
// Handling the 1 % inside the tier-1 kernel was measured and rejected (profiles/r2_epa_overflow.txt): with generic addressing and
// per-lane capacities the kernel executes 29 % more instructions, and a lane whose polytope lives in global memory slows its
// whole warp down (0.79 -> 1.49 ms).
#define EPAS_THREADS 64
#define EPAT2_THREADS 32
#ifndef NCB_EPAS_MINBLOCKS
#define NCB_EPAS_MINBLOCKS 6
#endif

NCB_HD void epa_rec_load(const uint32_t* q, uint32_t& p, int& sdim, CSOPoint* sv) {
    const float* f = reinterpret_cast<const float*>(q);
    p = q[0];
    sdim = (int)q[1];
    for (int i = 0; i < 4; ++i) {
        sv[i].orig1 = v3(f[2 + 6 * i + 0], f[2 + 6 * i + 1], f[2 + 6 * i + 2]);
        sv[i].orig2 = v3(f[2 + 6 * i + 3], f[2 + 6 * i + 4], f[2 + 6 * i + 5]);
        sv[i].point = sv[i].orig1 - sv[i].orig2;  // bit-identical to the value GJK computed (CSOPoint::new)
    }
}

template <int TIER>
struct EpaTierTraits;
template <>
struct EpaTierTraits<1> {
    enum { THREADS = EPAS_THREADS, MINBLOCKS = NCB_EPAS_MINBLOCKS };
    typedef EpaTier1<EPAS_THREADS> Store;
};
template <>
struct EpaTierTraits<2> {
    enum { THREADS = EPAT2_THREADS, MINBLOCKS = 3 };
    typedef EpaTier2<EPAT2_THREADS> Store;
};

template <bool PS, int TIER>
__global__ void __launch_bounds__(EpaTierTraits<TIER>::THREADS, EpaTierTraits<TIER>::MINBLOCKS) k_cc_epa_tier(NarrowArgs A) {
    extern __shared__ uint32_t epa_smem[];
    typedef typename EpaTierTraits<TIER>::Store Store;
    const int KEY = CCQ;
    // tier 1 walks the EPA queue itself; tier 2 the list of queue indices tier 1 deferred
    const uint32_t seg_end = TIER == 1 ? A.cnt->epa_cursor[KEY] : A.cnt->epa_long_n;
    uint32_t* fetch = TIER == 1 ? &A.cnt->epa_fetch[KEY] : &A.cnt->epa_long_fetch;
    if (TIER == 2 && seg_end == 0) return;
    const int lane = threadIdx.x & 31;
    // tier 2: a handful of pairs, long runs: spread them over the warps of the grid instead of filling the first warps
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const uint32_t lanes = TIER == 1 ? 32u : min(32u, (seg_end + nwarps - 1) / nwarps);
    const bool usable = (uint32_t)lane < lanes;
    const int refill_min = TIER == 1 ? A.epa_refill_min : 1;
    Store e;
    e.base = epa_smem + threadIdx.x;
    bool active = false, exhausted = false;
    uint32_t p = 0, wq = 0;
    Iso ma, mb;
    SupportS ga, gb;
    V3 p1, p2, n;
    for (;;) {
        int status = EPA_CONTINUE;
        uint32_t res_face = EPA_RES_DIRECT;
        unsigned idle = __ballot_sync(0xffffffffu, usable && !active);
        unsigned all_idle = __ballot_sync(0xffffffffu, !active);
        bool refill = !exhausted && idle != 0 && (all_idle == 0xffffffffu || __popc(idle) >= refill_min);
        if (refill) {  // warp-uniform
            uint32_t base = 0;
            int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(fetch, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + __popc(idle) >= seg_end) exhausted = true;  // nothing left after this batch
            if (usable && !active) {
                uint32_t w = base + __popc(idle & ((1u << lane) - 1));
                if (w < seg_end) {
                    wq = TIER == 1 ? w : __ldg(&A.epa_long[w]);
                    int sdim;
                    CSOPoint sv[4];
                    epa_rec_load(A.epa_queue + (size_t)wq * EPA_REC_WORDS, p, sdim, sv);
                    uint2 pr = __ldg(&A.pairs[p]);
                    uint32_t i1 = pr.x, i2 = pr.y;
                    uint32_t t1 = __ldg(&A.o.type[i1]), t2 = __ldg(&A.o.type[i2]);
                    ma = load_iso(A.o, i1), mb = load_iso(A.o, i2);
                    ga = load_slim_support(A.o, A.H, i1, t1), gb = load_slim_support(A.o, A.H, i2, t2);
                    active = true;
                    status = epa_init_t<true>(e, ma, ga, mb, gb, sdim, sv, p1, p2, n, res_face);
                }
            }
        } else if (active) {
            status = epa_step_t(e, ma, ga, mb, gb, res_face);
        }
        bool ok = active && status == EPA_DONE_OK;
        bool defer = active && status == EPA_DONE_FAIL && e.overflow;
        bool fail = active && status == EPA_DONE_FAIL && !e.overflow;
        if (ok && res_face != EPA_RES_DIRECT) epa_result_from_face(e, res_face, p1, p2, n);
        uint32_t slot = queue_append(&A.cnt->cp_cursor[KEY], ok);
        if (ok) {
            cp_store(A.cp_queue, slot, p, p1, p2, n);
            if constexpr (PS) A.ps.dir[A.pair_index ? __ldg(&A.pair_index[p]) : p] = make_float4(n.x, n.y, n.z, 1.f);
        }
        // tier 1 defers to tier 2 (A.epa_long), tier 2 to the last resort (the second half of the same array)
        slot = queue_append(TIER == 1 ? &A.cnt->epa_long_n : &A.cnt->epa_defer_n, defer);
        if (defer) A.epa_long[(TIER == 1 ? 0u : A.cap_pairs) + slot] = wq;
        if (fail) {
            if (e.panicked) atomicAdd(&A.cnt->ref_panics, 1u);
            uint32_t out_index = A.pair_index ? __ldg(&A.pair_index[p]) : p;
            if constexpr (PS) {  // NoIntersection(x axis) (contact_support_map_support_map.rs:76)
                A.ps.dir[out_index] = make_float4(1.f, 0.f, 0.f, 1.f);
                uint2 pr = __ldg(&A.pairs[p]);
                pm_age_only(A.ps, out_index, pr.x, pr.y);
            } else {
                A.manifold_start[out_index] = 0;
                A.manifold_count[out_index] = 0;
            }
        }
        if (ok || defer || fail) active = false;
        if (exhausted && __all_sync(0xffffffffu, !active)) break;
    }
}

// The last-resort queue on the big local-memory store (48 / 192 / 160): pairs beyond tier 2, segment simplices.  Empty on every
// scene measured so far (the kernel then returns at once).
#ifndef NCB_EPA_MINBLOCKS
#define NCB_EPA_MINBLOCKS 16
#endif
template <bool PS>
__global__ void __launch_bounds__(64, NCB_EPA_MINBLOCKS) k_cc_epa_big(NarrowArgs A) {
    const int KEY = CCQ;
    const uint32_t seg_end = A.cnt->epa_defer_n;
    if (seg_end == 0) return;
    uint32_t* fetch = &A.cnt->epa_defer_fetch;
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const uint32_t lanes = min(32u, (seg_end + nwarps - 1) / nwarps);
    const bool usable = (uint32_t)lane < lanes;
    EpaState e;
    bool active = false, exhausted = false;
    uint32_t p = 0;
    Iso ma, mb;
    Support ga, gb;
    V3 p1, p2, n;
    for (;;) {
        int status = EPA_CONTINUE;
        unsigned idle = __ballot_sync(0xffffffffu, usable && !active);
        bool refill = !exhausted && idle != 0;
        if (refill) {  // warp-uniform
            uint32_t base = 0;
            int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(fetch, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + __popc(idle) >= seg_end) exhausted = true;  // nothing left after this batch
            if (usable && !active) {
                uint32_t w = base + __popc(idle & ((1u << lane) - 1));
                if (w < seg_end) {
                    uint32_t wq = __ldg(&A.epa_long[A.cap_pairs + w]);
                    int sdim;
                    CSOPoint sv[4];
                    epa_rec_load(A.epa_queue + (size_t)wq * EPA_REC_WORDS, p, sdim, sv);
                    uint2 pr = __ldg(&A.pairs[p]);
                    uint32_t i1 = pr.x, i2 = pr.y;
                    uint32_t t1 = __ldg(&A.o.type[i1]), t2 = __ldg(&A.o.type[i2]);
                    ma = load_iso(A.o, i1), mb = load_iso(A.o, i2);
                    Shape a = load_shape(A.o, A.H, i1, t1), b = load_shape(A.o, A.H, i2, t2);
                    ga = as_support(a), gb = as_support(b);
                    active = true;
                    status = epa_init(e, ma, ga, mb, gb, sdim, sv, p1, p2, n);
                }
            }
        } else if (active) {
            status = epa_step(e, ma, ga, mb, gb, p1, p2, n);
        }
        bool ok = active && status == EPA_DONE_OK;
        bool fail = active && status == EPA_DONE_FAIL;
        uint32_t slot = queue_append(&A.cnt->cp_cursor[KEY], ok);
        if (ok) {
            cp_store(A.cp_queue, slot, p, p1, p2, n);
            if constexpr (PS) A.ps.dir[A.pair_index ? __ldg(&A.pair_index[p]) : p] = make_float4(n.x, n.y, n.z, 1.f);
        }
        if (fail) {
            if (e.overflow) atomicAdd(&A.cnt->epa_overflow, 1u);
            if (e.panicked) atomicAdd(&A.cnt->ref_panics, 1u);
            uint32_t out_index = A.pair_index ? __ldg(&A.pair_index[p]) : p;
            if constexpr (PS) {  // NoIntersection(x axis) (contact_support_map_support_map.rs:76)
                A.ps.dir[out_index] = make_float4(1.f, 0.f, 0.f, 1.f);
                uint2 pr = __ldg(&A.pairs[p]);
                pm_age_only(A.ps, out_index, pr.x, pr.y);
            } else {
                A.manifold_start[out_index] = 0;
                A.manifold_count[out_index] = 0;
            }
        }
        if (ok || fail) active = false;
        if (exhausted && __all_sync(0xffffffffu, !active)) break;
    }
}

__global__ void k_snap_manifold_split(DevCounters* cnt) { cnt->man_split = cnt->cp_cursor[CCQ]; }

#ifndef NCB_MAN_MINBLOCKS
#define NCB_MAN_MINBLOCKS 6
#endif
template <bool PS>
__global__ void __launch_bounds__(128, NCB_MAN_MINBLOCKS) k_cc_manifold(NarrowArgs A) {
    const int KEY = CCQ;
    // The queue has two ranges: [key_start, man_split) was complete when the first EPA tier finished and is processed in
    // man_parts parts (part index < man_parts) while the later EPA tiers still append beyond it; [man_split, cp_cursor) is the
    // tail they produced (part index == man_parts, launched after they finished).
    uint32_t seg_begin = A.cnt->key_start[KEY];
    uint32_t seg_end = A.cnt->man_split;
    if (A.man_part >= A.man_parts) {
        seg_begin = seg_end;
        seg_end = A.cnt->cp_cursor[KEY];
    } else if (A.man_parts > 1) {
        uint32_t len = seg_end - seg_begin, per = (len + A.man_parts - 1) / A.man_parts;
        uint32_t b = seg_begin + min(len, per * (uint32_t)A.man_part);
        seg_end = seg_begin + min(len, per * (uint32_t)(A.man_part + 1));
        seg_begin = b;
    }
    uint32_t stride = gridDim.x * blockDim.x;
    ManifoldT<PS> mf;
    Kin kin_side[PS ? 1 : MANIFOLD_MAX];  // only touched when kinematics are wanted
    if constexpr (!PS) mf.kin = A.kin_out ? kin_side : nullptr;
    for (uint32_t base = seg_begin + blockIdx.x * blockDim.x; base < seg_end; base += stride) {
        uint32_t w = base + threadIdx.x;
        bool valid = w < seg_end;
        mf.n = 0;
        mf.deepest = 0;
        uint32_t p = 0, i1 = 0, i2 = 0;
        if (valid) {
            const uint32_t* q = A.cp_queue + (size_t)w * CP_REC_WORDS;
            const float* f = reinterpret_cast<const float*>(q);
            p = q[0];
            V3 p1 = v3(f[1], f[2], f[3]), p2 = v3(f[4], f[5], f[6]), dir = v3(f[7], f[8], f[9]);
            uint2 pr = __ldg(&A.pairs[p]);
            i1 = pr.x, i2 = pr.y;
            if constexpr (PS) pm_load_and_age(A.ps, A.pair_index ? __ldg(&A.pair_index[p]) : p, mf);
            uint32_t t1 = __ldg(&A.o.type[i1]), t2 = __ldg(&A.o.type[i2]);
            Iso ma = load_iso(A.o, i1), mb = load_iso(A.o, i2);
            float linear = __ldg(&A.o.qlimit[i1]) + __ldg(&A.o.qlimit[i2]);
            Shape a = load_shape(A.o, A.H, i1, t1), b = load_shape(A.o, A.H, i2, t2);
            float2 ang1 = __ldg(&A.o.ang_cs[i1 * A.o.ang_stride]), ang2 = __ldg(&A.o.ang_cs[i2 * A.o.ang_stride]);
            Feature f1, f2;
            convex_convex_manifold(ma, a, mb, b, linear, ang1, ang2, p1, p2, dir, mf, f1, f2);
        }
        uint32_t out_index = valid ? (A.pair_index ? __ldg(&A.pair_index[p]) : p) : 0;
        if constexpr (PS) {
            if (valid) pm_store(A.ps, out_index, mf, i1, i2);
        } else {
            write_manifold(mf, valid, p, out_index, A.contacts, A.cap_contacts, A.manifold_start, A.manifold_count, A.cnt, A.kin_out);
        }
    }
}

// One persistent kernel per key; the segment bounds are read from the device counters.
template <int KEY, bool PS>
__global__ void __launch_bounds__(128) k_narrow(NarrowArgs A) {
    uint32_t seg_begin = A.cnt->key_start[KEY];
    uint32_t seg_end = seg_begin + A.cnt->key_hist[KEY];
    uint32_t stride = gridDim.x * blockDim.x;
    // local-memory working set, only instantiated for the keys that need it
    ManifoldT<PS> mf;
    Kin kin_side[PS ? 1 : MANIFOLD_MAX];  // only touched when kinematics are wanted
    if constexpr (!PS) mf.kin = A.kin_out ? kin_side : nullptr;
    for (uint32_t base = seg_begin + blockIdx.x * blockDim.x; base < seg_end; base += stride) {
        uint32_t p = base + threadIdx.x;
        bool valid = p < seg_end;
        bool deferred = false;
        Simplex bh_simplex;
        mf.n = 0;
        mf.deepest = 0;
        uint32_t i1 = 0, i2 = 0;
        if (valid) {
            uint2 pr = __ldg(&A.pairs[p]);
            i1 = pr.x, i2 = pr.y;
            if constexpr (PS) pm_load_and_age(A.ps, A.pair_index ? __ldg(&A.pair_index[p]) : p, mf);
            uint32_t t1 = __ldg(&A.o.type[i1]), t2 = __ldg(&A.o.type[i2]);
            Iso ma = load_iso(A.o, i1), mb = load_iso(A.o, i2);
            float linear = __ldg(&A.o.qlimit[i1]) + __ldg(&A.o.qlimit[i2]);
            if (KEY == K_BALL_BALL) {
                gen_ball_ball(ma, __ldg(&A.o.param[i1]).x, mb, __ldg(&A.o.param[i2]).x, linear, mf);
            } else if (KEY == K_PLANE_BALL) {
                Shape a = load_shape(A.o, A.H, i1, t1), b = load_shape(A.o, A.H, i2, t2);
                if (t1 == NCB_SHAPE_PLANE)
                    gen_plane_ball(ma, a.he, mb, b.radius, linear, false, mf);
                else
                    gen_plane_ball(mb, b.he, ma, a.radius, linear, true, mf);
            } else if (KEY == K_PLANE_CUBOID || KEY == K_PLANE_HULL) {
                Shape a = load_shape(A.o, A.H, i1, t1), b = load_shape(A.o, A.H, i2, t2);
                Feature feat;
                if (t1 == NCB_SHAPE_PLANE)
                    gen_plane_convex(ma, a.he, mb, b, linear, false, mf, feat);
                else
                    gen_plane_convex(mb, b.he, ma, a, linear, true, mf, feat);
            } else if (KEY == K_BALL_CUBOID) {
                Shape a = load_shape(A.o, A.H, i1, t1), b = load_shape(A.o, A.H, i2, t2);
                bool flip = t1 != NCB_SHAPE_BALL;
                const Shape& ball = flip ? b : a;
                const Shape& cp = flip ? a : b;
                const Iso& mball = flip ? mb : ma;
                const Iso& mcp = flip ? ma : mb;
                bool inside;
                V3 world2;
                uint32_t f2;
                cuboid_project_point_with_feature(cp.he, mcp, mball.t, inside, world2, f2);
                gen_ball_convex_finish(mball.t, ball.radius, cp, mcp, inside, world2, f2, linear, flip, mf);
            } else if (KEY == K_BALL_HULL) {
                Shape a = load_shape(A.o, A.H, i1, t1), b = load_shape(A.o, A.H, i2, t2);
                bool flip = t1 != NCB_SHAPE_BALL;
                const Shape& ball = flip ? b : a;
                const Shape& cp = flip ? a : b;
                const Iso& mball = flip ? mb : ma;
                const Iso& mcp = flip ? ma : mb;
                HullProjSetup u = hull_proj_setup(cp.hull, mcp, mball.t);
                V3 world2;
                if (hull_project_gjk(u, mball.t, bh_simplex, world2) == GJK_CLOSEST_POINTS) {
                    uint32_t f2 = hull_project_feature(cp.hull, mcp, mball.t, false, world2, A.one_degree_cs);
                    gen_ball_convex_finish(mball.t, ball.radius, cp, mcp, false, world2, f2, linear, flip, mf);
                } else {
                    deferred = true;  // ball centre inside the hull: EPA, in k_bh_epa
                }
            }
        }
        if (KEY == K_BALL_HULL) {
            uint32_t slot = queue_append(&A.cnt->epa_cursor[K_BALL_HULL], deferred);
            if (deferred) {
                uint32_t* q = A.epa_queue + (size_t)slot * EPA_REC_WORDS;
                float* f = reinterpret_cast<float*>(q);
                q[0] = p;
                q[1] = (uint32_t)bh_simplex.dim;
                for (int i = 0; i < 4; ++i) {
                    f[2 + 6 * i + 0] = bh_simplex.v[i].orig1.x, f[2 + 6 * i + 1] = bh_simplex.v[i].orig1.y, f[2 + 6 * i + 2] = bh_simplex.v[i].orig1.z;
                    f[2 + 6 * i + 3] = bh_simplex.v[i].orig2.x, f[2 + 6 * i + 4] = bh_simplex.v[i].orig2.y, f[2 + 6 * i + 5] = bh_simplex.v[i].orig2.z;
                }
            }
        }
        uint32_t out_index = valid ? (A.pair_index ? __ldg(&A.pair_index[p]) : p) : 0;
        if constexpr (PS) {
            if (valid && !deferred) pm_store(A.ps, out_index, mf, i1, i2);  // deferred pairs are loaded again by k_bh_epa
        } else {
            write_manifold(mf, valid && !deferred, p, out_index, A.contacts, A.cap_contacts, A.manifold_start, A.manifold_count, A.cnt, A.kin_out);
        }
    }
}

// Ball x hull pairs whose ball centre is inside the hull: EPA::project_origin + the rest of the generator.
template <bool PS>
__global__ void __launch_bounds__(64) k_bh_epa(NarrowArgs A) {
    uint32_t seg_begin = A.cnt->key_start[K_BALL_HULL];
    uint32_t seg_end = A.cnt->epa_cursor[K_BALL_HULL];
    uint32_t stride = gridDim.x * blockDim.x;
    EpaState e;
    ManifoldT<PS> mf;
    Kin kin_side[PS ? 1 : MANIFOLD_MAX];  // only touched when kinematics are wanted
    if constexpr (!PS) mf.kin = A.kin_out ? kin_side : nullptr;
    for (uint32_t base = seg_begin + blockIdx.x * blockDim.x; base < seg_end; base += stride) {
        uint32_t w = base + threadIdx.x;
        bool valid = w < seg_end;
        mf.n = 0;
        mf.deepest = 0;
        uint32_t p = 0, h1 = 0, h2 = 0;
        if (valid) {
            int sdim;
            CSOPoint sv[4];
            epa_rec_load(A.epa_queue + (size_t)w * EPA_REC_WORDS, p, sdim, sv);
            uint2 pr = __ldg(&A.pairs[p]);
            uint32_t i1 = pr.x, i2 = pr.y;
            h1 = i1, h2 = i2;
            if constexpr (PS) pm_load_and_age(A.ps, A.pair_index ? __ldg(&A.pair_index[p]) : p, mf);
            uint32_t t1 = __ldg(&A.o.type[i1]), t2 = __ldg(&A.o.type[i2]);
            Iso ma = load_iso(A.o, i1), mb = load_iso(A.o, i2);
            float linear = __ldg(&A.o.qlimit[i1]) + __ldg(&A.o.qlimit[i2]);
            Shape a = load_shape(A.o, A.H, i1, t1), b = load_shape(A.o, A.H, i2, t2);
            bool flip = t1 != NCB_SHAPE_BALL;
            const Shape& ball = flip ? b : a;
            const Shape& cp = flip ? a : b;
            const Iso& mball = flip ? mb : ma;
            const Iso& mcp = flip ? ma : mb;
            HullProjSetup u = hull_proj_setup(cp.hull, mcp, mball.t);
            Iso id = iso_id();
            V3 p1, p2, d, world2;
            if (epa_closest_points(e, u.m, u.shape, id, u.origin, sdim, sv, p1, p2, d))
                world2 = p1 + mball.t;
            else {
                if (e.overflow) atomicAdd(&A.cnt->epa_overflow, 1u);
                if (e.panicked) atomicAdd(&A.cnt->ref_panics, 1u);
                world2 = mball.t;
            }
            uint32_t f2 = hull_project_feature(cp.hull, mcp, mball.t, true, world2, A.one_degree_cs);
            gen_ball_convex_finish(mball.t, ball.radius, cp, mcp, true, world2, f2, linear, flip, mf);
        }
        uint32_t out_index = valid ? (A.pair_index ? __ldg(&A.pair_index[p]) : p) : 0;
        if constexpr (PS) {
            if (valid) pm_store(A.ps, out_index, mf, h1, h2);
        } else {
            write_manifold(mf, valid, p, out_index, A.contacts, A.cap_contacts, A.manifold_start, A.manifold_count, A.cnt, A.kin_out);
        }
    }
}

// K_NONE pairs (plane x plane) and pairs beyond a key segment get an empty manifold.
__global__ void __launch_bounds__(256) k_narrow_none(NarrowArgs A) {
    uint32_t seg_begin = A.cnt->key_start[K_NONE];
    uint32_t seg_end = seg_begin + A.cnt->key_hist[K_NONE];
    for (uint32_t p = seg_begin + blockIdx.x * blockDim.x + threadIdx.x; p < seg_end; p += gridDim.x * blockDim.x) {
        uint32_t out_index = A.pair_index ? A.pair_index[p] : p;
        A.manifold_start[out_index] = 0;
        A.manifold_count[out_index] = 0;
    }
}

// Pairs with a capsule: the five capsule key segments are adjacent, one launch walks them all (one pair per thread, the whole
// generator in the thread: capsule.cuh).  Worlds without capsules never launch it.
template <bool PS>
__global__ void __launch_bounds__(64) k_capsule(NarrowArgs A) {
    uint32_t seg_begin = A.cnt->key_start[K_CAPSULE_BALL];
    uint32_t seg_end = A.cnt->key_start[K_CAPSULE_HULL] + A.cnt->key_hist[K_CAPSULE_HULL];
    uint32_t stride = gridDim.x * blockDim.x;
    EpaState e;
    ManifoldT<PS> mf;
    Kin kin_side[PS ? 1 : MANIFOLD_MAX];  // only touched when kinematics are wanted
    if constexpr (!PS) mf.kin = A.kin_out ? kin_side : nullptr;
    for (uint32_t base = seg_begin + blockIdx.x * blockDim.x; base < seg_end; base += stride) {
        uint32_t p = base + threadIdx.x;
        bool valid = p < seg_end;
        mf.n = 0;
        mf.deepest = 0;
        uint32_t out_index = 0;
        if (valid) {
            uint2 pr = __ldg(&A.pairs[p]);
            out_index = A.pair_index ? __ldg(&A.pair_index[p]) : p;
            capsule_pair<PS>(A.o, A.H, A.ps, out_index, e, mf, pr.x, pr.y, &A.cnt->epa_overflow, &A.cnt->ref_panics);
        }
        if constexpr (!PS) write_manifold(mf, valid, p, out_index, A.contacts, A.cap_contacts, A.manifold_start, A.manifold_count, A.cnt, A.kin_out);
    }
}

template <bool PS>
static cudaError_t launch_narrow_phase_t(ncb_ctx* c, const DevObjects& o, const uint2* pairs, const uint32_t* pair_index, uint32_t cap_pairs,
                                         uint32_t cap_contacts, const PersistArgs* ps) {
    NarrowArgs A;
    memset(&A.ps, 0, sizeof A.ps);
    if (ps) A.ps = *ps;
    A.o = o;
    A.H = c->hulls;
    A.pairs = pairs;
    A.pair_index = pair_index;
    A.contacts = c->contacts.p;
    A.kin_out = (!PS && c->want_kinematics) ? c->kinematics.p : nullptr;
    A.cap_contacts = cap_contacts;
    A.manifold_start = c->manifold_start.p;
    A.manifold_count = c->manifold_count.p;
    A.cnt = c->counters.p;
    A.cap_pairs = cap_pairs;
    A.man_part = 0, A.man_parts = 1;
    static int refill_min = getenv("NCB_EPA_REFILL") ? atoi(getenv("NCB_EPA_REFILL")) : 16;
    A.epa_refill_min = refill_min;
    A.epa_queue = c->epa_queue.p;
    A.cp_queue = c->cp_queue.p;
    A.epa_long = c->epa_long.p;
    {
        float one_degree = (float)(3.14159265358979323846 / 180.0);
        A.one_degree_cs = make_float2(cosf(one_degree), sinf(one_degree));
    }
    cudaStream_t s = c->stream;
    int sm = c->sm_count;
    // tuning knobs (CTAs per SM of the persistent kernels); defaults chosen from ncu runs, see profiles/
    static int gjk_bpsm = getenv("NCB_GJK_BPSM") ? atoi(getenv("NCB_GJK_BPSM")) : 6;
    static int epa_bpsm = getenv("NCB_EPA_BPSM") ? atoi(getenv("NCB_EPA_BPSM")) : 2;  // overflow kernel (k_cc_epa_big): small on purpose
    static int man_bpsm = getenv("NCB_MAN_BPSM") ? atoi(getenv("NCB_MAN_BPSM")) : 6;
    // Two independent chains: the convex-convex phases on the context's stream, everything else on a side stream
    // (each persistent kernel alone leaves most issue slots idle; together they overlap).
    cudaStream_t s2 = c->side_stream ? c->side_stream : s;
    if (c->side_stream) {
        cudaEventRecord(c->ev_fork, s);
        cudaStreamWaitEvent(s2, c->ev_fork, 0);
    }
    k_narrow<K_BALL_HULL, PS><<<sm * 8, 128, 0, s2>>>(A);
    k_bh_epa<PS><<<sm * 4, 64, 0, s2>>>(A);
    k_narrow<K_BALL_CUBOID, PS><<<sm * 8, 128, 0, s2>>>(A);
    k_narrow<K_BALL_BALL, PS><<<sm * 8, 128, 0, s2>>>(A);
    k_narrow<K_PLANE_BALL, PS><<<sm * 2, 128, 0, s2>>>(A);
    k_narrow<K_PLANE_CUBOID, PS><<<sm * 2, 128, 0, s2>>>(A);
    k_narrow<K_PLANE_HULL, PS><<<sm * 2, 128, 0, s2>>>(A);
    if (!PS) k_narrow_none<<<sm, 256, 0, s2>>>(A);
    if (c->has_capsules) k_capsule<PS><<<sm * 8, 64, 0, s2>>>(A);
    if (!PS && c->has_prox && c->prox.p) launch_proximity_segments(c, o, pairs, pair_index, s2);  // sensor pairs (proximity.cu)
    if (c->side_stream) cudaEventRecord(c->ev_join, s2);
    k_cc_gjk<PS><<<sm * gjk_bpsm, 128, 0, s>>>(A);
    timer_mark(c, "cc_gjk", 1);
    // ncb_world_update (host buffers): the side chain joins here; from this point until k_cc_manifold no kernel allocates
    // contact slots, so a snapshot of the counters tells which contacts are final and they are copied to the host while
    // the EPA / manifold phases run.
    const bool early = !PS && c->early.active;
    if (early) {
        if (c->side_stream) cudaStreamWaitEvent(s, c->ev_join, 0);
        cudaMemcpyAsync(c->snap.p, &c->counters.p->n_contacts, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
        cudaEventRecord(c->ev_snap[0], s);
    }
    static int epas_bpsm = getenv("NCB_EPAS_BPSM") ? atoi(getenv("NCB_EPAS_BPSM")) : NCB_EPAS_MINBLOCKS;
    {
        const size_t smem1 = (size_t)EpaTierTraits<1>::Store::WORDS * EPAS_THREADS * sizeof(uint32_t);
        const size_t smem2 = (size_t)EpaTierTraits<2>::Store::WORDS * EPAT2_THREADS * sizeof(uint32_t);
        static bool attr_set[2] = {false, false};
        if (!attr_set[PS]) {
            cudaFuncSetAttribute(k_cc_epa_tier<PS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
            cudaFuncSetAttribute(k_cc_epa_tier<PS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            attr_set[PS] = true;
        }
        k_cc_epa_tier<PS, 1><<<sm * epas_bpsm, EPAS_THREADS, smem1, s>>>(A);
        // The later tiers hold ~1 % of the pairs in a few long runs (0.22 ms at 4 % occupancy).  NCB_EPA_TIERS_ASIDE=1 runs them on
        // the side stream, followed there by a short manifold launch for what they append, beside the main manifold kernel (which
        // works through what the queue held when tier 1 finished, man_split).  Measured and left off (profiles/r2_epa_tiers_aside.txt):
        // the persistent manifold grid fills every SM, so the side chain only starts when it drains, and every extra launch of
        // these kernels costs the latency of one pair's run (~0.15 ms): 3.77 ms per step instead of 3.65.
        static const bool aside_ok = getenv("NCB_EPA_TIERS_ASIDE") != nullptr;
        const bool aside = aside_ok && c->side_stream && !early;
        cudaStream_t st = aside ? c->side_stream : s;
        if (aside) {
            k_snap_manifold_split<<<1, 1, 0, s>>>(A.cnt);
            cudaEventRecord(c->ev_tier1, s);
            cudaStreamWaitEvent(st, c->ev_tier1, 0);
        }
        k_cc_epa_tier<PS, 2><<<sm * 3, EPAT2_THREADS, smem2, st>>>(A);
        k_cc_epa_big<PS><<<sm * epa_bpsm, 64, 0, st>>>(A);  // last resort, normally an empty queue
        if (aside) {
            A.man_part = 1, A.man_parts = 1;  // the tail: [man_split, cp_cursor)
            k_cc_manifold<PS><<<sm * 2, 128, 0, st>>>(A);
            cudaEventRecord(c->ev_tier2, st);
        } else {
            k_snap_manifold_split<<<1, 1, 0, s>>>(A.cnt);  // everything is in the queue: no tail
        }
        timer_mark(c, "cc_epa", aside ? 2 : 4);
        if (early) {
            // the manifold queue in NCB_MAN_PARTS parts, a snapshot of the contact counter after each: the copy stream ships the
            // contacts of a finished part while the next one computes (api.cu, update_after_aabbs)
            for (int part = 0; part < NCB_MAN_PARTS; ++part) {
                A.man_part = part, A.man_parts = NCB_MAN_PARTS;
                k_cc_manifold<PS><<<sm * man_bpsm, 128, 0, s>>>(A);
                cudaMemcpyAsync(c->snap.p + 1 + part, &c->counters.p->n_contacts, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
                cudaEventRecord(c->ev_snap[1 + part], s);
            }
            timer_mark(c, "cc_manifold", NCB_MAN_PARTS);
        } else {
            A.man_part = 0, A.man_parts = 1;
            k_cc_manifold<PS><<<sm * man_bpsm, 128, 0, s>>>(A);
            if (aside) cudaStreamWaitEvent(s, c->ev_tier2, 0);
            timer_mark(c, "cc_manifold", aside ? 4 : 1);
        }
    }
    if (!early && c->side_stream) cudaStreamWaitEvent(s, c->ev_join, 0);  // device-only updates: the side chain may run to the end
    timer_mark(c, "narrow_other_join", 7);
    return cudaGetLastError();
}
cudaError_t launch_narrow_phase(ncb_ctx* c, const DevObjects& o, const uint2* pairs, const uint32_t* pair_index, uint32_t cap_pairs,
                                uint32_t cap_contacts) {
    return launch_narrow_phase_t<false>(c, o, pairs, pair_index, cap_pairs, cap_contacts, nullptr);
}
// Stepping world: the manifolds / generator directions live in the persistent arrays of `ps`, indexed by pair_index[p].
cudaError_t launch_narrow_phase_persistent(ncb_ctx* c, const DevObjects& o, const uint2* pairs, const uint32_t* pair_index, uint32_t cap_pairs,
                                           const PersistArgs& ps) {
    return launch_narrow_phase_t<true>(c, o, pairs, pair_index, cap_pairs, 0, &ps);
}

// Classify caller-provided pairs (ncb_generate_contacts): key per pair from the two shape types.
__global__ void __launch_bounds__(256) k_classify_pairs(const uint2* __restrict__ pairs, uint32_t n, const uint32_t* __restrict__ type,
                                                        uint8_t* __restrict__ keys, DevCounters* cnt) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) cnt->n_pairs = n;
    if (p >= n) return;
    uint2 pr = pairs[p];
    keys[p] = (uint8_t)pair_key(type[pr.x], type[pr.y]);
}
cudaError_t launch_classify_pairs(ncb_ctx* c, const uint2* pairs, uint32_t n) {
    if (n == 0) return cudaSuccess;
    k_classify_pairs<<<(n + 255) / 256, 256, 0, c->stream>>>(pairs, n, c->type.p, c->keys_raw.p, c->counters.p);
    return cudaGetLastError();
}

#endif  // NCB_HOST_SHIM

}  // namespace ncb
