// Shape records as the kernels read them (shared by narrow.cu and query.cu).
#pragma once
#include "gjk.cuh"
#include "ncb_internal.h"
#include "vec.cuh"

namespace ncb {

#define FID(kind, id) ((((uint32_t)(kind)) << 30) | ((uint32_t)(id)&0x3fffffffu))
#define FID_KIND(f) ((f) >> 30)
#define FID_ID(f) ((f)&0x3fffffffu)
#define FID_UNKNOWN 0xc0000000u
#define FACE0 0x80000000u

struct Shape {
    uint32_t type;
    float radius;
    V3 he;  // cuboid half extents / plane normal
    HullView hull;
};

NCB_HD Shape load_shape(const DevObjects& o, const DevHulls& H, uint32_t i, uint32_t type) {
    Shape s;
    float4 p = __ldg(&o.param[i]);
    s.type = type;
    s.radius = p.x;
    s.he = v3(p.x, p.y, p.z);
    if (type == NCB_SHAPE_CONVEX_HULL) s.hull = hull_view(H, (uint32_t)p.x);
    return s;
}
NCB_HD Iso load_iso(const DevObjects& o, uint32_t i) {
    float4 q = __ldg(&o.rot[i]);
    Iso m;
    m.t = v3(__ldg(o.pos + 3 * i), __ldg(o.pos + 3 * i + 1), __ldg(o.pos + 3 * i + 2));
    m.q = Quat{q.x, q.y, q.z, q.w};
    return m;
}
NCB_HD Support as_support(const Shape& s) {
    Support g;
    g.kind = s.type == NCB_SHAPE_CUBOID ? 0 : 1;
    g.he = s.he;
    g.hull = s.hull;
    return g;
}

// The slim operand of the EPA kernels straight from the object arrays (no HullView with its 13 table pointers).
NCB_HD SupportS load_slim_support(const DevObjects& o, const DevHulls& H, uint32_t i, uint32_t type) {
    SupportS g;
    float4 p = __ldg(&o.param[i]);
    g.kind = type == NCB_SHAPE_CUBOID ? 0 : 1;
    g.he = v3(p.x, p.y, p.z);
    g.nverts = 0;
#ifndef NCB_HOST_SHIM
    g.pts4 = nullptr;
#else
    g.pts = nullptr;
#endif
    if (type == NCB_SHAPE_CONVEX_HULL) {
        uint32_t h = (uint32_t)p.x;
        uint32_t v0 = __ldg(H.vert_off + h);
        g.nverts = __ldg(H.vert_off + h + 1) - v0;
#ifndef NCB_HOST_SHIM
        g.pts4 = H.points4 + v0;
#else
        g.pts = H.points + 3 * (size_t)v0;
#endif
    }
    return g;
}

// ConvexHull::project_point_with_feature (point_support_map.rs:15-53 with solid = false, :97-116), in two parts so that
// the rare "point inside the hull" case (EPA) can be deferred to its own compacted kernel.
struct HullProjSetup {
    Iso m;  // Translation::from(-point) * m
    Support shape, origin;
};
NCB_HD HullProjSetup hull_proj_setup(const HullView& H, const Iso& m_in, V3 point) {
    HullProjSetup u;
    u.m = m_in;
    u.m.t = (-point) + m_in.t;
    u.shape.kind = 1;
    u.shape.hull = H;
    u.origin.kind = 2;
    return u;
}
NCB_HD Iso iso_id() {
    Iso id;
    id.t = v3(0.f, 0.f, 0.f);
    id.q = Quat{0.f, 0.f, 0.f, 1.f};
    return id;
}
// gjk::project_origin: GJK_CLOSEST_POINTS (outside, proj set) or GJK_INTERSECTION (inside: simplex s feeds EPA)
static __device__ __noinline__ int hull_project_gjk(const HullProjSetup& u, V3 point, Simplex& s, V3& proj) {
    Iso id = iso_id();
    V3 dir;
    if (!unit_try_new(-u.m.t, NCB_EPS, dir)) dir = v3(1.f, 0.f, 0.f);
    V3 p1, p2, d;
    int r = gjk_closest_points(u.m, u.shape, id, u.origin, NCB_FMAX, dir, s, p1, p2, d);
    if (r == GJK_CLOSEST_POINTS) proj = p1 + point;
    return r == GJK_CLOSEST_POINTS ? GJK_CLOSEST_POINTS : GJK_INTERSECTION;
}
}  // namespace ncb
