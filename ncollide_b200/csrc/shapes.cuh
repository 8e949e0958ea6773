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

}  // namespace ncb
