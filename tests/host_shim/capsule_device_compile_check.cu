// TEST INFRASTRUCTURE ONLY — device-compile check of the STAGED capsule functions (ncollide_b200/csrc/capsule.cuh): includes the narrow-phase
// translation unit, the staged header, and one kernel that instantiates every staged template for the fresh (P = false) and the
// persistent (P = true) manifold, so that nvcc (sm_100a, --fmad=false) sees them as device code.  Not linked into anything.
#include "narrow.cu"
#include "capsule.cuh"

namespace ncb {
template <bool P>
__device__ void capsule_touch(const Iso& ma, const Iso& mb, const CapOperand& a, const CapOperand& b, ManifoldT<P>& mf, EpaState& e, float2 ang) {
    Feature f1, f2;
    V3 p1, p2, dir, d0 = v3(1.f, 0.f, 0.f);
    Simplex s;
    Support ga = cap_support(a), gb = cap_support(b);
    int r = gjk_closest_points(ma, ga, mb, gb, 0.1f, d0, s, p1, p2, dir);
    if (r == GJK_INTERSECTION && epa_closest_points(e, ma, ga, mb, gb, s.dim, s.v, p1, p2, dir)) r = GJK_CLOSEST_POINTS;
    if (r == GJK_CLOSEST_POINTS) capsule_convex_manifold(ma, a, mb, b, 0.1f, ang, ang, p1, p2, dir, mf, f1, f2);
    gen_ball_segment(ma, 0.5f, mb, b.hh, 0.1f, false, b.pre, mf);
    gen_plane_segment(ma, v3(0.f, 1.f, 0.f), mb, b.hh, 0.1f, true, b.pre, mf, f1);
}
__global__ void k_capsule_compile_check(const float* seg_pts, float* out, PersistArgs ps) {
    Iso ma = iso_id(), mb = iso_id();
    mb.t = v3(0.3f, 0.1f, 0.f);
    CapOperand a, b;
    a.is_segment = b.is_segment = true;
    a.hh = b.hh = 0.5f;
    a.seg_pts = seg_pts, b.seg_pts = seg_pts + 6;
    a.pre = CapsulePre{true, 0.2f}, b.pre = CapsulePre{true, 0.25f};
    a.shape.type = b.shape.type = NCB_SHAPE_CUBOID;
    EpaState e;
    Manifold mf;
    mf.n = 0, mf.deepest = 0;
    capsule_touch<false>(ma, mb, a, b, mf, e, make_float2(1.f, 0.f));
    PManifold pm;
    pm_load_and_age(ps, 0, pm);
    capsule_touch<true>(ma, mb, a, b, pm, e, make_float2(1.f, 0.f));
    pm_store(ps, 0, pm, 0, 1);
    V3 lo, hi;
    capsule_aabb(ma, 0.5f, 0.2f, lo, hi);
    out[0] = (float)mf.n + lo.x + hi.y + (float)pm.n;
}
}  // namespace ncb
