// Capsules, per pair: the CapsuleCapsule / CapsuleShape contact generators as device functions, run by k_capsule
// (narrow.cu) over the five capsule key segments of the sorted pair array, one pair per thread.  The per-pair functions are also
// compiled for the host and checked bit for bit against the oracle (tests/host_shim/narrow_host.cpp).
// Include after the definitions of narrow.cu (Feature, clip, manifold_push, gjk.cuh).
//
// Replaces (reference, file:line):
//   CapsuleCapsuleManifoldGenerator     contact_generator/capsule_capsule_manifold_generator.rs:24-55
//   CapsuleShapeManifoldGenerator       contact_generator/capsule_shape_manifold_generator.rs:23-75 (the sub-detector the dispatcher picks
//                                       for (segment, other): ball / plane / convex polyhedron generators)
//   Capsule::segment, contact_preprocessor   shape/capsule.rs:53-61,87-132
//   Segment as SupportMap / ConvexPolyhedron shape/segment.rs:182-190,237-341; query/point/point_segment.rs:14-91
//
// The capsule is replaced by its segment, the linear prediction grows by the radius, and every contact the sub-detector produces
// goes through the capsule's ContactPreprocessor before ContactManifold::push.  GJK / EPA need no new support kind: the segment is
// the 2-point "hull" [b, a] (in that order the hull scan's first-maximum rule returns `a` exactly when a.dir > b.dir, which is
// Segment::local_support_point); `seg_pts` points at those six floats (b then a) in GLOBAL memory (HullView reads through __ldg):
// DevObjects::cap_pts holds them for every object of a world with capsules (k_fill_cap_pts).
#pragma once

namespace ncb {

// ContactPreprocessor of a capsule side (capsule.rs:98-132); active = false: no preprocessor on that side
struct CapsulePre {
    bool active;
    float radius;
};
// returns false when the contact must be ignored (FeatureId::Unknown)
NCB_HD bool capsule_preprocess(const CapsulePre& pp, V3& w1, V3& w2, V3 n, float& depth, uint32_t& f1, uint32_t& f2, bool is_first, Kin* kin) {
    if (!pp.active) return true;
    uint32_t f = is_first ? f1 : f2, actual;
    uint32_t kind = FID_KIND(f);
    if (kind == NCB_FEATURE_VERTEX)
        actual = FID(NCB_FEATURE_FACE, FID_ID(f));
    else if (kind == NCB_FEATURE_EDGE)
        actual = FID(NCB_FEATURE_FACE, 2);
    else if (kind == NCB_FEATURE_FACE)
        actual = FID(NCB_FEATURE_FACE, 2 + FID_ID(f));
    else
        return false;
    if (is_first) {
        f1 = actual;
        if (kin) kin->dil1 = pp.radius;  // kinematic.set_dilation1
        w1 = w1 + n * pp.radius;
        depth += pp.radius;
    } else {
        f2 = actual;
        if (kin) kin->dil2 = pp.radius;
        w2 = w2 - n * pp.radius;
        depth += pp.radius;
    }
    return true;
}
// ContactManifold::push(contact, kinematic, tracking_pt, preprocessor1, preprocessor2) (contact_manifold.rs:165-181)
template <bool P>
NCB_HD void manifold_push_pp(ManifoldT<P>& mf, V3 w1, V3 w2, V3 n, float depth, uint32_t f1, uint32_t f2, V3 tracking_pt, const CapsulePre& pp1,
                             const CapsulePre& pp2, Kin* kin = nullptr) {
    if (!capsule_preprocess(pp1, w1, w2, n, depth, f1, f2, true, kin)) return;
    if (!capsule_preprocess(pp2, w1, w2, n, depth, f1, f2, false, kin)) return;
    manifold_push(mf, w1, w2, n, depth, f1, f2, tracking_pt, kin);
}

// ---- Segment a = (0, -hh, 0), b = (0, hh, 0) as a ConvexPolyhedron (segment.rs, dim3) -----------------------------------------
NCB_HD void segment_support_face_toward(float hh, const Iso& m, Feature& out) {  // :286-299
    feat_clear(out);
    feat_push(out, v3(0.f, -hh, 0.f), FID(NCB_FEATURE_VERTEX, 0));
    feat_push(out, v3(0.f, hh, 0.f), FID(NCB_FEATURE_VERTEX, 1));
    feat_push_edge(out, FID(NCB_FEATURE_EDGE, 0));
    out.feature_id = FID(NCB_FEATURE_EDGE, 0);
    feat_transform(out, m);
}
// ang_cs = (cos, sin) of the angular prediction, evaluated on the host with libm like every angle on this path
NCB_HD void segment_support_feature_toward(float hh, const Iso& m, V3 dir, float2 ang_cs, Feature& out) {  // :301-341
    feat_clear(out);
    V3 a = iso_mul_point(m, v3(0.f, -hh, 0.f)), b = iso_mul_point(m, v3(0.f, hh, 0.f));
    float ceps = ang_cs.y;
    V3 seg_dir;
    if (!unit_try_new(b - a, NCB_EPS, seg_dir)) return;
    float cang = dot(dir, seg_dir);
    if (cang > ceps) {
        out.feature_id = FID(NCB_FEATURE_VERTEX, 1);
        feat_push(out, b, FID(NCB_FEATURE_VERTEX, 1));
    } else if (cang < -ceps) {
        out.feature_id = FID(NCB_FEATURE_VERTEX, 0);
        feat_push(out, a, FID(NCB_FEATURE_VERTEX, 0));
    } else {
        feat_push(out, a, FID(NCB_FEATURE_VERTEX, 0));
        feat_push(out, b, FID(NCB_FEATURE_VERTEX, 1));
        feat_push_edge(out, FID(NCB_FEATURE_EDGE, 0));
        out.feature_id = FID(NCB_FEATURE_EDGE, 0);
    }
}
NCB_HD V3 segment_feature_normal(float hh, uint32_t f) {  // :237-284
    V3 direction;
    if (!unit_try_new(v3(0.f, hh, 0.f) - v3(0.f, -hh, 0.f), NCB_EPS, direction)) return v3(0.f, 1.f, 0.f);
    uint32_t kind = FID_KIND(f);
    if (kind == NCB_FEATURE_VERTEX) return FID_ID(f) == 0 ? direction : -direction;
    if (kind == NCB_FEATURE_EDGE) {
        int iamin = 0;  // first component of smallest absolute value
        for (int k = 1; k < 3; ++k)
            if (fabsf(vget(direction, k)) < fabsf(vget(direction, iamin))) iamin = k;
        V3 normal = v3(0.f, 0.f, 0.f);
        vset(normal, iamin, 1.f);
        normal = normal - direction * vget(direction, iamin);
        return normalize(normal);
    }
    return FID_ID(f) == 0 ? v3(direction.y, -direction.x, 0.f) : v3(-direction.y, direction.x, 0.f);
}
// PointQuery::project_point_with_feature for Segment (point_segment.rs:14-40,50-91)
NCB_HD void segment_project_point_with_feature(float hh, const Iso& m, V3 pt, bool& inside, V3& proj_out, uint32_t& feature) {
    V3 a = v3(0.f, -hh, 0.f), b = v3(0.f, hh, 0.f);
    V3 ls_pt = iso_inv_point(m, pt);
    V3 ab = b - a, ap = ls_pt - a;
    float ab_ap = dot(ab, ap), sqnab = norm_squared(ab);
    V3 proj;
    if (ab_ap <= 0.f) {
        feature = FID(NCB_FEATURE_VERTEX, 0);
        proj = iso_mul_point(m, a);
    } else if (ab_ap >= sqnab) {
        feature = FID(NCB_FEATURE_VERTEX, 1);
        proj = iso_mul_point(m, b);
    } else {
        float u = ab_ap / sqnab;
        feature = FID(NCB_FEATURE_EDGE, 0);
        proj = iso_mul_point(m, a + ab * u);
    }
    inside = relative_eq(proj.x, pt.x) && relative_eq(proj.y, pt.y) && relative_eq(proj.z, pt.z);
    proj_out = proj;
}

// A convex polyhedron operand of the capsule generators: a cuboid / hull Shape, or the segment of a capsule
struct CapOperand {
    bool is_segment;
    float hh;              // segment half height
    const float* seg_pts;  // the 2-point hull [b, a] of the segment, 6 floats in GLOBAL memory
    Shape shape;           // cuboid / hull when !is_segment
    CapsulePre pre;
};
NCB_HD Support cap_support(const CapOperand& c) {
    if (!c.is_segment) return as_support(c.shape);
    Support g;
    g.kind = 1;
    g.he = v3(0.f, 0.f, 0.f);
    g.hull = c.shape.hull;  // irrelevant fields
    g.hull.nv = 2;
    g.hull.pts = c.seg_pts;
    return g;
}
NCB_HD void cap_support_face_toward(const CapOperand& c, const Iso& m, V3 dir, Feature& out) {
    if (c.is_segment)
        segment_support_face_toward(c.hh, m, out);
    else
        support_face_toward(c.shape, m, dir, out);
}
NCB_HD void cap_support_feature_toward(const CapOperand& c, const Iso& m, V3 dir, float2 ang, Feature& out) {
    if (c.is_segment)
        segment_support_feature_toward(c.hh, m, dir, ang, out);
    else
        support_feature_toward(c.shape, m, dir, ang, out);
}

// clip_flush with the preprocessors (narrow.cu's clip_flush pushes without them)
template <bool P>
NCB_HD void clip_flush_pp(ClipCtxT<P>& cc, const CapsulePre& pp1, const CapsulePre& pp2) {
    for (int k = 0; k < cc.n_buf; ++k) {
        const ClipCand& c = cc.buf[k];
        if (!feature_ok_for_manifold(*cc.m1, c.f1)) continue;
        if (!feature_ok_for_manifold(*cc.m2, c.f2)) continue;
        float depth = -dot(cc.normal, c.w2 - c.w1);
        V3 local1 = iso_inv_point(*cc.ma, c.w1);
        Kin kin = kin_zero();
        if (cc.mf->wants_kin()) kin = kin_from_features(*cc.m1, *cc.m2, *cc.ma, *cc.mb, c.w1, c.w2, c.f1, c.f2, local1);
        manifold_push_pp(*cc.mf, c.w1, c.w2, cc.normal, depth, c.f1, c.f2, local1, pp1, pp2, &kin);
    }
    cc.n_buf = 0;
}
// ConvexPolyhedronConvexPolyhedronManifoldGenerator after GJK / EPA, for operands that may be capsule segments (the sub-detector
// of CapsuleCapsule / CapsuleShape).  `linear` already includes the capsule radii.  Returns false if clip's candidate buffer spilled
// (impossible with <= 2-vertex segment features against <= 16-vertex faces; reported, never silent).
template <bool P>
__device__ __noinline__ bool capsule_convex_manifold(const Iso& ma, const CapOperand& a, const Iso& mb, const CapOperand& b, float linear, float2 ang1,
                                                     float2 ang2, V3 p1, V3 p2, V3 dir, ManifoldT<P>& mf, Feature& m1, Feature& m2) {
    float depth = -dot(dir, p2 - p1);
    if (depth > 0.f) {
        cap_support_face_toward(a, ma, dir, m1);
        cap_support_face_toward(b, mb, -dir, m2);
    } else {
        cap_support_feature_toward(a, ma, dir, ang1, m1);
        cap_support_feature_toward(b, mb, -dir, ang2, m2);
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
    bool spilled = cc.n_new != cc.n_buf;
    clip_flush_pp(cc, a.pre, b.pre);
    if (cc.n_new == 0) {
        if (feature_ok_for_manifold(m1, m1.feature_id) && feature_ok_for_manifold(m2, m2.feature_id)) {
            V3 local1 = iso_inv_point(ma, p1);
            Kin k = kin_zero();
            if (mf.wants_kin()) k = kin_from_features(m1, m2, ma, mb, p1, p2, m1.feature_id, m2.feature_id, local1);
            manifold_push_pp(mf, p1, p2, dir, depth, m1.feature_id, m2.feature_id, local1, a.pre, b.pre, &k);
        }
    }
    return !spilled;
}

// BallConvexPolyhedronManifoldGenerator with the capsule's segment as the polyhedron (ball_convex_polyhedron_manifold_generator.rs:
// 28-122).  (mball, radius) the ball, (mseg, hh) the segment, seg_pre the capsule's preprocessor; flip: the segment is object 1.
template <bool P>
NCB_HD void gen_ball_segment(const Iso& mball, float radius, const Iso& mseg, float hh, float prediction, bool flip, const CapsulePre& seg_pre,
                             ManifoldT<P>& mf) {
    const CapsulePre none = {false, 0.f};
    V3 ball_center = mball.t;
    bool inside;
    V3 world2;
    uint32_t f2;
    segment_project_point_with_feature(hh, mseg, ball_center, inside, world2, f2);
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
        normal = -segment_feature_normal(hh, f2);
    }
    if (depth >= -prediction) {
        V3 world1 = ball_center + normal * radius;
        Kin k = kin_zero();
        if (mf.wants_kin()) k = kin_ball_polyhedron(radius, mseg, world2, normal, f2, v3(0.f, -hh, 0.f), v3(0.f, hh, 0.f), flip);  // Segment::edge = (a, b)
        if (!flip)
            manifold_push_pp(mf, world1, world2, normal, depth, FACE0, f2, v3(0.f, 0.f, 0.f), none, seg_pre, &k);
        else
            manifold_push_pp(mf, world2, world1, -normal, depth, f2, FACE0, v3(0.f, 0.f, 0.f), seg_pre, none, &k);
    }
}
// PlaneConvexPolyhedronManifoldGenerator with the capsule's segment (plane_convex_polyhedron_manifold_generator.rs:29-81)
template <bool P>
NCB_HD void gen_plane_segment(const Iso& mplane, V3 plane_n, const Iso& mseg, float hh, float prediction, bool flip, const CapsulePre& seg_pre,
                              ManifoldT<P>& mf, Feature& feat) {
    const CapsulePre none = {false, 0.f};
    V3 n = iso_mul_vec(mplane, plane_n);
    V3 pc = mplane.t;
    segment_support_face_toward(hh, mseg, feat);
    for (int i = 0; i < feat.nv; ++i) {
        V3 world2 = feat.v[i];
        float dist = dot(world2 - pc, n);
        if (dist <= prediction) {
            V3 world1 = world2 + (-n * dist);
            V3 local2 = iso_inv_point(mseg, world2);
            uint32_t f2 = feat.vid[i];
            Kin k = kin_zero();
            if (mf.wants_kin()) {
                V3 local1 = iso_inv_point(mplane, world1);
                if (!flip)
                    k.local1 = local1, k.g1 = G_PLANE, k.dir1 = plane_n, k.local2 = local2;
                else
                    k.local1 = local2, k.local2 = local1, k.g2 = G_PLANE, k.dir2 = plane_n;
            }
            if (!flip)
                manifold_push_pp(mf, world1, world2, n, -dist, FACE0, f2, local2, none, seg_pre, &k);
            else
                manifold_push_pp(mf, world2, world1, -n, -dist, f2, FACE0, local2, seg_pre, none, &k);
        }
    }
}

// One pair with at least one capsule, from the object arrays to its manifold: what k_capsule runs per thread.
//   PS = false: `mf` (cleared by the caller) receives the contacts; the caller writes it out.
//   PS = true : the stepping world's state slot `slot` is loaded (aged), updated and stored here (warm-started GJK direction incl.).
// CapsuleCapsule: both operands become segments; CapsuleShape: the capsule becomes a segment and the sub-detector is the one the
// dispatcher picks for (segment, other) — ball / plane / convex polyhedron generators — with the capsule's ContactPreprocessor on
// the segment's side and the linear prediction grown by the radius (capsule_shape_manifold_generator.rs:37-75).
template <bool PS>
__device__ __noinline__ void capsule_pair(const DevObjects& o, const DevHulls& H, const PersistArgs& ps, uint32_t slot, EpaState& e, ManifoldT<PS>& mf,
                                          uint32_t i1, uint32_t i2, uint32_t* epa_overflow, uint32_t* ref_panics) {
    uint32_t t1 = __ldg(&o.type[i1]) & NCB_TYPE_MASK, t2 = __ldg(&o.type[i2]) & NCB_TYPE_MASK;
    bool a_cap = t1 == NCB_SHAPE_CAPSULE, b_cap = t2 == NCB_SHAPE_CAPSULE;
    Iso ma = load_iso(o, i1), mb = load_iso(o, i2);
    float linear = __ldg(&o.qlimit[i1]) + __ldg(&o.qlimit[i2]);
    CapOperand a, b;
    a.shape = Shape{}, b.shape = Shape{};
    a.is_segment = a_cap, b.is_segment = b_cap;
    a.hh = b.hh = 0.f;
    a.seg_pts = b.seg_pts = nullptr;
    a.pre.active = b.pre.active = false;
    a.pre.radius = b.pre.radius = 0.f;
    if (a_cap) {
        float4 p = __ldg(&o.param[i1]);
        a.hh = p.x, a.seg_pts = o.cap_pts + 6 * (size_t)i1, a.pre.active = true, a.pre.radius = p.y;
        a.shape.type = NCB_SHAPE_CAPSULE;
        linear = linear + a.pre.radius;
    } else
        a.shape = load_shape(o, H, i1, t1);
    if (b_cap) {
        float4 p = __ldg(&o.param[i2]);
        b.hh = p.x, b.seg_pts = o.cap_pts + 6 * (size_t)i2, b.pre.active = true, b.pre.radius = p.y;
        b.shape.type = NCB_SHAPE_CAPSULE;
        linear = linear + b.pre.radius;
    } else
        b.shape = load_shape(o, H, i2, t2);
    bool simple = (!a_cap && (t1 == NCB_SHAPE_BALL || t1 == NCB_SHAPE_PLANE)) || (!b_cap && (t2 == NCB_SHAPE_BALL || t2 == NCB_SHAPE_PLANE));
    if (simple) {
        if constexpr (PS) pm_load_and_age(ps, slot, mf);
        Feature feat;
        if (!a_cap && t1 == NCB_SHAPE_BALL)
            gen_ball_segment(ma, a.shape.radius, mb, b.hh, linear, false, b.pre, mf);
        else if (!b_cap && t2 == NCB_SHAPE_BALL)
            gen_ball_segment(mb, b.shape.radius, ma, a.hh, linear, true, a.pre, mf);
        else if (!a_cap)
            gen_plane_segment(ma, a.shape.he, mb, b.hh, linear, false, b.pre, mf, feat);
        else
            gen_plane_segment(mb, b.shape.he, ma, a.hh, linear, true, a.pre, mf, feat);
        if constexpr (PS) pm_store(ps, slot, mf, i1, i2);
        return;
    }
    Support ga = cap_support(a), gb = cap_support(b);
    V3 d0;
    bool warm = false;
    if constexpr (PS) {
        float4 pd = ps.dir[slot];
        if (pd.w != 0.f) d0 = v3(pd.x, pd.y, pd.z), warm = true;
    }
    if (!warm && !unit_try_new(mb.t - ma.t, NCB_EPS, d0)) d0 = v3(1.f, 0.f, 0.f);
    V3 p1, p2, dir;
    Simplex s;
    int r = gjk_closest_points(ma, ga, mb, gb, linear, d0, s, p1, p2, dir);
    if constexpr (PS) {
        if (r != GJK_INTERSECTION) ps.dir[slot] = make_float4(dir.x, dir.y, dir.z, 1.f);
    }
    if (r == GJK_INTERSECTION) {
        if (epa_closest_points(e, ma, ga, mb, gb, s.dim, s.v, p1, p2, dir)) {
            r = GJK_CLOSEST_POINTS;
            if constexpr (PS) ps.dir[slot] = make_float4(dir.x, dir.y, dir.z, 1.f);
        } else {
            if (e.overflow) atomicAdd(epa_overflow, 1u);
            if (e.panicked) atomicAdd(ref_panics, 1u);
            r = GJK_NO_INTERSECTION;
            if constexpr (PS) ps.dir[slot] = make_float4(1.f, 0.f, 0.f, 1.f);
        }
    }
    if (r == GJK_NO_INTERSECTION) {
        if constexpr (PS) pm_age_only(ps, slot, i1, i2);
        return;
    }
    if constexpr (PS) pm_load_and_age(ps, slot, mf);
    float2 ang1 = __ldg(&o.ang_cs[i1 * o.ang_stride]), ang2 = __ldg(&o.ang_cs[i2 * o.ang_stride]);
    Feature f1, f2;
    if (!capsule_convex_manifold(ma, a, mb, b, linear, ang1, ang2, p1, p2, dir, mf, f1, f2)) atomicAdd(epa_overflow, 1u);
    if constexpr (PS) pm_store(ps, slot, mf, i1, i2);
}

}  // namespace ncb
