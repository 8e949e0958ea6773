// TEST INFRASTRUCTURE ONLY — compiles the per-pair device functions of ncollide_b200/csrc/narrow.cu (features, clipping, manifold,
// the five contact generators; with gjk.cuh / shapes.cuh) for the host through tests/host_shim/cuda_runtime.h and runs the
// fresh-world narrow phase one pair after the other: the dispatch below mirrors the bodies of k_narrow<KEY>, k_bh_epa, k_cc_gjk ->
// k_cc_epa -> k_cc_manifold (same calls in the same order; the work queues between the phases only carry these values across
// kernels).  Compared against the oracle by tests/test_device_source_on_host.py.  The product never links this.
#include "narrow.cu"
#include "capsule.cuh"

using namespace ncb;

static DevHulls hulls_from(const ncb_hull_library* L) {
    DevHulls H;
    std::memset(&H, 0, sizeof H);
    if (!L) return H;
    H.n_hulls = L->n_hulls;
    H.vert_off = L->vert_off, H.face_off = L->face_off, H.edge_off = L->edge_off, H.fadj_off = L->fadj_off, H.vadj_off = L->vadj_off;
    H.points = L->points;
    H.vert_first_adj = L->vert_first_adj, H.vert_num_adj = L->vert_num_adj;
    H.face_first = L->face_first, H.face_num = L->face_num;
    H.face_normal = L->face_normal;
    H.vaf = L->vertices_adj_to_face, H.eaf = L->edges_adj_to_face;
    H.edge_vertices = L->edge_vertices, H.edge_faces = L->edge_faces;
    H.edge_dir = L->edge_dir;
    H.fav = L->faces_adj_to_vertex, H.eav = L->edges_adj_to_vertex;
    return H;
}

// One fresh-world pair without capsules: the bodies of k_narrow<KEY>, k_bh_epa, k_cc_gjk -> k_cc_epa -> k_cc_manifold.
static uint8_t fresh_pair(const DevObjects& o, const DevHulls& H, const ncb_objects* objs, float2 one_degree_cs, EpaState* e, Manifold& mf, uint32_t* flags,
                          uint32_t i1, uint32_t i2) {
    static const uint8_t algo_of[4][4] = {{NCB_ALGO_BALL_BALL, NCB_ALGO_BALL_CONVEX, NCB_ALGO_BALL_CONVEX, NCB_ALGO_PLANE_BALL},
                                          {NCB_ALGO_BALL_CONVEX, NCB_ALGO_CONVEX_CONVEX, NCB_ALGO_CONVEX_CONVEX, NCB_ALGO_PLANE_CONVEX},
                                          {NCB_ALGO_BALL_CONVEX, NCB_ALGO_CONVEX_CONVEX, NCB_ALGO_CONVEX_CONVEX, NCB_ALGO_PLANE_CONVEX},
                                          {NCB_ALGO_PLANE_BALL, NCB_ALGO_PLANE_CONVEX, NCB_ALGO_PLANE_CONVEX, NCB_ALGO_NONE}};
    uint32_t t1 = o.type[i1] & 3u, t2 = o.type[i2] & 3u;
    Iso ma = load_iso(o, i1), mb = load_iso(o, i2);
    float linear = o.qlimit[i1] + o.qlimit[i2];
    Shape a = load_shape(o, H, i1, t1), b = load_shape(o, H, i2, t2);
    uint8_t al = algo_of[t1][t2];
    if (al == NCB_ALGO_BALL_BALL) {
        gen_ball_ball(ma, a.radius, mb, b.radius, linear, mf);
    } else if (al == NCB_ALGO_PLANE_BALL) {
        if (t1 == NCB_SHAPE_PLANE)
            gen_plane_ball(ma, a.he, mb, b.radius, linear, false, mf);
        else
            gen_plane_ball(mb, b.he, ma, a.radius, linear, true, mf);
    } else if (al == NCB_ALGO_PLANE_CONVEX) {
        Feature feat;
        if (t1 == NCB_SHAPE_PLANE)
            gen_plane_convex(ma, a.he, mb, b, linear, false, mf, feat);
        else
            gen_plane_convex(mb, b.he, ma, a, linear, true, mf, feat);
    } else if (al == NCB_ALGO_BALL_CONVEX) {
        bool flip = t1 != NCB_SHAPE_BALL;
        const Shape& ball = flip ? b : a;
        const Shape& cp = flip ? a : b;
        const Iso& mball = flip ? mb : ma;
        const Iso& mcp = flip ? ma : mb;
        if (cp.type == NCB_SHAPE_CUBOID) {
            bool inside;
            V3 world2;
            uint32_t f2;
            cuboid_project_point_with_feature(cp.he, mcp, mball.t, inside, world2, f2);
            gen_ball_convex_finish(mball.t, ball.radius, cp, mcp, inside, world2, f2, linear, flip, mf);
        } else {
            HullProjSetup u = hull_proj_setup(cp.hull, mcp, mball.t);
            V3 world2;
            Simplex s;
            if (hull_project_gjk(u, mball.t, s, world2) == GJK_CLOSEST_POINTS) {
                uint32_t f2 = hull_project_feature(cp.hull, mcp, mball.t, false, world2, one_degree_cs);
                gen_ball_convex_finish(mball.t, ball.radius, cp, mcp, false, world2, f2, linear, flip, mf);
            } else {  // k_bh_epa: the ball centre is inside the hull
                Iso id = iso_id();
                V3 p1, p2, d;
                if (epa_closest_points(*e, u.m, u.shape, id, u.origin, s.dim, s.v, p1, p2, d))
                    world2 = p1 + mball.t;
                else {
                    flags[0] += e->overflow, flags[1] += e->panicked;
                    world2 = mball.t;
                }
                uint32_t f2 = hull_project_feature(cp.hull, mcp, mball.t, true, world2, one_degree_cs);
                gen_ball_convex_finish(mball.t, ball.radius, cp, mcp, true, world2, f2, linear, flip, mf);
            }
        }
    } else if (al == NCB_ALGO_CONVEX_CONVEX) {
        Support ga = as_support(a), gb = as_support(b);
        V3 d0;
        if (!unit_try_new(mb.t - ma.t, NCB_EPS, d0)) d0 = v3(1.f, 0.f, 0.f);
        V3 p1, p2, dir;
        Simplex s;
        int r = gjk_closest_points(ma, ga, mb, gb, linear, d0, s, p1, p2, dir);  // k_cc_gjk
        if (r == GJK_INTERSECTION) {                                              // k_cc_epa
            if (epa_closest_points(*e, ma, ga, mb, gb, s.dim, s.v, p1, p2, dir))
                r = GJK_CLOSEST_POINTS;
            else {
                flags[0] += e->overflow, flags[1] += e->panicked;
                r = GJK_NO_INTERSECTION;
            }
        }
        if (r == GJK_CLOSEST_POINTS) {                                            // k_cc_manifold
            float a1 = objs->ang_pred[i1], a2 = objs->ang_pred[i2];
            float2 ang1 = make_float2(cosf(a1), sinf(a1)), ang2 = make_float2(cosf(a2), sinf(a2));
            Feature f1, f2;
            convex_convex_manifold(ma, a, mb, b, linear, ang1, ang2, p1, p2, dir, mf, f1, f2);
        }
    }
    return al;
}

// A pair with at least one capsule (shape_type 4, param = half_height, radius): capsule_pair<false> of capsule.cuh, the function
// k_capsule runs per thread.  seg_pts: 6 floats per OBJECT (b then a of the capsule's segment; unused for other shapes).
static uint8_t capsule_pair_host(DevObjects o, const DevHulls& H, const ncb_objects* objs, const float* seg_pts, const float2* ang_cs, EpaState* e,
                                 Manifold& mf, uint32_t* flags, uint32_t i1, uint32_t i2) {
    (void)objs;
    o.cap_pts = seg_pts;
    o.ang_cs = ang_cs;
    o.ang_stride = 1;
    PersistArgs none;
    std::memset(&none, 0, sizeof none);
    capsule_pair<false>(o, H, none, 0, *e, mf, i1, i2, &flags[0], &flags[1]);
    return (o.type[i1] == 4 && o.type[i2] == 4) ? 7 : 8;  // CapsuleCapsule / CapsuleShape
}

extern "C" {
// Returns the number of contacts (may exceed cap).  manifold_off[n_pairs + 1]; algo[p] = NCB_ALGO_* (7 / 8: the capsule generators);
// flags[0] += EPA capacity overflows / dropped contacts, flags[1] += reference panics.  seg_pts: NULL, or 6 floats per object for
// worlds with capsules (shape_type 4).
// kin_out (optional): the ContactKinematic of every contact (ncb_set_kinematics), aligned with out
uint64_t shim_narrow_phase_kin(const ncb_objects* objs, const ncb_hull_library* lib, const float* seg_pts, uint64_t n_pairs, const uint32_t* pairs,
                               ncb_contact* out, ncb_kinematic* kin_out, uint64_t cap, uint32_t* manifold_off, uint8_t* algo, uint32_t* flags) {
    DevObjects o;
    std::memset(&o, 0, sizeof o);
    o.n = objs->n;
    o.pos = objs->pos;
    o.rot = reinterpret_cast<const float4*>(objs->rot);
    o.type = objs->shape_type;
    o.param = reinterpret_cast<const float4*>(objs->shape_param);
    o.qlimit = objs->query_limit;
    DevHulls H = hulls_from(lib);
    // what ncb_set_objects / launch_narrow_phase_t prepare on the host with libm
    const float one_degree = (float)(3.14159265358979323846 / 180.0);
    const float2 one_degree_cs = make_float2(cosf(one_degree), sinf(one_degree));
    EpaState* e = new EpaState;
    Manifold* mfp = new Manifold;
    Manifold& mf = *mfp;
    Kin* kin_side = new Kin[MANIFOLD_MAX];
    mf.kin = kin_out ? kin_side : nullptr;
    uint64_t nc = 0;
    float2* ang_cs = new float2[objs->n ? objs->n : 1];  // ncb_set_objects' (cos, sin) table of the angular predictions
    for (uint32_t i = 0; i < objs->n; ++i) ang_cs[i] = make_float2(cosf(objs->ang_pred[i]), sinf(objs->ang_pred[i]));
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        mf.n = 0;
        mf.deepest = 0;
        uint8_t al = (o.type[i1] == 4 || o.type[i2] == 4) ? capsule_pair_host(o, H, objs, seg_pts, ang_cs, e, mf, flags, i1, i2)
                                                          : fresh_pair(o, H, objs, one_degree_cs, e, mf, flags, i1, i2);
        if (mf.deepest < 0) flags[0] += 1;  // more than MANIFOLD_MAX distinct contacts
        algo[p] = al;
        manifold_off[p] = (uint32_t)nc;
        for (int k = 0; k < mf.n; ++k, ++nc) {
            if (nc >= cap) continue;
            const ManifoldContact& c = mf.c[k];
            ncb_contact& w = out[nc];
            w.world1[0] = c.w1.x, w.world1[1] = c.w1.y, w.world1[2] = c.w1.z;
            w.world2[0] = c.w2.x, w.world2[1] = c.w2.y, w.world2[2] = c.w2.z;
            w.normal[0] = c.n.x, w.normal[1] = c.n.y, w.normal[2] = c.n.z;
            w.depth = c.depth;
            w.f1 = c.f1, w.f2 = c.f2;
            w.pair = (uint32_t)p;
            if (kin_out) {
                const Kin& q = kin_side[k];
                ncb_kinematic& z = kin_out[nc];
                z.local1[0] = q.local1.x, z.local1[1] = q.local1.y, z.local1[2] = q.local1.z;
                z.local2[0] = q.local2.x, z.local2[1] = q.local2.y, z.local2[2] = q.local2.z;
                z.dir1[0] = q.dir1.x, z.dir1[1] = q.dir1.y, z.dir1[2] = q.dir1.z;
                z.dir2[0] = q.dir2.x, z.dir2[1] = q.dir2.y, z.dir2[2] = q.dir2.z;
                z.dilation1 = q.dil1, z.dilation2 = q.dil2;
                z.geometry1 = q.g1, z.geometry2 = q.g2;
            }
        }
    }
    manifold_off[n_pairs] = (uint32_t)nc;
    delete e;
    delete mfp;
    delete[] kin_side;
    delete[] ang_cs;
    return nc;
}
uint64_t shim_narrow_phase_ex(const ncb_objects* objs, const ncb_hull_library* lib, const float* seg_pts, uint64_t n_pairs, const uint32_t* pairs,
                              ncb_contact* out, uint64_t cap, uint32_t* manifold_off, uint8_t* algo, uint32_t* flags) {
    return shim_narrow_phase_kin(objs, lib, seg_pts, n_pairs, pairs, out, nullptr, cap, manifold_off, algo, flags);
}
uint64_t shim_narrow_phase(const ncb_objects* objs, const ncb_hull_library* lib, uint64_t n_pairs, const uint32_t* pairs, ncb_contact* out,
                           uint64_t cap, uint32_t* manifold_off, uint8_t* algo, uint32_t* flags) {
    return shim_narrow_phase_ex(objs, lib, nullptr, n_pairs, pairs, out, cap, manifold_off, algo, flags);
}

// capsule_aabb of capsule.cuh for the capsules of a scene (mode 0: the shape's AABB); out: 6 floats per object, untouched for other shapes
void shim_capsule_aabbs(const ncb_objects* objs, float* out) {
    for (uint32_t i = 0; i < objs->n; ++i) {
        if (objs->shape_type[i] != 4) continue;
        Iso m;
        m.t = v3(objs->pos[3 * i], objs->pos[3 * i + 1], objs->pos[3 * i + 2]);
        m.q = Quat{objs->rot[4 * i], objs->rot[4 * i + 1], objs->rot[4 * i + 2], objs->rot[4 * i + 3]};
        V3 mins, maxs;
        capsule_aabb(m, objs->shape_param[4 * i], objs->shape_param[4 * i + 1], mins, maxs);
        float* d = out + 6 * (size_t)i;
        d[0] = mins.x, d[1] = mins.y, d[2] = mins.z, d[3] = maxs.x, d[4] = maxs.y, d[5] = maxs.z;
    }
}

// Stepping world, per pair: what k_narrow<KEY, true>, k_bh_epa<true>, k_cc_gjk<true> -> k_cc_epa<true> -> k_cc_manifold<true> do for ONE
// updated pair whose persistent state lives in slot `slots[k]` of dir / pm_hdr / pm_entry (same calls, same order: load + age the
// manifold cache, warm-started GJK, generate, store back, contact events).
void shim_persist_update_ex(const ncb_objects* objs, const ncb_hull_library* lib, const float* seg_pts, uint64_t n_update, const uint32_t* pairs,
                            const uint32_t* slots,
                         float* dir, uint32_t* pm_hdr, float* pm_entry, unsigned long long* events, uint32_t* n_events, uint32_t cap_events,
                         uint32_t* pm_overflow, uint32_t* flags) {
    DevObjects o;
    std::memset(&o, 0, sizeof o);
    o.n = objs->n;
    o.pos = objs->pos;
    o.rot = reinterpret_cast<const float4*>(objs->rot);
    o.type = objs->shape_type;
    o.param = reinterpret_cast<const float4*>(objs->shape_param);
    o.qlimit = objs->query_limit;
    DevHulls H = hulls_from(lib);
    const float one_degree = (float)(3.14159265358979323846 / 180.0);
    const float2 one_degree_cs = make_float2(cosf(one_degree), sinf(one_degree));
    float2* ang_cs_tab = new float2[objs->n ? objs->n : 1];
    for (uint32_t i = 0; i < objs->n; ++i) ang_cs_tab[i] = make_float2(cosf(objs->ang_pred[i]), sinf(objs->ang_pred[i]));
    o.cap_pts = seg_pts;
    o.ang_cs = ang_cs_tab;
    o.ang_stride = 1;
    PersistArgs ps;
    ps.dir = reinterpret_cast<float4*>(dir);
    ps.pm_hdr = pm_hdr;
    ps.pm_entry = reinterpret_cast<float4*>(pm_entry);
    ps.events = events;
    ps.n_events = n_events;
    ps.cap_events = cap_events;
    ps.pm_overflow = pm_overflow;
    EpaState* e = new EpaState;
    PManifold* mfp = new PManifold;
    PManifold& mf = *mfp;
    for (uint64_t k = 0; k < n_update; ++k) {
        uint32_t i1 = pairs[2 * k], i2 = pairs[2 * k + 1], slot = slots[k];
        if (o.type[i1] == 4 || o.type[i2] == 4) {  // a capsule pair: capsule_pair<true>, the function k_capsule<true> runs per thread
            capsule_pair<true>(o, H, ps, slot, *e, mf, i1, i2, &flags[0], &flags[1]);
            continue;
        }
        uint32_t t1 = o.type[i1] & 3u, t2 = o.type[i2] & 3u;
        Iso ma = load_iso(o, i1), mb = load_iso(o, i2);
        float linear = o.qlimit[i1] + o.qlimit[i2];
        Shape a = load_shape(o, H, i1, t1), b = load_shape(o, H, i2, t2);
        bool a_ball = t1 == NCB_SHAPE_BALL, b_ball = t2 == NCB_SHAPE_BALL, a_plane = t1 == NCB_SHAPE_PLANE, b_plane = t2 == NCB_SHAPE_PLANE;
        if (a_plane && b_plane) continue;  // K_NONE: no edge
        bool convex_convex = !a_ball && !b_ball && !a_plane && !b_plane;
        if (!convex_convex) {  // k_narrow<KEY, true> (+ k_bh_epa<true>)
            pm_load_and_age(ps, slot, mf);
            if (a_ball && b_ball) {
                gen_ball_ball(ma, a.radius, mb, b.radius, linear, mf);
            } else if ((a_plane && b_ball) || (a_ball && b_plane)) {
                if (a_plane)
                    gen_plane_ball(ma, a.he, mb, b.radius, linear, false, mf);
                else
                    gen_plane_ball(mb, b.he, ma, a.radius, linear, true, mf);
            } else if (a_plane || b_plane) {
                Feature feat;
                if (a_plane)
                    gen_plane_convex(ma, a.he, mb, b, linear, false, mf, feat);
                else
                    gen_plane_convex(mb, b.he, ma, a, linear, true, mf, feat);
            } else {
                bool flip = !a_ball;
                const Shape& ball = flip ? b : a;
                const Shape& cp = flip ? a : b;
                const Iso& mball = flip ? mb : ma;
                const Iso& mcp = flip ? ma : mb;
                if (cp.type == NCB_SHAPE_CUBOID) {
                    bool inside;
                    V3 world2;
                    uint32_t f2;
                    cuboid_project_point_with_feature(cp.he, mcp, mball.t, inside, world2, f2);
                    gen_ball_convex_finish(mball.t, ball.radius, cp, mcp, inside, world2, f2, linear, flip, mf);
                } else {
                    HullProjSetup u = hull_proj_setup(cp.hull, mcp, mball.t);
                    V3 world2;
                    Simplex s;
                    if (hull_project_gjk(u, mball.t, s, world2) == GJK_CLOSEST_POINTS) {
                        uint32_t f2 = hull_project_feature(cp.hull, mcp, mball.t, false, world2, one_degree_cs);
                        gen_ball_convex_finish(mball.t, ball.radius, cp, mcp, false, world2, f2, linear, flip, mf);
                    } else {
                        Iso id = iso_id();
                        V3 p1, p2, d;
                        if (epa_closest_points(*e, u.m, u.shape, id, u.origin, s.dim, s.v, p1, p2, d))
                            world2 = p1 + mball.t;
                        else {
                            flags[0] += e->overflow, flags[1] += e->panicked;
                            world2 = mball.t;
                        }
                        uint32_t f2 = hull_project_feature(cp.hull, mcp, mball.t, true, world2, one_degree_cs);
                        gen_ball_convex_finish(mball.t, ball.radius, cp, mcp, true, world2, f2, linear, flip, mf);
                    }
                }
            }
            pm_store(ps, slot, mf, i1, i2);
            continue;
        }
        // k_cc_gjk<true>
        Support ga = as_support(a), gb = as_support(b);
        V3 d0;
        bool warm = false;
        float4 pd = ps.dir[slot];
        if (pd.w != 0.f) d0 = v3(pd.x, pd.y, pd.z), warm = true;
        if (!warm && !unit_try_new(mb.t - ma.t, NCB_EPS, d0)) d0 = v3(1.f, 0.f, 0.f);
        V3 p1, p2, dirv;
        Simplex s;
        int r = gjk_closest_points(ma, ga, mb, gb, linear, d0, s, p1, p2, dirv);
        if (r != GJK_INTERSECTION) ps.dir[slot] = make_float4(dirv.x, dirv.y, dirv.z, 1.f);
        if (r == GJK_NO_INTERSECTION) {
            pm_age_only(ps, slot, i1, i2);
            continue;
        }
        if (r == GJK_INTERSECTION) {  // k_cc_epa<true>
            if (epa_closest_points(*e, ma, ga, mb, gb, s.dim, s.v, p1, p2, dirv)) {
                ps.dir[slot] = make_float4(dirv.x, dirv.y, dirv.z, 1.f);
            } else {
                flags[0] += e->overflow, flags[1] += e->panicked;
                ps.dir[slot] = make_float4(1.f, 0.f, 0.f, 1.f);
                pm_age_only(ps, slot, i1, i2);
                continue;
            }
        }
        // k_cc_manifold<true>
        pm_load_and_age(ps, slot, mf);
        float a1 = objs->ang_pred[i1], a2 = objs->ang_pred[i2];
        float2 ang1 = make_float2(cosf(a1), sinf(a1)), ang2 = make_float2(cosf(a2), sinf(a2));
        Feature f1, f2;
        convex_convex_manifold(ma, a, mb, b, linear, ang1, ang2, p1, p2, dirv, mf, f1, f2);
        pm_store(ps, slot, mf, i1, i2);
    }
    delete e;
    delete mfp;
    delete[] ang_cs_tab;
}

void shim_persist_update(const ncb_objects* objs, const ncb_hull_library* lib, uint64_t n_update, const uint32_t* pairs, const uint32_t* slots,
                         float* dir, uint32_t* pm_hdr, float* pm_entry, unsigned long long* events, uint32_t* n_events, uint32_t cap_events,
                         uint32_t* pm_overflow, uint32_t* flags) {
    shim_persist_update_ex(objs, lib, nullptr, n_update, pairs, slots, dir, pm_hdr, pm_entry, events, n_events, cap_events, pm_overflow, flags);
}

// k_sim_export for the listed slots: live contacts in slab order, ids = insertion counter << 8 | slab slot.
uint64_t shim_persist_export(uint64_t n, const uint32_t* slots, const uint32_t* pm_hdr, const float* pm_entry_f, uint32_t* manifold_off,
                             ncb_contact* contacts, uint32_t* ids, uint64_t cap) {
    const float4* pm_entry = reinterpret_cast<const float4*>(pm_entry_f);
    uint64_t nc = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t slot = slots[i];
        manifold_off[i] = (uint32_t)nc;
        int n0 = (int)(pm_hdr[(size_t)slot * PM_HDR_WORDS] & 0xffu);
        const float4* e = pm_entry + (size_t)slot * PM_CAP * PM_ENTRY_F4;
        uint32_t done = 0;
        for (;;) {
            int best = -1;
            uint32_t best_slot = 0xffffffffu;
            for (int k = 0; k < n0; ++k) {
                uint32_t meta = __float_as_uint(e[4 * k + 3].w);
                if (!((meta >> 8) & 1u) || ((done >> k) & 1u)) continue;
                if ((meta & 0xffu) < best_slot) best_slot = meta & 0xffu, best = k;
            }
            if (best < 0) break;
            done |= 1u << best;
            if (nc < cap) {
                float4 a = e[4 * best], b = e[4 * best + 1], c = e[4 * best + 2];
                uint32_t meta = __float_as_uint(e[4 * best + 3].w);
                ncb_contact& w = contacts[nc];
                w.world1[0] = a.x, w.world1[1] = a.y, w.world1[2] = a.z;
                w.world2[0] = b.x, w.world2[1] = b.y, w.world2[2] = b.z;
                w.normal[0] = c.x, w.normal[1] = c.y, w.normal[2] = c.z;
                w.depth = a.w;
                w.f1 = __float_as_uint(b.w), w.f2 = __float_as_uint(c.w);
                w.pair = (uint32_t)i;
                ids[nc] = ((meta >> 9) << 8) | (meta & 0xffu);
            }
            nc++;
        }
    }
    manifold_off[n] = (uint32_t)nc;
    return nc;
}
}
