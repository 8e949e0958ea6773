// TEST INFRASTRUCTURE ONLY — compiles the per-pair device functions of ncollide_b200/csrc/narrow.cu (features, clipping, manifold,
// the five contact generators; with gjk.cuh / shapes.cuh) for the host through tests/host_shim/cuda_runtime.h and runs the
// fresh-world narrow phase one pair after the other: the dispatch below mirrors the bodies of k_narrow<KEY>, k_bh_epa, k_cc_gjk ->
// k_cc_epa -> k_cc_manifold (same calls in the same order; the work queues between the phases only carry these values across
// kernels).  Compared against the oracle by tests/test_device_source_on_host.py.  The product never links this.
#include "narrow.cu"

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

extern "C" {
// Returns the number of contacts (may exceed cap).  manifold_off[n_pairs + 1]; algo[p] = NCB_ALGO_*; flags[0] += EPA capacity
// overflows / dropped contacts, flags[1] += reference panics.
uint64_t shim_narrow_phase(const ncb_objects* objs, const ncb_hull_library* lib, uint64_t n_pairs, const uint32_t* pairs, ncb_contact* out,
                           uint64_t cap, uint32_t* manifold_off, uint8_t* algo, uint32_t* flags) {
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
    static const uint8_t algo_of[4][4] = {{NCB_ALGO_BALL_BALL, NCB_ALGO_BALL_CONVEX, NCB_ALGO_BALL_CONVEX, NCB_ALGO_PLANE_BALL},
                                          {NCB_ALGO_BALL_CONVEX, NCB_ALGO_CONVEX_CONVEX, NCB_ALGO_CONVEX_CONVEX, NCB_ALGO_PLANE_CONVEX},
                                          {NCB_ALGO_BALL_CONVEX, NCB_ALGO_CONVEX_CONVEX, NCB_ALGO_CONVEX_CONVEX, NCB_ALGO_PLANE_CONVEX},
                                          {NCB_ALGO_PLANE_BALL, NCB_ALGO_PLANE_CONVEX, NCB_ALGO_PLANE_CONVEX, NCB_ALGO_NONE}};
    EpaState* e = new EpaState;
    Manifold* mfp = new Manifold;
    Manifold& mf = *mfp;
    uint64_t nc = 0;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        uint32_t t1 = o.type[i1] & 3u, t2 = o.type[i2] & 3u;
        mf.n = 0;
        mf.deepest = 0;
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
                gen_ball_convex_finish(mball.t, ball.radius, cp, inside, world2, f2, linear, flip, mf);
            } else {
                HullProjSetup u = hull_proj_setup(cp.hull, mcp, mball.t);
                V3 world2;
                Simplex s;
                if (hull_project_gjk(u, mball.t, s, world2) == GJK_CLOSEST_POINTS) {
                    uint32_t f2 = hull_project_feature(cp.hull, mcp, mball.t, false, world2, one_degree_cs);
                    gen_ball_convex_finish(mball.t, ball.radius, cp, false, world2, f2, linear, flip, mf);
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
                    gen_ball_convex_finish(mball.t, ball.radius, cp, true, world2, f2, linear, flip, mf);
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
        }
    }
    manifold_off[n_pairs] = (uint32_t)nc;
    delete e;
    delete mfp;
    return nc;
}
}
