// TEST INFRASTRUCTURE ONLY — compiles the device GJK / EPA of ncollide_b200/csrc/gjk.cuh (the functions k_cc_gjk / k_cc_epa /
// k_bh_epa run per pair: gjk_closest_points, epa_init, epa_step, the Voronoi simplex) for the host through
// tests/host_shim/cuda_runtime.h, so that their logic and f32 operation order can be checked against the oracle without a GPU.
#define NCB_EPA_STATS  // peak heap / silhouette / flood-stack sizes for scripts/epa_work_stats.py (this test library only)
#include "shapes.cuh"

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

template <class Store>
static int run_tier(Store& e, const Iso& ma, const SupportS& sa, const Iso& mb, const SupportS& sb, const Simplex& s, V3& p1, V3& p2, V3& n) {
    uint32_t res_face;
    int st = epa_init_t<true>(e, ma, sa, mb, sb, s.dim, s.v, p1, p2, n, res_face);
    while (st == EPA_CONTINUE) st = epa_step_t(e, ma, sa, mb, sb, res_face);
    if (st == EPA_DONE_OK && res_face != EPA_RES_DIRECT) epa_result_from_face(e, res_face, p1, p2, n);
    return st;
}

extern "C" {
// contact_sm_sm (GJK, then EPA when the origin is inside the CSO) for cuboid / hull pairs.
// out[10 p] = p1, p2, normal, found flag; flags[0] += EPA capacity overflows, flags[1] += reference panics, flags[2] = EPA calls.
void shim_contact_sm_sm(const ncb_objects* objs, const ncb_hull_library* lib, uint64_t n_pairs, const uint32_t* pairs, const float* predictions,
                        float* out, uint32_t* flags) {
    DevObjects o;
    std::memset(&o, 0, sizeof o);
    o.n = objs->n;
    o.pos = objs->pos;
    o.rot = reinterpret_cast<const float4*>(objs->rot);
    o.type = objs->shape_type;
    o.param = reinterpret_cast<const float4*>(objs->shape_param);
    o.qlimit = objs->query_limit;
    DevHulls H = hulls_from(lib);
    EpaState* e = new EpaState;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        Shape a = load_shape(o, H, i1, o.type[i1]), b = load_shape(o, H, i2, o.type[i2]);
        Iso ma = load_iso(o, i1), mb = load_iso(o, i2);
        float prediction = predictions ? predictions[p] : o.qlimit[i1] + o.qlimit[i2];
        V3 p1 = v3(0.f, 0.f, 0.f), p2 = p1, n = p1;
        uint32_t before = flags[0] + flags[1];
        (void)before;
        int r = contact_sm_sm(*e, ma, as_support(a), mb, as_support(b), prediction, p1, p2, n, &flags[0], &flags[1]);
        float* d = out + 10 * p;
        for (int k = 0; k < 10; ++k) d[k] = 0.f;
        if (r == GJK_CLOSEST_POINTS) {
            d[0] = p1.x, d[1] = p1.y, d[2] = p1.z, d[3] = p2.x, d[4] = p2.y, d[5] = p2.z, d[6] = n.x, d[7] = n.y, d[8] = n.z;
            d[9] = 1.f;
        }
    }
    delete e;
}

// Work statistics of the device EPA per penetrating pair (design data for kernel restructuring, scripts/epa_work_stats.py):
// stats[8 k] = expansion steps, vertices, faces (incl. deleted), heap entries at the end of pair k's EPA run, then the PEAK heap size,
// silhouette length and flood-stack depth, and the simplex dimension + 1; 0s when GJK did not reach EPA.
void shim_epa_work_stats(const ncb_objects* objs, const ncb_hull_library* lib, uint64_t n_pairs, const uint32_t* pairs, uint32_t* stats) {
    DevObjects o;
    std::memset(&o, 0, sizeof o);
    o.n = objs->n;
    o.pos = objs->pos;
    o.rot = reinterpret_cast<const float4*>(objs->rot);
    o.type = objs->shape_type;
    o.param = reinterpret_cast<const float4*>(objs->shape_param);
    o.qlimit = objs->query_limit;
    DevHulls H = hulls_from(lib);
    EpaState* e = new EpaState;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        Shape a = load_shape(o, H, i1, o.type[i1]), b = load_shape(o, H, i2, o.type[i2]);
        Iso ma = load_iso(o, i1), mb = load_iso(o, i2);
        Support ga = as_support(a), gb = as_support(b);
        V3 d0;
        if (!unit_try_new(mb.t - ma.t, NCB_EPS, d0)) d0 = v3(1.f, 0.f, 0.f);
        V3 p1, p2, dir;
        Simplex s;
        uint32_t* st = stats + 8 * p;
        for (int k = 0; k < 8; ++k) st[k] = 0;
        if (gjk_closest_points(ma, ga, mb, gb, o.qlimit[i1] + o.qlimit[i2], d0, s, p1, p2, dir) != GJK_INTERSECTION) continue;
        int status = epa_init(*e, ma, ga, mb, gb, s.dim, s.v, p1, p2, dir);
        uint32_t steps = 0;
        while (status == EPA_CONTINUE) status = epa_step(*e, ma, ga, mb, gb, p1, p2, dir), steps++;
        st[0] = steps + 1, st[1] = (uint32_t)e->nverts, st[2] = (uint32_t)e->nfaces, st[3] = (uint32_t)e->nheap;
        st[4] = (uint32_t)e->peak_heap, st[5] = (uint32_t)e->peak_sil, st[6] = (uint32_t)e->peak_stk, st[7] = (uint32_t)s.dim + 1;
    }
    delete e;
}

// The same as shim_contact_sm_sm, but EPA runs the way the tiered kernels run it: on the shared-memory polytope store of epa.cuh
// (first tier: the words of one lane of a 64-thread CTA, here a host array with the kernel's lane stride; second tier: one lane of a
// 32-thread CTA) with the slim operands; a pair that exceeds a capacity restarts on the next tier, and beyond the second one on the
// local-memory store, exactly like the kernels.  flags[3] += pairs that restarted on the second tier, flags[2] = EPA calls.
void shim_contact_sm_sm_compact(const ncb_objects* objs, const ncb_hull_library* lib, uint64_t n_pairs, const uint32_t* pairs,
                                const float* predictions, float* out, uint32_t* flags) {
    DevObjects o;
    std::memset(&o, 0, sizeof o);
    o.n = objs->n;
    o.pos = objs->pos;
    o.rot = reinterpret_cast<const float4*>(objs->rot);
    o.type = objs->shape_type;
    o.param = reinterpret_cast<const float4*>(objs->shape_param);
    o.qlimit = objs->query_limit;
    DevHulls H = hulls_from(lib);
    EpaState* last = new EpaState;
    typedef EpaTier1<64> T1;
    typedef EpaTier2<32> T2;
    uint32_t* words1 = new uint32_t[T1::WORDS * 64];
    uint32_t* words2 = new uint32_t[T2::WORDS * 32];
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        Shape a = load_shape(o, H, i1, o.type[i1]), b = load_shape(o, H, i2, o.type[i2]);
        Iso ma = load_iso(o, i1), mb = load_iso(o, i2);
        Support ga = as_support(a), gb = as_support(b);
        float prediction = predictions ? predictions[p] : o.qlimit[i1] + o.qlimit[i2];
        V3 p1 = v3(0.f, 0.f, 0.f), p2 = p1, n = p1;
        V3 dir;
        if (!unit_try_new(mb.t - ma.t, NCB_EPS, dir)) dir = v3(1.f, 0.f, 0.f);
        Simplex s;
        int r = gjk_closest_points(ma, ga, mb, gb, prediction, dir, s, p1, p2, n);
        if (r == GJK_INTERSECTION) {
            flags[2]++;
            SupportS sa = slim_support(ga), sb = slim_support(gb);
            T1 e1;
            e1.base = words1 + (p % 64);  // any lane of the CTA
            int st = run_tier(e1, ma, sa, mb, sb, s, p1, p2, n);
            bool overflow = st == EPA_DONE_FAIL && e1.overflow, panicked = e1.panicked;
            if (overflow) {
                flags[3]++;
                T2 e2;
                e2.base = words2 + (p % 32);
                st = run_tier(e2, ma, sa, mb, sb, s, p1, p2, n);
                overflow = st == EPA_DONE_FAIL && e2.overflow, panicked = e2.panicked;
            }
            if (st == EPA_DONE_OK) {
                r = GJK_CLOSEST_POINTS;
            } else if (overflow) {  // last resort (k_cc_epa_big)
                if (epa_closest_points(*last, ma, ga, mb, gb, s.dim, s.v, p1, p2, n))
                    r = GJK_CLOSEST_POINTS;
                else {
                    if (last->overflow) flags[0]++;
                    if (last->panicked) flags[1]++;
                    r = GJK_NO_INTERSECTION;
                }
            } else {
                if (panicked) flags[1]++;
                r = GJK_NO_INTERSECTION;
            }
        }
        float* d = out + 10 * p;
        for (int k = 0; k < 10; ++k) d[k] = 0.f;
        if (r == GJK_CLOSEST_POINTS) {
            d[0] = p1.x, d[1] = p1.y, d[2] = p1.z, d[3] = p2.x, d[4] = p2.y, d[5] = p2.z, d[6] = n.x, d[7] = n.y, d[8] = n.z;
            d[9] = 1.f;
        }
    }
    delete last;
    delete[] words1;
    delete[] words2;
}

// GJK work per convex pair: stats[3 k] = support-point evaluations (loop turns incl. the initial one), exit kind (GJK_*), simplex
// dimension + 1 at the exit.  Design data for restructuring k_cc_gjk (scripts/gjk_work_stats.py).
void shim_gjk_work_stats(const ncb_objects* objs, const ncb_hull_library* lib, uint64_t n_pairs, const uint32_t* pairs, uint32_t* stats) {
    DevObjects o;
    std::memset(&o, 0, sizeof o);
    o.n = objs->n;
    o.pos = objs->pos;
    o.rot = reinterpret_cast<const float4*>(objs->rot);
    o.type = objs->shape_type;
    o.param = reinterpret_cast<const float4*>(objs->shape_param);
    o.qlimit = objs->query_limit;
    DevHulls H = hulls_from(lib);
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        Shape a = load_shape(o, H, i1, o.type[i1]), b = load_shape(o, H, i2, o.type[i2]);
        Iso ma = load_iso(o, i1), mb = load_iso(o, i2);
        Support ga = as_support(a), gb = as_support(b);
        V3 d0;
        if (!unit_try_new(mb.t - ma.t, NCB_EPS, d0)) d0 = v3(1.f, 0.f, 0.f);
        V3 p1, p2, dir;
        Simplex s;
        g_gjk_support_evals = 0;
        int r = gjk_closest_points(ma, ga, mb, gb, o.qlimit[i1] + o.qlimit[i2], d0, s, p1, p2, dir);
        stats[3 * p] = (uint32_t)g_gjk_support_evals, stats[3 * p + 1] = (uint32_t)r, stats[3 * p + 2] = (uint32_t)s.dim + 1;
    }
}
}
