// TEST INFRASTRUCTURE ONLY — compiles the per-shape ray casts of ncollide_b200/csrc/query.cu (ball, cuboid, plane, convex hull through
// the GJK ray cast) and the shapes' point containment for the host through tests/host_shim/cuda_runtime.h; the dispatch mirrors
// visit_leaf / visit_leaf_q of the world-query kernels.  Compared with the oracle by tests/test_device_source_on_host.py.
#include "query.cu"

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
static DevObjects objects_from(const ncb_objects* objs) {
    DevObjects o;
    std::memset(&o, 0, sizeof o);
    o.n = objs->n;
    o.pos = objs->pos;
    o.rot = reinterpret_cast<const float4*>(objs->rot);
    o.type = objs->shape_type;
    o.param = reinterpret_cast<const float4*>(objs->shape_param);
    o.qlimit = objs->query_limit;
    return o;
}

extern "C" {
// RayCast::toi_and_normal_with_ray(position, ray, max_toi, solid = true) of object which[k] for ray k (7 floats: origin, dir, max_toi).
// out[4 k] = toi, normal; feat[k]; hit[k].
void shim_shape_ray_cast(const ncb_objects* objs, const ncb_hull_library* lib, uint64_t n, const uint32_t* which, const float* rays, float* out,
                         uint32_t* feat, uint8_t* hit) {
    DevObjects o = objects_from(objs);
    DevHulls H = hulls_from(lib);
    for (uint64_t k = 0; k < n; ++k) {
        uint32_t handle = which[k];
        const float* q = rays + 7 * k;
        uint32_t type = o.type[handle];
        Shape sh = load_shape(o, H, handle, type);
        Iso m = load_iso(o, handle);
        V3 ro = v3(q[0], q[1], q[2]), rd = v3(q[3], q[4], q[5]);
        RayHit h;
        if (type == NCB_SHAPE_BALL)
            h = ray_cast_ball(m.t, sh.radius, ro, rd, q[6]);
        else if (type == NCB_SHAPE_CUBOID)
            h = ray_cast_cuboid(sh.he, m, ro, rd, q[6]);
        else if (type == NCB_SHAPE_CONVEX_HULL)
            h = ray_cast_hull(sh.hull, m, ro, rd, q[6]);
        else
            h = ray_cast_plane(sh.he, m, ro, rd, q[6]);
        hit[k] = h.hit ? 1 : 0;
        out[4 * k] = h.hit ? h.toi : 0.f;
        out[4 * k + 1] = h.hit ? h.normal.x : 0.f, out[4 * k + 2] = h.hit ? h.normal.y : 0.f, out[4 * k + 3] = h.hit ? h.normal.z : 0.f;
        feat[k] = h.hit ? h.feature : 0u;
    }
}
// PointQuery::contains_point of object which[k] for point k (visit_leaf_q, KIND == 2)
void shim_shape_contains_point(const ncb_objects* objs, const ncb_hull_library* lib, uint64_t n, const uint32_t* which, const float* pts, uint8_t* inside_out) {
    DevObjects o = objects_from(objs);
    DevHulls H = hulls_from(lib);
    for (uint64_t k = 0; k < n; ++k) {
        uint32_t handle = which[k];
        uint32_t type = o.type[handle];
        Shape sh = load_shape(o, H, handle, type);
        Iso m = load_iso(o, handle);
        V3 pt = v3(pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]);
        bool inside;
        if (type == NCB_SHAPE_BALL) {
            inside = norm_squared(iso_inv_point(m, pt)) <= sh.radius * sh.radius;
        } else if (type == NCB_SHAPE_CUBOID) {
            V3 l = iso_inv_point(m, pt);
            inside = !(l.x < -sh.he.x || l.x > sh.he.x || l.y < -sh.he.y || l.y > sh.he.y || l.z < -sh.he.z || l.z > sh.he.z);
        } else if (type == NCB_SHAPE_CONVEX_HULL) {
            HullProjSetup u = hull_proj_setup(sh.hull, m, pt);
            Simplex s;
            V3 proj;
            inside = hull_project_gjk(u, pt, s, proj) != GJK_CLOSEST_POINTS;
        } else {
            inside = dot(sh.he, iso_inv_point(m, pt)) <= 0.f;
        }
        inside_out[k] = inside ? 1 : 0;
    }
}
}
