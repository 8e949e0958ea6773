// TEST INFRASTRUCTURE ONLY — compiles the device functions of ncollide_b200/csrc/proximity.cu (and the GJK / simplex code of
// gjk.cuh they call) for the host through tests/host_shim/cuda_runtime.h and exposes them to the tests.  The kernels and
// launchers are compiled out (NCB_HOST_SHIM); what runs here is the same per-pair source the GPU threads execute.
#include "proximity.cu"

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
// proximity_pair for a batch; axis_io (optional, 4 floats per pair: xyz + valid flag) carries the detector's sep_axis in and out
void shim_proximity(const ncb_objects* objs, const ncb_hull_library* lib, uint64_t n_pairs, const uint32_t* pairs, const float* margins,
                    float* axis_io, uint8_t* out) {
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
        float margin = margins ? margins[p] : o.qlimit[i1] + o.qlimit[i2];
        V3 axis = v3(0.f, 0.f, 0.f);
        bool has_axis = false;
        if (axis_io) axis = v3(axis_io[4 * p], axis_io[4 * p + 1], axis_io[4 * p + 2]), has_axis = axis_io[4 * p + 3] != 0.f;
        out[p] = proximity_pair(o, H, i1, i2, margin, axis, has_axis);
        if (axis_io) axis_io[4 * p] = axis.x, axis_io[4 * p + 1] = axis.y, axis_io[4 * p + 2] = axis.z, axis_io[4 * p + 3] = has_axis ? 1.f : 0.f;
    }
}
}
