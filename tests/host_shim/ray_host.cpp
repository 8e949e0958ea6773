// TEST INFRASTRUCTURE ONLY — compiles the per-ray primitives of ncollide_b200/csrc/ray.cu (slab_toi, ray_triangle) for the host and
// evaluates the DEVICE SEMANTICS of the TriMesh ray cast by brute force: over all triangles whose AABB the ray enters, the hit with the
// smallest toi (<= max_toi), ties -> smallest face index; back-face hits are reported as face + n_tris; the winning normal is
// normalised (k_ray_cast's test_leaf / epilogue, without the tree).  Compared with the oracle's brute-force mode.
#include "ray.cu"

using namespace ncb;

extern "C" {
void shim_trimesh_ray_cast(uint32_t n_tris, const float* verts, const uint32_t* tris, const float* pose_tq, uint64_t n_rays, const float* origins,
                           const float* dirs, float max_toi, float* toi_out, uint32_t* face_out, float* normal_out) {
    Iso pose;
    bool has_pose = pose_tq != nullptr;
    if (has_pose) pose.t = v3(pose_tq[0], pose_tq[1], pose_tq[2]), pose.q = Quat{pose_tq[3], pose_tq[4], pose_tq[5], pose_tq[6]};
    for (uint64_t r = 0; r < n_rays; ++r) {
        V3 o = v3(origins[3 * r], origins[3 * r + 1], origins[3 * r + 2]);
        V3 d = v3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
        if (has_pose) {
            o = iso_inv_point(pose, o);
            d = iso_inv_vec(pose, d);
        }
        const V3 inv = v3(1.f / d.x, 1.f / d.y, 1.f / d.z);
        float best = NCB_FMAX;
        uint32_t best_face = 0xffffffffu;
        int best_side = 0;
        V3 best_n = v3(0.f, 0.f, 0.f);
        bool have = false;
        for (uint32_t t = 0; t < n_tris; ++t) {
            const float* pa = verts + 3 * (size_t)tris[3 * t];
            const float* pb = verts + 3 * (size_t)tris[3 * t + 1];
            const float* pc = verts + 3 * (size_t)tris[3 * t + 2];
            V3 a = v3(pa[0], pa[1], pa[2]), b = v3(pb[0], pb[1], pb[2]), c = v3(pc[0], pc[1], pc[2]);
            // k_tri_aabb: the leaf box
            float4 lo = make_float4(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z), 0.f);
            float4 hi = make_float4(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z), 0.f);
            if (!(slab_toi(lo, hi, o, d, inv, max_toi) >= 0.f)) continue;
            float toi;
            V3 n;
            int side;
            if (ray_triangle(a, b, c, o, d, toi, n, side) && toi <= max_toi) {
                if (!have || toi < best || (toi == best && t < best_face)) have = true, best = toi, best_face = t, best_side = side, best_n = n;
            }
        }
        if (have) {
            toi_out[r] = best;
            face_out[r] = best_side == 1 ? best_face + n_tris : best_face;
            V3 n = normalize(best_n);
            if (best_n.x == 0.f && best_n.y == 0.f && best_n.z == 0.f) n = best_n;
            if (has_pose) n = iso_mul_vec(pose, n);
            normal_out[3 * r] = n.x, normal_out[3 * r + 1] = n.y, normal_out[3 * r + 2] = n.z;
        } else {
            toi_out[r] = -1.f;
            face_out[r] = 0xffffffffu;
            normal_out[3 * r] = normal_out[3 * r + 1] = normal_out[3 * r + 2] = 0.f;
        }
    }
}

// ncollide2d Polyline: the device semantics by brute force (k_seg_aabb's leaf box, slab_toi2, segment_ray2, k_ray_cast4_polyline's
// epilogue): minimum toi over the edges whose box is entered and whose segment is hit, ties -> smallest edge.
void shim2_polyline_ray_cast(uint32_t n_edges, const float* pts, const uint32_t* edges, const float* pose, uint64_t n_rays, const float* origins,
                             const float* dirs, float max_toi_all, const float* max_tois, float* toi_out, uint32_t* feature_out, float* normal_out) {
    for (uint64_t r = 0; r < n_rays; ++r) {
        float ox = origins[2 * r], oy = origins[2 * r + 1], dx = dirs[2 * r], dy = dirs[2 * r + 1];
        if (pose) {
            float px = ox - pose[0], py = oy - pose[1], re = pose[2], im = pose[3];
            ox = re * px + im * py, oy = -im * px + re * py;
            float qx = dx, qy = dy;
            dx = re * qx + im * qy, dy = -im * qx + re * qy;
        }
        const float max_toi = max_tois ? max_tois[r] : max_toi_all;
        const float ivx = 1.f / dx, ivy = 1.f / dy;
        bool have = false;
        float best = NCB_FMAX, bnx = 0.f, bny = 0.f;
        uint32_t best_edge = 0xffffffffu;
        int best_face1 = 0;
        for (uint32_t e = 0; e < n_edges; ++e) {
            const float* a = pts + 2 * (size_t)edges[2 * e];
            const float* b = pts + 2 * (size_t)edges[2 * e + 1];
            float lox = -a[0] > -b[0] ? a[0] : b[0], loy = -a[1] > -b[1] ? a[1] : b[1];
            float hix = a[0] > b[0] ? a[0] : b[0], hiy = a[1] > b[1] ? a[1] : b[1];
            if (!(slab_toi2(lox, loy, hix, hiy, ox, oy, dx, dy, ivx, ivy, max_toi) >= 0.f)) continue;
            float toi, nx, ny;
            int face1;
            if (segment_ray2(a[0], a[1], b[0], b[1], ox, oy, dx, dy, toi, nx, ny, face1))
                if (!have || toi < best || (toi == best && e < best_edge)) have = true, best = toi, best_edge = e, bnx = nx, bny = ny, best_face1 = face1;
        }
        if (have) {
            toi_out[r] = best;
            feature_out[r] = best_face1 ? best_edge + n_edges : best_edge;
            float nx = bnx, ny = bny;
            if (pose) nx = pose[2] * bnx - pose[3] * bny, ny = pose[3] * bnx + pose[2] * bny;
            normal_out[2 * r] = nx, normal_out[2 * r + 1] = ny;
        } else {
            toi_out[r] = -1.f;
            feature_out[r] = 0xffffffffu;
            normal_out[2 * r] = normal_out[2 * r + 1] = 0.f;
        }
    }
}
}
