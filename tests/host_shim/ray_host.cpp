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
}
