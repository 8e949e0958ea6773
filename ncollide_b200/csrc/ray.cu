// Batched first-hit ray casting against a TriMesh: device LBVH over the per-triangle AABBs + one thread per ray,
// ordered (near child first) stack traversal with best-hit pruning.
//
// Replaces (reference, file:line):
//   TriMesh::new / BVT::new_balanced      shape/trimesh.rs:100-144, partitioning/bvt.rs:281-404 (median-split BVT -> LBVH)
//   Triangle::local_aabb                  bounding_volume/aabb_triangle.rs:27-41
//   RayCast for TriMesh                   query/ray/ray_trimesh.rs:22-50,150-190
//   BVH::best_first_search                partitioning/bvh.rs:101-160 (BinaryHeap -> per-thread stack)
//   AABB::toi_with_ray                    query/ray/ray_aabb.rs:13-50
//   ray_intersection_with_triangle       query/ray/ray_triangle.rs:32-114
// Semantics: the slab test is monotone under box inclusion, so the set of accepted hits
// (leaf AABB hit AND triangle hit with toi <= max_toi) is tree independent; the device returns the minimum toi
// over that set, ties -> smallest face index.
#include <cstdio>
#include <cstring>
#include <string>
#include "ncb_internal.h"
#include "vec.cuh"

struct ncb_mesh {
    ncb_ctx* owner = nullptr;
    ncb_ctx* bvh = nullptr;  // private context holding the LBVH buffers of this mesh
    uint32_t n_verts = 0, n_tris = 0;
    ncb::DevBuf<float> verts;
    ncb::DevBuf<uint32_t> tris;
    ncb::DevBuf<float> d_in;   // staging for the host-buffer entry point
    ncb::DevBuf<float> d_out;
};

namespace ncb {

#define LEAF_BIT 0x80000000u

__global__ void __launch_bounds__(256) k_tri_aabb(const float* __restrict__ verts, const uint32_t* __restrict__ tris, uint32_t nt,
                                                  float4* __restrict__ lo, float4* __restrict__ hi) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    uint32_t ia = __ldg(tris + 3 * t), ib = __ldg(tris + 3 * t + 1), ic = __ldg(tris + 3 * t + 2);
    V3 a = v3(__ldg(verts + 3 * ia), __ldg(verts + 3 * ia + 1), __ldg(verts + 3 * ia + 2));
    V3 b = v3(__ldg(verts + 3 * ib), __ldg(verts + 3 * ib + 1), __ldg(verts + 3 * ib + 2));
    V3 c = v3(__ldg(verts + 3 * ic), __ldg(verts + 3 * ic + 1), __ldg(verts + 3 * ic + 2));
    lo[t] = make_float4(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z), 0.f);
    hi[t] = make_float4(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z), 0.f);
}

// AABB::toi_with_ray(identity, ray, max_toi, solid = true): returns tmin or -1 (miss)
NCB_HD float slab_toi(float4 lo, float4 hi, V3 o, V3 d, float max_toi) {
    float tmin = 0.f, tmax = max_toi;
    // x
    if (d.x == 0.f) {
        if (o.x < lo.x || o.x > hi.x) return -1.f;
    } else {
        float denom = 1.f / d.x;
        float n = (lo.x - o.x) * denom, f = (hi.x - o.x) * denom;
        if (n > f) {
            float t = n;
            n = f;
            f = t;
        }
        tmin = fmaxf(tmin, n);
        tmax = fminf(tmax, f);
        if (tmin > tmax) return -1.f;
    }
    if (d.y == 0.f) {
        if (o.y < lo.y || o.y > hi.y) return -1.f;
    } else {
        float denom = 1.f / d.y;
        float n = (lo.y - o.y) * denom, f = (hi.y - o.y) * denom;
        if (n > f) {
            float t = n;
            n = f;
            f = t;
        }
        tmin = fmaxf(tmin, n);
        tmax = fminf(tmax, f);
        if (tmin > tmax) return -1.f;
    }
    if (d.z == 0.f) {
        if (o.z < lo.z || o.z > hi.z) return -1.f;
    } else {
        float denom = 1.f / d.z;
        float n = (lo.z - o.z) * denom, f = (hi.z - o.z) * denom;
        if (n > f) {
            float t = n;
            n = f;
            f = t;
        }
        tmin = fmaxf(tmin, n);
        tmax = fminf(tmax, f);
        if (tmin > tmax) return -1.f;
    }
    return tmin;
}

// ray_intersection_with_triangle; side: 0 front, 1 back
NCB_HD bool ray_triangle(V3 a, V3 b, V3 c, V3 o, V3 dir, float& toi, V3& n_out, int& side) {
    V3 ab = b - a, ac = c - a;
    V3 n = cross(ab, ac);
    float d = dot(n, dir);
    if (d == 0.f) return false;
    V3 ap = o - a;
    float t = dot(ap, n);
    if ((t < 0.f && d < 0.f) || (t > 0.f && d > 0.f)) return false;
    side = d < 0.f ? 0 : 1;
    d = fabsf(d);
    V3 e = -cross(dir, ap);
    float v, w;
    if (t < 0.f) {
        v = -dot(ac, e);
        if (v < 0.f || v > d) return false;
        w = dot(ab, e);
        if (w < 0.f || v + w > d) return false;
        float invd = 1.f / d;
        toi = -t * invd;
        n_out = -n;  // normalised by the caller only for the winning hit
    } else {
        v = dot(ac, e);
        if (v < 0.f || v > d) return false;
        w = -dot(ab, e);
        if (w < 0.f || v + w > d) return false;
        float invd = 1.f / d;
        toi = t * invd;
        n_out = n;
    }
    return true;
}

struct RayArgs {
    const float4* nodes;
    const float4* leaf_lo;
    const float4* leaf_hi;
    const float* verts;
    const uint32_t* tris;
    uint32_t n_tris;
    Iso pose;
    int has_pose;
    const float* origins;
    const float* dirs;
    uint32_t n_rays;
    float max_toi;
    float* toi;
    uint32_t* face;
    float* normal;
};

__global__ void __launch_bounds__(128) k_ray_cast(RayArgs A) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.n_rays) return;
    V3 o = v3(__ldg(A.origins + 3 * r), __ldg(A.origins + 3 * r + 1), __ldg(A.origins + 3 * r + 2));
    V3 d = v3(__ldg(A.dirs + 3 * r), __ldg(A.dirs + 3 * r + 1), __ldg(A.dirs + 3 * r + 2));
    if (A.has_pose) {  // ray.inverse_transform_by(m)
        o = iso_inv_point(A.pose, o);
        d = iso_inv_vec(A.pose, d);
    }
    const float max_toi = A.max_toi;
    float best = NCB_FMAX;  // best accepted toi so far (bounded by max_toi through the triangle test)
    uint32_t best_face = 0xffffffffu;
    int best_side = 0;
    V3 best_n = v3(0.f, 0.f, 0.f);
    bool have = false;

    auto test_leaf = [&](uint32_t leaf_pos) {
        uint32_t t = __float_as_uint(__ldg(&A.leaf_lo[leaf_pos].w));
        uint32_t ia = __ldg(A.tris + 3 * t), ib = __ldg(A.tris + 3 * t + 1), ic = __ldg(A.tris + 3 * t + 2);
        V3 a = v3(__ldg(A.verts + 3 * ia), __ldg(A.verts + 3 * ia + 1), __ldg(A.verts + 3 * ia + 2));
        V3 b = v3(__ldg(A.verts + 3 * ib), __ldg(A.verts + 3 * ib + 1), __ldg(A.verts + 3 * ib + 2));
        V3 c = v3(__ldg(A.verts + 3 * ic), __ldg(A.verts + 3 * ic + 1), __ldg(A.verts + 3 * ic + 2));
        float toi;
        V3 n;
        int side;
        if (ray_triangle(a, b, c, o, d, toi, n, side) && toi <= max_toi) {
            if (!have || toi < best || (toi == best && t < best_face)) {
                have = true;
                best = toi;
                best_face = t;
                best_side = side;
                best_n = n;
            }
        }
    };

    if (A.n_tris == 1) {
        float4 lo = __ldg(&A.leaf_lo[0]), hi = __ldg(&A.leaf_hi[0]);
        if (slab_toi(lo, hi, o, d, max_toi) >= 0.f) test_leaf(0);
    } else if (A.n_tris >= 2) {
        uint32_t stack[64];
        float stack_t[64];
        int sp = 0;
        uint32_t node = 0;
        for (;;) {
            const float4* rec = A.nodes + 4 * (size_t)node;
            float4 Llo = __ldg(rec + 0), Lhi = __ldg(rec + 1), Rlo = __ldg(rec + 2), Rhi = __ldg(rec + 3);
            uint32_t left = __float_as_uint(Llo.w), right = __float_as_uint(Lhi.w);
            float tl = slab_toi(Llo, Lhi, o, d, max_toi);
            float tr = slab_toi(Rlo, Rhi, o, d, max_toi);
            bool goL = tl >= 0.f && !(have && tl > best);
            bool goR = tr >= 0.f && !(have && tr > best);
            if (goL && (left & LEAF_BIT)) {
                test_leaf(left & ~LEAF_BIT);
                goL = false;
                goR = goR && !(have && tr > best);
            }
            if (goR && (right & LEAF_BIT)) {
                test_leaf(right & ~LEAF_BIT);
                goR = false;
                goL = goL && !(have && tl > best);
            }
            if (goL && goR) {
                // near child first, far child on the stack with its entry distance
                bool left_first = tl <= tr;
                uint32_t nearn = left_first ? left : right, farn = left_first ? right : left;
                float fart = left_first ? tr : tl;
                if (sp < 64) {
                    stack[sp] = farn;
                    stack_t[sp] = fart;
                    sp++;
                }
                node = nearn;
            } else if (goL) {
                node = left;
            } else if (goR) {
                node = right;
            } else {
                bool found = false;
                while (sp > 0) {
                    sp--;
                    if (!(have && stack_t[sp] > best)) {
                        node = stack[sp];
                        found = true;
                        break;
                    }
                }
                if (!found) break;
            }
        }
    }
    if (have) {
        A.toi[r] = best;
        A.face[r] = best_side == 1 ? best_face + A.n_tris : best_face;  // ray_trimesh.rs:41-45
        if (A.normal) {
            V3 n = normalize(best_n);
            if (best_n.x == 0.f && best_n.y == 0.f && best_n.z == 0.f) n = best_n;
            // -n.normalize() == (-n).normalize() component-wise (division by the same norm)
            if (A.has_pose) n = iso_mul_vec(A.pose, n);
            A.normal[3 * r] = n.x, A.normal[3 * r + 1] = n.y, A.normal[3 * r + 2] = n.z;
        }
    } else {
        A.toi[r] = -1.f;
        A.face[r] = 0xffffffffu;
        if (A.normal) A.normal[3 * r] = A.normal[3 * r + 1] = A.normal[3 * r + 2] = 0.f;
    }
}

}  // namespace ncb

using namespace ncb;

#define CKM(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            char b__[512];                                                                                \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            ctx->err = b__;                                                                               \
            return NCB_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

extern "C" {

int ncb_trimesh_create(ncb_ctx* ctx, uint32_t n_verts, const float* xyz, uint32_t n_tris, const uint32_t* idx, ncb_mesh** out) {
    if (!ctx || !out || (n_verts && !xyz) || (n_tris && !idx)) return NCB_ERR_ARG;
    *out = nullptr;
    CKM(cudaSetDevice(ctx->device));
    for (size_t k = 0; k < 3 * (size_t)n_tris; ++k)
        if (idx[k] >= n_verts) {
            ctx->err = "ncb_trimesh_create: vertex index out of range";
            return NCB_ERR_ARG;
        }
    ncb_mesh* m = new ncb_mesh;
    m->owner = ctx;
    m->n_verts = n_verts;
    m->n_tris = n_tris;
    ncb_ctx* b = new ncb_ctx;
    b->device = ctx->device;
    b->stream = ctx->stream;
    b->sm_count = ctx->sm_count;
    m->bvh = b;
    cudaStream_t s = ctx->stream;
    cudaError_t e = cudaSuccess;
    auto fail = [&](const char* what) {
        ctx->err = std::string("ncb_trimesh_create: ") + what + ": " + cudaGetErrorString(e);
        ncb_trimesh_destroy(m);
        return NCB_ERR_CUDA;
    };
    if ((e = m->verts.reserve(3 * (size_t)n_verts + 3)) != cudaSuccess) return fail("alloc verts");
    if ((e = m->tris.reserve(3 * (size_t)n_tris + 3)) != cudaSuccess) return fail("alloc tris");
    if (n_verts && (e = cudaMemcpyAsync(m->verts.p, xyz, 12 * (size_t)n_verts, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail("copy");
    if (n_tris && (e = cudaMemcpyAsync(m->tris.p, idx, 12 * (size_t)n_tris, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail("copy");
    if (n_tris) {
        uint32_t n = n_tris;
        if ((e = b->aabb_lo.reserve(n)) != cudaSuccess || (e = b->aabb_hi.reserve(n)) != cudaSuccess || (e = b->keys_a.reserve(n)) != cudaSuccess ||
            (e = b->keys_b.reserve(n)) != cudaSuccess || (e = b->idx_a.reserve(n)) != cudaSuccess || (e = b->idx_b.reserve(n)) != cudaSuccess ||
            (e = b->leaf_lo.reserve(n)) != cudaSuccess || (e = b->leaf_hi.reserve(n)) != cudaSuccess ||
            (e = b->nodes.reserve(4 * (size_t)n)) != cudaSuccess || (e = b->parent.reserve(2 * (size_t)n)) != cudaSuccess ||
            (e = b->flags.reserve(n)) != cudaSuccess || (e = b->cub_tmp.reserve(lbvh_temp_bytes(n) + 256)) != cudaSuccess ||
            (e = b->counters.reserve(1)) != cudaSuccess)
            return fail("alloc bvh");
        DevCounters z;
        memset(&z, 0, sizeof z);
        for (int k = 0; k < 3; ++k) {
            z.bounds[k] = 0x7f7fffff;
            z.bounds[3 + k] = (int)0x80800000;
        }
        if ((e = cudaMemcpyAsync(b->counters.p, &z, sizeof z, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail("counters");
        k_tri_aabb<<<(n + 255) / 256, 256, 0, s>>>(m->verts.p, m->tris.p, n, b->aabb_lo.p, b->aabb_hi.p);
        if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_tri_aabb");
        if ((e = launch_lbvh_build(b, n, nullptr)) != cudaSuccess) return fail("lbvh build");
    }
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail("sync");
    if (n_tris) {
        // the sort scratch is not needed after the build
        b->keys_a.release(), b->keys_b.release(), b->idx_a.release(), b->idx_b.release(), b->cub_tmp.release();
        b->aabb_lo.release(), b->aabb_hi.release(), b->parent.release(), b->flags.release();
    }
    *out = m;
    return NCB_OK;
}

void ncb_trimesh_destroy(ncb_mesh* m) {
    if (!m) return;
    if (m->owner) {
        cudaSetDevice(m->owner->device);
        cudaStreamSynchronize(m->owner->stream);
    }
    if (m->bvh) {
        ncb_ctx* b = m->bvh;
        b->aabb_lo.release(), b->aabb_hi.release(), b->keys_a.release(), b->keys_b.release(), b->idx_a.release(), b->idx_b.release();
        b->cub_tmp.release(), b->leaf_lo.release(), b->leaf_hi.release(), b->nodes.release(), b->parent.release(), b->flags.release();
        b->counters.release();
        delete b;
    }
    m->verts.release(), m->tris.release(), m->d_in.release(), m->d_out.release();
    delete m;
}

int ncb_trimesh_ray_cast_device(ncb_mesh* m, const float* pose, uint32_t n_rays, const float* d_origins, const float* d_dirs, float max_toi,
                                float* d_toi, uint32_t* d_face, float* d_normal) {
    if (!m || (n_rays && (!d_origins || !d_dirs || !d_toi || !d_face))) return NCB_ERR_ARG;
    ncb_ctx* ctx = m->owner;
    CKM(cudaSetDevice(ctx->device));
    if (n_rays == 0) return NCB_OK;
    RayArgs A;
    A.nodes = m->bvh->nodes.p;
    A.leaf_lo = m->bvh->leaf_lo.p;
    A.leaf_hi = m->bvh->leaf_hi.p;
    A.verts = m->verts.p;
    A.tris = m->tris.p;
    A.n_tris = m->n_tris;
    A.has_pose = pose != nullptr;
    if (pose)
        A.pose = Iso{V3{pose[0], pose[1], pose[2]}, Quat{pose[3], pose[4], pose[5], pose[6]}};
    else
        A.pose = Iso{V3{0, 0, 0}, Quat{0, 0, 0, 1}};
    A.origins = d_origins;
    A.dirs = d_dirs;
    A.n_rays = n_rays;
    A.max_toi = max_toi;
    A.toi = d_toi;
    A.face = d_face;
    A.normal = d_normal;
    k_ray_cast<<<(n_rays + 127) / 128, 128, 0, ctx->stream>>>(A);
    CKM(cudaGetLastError());
    return NCB_OK;
}

int ncb_trimesh_ray_cast(ncb_mesh* m, const float* pose, uint32_t n_rays, const float* origins, const float* dirs, float max_toi, float* toi,
                         uint32_t* face, float* normal) {
    if (!m || (n_rays && (!origins || !dirs || !toi || !face))) return NCB_ERR_ARG;
    ncb_ctx* ctx = m->owner;
    CKM(cudaSetDevice(ctx->device));
    if (n_rays == 0) return NCB_OK;
    cudaStream_t s = ctx->stream;
    size_t n = n_rays;
    CKM(m->d_in.reserve(6 * n));
    CKM(m->d_out.reserve(5 * n));
    float* d_o = m->d_in.p;
    float* d_d = m->d_in.p + 3 * n;
    float* d_toi = m->d_out.p;
    uint32_t* d_face = reinterpret_cast<uint32_t*>(m->d_out.p + n);
    float* d_n = m->d_out.p + 2 * n;
    CKM(cudaMemcpyAsync(d_o, origins, 12 * n, cudaMemcpyHostToDevice, s));
    CKM(cudaMemcpyAsync(d_d, dirs, 12 * n, cudaMemcpyHostToDevice, s));
    int r = ncb_trimesh_ray_cast_device(m, pose, n_rays, d_o, d_d, max_toi, d_toi, d_face, normal ? d_n : nullptr);
    if (r) return r;
    CKM(cudaMemcpyAsync(toi, d_toi, 4 * n, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(face, d_face, 4 * n, cudaMemcpyDeviceToHost, s));
    if (normal) CKM(cudaMemcpyAsync(normal, d_n, 12 * n, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return NCB_OK;
}

}  // extern "C"
