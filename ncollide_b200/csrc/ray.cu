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
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#ifndef NCB_HOST_SHIM  // tests/host_shim compiles slab_toi / ray_triangle for the host (test infrastructure)
#include <cub/cub.cuh>
#endif
#include "ncb_internal.h"
#include "vec.cuh"

struct ncb_mesh {
    ncb_ctx* owner = nullptr;
    ncb_ctx* bvh = nullptr;  // private context holding the LBVH buffers of this mesh
    uint32_t n_verts = 0, n_tris = 0;
    ncb::DevBuf<float> verts;
    ncb::DevBuf<uint32_t> tris;
    ncb::DevBuf<float4> tri_packed;  // 3 float4 per leaf, Morton order
    ncb::DevBuf<float4> top_tile;    // top TOP_DEPTH levels of the BVH as an implicit complete tree (TMA-staged per CTA)
    ncb::DevBuf<uint32_t> rkeys_a, rkeys_b, ridx_a, ridx_b;  // ray sorting scratch
    ncb::DevBuf<uint8_t> rsort_tmp;
    float bounds[6] = {0, 0, 0, 0, 0, 0};
    int use_tile = 0, sort_rays = 0;
    ncb::DevBuf<float> d_in;   // staging for the host-buffer entry point
    ncb::DevBuf<float> d_out;
    ncb::DevBuf<float4> nodes4;  // 4-wide BVH: 8 float4 per binary node id (children of its two children), see k_build_bvh4
    ncb::DevBuf<float> uvs;      // per-vertex texture coordinates (TriMesh::uvs), optional
    int use_wide = 1;
    int dim2 = 0;                    // 1: a ncollide2d Polyline (n_verts points, n_tris edges; verts / tris hold 2 words per entry)
    ncb::DevBuf<float4> seg_packed;  // Polyline: a.x a.y b.x b.y per leaf, Morton order
    // host-buffer entry: the batch is cut into chunks that run [upload, cast, download] on a small pool of streams, so the copy
    // engines and the SMs work on different chunks at the same time and the casts of neighbouring chunks overlap
    static const int N_STREAMS = 4;
    cudaStream_t chunk_stream[N_STREAMS] = {};
    cudaEvent_t ev_begin = nullptr, ev_end[N_STREAMS] = {};
    bool pipeline_ready = false;
};

namespace ncb {

#define LEAF_BIT 0x80000000u
#ifndef NCB_HOST_SHIM  // kernels and TMA helpers: CUDA only

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

// ---- top-of-tree tile ------------------------------------------------------------------------------------------------
// Every ray walks the same first levels of the BVH.  They are copied once per mesh into an implicit complete binary
// tree of TOP_DEPTH levels (slot i -> children 2i+1, 2i+2; 64 B records, same format as the global nodes) and each CTA
// stages that tile into shared memory with ONE TMA bulk copy (cp.async.bulk + mbarrier).  Child words that stay inside
// the tile carry TILE_BIT | slot; the others keep their global meaning (LEAF_BIT | leaf position, or node index).
#define TOP_DEPTH 8
#define TOP_SLOTS ((1 << TOP_DEPTH) - 1)
#define TILE_BIT 0x40000000u

__global__ void __launch_bounds__(256) k_build_top_tile(const float4* __restrict__ nodes, uint32_t n_tris, float4* __restrict__ tile) {
    __shared__ uint32_t src[TOP_SLOTS];
    for (int i = threadIdx.x; i < TOP_SLOTS; i += blockDim.x) src[i] = 0xffffffffu;
    __syncthreads();
    if (threadIdx.x == 0) src[0] = 0;
    __syncthreads();
    for (int level = 0; level < TOP_DEPTH; ++level) {
        int first = (1 << level) - 1, count = 1 << level;
        for (int k = threadIdx.x; k < count; k += blockDim.x) {
            int i = first + k;
            uint32_t g = src[i];
            float4 r0 = make_float4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0;
            if (g != 0xffffffffu) {
                r0 = nodes[4 * (size_t)g + 0], r1 = nodes[4 * (size_t)g + 1], r2 = nodes[4 * (size_t)g + 2], r3 = nodes[4 * (size_t)g + 3];
                uint32_t left = __float_as_uint(r0.w), right = __float_as_uint(r1.w);
                if (level + 1 < TOP_DEPTH) {
                    if (!(left & LEAF_BIT)) {
                        src[2 * i + 1] = left;
                        r0.w = __uint_as_float(TILE_BIT | (uint32_t)(2 * i + 1));
                    }
                    if (!(right & LEAF_BIT)) {
                        src[2 * i + 2] = right;
                        r1.w = __uint_as_float(TILE_BIT | (uint32_t)(2 * i + 2));
                    }
                }
            }
            tile[4 * i + 0] = r0, tile[4 * i + 1] = r1, tile[4 * i + 2] = r2, tile[4 * i + 3] = r3;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
    uint32_t bar = smem_u32(mbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t done = 0, bar = smem_u32(mbar);
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    }
}

// ---- ray ordering ----------------------------------------------------------------------------------------------------
// Rays are traversed in Morton order of their origin (+ two direction sign bits): neighbouring lanes then walk the same
// part of the tree.  Results are written back at the ray's own index, so the order is invisible to the caller.
__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__global__ void __launch_bounds__(256) k_ray_keys(const float* __restrict__ origins, const float* __restrict__ dirs, uint32_t n, Iso pose,
                                                  int has_pose, float bx, float by, float bz, float sx, float sy, float sz,
                                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    V3 o = v3(origins[3 * r], origins[3 * r + 1], origins[3 * r + 2]);
    V3 d = v3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
    if (has_pose) {
        o = iso_inv_point(pose, o);
        d = iso_inv_vec(pose, d);
    }
    uint32_t ux = (uint32_t)fminf(fmaxf((o.x - bx) * sx, 0.f), 1023.f);
    uint32_t uy = (uint32_t)fminf(fmaxf((o.y - by) * sy, 0.f), 1023.f);
    uint32_t uz = (uint32_t)fminf(fmaxf((o.z - bz) * sz, 0.f), 1023.f);
    uint32_t m = (expand10(ux) << 2) | (expand10(uy) << 1) | expand10(uz);
    keys[r] = (m << 2) | (d.x < 0.f ? 2u : 0u) | (d.y < 0.f ? 1u : 0u);
    idx[r] = r;
}

__global__ void __launch_bounds__(256) k_pack_tris(const float* __restrict__ verts, const uint32_t* __restrict__ tris,
                                                   const float4* __restrict__ leaf_lo, uint32_t nt, float4* __restrict__ out) {
    uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= nt) return;
    uint32_t t = __float_as_uint(__ldg(&leaf_lo[pos].w));
    uint32_t ia = __ldg(tris + 3 * t), ib = __ldg(tris + 3 * t + 1), ic = __ldg(tris + 3 * t + 2);
    out[3 * (size_t)pos + 0] = make_float4(verts[3 * ia], verts[3 * ia + 1], verts[3 * ia + 2], __uint_as_float(t));
    out[3 * (size_t)pos + 1] = make_float4(verts[3 * ib], verts[3 * ib + 1], verts[3 * ib + 2], 0.f);
    out[3 * (size_t)pos + 2] = make_float4(verts[3 * ic], verts[3 * ic + 1], verts[3 * ic + 2], 0.f);
}

// Polyline::new (shape/polyline.rs:78-90): Segment::local_aabb = local_support_map_aabb (aabb_utils.rs:34-56) over
// Segment::local_support_point (segment.rs:182-191: a if a . dir > b . dir, else b); boxes get z = 0 for the 3-D LBVH build.
__global__ void __launch_bounds__(256) k_seg_aabb(const float* __restrict__ pts, const uint32_t* __restrict__ edges, uint32_t ne,
                                                  float4* __restrict__ lo, float4* __restrict__ hi) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    uint32_t ia = __ldg(edges + 2 * e), ib = __ldg(edges + 2 * e + 1);
    float ax = __ldg(pts + 2 * ia), ay = __ldg(pts + 2 * ia + 1), bx = __ldg(pts + 2 * ib), by = __ldg(pts + 2 * ib + 1);
    lo[e] = make_float4(-ax > -bx ? ax : bx, -ay > -by ? ay : by, 0.f, 0.f);
    hi[e] = make_float4(ax > bx ? ax : bx, ay > by ? ay : by, 0.f, 0.f);
}
__global__ void __launch_bounds__(256) k_pack_segs(const float* __restrict__ pts, const uint32_t* __restrict__ edges,
                                                   const float4* __restrict__ leaf_lo, uint32_t ne, float4* __restrict__ out) {
    uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= ne) return;
    uint32_t e = __float_as_uint(__ldg(&leaf_lo[pos].w));
    uint32_t ia = __ldg(edges + 2 * e), ib = __ldg(edges + 2 * e + 1);
    out[pos] = make_float4(pts[2 * ia], pts[2 * ia + 1], pts[2 * ib], pts[2 * ib + 1]);
}

#endif  // NCB_HOST_SHIM

// AABB::toi_with_ray(identity, ray, max_toi, solid = true) (ray_aabb.rs:13-50): returns tmin or -1 (miss).
// `inv` holds 1 / dir per axis, computed once per ray: the reference recomputes the same IEEE quotient for every box.
NCB_HD float slab_toi(float4 lo, float4 hi, V3 o, V3 d, V3 inv, float max_toi) {
    float tmin = 0.f, tmax = max_toi;
    if (d.x == 0.f) {
        if (o.x < lo.x || o.x > hi.x) return -1.f;
    } else {
        float n = (lo.x - o.x) * inv.x, f = (hi.x - o.x) * inv.x;
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
        float n = (lo.y - o.y) * inv.y, f = (hi.y - o.y) * inv.y;
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
        float n = (lo.z - o.z) * inv.z, f = (hi.z - o.z) * inv.z;
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
NCB_HD bool ray_triangle(V3 a, V3 b, V3 c, V3 o, V3 dir, float& toi, V3& n_out, int& side, float* vw = nullptr) {
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
        if (vw) vw[0] = v * invd, vw[1] = w * invd;
    } else {
        v = dot(ac, e);
        if (v < 0.f || v > d) return false;
        w = -dot(ab, e);
        if (w < 0.f || v + w > d) return false;
        float invd = 1.f / d;
        toi = t * invd;
        n_out = n;
        if (vw) vw[0] = v * invd, vw[1] = w * invd;
    }
    return true;
}

// ---- ncollide2d: Polyline ray casting (query/ray/ray_polyline.rs) ------------------------------------------------------------------
// AABB::toi_with_ray with DIM = 2 (ray_aabb.rs:13-50)
NCB_HD float slab_toi2(float lox, float loy, float hix, float hiy, float ox, float oy, float dx, float dy, float ivx, float ivy, float max_toi) {
    float tmin = 0.f, tmax = max_toi;
    if (dx == 0.f) {
        if (ox < lox || ox > hix) return -1.f;
    } else {
        float n = (lox - ox) * ivx, f = (hix - ox) * ivx;
        if (n > f) {
            float t = n;
            n = f;
            f = t;
        }
        tmin = fmaxf(tmin, n);
        tmax = fminf(tmax, f);
        if (tmin > tmax) return -1.f;
    }
    if (dy == 0.f) {
        if (oy < loy || oy > hiy) return -1.f;
    } else {
        float n = (loy - oy) * ivy, f = (hiy - oy) * ivy;
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

// RayCast for Segment::toi_and_normal_with_ray, dim2 (query/ray/ray_support_map.rs:219-293) with the segment in the ray's frame:
// closest_points_line_line_parameters_eps (closest_points_line_line.rs:27-70), then the collinear / crossing cases.  The normal is
// the segment's SCALED normal and max_toi is not applied, as in the reference.  feature: 0 Face(0) or a Vertex, 1 Face(1).
NCB_HD bool segment_ray2(float ax, float ay, float bx, float by, float ox, float oy, float dx, float dy, float& toi, float& nx, float& ny,
                         int& face1) {
    const float eps = NCB_EPS;
    float sdx = bx - ax, sdy = by - ay;
    float rx = ox - ax, ry = oy - ay;
    float a = dx * dx + dy * dy, e = sdx * sdx + sdy * sdy, f = sdx * rx + sdy * ry;
    float s, t;
    bool parallel = false;
    if (a <= eps && e <= eps) {
        s = 0.f, t = 0.f;
    } else if (a <= eps) {
        s = 0.f, t = f / e;
    } else {
        float c = dx * rx + dy * ry;
        if (e <= eps) {
            s = -c / a, t = 0.f;
        } else {
            float b = dx * sdx + dy * sdy;
            float ae = a * e, bb = b * b, denom = ae - bb;
            parallel = denom <= eps || ulps_eq(ae, bb);
            s = !parallel ? (b * f - c * e) / denom : 0.f;
            t = (b * s + f) / e;
        }
    }
    nx = sdy, ny = -sdx;
    face1 = 0;
    if (parallel) {
        float px = ax - ox, py = ay - oy;
        if (!(fabsf(px * nx + py * ny) < eps)) return false;
        float dist1 = px * dx + py * dy;
        float dist2 = dist1 + (sdx * dx + sdy * dy);
        bool p1 = dist1 >= 0.f, p2 = dist2 >= 0.f;
        if (p1 && p2) {
            toi = (dist1 <= dist2 ? dist1 : dist2) / (dx * dx + dy * dy);
            return true;
        }
        if (p1 || p2) {
            toi = 0.f;
            return true;
        }
        return false;
    }
    if (s >= 0.f && t >= 0.f && t <= 1.f) {
        toi = s;
        if (nx * dx + ny * dy > 0.f) nx = -nx, ny = -ny, face1 = 1;
        return true;
    }
    return false;
}

#ifndef NCB_HOST_SHIM  // the traversal kernel and the host entry points: CUDA only
struct RayArgs {
    const float4* nodes;
    const float4* leaf_lo;
    const float4* leaf_hi;
    const float4* tri_packed;
    const float4* top_tile;   // nullptr: start at the global root
    const uint32_t* perm;     // nullptr: natural ray order
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
    uint32_t* trav_overflow;
    const float4* nodes4;     // 4-wide nodes (k_ray_cast4)
    const float* max_tois;    // nullptr: max_toi for every ray
    const float* uvs;         // per-vertex uv (with tris): toi_and_normal_and_uv_with_ray
    const uint32_t* tris;
    float* uv_out;            // 2 floats per ray
};

// The best hit of a ray while its traversal runs, and the triangle test of a leaf (shared by the binary and the 4-wide kernel).
struct RayBest {
    float toi;
    uint32_t face;
    int side;
    V3 n;
    float v, w;
    bool have;
};
template <bool UV>
__device__ __forceinline__ void ray_test_leaf(const float4* __restrict__ tri_packed, uint32_t leaf_pos, V3 o, V3 d, float max_toi, RayBest& B) {
    // one 48 B record per leaf, in leaf (Morton) order: a.xyz | face id, b.xyz, c.xyz  (no index -> vertex chain)
    const float4* tp = tri_packed + 3 * (size_t)leaf_pos;
    float4 pa = __ldg(tp), pb = __ldg(tp + 1), pc = __ldg(tp + 2);
    uint32_t t = __float_as_uint(pa.w);
    V3 a = v3(pa.x, pa.y, pa.z), b = v3(pb.x, pb.y, pb.z), c = v3(pc.x, pc.y, pc.z);
    float toi, vw[2];
    V3 n;
    int side;
    if (ray_triangle(a, b, c, o, d, toi, n, side, UV ? vw : nullptr) && toi <= max_toi) {
        if (!B.have || toi < B.toi || (toi == B.toi && t < B.face)) {
            B.have = true, B.toi = toi, B.face = t, B.side = side, B.n = n;
            if (UV) B.v = vw[0], B.w = vw[1];
        }
    }
}
// RayCast for TriMesh, the tail of toi_and_normal(_and_uv)_with_ray (ray_trimesh.rs:39-49, 74-93)
__device__ __forceinline__ void ray_write_hit(const RayArgs& A, uint32_t r, const RayBest& B) {
    if (B.have) {
        A.toi[r] = B.toi;
        A.face[r] = B.side == 1 ? B.face + A.n_tris : B.face;  // ray_trimesh.rs:41-45
        if (A.normal) {
            V3 n = normalize(B.n);
            if (B.n.x == 0.f && B.n.y == 0.f && B.n.z == 0.f) n = B.n;
            // -n.normalize() == (-n).normalize() component-wise (division by the same norm)
            if (A.has_pose) n = iso_mul_vec(A.pose, n);
            A.normal[3 * r] = n.x, A.normal[3 * r + 1] = n.y, A.normal[3 * r + 2] = n.z;
        }
        if (A.uv_out) {
            float ux = 0.f, uy = 0.f;
            if (A.uvs) {  // uv1 * (1 - v - w) + uv2 * v + uv3 * w, evaluated like ray_trimesh.rs:83-84 / ray_triangle.rs:113
                uint32_t i0 = __ldg(A.tris + 3 * (size_t)B.face), i1 = __ldg(A.tris + 3 * (size_t)B.face + 1), i2 = __ldg(A.tris + 3 * (size_t)B.face + 2);
                float b0 = (-B.v - B.w) + 1.f, b1 = B.v, b2 = B.w;
                ux = (__ldg(A.uvs + 2 * i0) * b0 + __ldg(A.uvs + 2 * i1) * b1) + __ldg(A.uvs + 2 * i2) * b2;
                uy = (__ldg(A.uvs + 2 * i0 + 1) * b0 + __ldg(A.uvs + 2 * i1 + 1) * b1) + __ldg(A.uvs + 2 * i2 + 1) * b2;
            }
            A.uv_out[2 * r] = ux, A.uv_out[2 * r + 1] = uy;
        }
    } else {
        A.toi[r] = -1.f;
        A.face[r] = 0xffffffffu;
        if (A.normal) A.normal[3 * r] = A.normal[3 * r + 1] = A.normal[3 * r + 2] = 0.f;
        if (A.uv_out) A.uv_out[2 * r] = A.uv_out[2 * r + 1] = 0.f;
    }
}

// ---- 4-wide BVH ------------------------------------------------------------------------------------------------------
// One 128 B record per binary node id b: the (up to four) children of b's two children, as structure-of-arrays
// [lo.x x4][lo.y x4][lo.z x4][hi.x x4][hi.y x4][hi.z x4][ids x4][unused]; a child of b that is a leaf takes one slot itself.
// A ray visits half as many (dependent) node fetches as in the binary tree and tests four boxes per fetch.  The boxes that are
// tested are a subset of the binary tree's boxes and the slab test is monotone under box inclusion, so the set of leaves whose
// triangle is tested after pruning, and with it the hit (min toi, ties -> smallest face), is the same as before.
// Only records of nodes reachable by grandchild steps from the root are ever read.
#define EMPTY_CHILD 0xffffffffu
__global__ void __launch_bounds__(256) k_build_bvh4(const float4* __restrict__ nodes, uint32_t n_tris, float4* __restrict__ nodes4) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b + 1 >= n_tris) return;  // n_tris - 1 internal nodes
    float lo[4][3], hi[4][3];
    uint32_t id[4];
    int k = 0;
    const float4* rec = nodes + 4 * (size_t)b;
    float4 c_lo[2] = {rec[0], rec[2]}, c_hi[2] = {rec[1], rec[3]};
    uint32_t c_id[2] = {__float_as_uint(rec[0].w), __float_as_uint(rec[1].w)};
    for (int c = 0; c < 2; ++c) {
        if (c_id[c] & LEAF_BIT) {
            lo[k][0] = c_lo[c].x, lo[k][1] = c_lo[c].y, lo[k][2] = c_lo[c].z, hi[k][0] = c_hi[c].x, hi[k][1] = c_hi[c].y, hi[k][2] = c_hi[c].z;
            id[k++] = c_id[c];
        } else {
            const float4* g = nodes + 4 * (size_t)c_id[c];
            float4 g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
            lo[k][0] = g0.x, lo[k][1] = g0.y, lo[k][2] = g0.z, hi[k][0] = g1.x, hi[k][1] = g1.y, hi[k][2] = g1.z;
            id[k++] = __float_as_uint(g0.w);
            lo[k][0] = g2.x, lo[k][1] = g2.y, lo[k][2] = g2.z, hi[k][0] = g3.x, hi[k][1] = g3.y, hi[k][2] = g3.z;
            id[k++] = __float_as_uint(g1.w);
        }
    }
    for (; k < 4; ++k) {
        for (int a = 0; a < 3; ++a) lo[k][a] = NCB_FMAX, hi[k][a] = -NCB_FMAX;
        id[k] = EMPTY_CHILD;
    }
    float4* out = nodes4 + 8 * (size_t)b;
    for (int a = 0; a < 3; ++a) {
        out[a] = make_float4(lo[0][a], lo[1][a], lo[2][a], lo[3][a]);
        out[3 + a] = make_float4(hi[0][a], hi[1][a], hi[2][a], hi[3][a]);
    }
    out[6] = make_float4(__uint_as_float(id[0]), __uint_as_float(id[1]), __uint_as_float(id[2]), __uint_as_float(id[3]));
    out[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// The walk over the 4-wide tree, shared by the TriMesh (DIM 3) and the Polyline (DIM 2: the z planes are never read) kernels.
// B carries the best hit so far (B.have, B.toi); leaf(position) tests one primitive and may lower it.
template <int DIM, class BT, class LeafFn>
__device__ __forceinline__ void traverse_bvh4(const float4* __restrict__ nodes4, V3 o, V3 d, V3 inv, float max_toi, uint32_t* trav_overflow, BT& B,
                                              LeafFn leaf) {
    uint32_t stack[64];
    float stack_t[64];
    int sp = 0;
    uint32_t node = 0;
    for (;;) {
        const float4* rec = nodes4 + 8 * (size_t)node;
        float4 lx = __ldg(rec), ly = __ldg(rec + 1), hx = __ldg(rec + 3), hy = __ldg(rec + 4);
        float4 idw = __ldg(rec + 6);
        uint32_t id[4] = {__float_as_uint(idw.x), __float_as_uint(idw.y), __float_as_uint(idw.z), __float_as_uint(idw.w)};
        float t[4];
        if (DIM == 3) {
            float4 lz = __ldg(rec + 2), hz = __ldg(rec + 5);
            t[0] = slab_toi(make_float4(lx.x, ly.x, lz.x, 0.f), make_float4(hx.x, hy.x, hz.x, 0.f), o, d, inv, max_toi);
            t[1] = slab_toi(make_float4(lx.y, ly.y, lz.y, 0.f), make_float4(hx.y, hy.y, hz.y, 0.f), o, d, inv, max_toi);
            t[2] = id[2] == EMPTY_CHILD ? -1.f : slab_toi(make_float4(lx.z, ly.z, lz.z, 0.f), make_float4(hx.z, hy.z, hz.z, 0.f), o, d, inv, max_toi);
            t[3] = id[3] == EMPTY_CHILD ? -1.f : slab_toi(make_float4(lx.w, ly.w, lz.w, 0.f), make_float4(hx.w, hy.w, hz.w, 0.f), o, d, inv, max_toi);
        } else {
            t[0] = slab_toi2(lx.x, ly.x, hx.x, hy.x, o.x, o.y, d.x, d.y, inv.x, inv.y, max_toi);
            t[1] = slab_toi2(lx.y, ly.y, hx.y, hy.y, o.x, o.y, d.x, d.y, inv.x, inv.y, max_toi);
            t[2] = id[2] == EMPTY_CHILD ? -1.f : slab_toi2(lx.z, ly.z, hx.z, hy.z, o.x, o.y, d.x, d.y, inv.x, inv.y, max_toi);
            t[3] = id[3] == EMPTY_CHILD ? -1.f : slab_toi2(lx.w, ly.w, hx.w, hy.w, o.x, o.y, d.x, d.y, inv.x, inv.y, max_toi);
        }
        // leaves at once (their primitives may lower the bound for the internal children)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (t[k] >= 0.f && (id[k] & LEAF_BIT)) {
                if (!(B.have && t[k] > B.toi)) leaf(id[k] & ~LEAF_BIT);
                t[k] = -1.f;
            }
        // internal children still worth a visit, nearest first
        uint32_t cn[4];
        float ct[4];
        int nc = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (t[k] >= 0.f && !(B.have && t[k] > B.toi)) {
                int j = nc++;
                while (j > 0 && ct[j - 1] > t[k]) {
                    ct[j] = ct[j - 1], cn[j] = cn[j - 1];
                    --j;
                }
                ct[j] = t[k], cn[j] = id[k];
            }
        if (nc > 0) {
            for (int j = nc - 1; j >= 1; --j) {  // farthest first, so that the nearest of them is popped first
                if (sp < 64) {
                    stack[sp] = cn[j], stack_t[sp] = ct[j];
                    sp++;
                } else {
                    atomicAdd(trav_overflow, 1u);
                }
            }
            node = cn[0];
        } else {
            bool found = false;
            while (sp > 0) {
                sp--;
                if (!(B.have && stack_t[sp] > B.toi)) {
                    node = stack[sp];
                    found = true;
                    break;
                }
            }
            if (!found) break;
        }
    }
}

template <bool UV>
__global__ void __launch_bounds__(128) k_ray_cast4(RayArgs A) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.n_rays) return;
    V3 o = v3(__ldg(A.origins + 3 * r), __ldg(A.origins + 3 * r + 1), __ldg(A.origins + 3 * r + 2));
    V3 d = v3(__ldg(A.dirs + 3 * r), __ldg(A.dirs + 3 * r + 1), __ldg(A.dirs + 3 * r + 2));
    if (A.has_pose) {  // ray.inverse_transform_by(m)
        o = iso_inv_point(A.pose, o);
        d = iso_inv_vec(A.pose, d);
    }
    const float max_toi = A.max_tois ? __ldg(A.max_tois + r) : A.max_toi;
    const V3 inv = v3(1.f / d.x, 1.f / d.y, 1.f / d.z);  // only read on axes where d != 0
    RayBest B;
    B.toi = NCB_FMAX, B.face = 0xffffffffu, B.side = 0, B.n = v3(0.f, 0.f, 0.f), B.v = B.w = 0.f, B.have = false;
    if (A.n_tris == 1) {
        float4 lo = __ldg(&A.leaf_lo[0]), hi = __ldg(&A.leaf_hi[0]);
        if (slab_toi(lo, hi, o, d, inv, max_toi) >= 0.f) ray_test_leaf<UV>(A.tri_packed, 0, o, d, max_toi, B);
    } else if (A.n_tris >= 2) {
        traverse_bvh4<3>(A.nodes4, o, d, inv, max_toi, A.trav_overflow, B,
                         [&](uint32_t leaf_pos) { ray_test_leaf<UV>(A.tri_packed, leaf_pos, o, d, max_toi, B); });
    }
    ray_write_hit(A, r, B);
}

// RayCast for Polyline::toi_and_normal_with_ray (ray_polyline.rs:22-50,113-146): one thread per ray over the 4-wide tree of the
// edges' AABBs.  Accepted hits = edges whose AABB passes the slab test (with max_toi) and whose segment is hit (max_toi NOT applied,
// like RayCast for Segment in 2-D); the minimum toi wins, ties -> smallest edge.
struct Ray2Args {
    const float4* nodes4;
    const float4* leaf_lo;   // .w = edge id of the leaf
    const float4* leaf_hi;
    const float4* seg_packed;  // a.x a.y b.x b.y per leaf, leaf order
    uint32_t n_edges;
    float pose[4];           // x y re im
    int has_pose;
    const float* origins;    // 2 floats per ray
    const float* dirs;
    uint32_t n_rays;
    float max_toi;
    const float* max_tois;
    float* toi;
    uint32_t* feature;
    float* normal;           // 2 floats per ray (nullable)
    uint32_t* trav_overflow;
};
struct Ray2Best {
    float toi, nx, ny;
    uint32_t edge;
    int face1;
    bool have;
};
__device__ __forceinline__ void ray2_test_leaf(const Ray2Args& A, uint32_t leaf_pos, float ox, float oy, float dx, float dy, Ray2Best& B) {
    float4 sg = __ldg(A.seg_packed + leaf_pos);
    float toi, nx, ny;
    int face1;
    if (segment_ray2(sg.x, sg.y, sg.z, sg.w, ox, oy, dx, dy, toi, nx, ny, face1)) {
        uint32_t e = __float_as_uint(__ldg(&A.leaf_lo[leaf_pos].w));
        if (!B.have || toi < B.toi || (toi == B.toi && e < B.edge)) B.have = true, B.toi = toi, B.edge = e, B.nx = nx, B.ny = ny, B.face1 = face1;
    }
}
__global__ void __launch_bounds__(128) k_ray_cast4_polyline(Ray2Args A) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.n_rays) return;
    float2 o2 = __ldg(reinterpret_cast<const float2*>(A.origins) + r), d2 = __ldg(reinterpret_cast<const float2*>(A.dirs) + r);
    float ox = o2.x, oy = o2.y, dx = d2.x, dy = d2.y;
    if (A.has_pose) {  // ray.inverse_transform_by(m): the conjugate rotation of (origin - translation) and of dir
        float px = ox - A.pose[0], py = oy - A.pose[1], re = A.pose[2], im = A.pose[3];
        ox = re * px + im * py, oy = -im * px + re * py;
        float qx = dx, qy = dy;
        dx = re * qx + im * qy, dy = -im * qx + re * qy;
    }
    const float max_toi = A.max_tois ? __ldg(A.max_tois + r) : A.max_toi;
    Ray2Best B;
    B.toi = NCB_FMAX, B.nx = B.ny = 0.f, B.edge = 0xffffffffu, B.face1 = 0, B.have = false;
    const V3 o = v3(ox, oy, 0.f), d = v3(dx, dy, 0.f), inv = v3(1.f / dx, 1.f / dy, 0.f);
    if (A.n_edges == 1) {
        float4 lo = __ldg(&A.leaf_lo[0]), hi = __ldg(&A.leaf_hi[0]);
        if (slab_toi2(lo.x, lo.y, hi.x, hi.y, ox, oy, dx, dy, inv.x, inv.y, max_toi) >= 0.f) ray2_test_leaf(A, 0, ox, oy, dx, dy, B);
    } else if (A.n_edges >= 2) {
        traverse_bvh4<2>(A.nodes4, o, d, inv, max_toi, A.trav_overflow, B, [&](uint32_t leaf_pos) { ray2_test_leaf(A, leaf_pos, ox, oy, dx, dy, B); });
    }
    if (B.have) {
        A.toi[r] = B.toi;
        A.feature[r] = B.face1 ? B.edge + A.n_edges : B.edge;  // ray_polyline.rs:41-45
        if (A.normal) {  // m * res.normal
            float nx = B.nx, ny = B.ny;
            if (A.has_pose) {
                float re = A.pose[2], im = A.pose[3];
                nx = re * B.nx - im * B.ny, ny = im * B.nx + re * B.ny;
            }
            reinterpret_cast<float2*>(A.normal)[r] = make_float2(nx, ny);
        }
    } else {
        A.toi[r] = -1.f;
        A.feature[r] = 0xffffffffu;
        if (A.normal) reinterpret_cast<float2*>(A.normal)[r] = make_float2(0.f, 0.f);
    }
}

template <bool TILE>
__global__ void __launch_bounds__(128) k_ray_cast(RayArgs A) {
    __shared__ __align__(128) float4 s_tile[TILE ? TOP_SLOTS * 4 : 1];
    __shared__ __align__(8) uint64_t s_bar;
    if (TILE) {
        if (threadIdx.x == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) tma_load_1d(s_tile, A.top_tile, TOP_SLOTS * 64, &s_bar);
    }
    if (TILE) mbar_wait(&s_bar, 0);  // every thread observes the completed transaction before reading the tile
    // persistent CTAs: the tile is staged once per CTA, then the CTA walks the ray array with a grid stride
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n_rays; i += gridDim.x * blockDim.x) {
    uint32_t r = A.perm ? __ldg(A.perm + i) : i;
    V3 o = v3(__ldg(A.origins + 3 * r), __ldg(A.origins + 3 * r + 1), __ldg(A.origins + 3 * r + 2));
    V3 d = v3(__ldg(A.dirs + 3 * r), __ldg(A.dirs + 3 * r + 1), __ldg(A.dirs + 3 * r + 2));
    if (A.has_pose) {  // ray.inverse_transform_by(m)
        o = iso_inv_point(A.pose, o);
        d = iso_inv_vec(A.pose, d);
    }
    const float max_toi = A.max_tois ? __ldg(A.max_tois + r) : A.max_toi;
    const V3 inv = v3(1.f / d.x, 1.f / d.y, 1.f / d.z);  // only read on axes where d != 0
    RayBest B;
    B.toi = NCB_FMAX, B.face = 0xffffffffu, B.side = 0, B.n = v3(0.f, 0.f, 0.f), B.v = B.w = 0.f, B.have = false;
    float& best = B.toi;  // best accepted toi so far (bounded by max_toi through the triangle test)
    bool& have = B.have;
    auto test_leaf = [&](uint32_t leaf_pos) { ray_test_leaf<true>(A.tri_packed, leaf_pos, o, d, max_toi, B); };

    if (A.n_tris == 1) {
        float4 lo = __ldg(&A.leaf_lo[0]), hi = __ldg(&A.leaf_hi[0]);
        if (slab_toi(lo, hi, o, d, inv, max_toi) >= 0.f) test_leaf(0);
    } else if (A.n_tris >= 2) {
        uint32_t stack[64];
        float stack_t[64];
        int sp = 0;
        uint32_t node = TILE ? TILE_BIT : 0;
        for (;;) {
            float4 Llo, Lhi, Rlo, Rhi;
            if (TILE && (node & TILE_BIT)) {
                const float4* rec = s_tile + 4 * (node & ~TILE_BIT);
                Llo = rec[0], Lhi = rec[1], Rlo = rec[2], Rhi = rec[3];
            } else {
                const float4* rec = A.nodes + 4 * (size_t)node;
                Llo = __ldg(rec + 0), Lhi = __ldg(rec + 1), Rlo = __ldg(rec + 2), Rhi = __ldg(rec + 3);
            }
            uint32_t left = __float_as_uint(Llo.w), right = __float_as_uint(Lhi.w);
            float tl = slab_toi(Llo, Lhi, o, d, inv, max_toi);
            float tr = slab_toi(Rlo, Rhi, o, d, inv, max_toi);
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
                } else {
                    atomicAdd(A.trav_overflow, 1u);
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
    ray_write_hit(A, r, B);
    }  // ray loop
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

// TriMesh::new (dim2 == 0: 3 floats per vertex, 3 indices per triangle) and Polyline::new (dim2 == 1: 2 floats per point, 2 indices per edge)
static int mesh_create(ncb_ctx* ctx, int dim2, uint32_t n_verts, const float* xyz, uint32_t n_tris, const uint32_t* idx, ncb_mesh** out) {
    const size_t W = dim2 ? 2 : 3;  // words per vertex and per primitive
    const char* who = dim2 ? "ncb2d_polyline_create: " : "ncb_trimesh_create: ";
    if (!ctx || !out || (n_verts && !xyz) || (n_tris && !idx)) return NCB_ERR_ARG;
    *out = nullptr;
    CKM(cudaSetDevice(ctx->device));
    for (size_t k = 0; k < W * (size_t)n_tris; ++k)
        if (idx[k] >= n_verts) {
            ctx->err = std::string(who) + "vertex index out of range";
            return NCB_ERR_ARG;
        }
    ncb_mesh* m = new ncb_mesh;
    m->owner = ctx;
    m->n_verts = n_verts;
    m->n_tris = n_tris;
    m->dim2 = dim2;
    ncb_ctx* b = new ncb_ctx;
    b->device = ctx->device;
    b->stream = ctx->stream;
    b->sm_count = ctx->sm_count;
    m->bvh = b;
    cudaStream_t s = ctx->stream;
    cudaError_t e = cudaSuccess;
    auto fail = [&](const char* what) {
        ctx->err = std::string(who) + what + ": " + cudaGetErrorString(e);
        ncb_trimesh_destroy(m);
        return NCB_ERR_CUDA;
    };
    if ((e = m->verts.reserve(W * (size_t)n_verts + 3)) != cudaSuccess) return fail("alloc verts");
    if ((e = m->tris.reserve(W * (size_t)n_tris + 3)) != cudaSuccess) return fail("alloc tris");
    if (n_verts && (e = cudaMemcpyAsync(m->verts.p, xyz, 4 * W * (size_t)n_verts, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail("copy");
    if (n_tris && (e = cudaMemcpyAsync(m->tris.p, idx, 4 * W * (size_t)n_tris, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail("copy");
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
        if (dim2)
            k_seg_aabb<<<(n + 255) / 256, 256, 0, s>>>(m->verts.p, m->tris.p, n, b->aabb_lo.p, b->aabb_hi.p);
        else
            k_tri_aabb<<<(n + 255) / 256, 256, 0, s>>>(m->verts.p, m->tris.p, n, b->aabb_lo.p, b->aabb_hi.p);
        if ((e = cudaGetLastError()) != cudaSuccess) return fail("primitive AABBs");
        if ((e = launch_lbvh_build(b, n, nullptr)) != cudaSuccess) return fail("lbvh build");
        if (dim2) {
            if ((e = m->seg_packed.reserve(n)) != cudaSuccess) return fail("alloc packed segments");
            k_pack_segs<<<(n + 255) / 256, 256, 0, s>>>(m->verts.p, m->tris.p, b->leaf_lo.p, n, m->seg_packed.p);
        } else {
            if ((e = m->tri_packed.reserve(3 * (size_t)n)) != cudaSuccess) return fail("alloc packed triangles");
            k_pack_tris<<<(n + 255) / 256, 256, 0, s>>>(m->verts.p, m->tris.p, b->leaf_lo.p, n, m->tri_packed.p);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return fail("pack primitives");
        if (n >= 2) {
            if ((e = m->nodes4.reserve(8 * (size_t)(n - 1))) != cudaSuccess) return fail("alloc 4-wide nodes");
            k_build_bvh4<<<(n + 255) / 256, 256, 0, s>>>(b->nodes.p, n, m->nodes4.p);
            if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_build_bvh4");
        }
        if ((e = m->top_tile.reserve(TOP_SLOTS * 4)) != cudaSuccess) return fail("alloc top tile");
        if (n >= 2 && !dim2) k_build_top_tile<<<1, 256, 0, s>>>(b->nodes.p, n, m->top_tile.p);
        if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_build_top_tile");
        // mesh bounds = union of the root's two child boxes (or the single leaf box)
        float4 rec[4];
        if (n >= 2) {
            if ((e = cudaMemcpyAsync(rec, b->nodes.p, 64, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail("bounds");
        } else {
            if ((e = cudaMemcpyAsync(&rec[0], b->leaf_lo.p, 16, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail("bounds");
            if ((e = cudaMemcpyAsync(&rec[1], b->leaf_hi.p, 16, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail("bounds");
            rec[2] = rec[0], rec[3] = rec[1];
        }
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail("sync");
        m->bounds[0] = fminf(rec[0].x, rec[2].x), m->bounds[1] = fminf(rec[0].y, rec[2].y), m->bounds[2] = fminf(rec[0].z, rec[2].z);
        m->bounds[3] = fmaxf(rec[1].x, rec[3].x), m->bounds[4] = fmaxf(rec[1].y, rec[3].y), m->bounds[5] = fmaxf(rec[1].z, rec[3].z);
        if (const char* v = getenv("NCB_RAY_TILE")) m->use_tile = atoi(v);
        if (const char* v = getenv("NCB_RAY_SORT")) m->sort_rays = atoi(v);
        if (const char* v = getenv("NCB_RAY_WIDE")) m->use_wide = atoi(v);
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

int ncb_trimesh_create(ncb_ctx* ctx, uint32_t n_verts, const float* xyz, uint32_t n_tris, const uint32_t* idx, ncb_mesh** out) {
    return mesh_create(ctx, 0, n_verts, xyz, n_tris, idx, out);
}

// Polyline::new(points, Some(indices)) (shape/polyline.rs:57-120); edges == NULL: the line strip 0-1, 1-2, ... (n_edges is ignored)
int ncb2d_polyline_create(ncb_ctx* ctx, uint32_t n_points, const float* xy, uint32_t n_edges, const uint32_t* edges, ncb_mesh** out) {
    if (edges || !ctx || !out) return mesh_create(ctx, 1, n_points, xy, n_edges, edges, out);
    uint32_t ne = n_points ? n_points - 1 : 0;
    std::vector<uint32_t> strip(2 * (size_t)ne);
    for (uint32_t i = 0; i < ne; ++i) strip[2 * i] = i, strip[2 * i + 1] = i + 1;
    return mesh_create(ctx, 1, n_points, xy, ne, strip.data(), out);
}
void ncb2d_polyline_destroy(ncb_mesh* m) { ncb_trimesh_destroy(m); }

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
    m->verts.release(), m->tris.release(), m->d_in.release(), m->d_out.release(), m->tri_packed.release(), m->top_tile.release();
    m->nodes4.release(), m->uvs.release(), m->seg_packed.release();
    for (int k = 0; k < ncb_mesh::N_STREAMS; ++k) {
        if (m->chunk_stream[k]) cudaStreamDestroy(m->chunk_stream[k]);
        if (m->ev_end[k]) cudaEventDestroy(m->ev_end[k]);
    }
    if (m->ev_begin) cudaEventDestroy(m->ev_begin);
    m->rkeys_a.release(), m->rkeys_b.release(), m->ridx_a.release(), m->ridx_b.release(), m->rsort_tmp.release();
    delete m;
}

// TriMesh::uvs (shape/trimesh.rs: `uvs: Option<Vec<Point2<N>>>`): per-vertex texture coordinates, or NULL to clear them.
int ncb_trimesh_set_uvs(ncb_mesh* m, const float* uvs) {
    if (!m || m->dim2) return NCB_ERR_ARG;
    ncb_ctx* ctx = m->owner;
    CKM(cudaSetDevice(ctx->device));
    if (!uvs) {
        m->uvs.release();
        return NCB_OK;
    }
    CKM(m->uvs.reserve(2 * (size_t)m->n_verts + 2));
    CKM(cudaMemcpyAsync(m->uvs.p, uvs, 8 * (size_t)m->n_verts, cudaMemcpyHostToDevice, ctx->stream));
    CKM(cudaStreamSynchronize(ctx->stream));
    return NCB_OK;
}

static int ensure_chunk_streams(ncb_mesh* m) {
    ncb_ctx* ctx = m->owner;
    if (m->pipeline_ready) return NCB_OK;
    for (int k = 0; k < ncb_mesh::N_STREAMS; ++k) {
        CKM(cudaStreamCreateWithFlags(&m->chunk_stream[k], cudaStreamNonBlocking));
        CKM(cudaEventCreateWithFlags(&m->ev_end[k], cudaEventDisableTiming));
    }
    CKM(cudaEventCreateWithFlags(&m->ev_begin, cudaEventDisableTiming));
    m->pipeline_ready = true;
    return NCB_OK;
}

// One kernel launch over device-resident rays [0, n_rays) on `stream`.
static int launch_ray_cast(ncb_mesh* m, const float* pose, uint32_t n_rays, const float* d_origins, const float* d_dirs, float max_toi,
                           const float* d_max_tois, float* d_toi, uint32_t* d_face, float* d_normal, float* d_uv, cudaStream_t stream,
                           bool allow_sort) {
    ncb_ctx* ctx = m->owner;
    RayArgs A;
    A.nodes = m->bvh->nodes.p;
    A.leaf_lo = m->bvh->leaf_lo.p;
    A.leaf_hi = m->bvh->leaf_hi.p;
    A.tri_packed = m->tri_packed.p;
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
    A.trav_overflow = trav_overflow_counter(ctx);
    A.top_tile = (m->use_tile && m->n_tris >= 2) ? m->top_tile.p : nullptr;
    A.perm = nullptr;
    A.nodes4 = m->nodes4.p;
    A.max_tois = d_max_tois;
    A.uvs = m->uvs.p;
    A.tris = m->tris.p;
    A.uv_out = d_uv;
    if (m->use_wide && !A.top_tile && !(allow_sort && m->sort_rays)) {
        // Default: the 4-wide tree, one CTA per 128 rays (profiles/r2_ray_variants.txt).
        uint32_t grid = (n_rays + 127) / 128;
        if (d_uv)
            k_ray_cast4<true><<<grid, 128, 0, stream>>>(A);
        else
            k_ray_cast4<false><<<grid, 128, 0, stream>>>(A);
        CKM(cudaGetLastError());
        return NCB_OK;
    }
    if (allow_sort && m->sort_rays && n_rays >= 4096) {
        size_t n = n_rays;
        CKM(m->rkeys_a.reserve(n));
        CKM(m->rkeys_b.reserve(n));
        CKM(m->ridx_a.reserve(n));
        CKM(m->ridx_b.reserve(n));
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                        (int)n_rays, 0, 32);
        CKM(m->rsort_tmp.reserve(bytes + 256));
        // quantise origins over the mesh bounds grown by half their size on each side
        float ex = m->bounds[3] - m->bounds[0], ey = m->bounds[4] - m->bounds[1], ez = m->bounds[5] - m->bounds[2];
        float bx = m->bounds[0] - 0.5f * ex, by = m->bounds[1] - 0.5f * ey, bz = m->bounds[2] - 0.5f * ez;
        float sx = 1023.f / fmaxf(2.f * ex, 1e-20f), sy = 1023.f / fmaxf(2.f * ey, 1e-20f), sz = 1023.f / fmaxf(2.f * ez, 1e-20f);
        k_ray_keys<<<(n_rays + 255) / 256, 256, 0, stream>>>(d_origins, d_dirs, n_rays, A.pose, A.has_pose, bx, by, bz, sx, sy, sz,
                                                            m->rkeys_a.p, m->ridx_a.p);
        bytes = m->rsort_tmp.cap;
        CKM(cub::DeviceRadixSort::SortPairs(m->rsort_tmp.p, bytes, m->rkeys_a.p, m->rkeys_b.p, m->ridx_a.p, m->ridx_b.p, (int)n_rays, 0, 32, stream));
        A.perm = m->ridx_b.p;
    }
    // Binary-tree variants, measured on B200, 1M rays vs 1M-triangle terrain (profiles/r1_ray_variants.txt): one CTA per 128 rays
    // 0.753 ms; persistent CTAs (12 / SM, grid stride) 0.951 ms (ray costs vary, the hardware CTA scheduler balances better);
    // TMA-staged top tile 0.827 ms (the top levels are L1-resident anyway); Morton-sorted rays 0.79 ms (sort not repaid).
    // NCB_RAY_WIDE=0 / NCB_RAY_BPSM / NCB_RAY_TILE / NCB_RAY_SORT select them.
    static int ray_bpsm = getenv("NCB_RAY_BPSM") ? atoi(getenv("NCB_RAY_BPSM")) : 0;
    uint32_t need = (n_rays + 127) / 128;
    uint32_t grid = ray_bpsm > 0 ? (uint32_t)(ctx->sm_count * ray_bpsm) : need;
    if (grid > need) grid = need;
    if (A.top_tile)
        k_ray_cast<true><<<grid, 128, 0, stream>>>(A);
    else
        k_ray_cast<false><<<grid, 128, 0, stream>>>(A);
    CKM(cudaGetLastError());
    return NCB_OK;
}

int ncb_trimesh_ray_cast_device(ncb_mesh* m, const float* pose, uint32_t n_rays, const float* d_origins, const float* d_dirs, float max_toi,
                                float* d_toi, uint32_t* d_face, float* d_normal) {
    if (!m || m->dim2 || (n_rays && (!d_origins || !d_dirs || !d_toi || !d_face))) return NCB_ERR_ARG;
    ncb_ctx* ctx = m->owner;
    CKM(cudaSetDevice(ctx->device));
    if (n_rays == 0) return NCB_OK;
    return launch_ray_cast(m, pose, n_rays, d_origins, d_dirs, max_toi, nullptr, d_toi, d_face, d_normal, nullptr, ctx->stream, true);
}

// Host buffers in, host buffers out (RayCast::toi_and_normal_and_uv_with_ray for a batch; uv NULL = toi_and_normal_with_ray).
// The batch runs as a pipeline of up to 16 chunks on three streams: while chunk k is cast, chunk k + 1 is uploaded and the results of
// chunk k - 1 are downloaded (PCIe is full duplex), so with pinned host buffers the call costs about max(copies, kernel) instead of
// their sum.  max_tois: per-ray limits (NULL: max_toi for every ray).
int ncb_trimesh_ray_cast_uv(ncb_mesh* m, const float* pose, uint32_t n_rays, const float* origins, const float* dirs, float max_toi,
                            const float* max_tois, float* toi, uint32_t* face, float* normal, float* uv) {
    if (!m || m->dim2 || (n_rays && (!origins || !dirs || !toi || !face))) return NCB_ERR_ARG;
    ncb_ctx* ctx = m->owner;
    CKM(cudaSetDevice(ctx->device));
    if (n_rays == 0) return NCB_OK;
    cudaStream_t s = ctx->stream;
    size_t n = n_rays;
    CKM(m->d_in.reserve(7 * n));
    CKM(m->d_out.reserve(7 * n));
    if (int rc = ensure_chunk_streams(m)) return rc;
    float* d_o = m->d_in.p;
    float* d_d = m->d_in.p + 3 * n;
    float* d_mt = m->d_in.p + 6 * n;
    float* d_toi = m->d_out.p;
    uint32_t* d_face = reinterpret_cast<uint32_t*>(m->d_out.p + n);
    float* d_n = m->d_out.p + 2 * n;
    float* d_uv = m->d_out.p + 5 * n;
    // A cast has a latency floor (a ray is a chain of dependent node fetches), so chunks must be large enough to fill the GPU and
    // their casts must overlap: chunks of 256 k rays on 4 streams (profiles/r2_ray_variants.txt).
    static int chunk_rays = getenv("NCB_RAY_CHUNK") ? atoi(getenv("NCB_RAY_CHUNK")) : 262144;
    uint32_t per = (uint32_t)(chunk_rays > 0 ? chunk_rays : 262144);
    per = (per + 127) & ~127u;
    uint32_t n_chunks = (n_rays + per - 1) / per;
    if (n_chunks > 32) {
        per = ((n_rays + 31) / 32 + 127) & ~127u;
        n_chunks = (n_rays + per - 1) / per;
    }
    // the staging buffers may still be read by earlier work of the context's stream (ncb_trimesh_ray_cast_device is asynchronous)
    CKM(cudaEventRecord(m->ev_begin, s));
    int used = (int)std::min<uint32_t>(n_chunks, ncb_mesh::N_STREAMS);
    for (int j = 0; j < used; ++j) CKM(cudaStreamWaitEvent(m->chunk_stream[j], m->ev_begin, 0));
    int rc = NCB_OK;
    for (uint32_t k = 0; k < n_chunks && rc == NCB_OK; ++k) {
        cudaStream_t cs = m->chunk_stream[k % ncb_mesh::N_STREAMS];
        size_t off = (size_t)k * per, cnt = std::min<size_t>(per, n - off);
        CKM(cudaMemcpyAsync(d_o + 3 * off, origins + 3 * off, 12 * cnt, cudaMemcpyHostToDevice, cs));
        CKM(cudaMemcpyAsync(d_d + 3 * off, dirs + 3 * off, 12 * cnt, cudaMemcpyHostToDevice, cs));
        if (max_tois) CKM(cudaMemcpyAsync(d_mt + off, max_tois + off, 4 * cnt, cudaMemcpyHostToDevice, cs));
        rc = launch_ray_cast(m, pose, (uint32_t)cnt, d_o + 3 * off, d_d + 3 * off, max_toi, max_tois ? d_mt + off : nullptr, d_toi + off,
                             d_face + off, normal ? d_n + 3 * off : nullptr, uv ? d_uv + 2 * off : nullptr, cs, false);
        if (rc) break;
        CKM(cudaMemcpyAsync(toi + off, d_toi + off, 4 * cnt, cudaMemcpyDeviceToHost, cs));
        CKM(cudaMemcpyAsync(face + off, d_face + off, 4 * cnt, cudaMemcpyDeviceToHost, cs));
        if (normal) CKM(cudaMemcpyAsync(normal + 3 * off, d_n + 3 * off, 12 * cnt, cudaMemcpyDeviceToHost, cs));
        if (uv) CKM(cudaMemcpyAsync(uv + 2 * off, d_uv + 2 * off, 8 * cnt, cudaMemcpyDeviceToHost, cs));
    }
    for (int j = 0; j < used; ++j) {  // the context's stream continues after the chunks (later device casts reuse nothing of this call)
        cudaStreamSynchronize(m->chunk_stream[j]);
    }
    if (rc) return rc;
    CKM(cudaGetLastError());
    return NCB_OK;
}

int ncb_trimesh_ray_cast(ncb_mesh* m, const float* pose, uint32_t n_rays, const float* origins, const float* dirs, float max_toi, float* toi,
                         uint32_t* face, float* normal) {
    return ncb_trimesh_ray_cast_uv(m, pose, n_rays, origins, dirs, max_toi, nullptr, toi, face, normal, nullptr);
}

// ---- ncollide2d Polyline -----------------------------------------------------------------------------------------------------------
static int launch_ray_cast2(ncb_mesh* m, const float* pose, uint32_t n_rays, const float* d_origins, const float* d_dirs, float max_toi,
                            const float* d_max_tois, float* d_toi, uint32_t* d_feature, float* d_normal, cudaStream_t stream) {
    ncb_ctx* ctx = m->owner;
    Ray2Args A;
    A.nodes4 = m->nodes4.p;
    A.leaf_lo = m->bvh->leaf_lo.p;
    A.leaf_hi = m->bvh->leaf_hi.p;
    A.seg_packed = m->seg_packed.p;
    A.n_edges = m->n_tris;
    A.has_pose = pose != nullptr;
    A.pose[0] = pose ? pose[0] : 0.f, A.pose[1] = pose ? pose[1] : 0.f, A.pose[2] = pose ? pose[2] : 1.f, A.pose[3] = pose ? pose[3] : 0.f;
    A.origins = d_origins;
    A.dirs = d_dirs;
    A.n_rays = n_rays;
    A.max_toi = max_toi;
    A.max_tois = d_max_tois;
    A.toi = d_toi;
    A.feature = d_feature;
    A.normal = d_normal;
    A.trav_overflow = trav_overflow_counter(ctx);
    k_ray_cast4_polyline<<<(n_rays + 127) / 128, 128, 0, stream>>>(A);
    CKM(cudaGetLastError());
    return NCB_OK;
}

int ncb2d_polyline_ray_cast_device(ncb_mesh* m, const float* pose, uint32_t n_rays, const float* d_origins, const float* d_dirs, float max_toi,
                                   const float* d_max_tois, float* d_toi, uint32_t* d_feature, float* d_normal) {
    if (!m || !m->dim2 || (n_rays && (!d_origins || !d_dirs || !d_toi || !d_feature))) return NCB_ERR_ARG;
    ncb_ctx* ctx = m->owner;
    CKM(cudaSetDevice(ctx->device));
    if (n_rays == 0) return NCB_OK;
    return launch_ray_cast2(m, pose, n_rays, d_origins, d_dirs, max_toi, d_max_tois, d_toi, d_feature, d_normal, ctx->stream);
}

// Host buffers in, host buffers out: the chunk pipeline of ncb_trimesh_ray_cast_uv (16-20 B per ray in, 8-16 B out).
int ncb2d_polyline_ray_cast(ncb_mesh* m, const float* pose, uint32_t n_rays, const float* origins, const float* dirs, float max_toi,
                            const float* max_tois, float* toi, uint32_t* feature, float* normal) {
    if (!m || !m->dim2 || (n_rays && (!origins || !dirs || !toi || !feature))) return NCB_ERR_ARG;
    ncb_ctx* ctx = m->owner;
    CKM(cudaSetDevice(ctx->device));
    if (n_rays == 0) return NCB_OK;
    cudaStream_t s = ctx->stream;
    size_t n = n_rays;
    CKM(m->d_in.reserve(5 * n + 8));
    CKM(m->d_out.reserve(4 * n + 8));
    if (int rc = ensure_chunk_streams(m)) return rc;
    float* d_o = m->d_in.p;
    float* d_d = m->d_in.p + 2 * n;
    float* d_mt = m->d_in.p + 4 * n;
    float* d_n = m->d_out.p;  // float2 per ray first: 8-byte aligned
    float* d_toi = m->d_out.p + 2 * n;
    uint32_t* d_feat = reinterpret_cast<uint32_t*>(m->d_out.p + 3 * n);
    static int chunk_rays = getenv("NCB_RAY_CHUNK") ? atoi(getenv("NCB_RAY_CHUNK")) : 262144;
    uint32_t per = (uint32_t)(chunk_rays > 0 ? chunk_rays : 262144);
    per = (per + 127) & ~127u;
    uint32_t n_chunks = (n_rays + per - 1) / per;
    if (n_chunks > 32) {
        per = ((n_rays + 31) / 32 + 127) & ~127u;
        n_chunks = (n_rays + per - 1) / per;
    }
    CKM(cudaEventRecord(m->ev_begin, s));
    int used = (int)std::min<uint32_t>(n_chunks, ncb_mesh::N_STREAMS);
    for (int j = 0; j < used; ++j) CKM(cudaStreamWaitEvent(m->chunk_stream[j], m->ev_begin, 0));
    int rc = NCB_OK;
    for (uint32_t k = 0; k < n_chunks && rc == NCB_OK; ++k) {
        cudaStream_t cs = m->chunk_stream[k % ncb_mesh::N_STREAMS];
        size_t off = (size_t)k * per, cnt = std::min<size_t>(per, n - off);
        CKM(cudaMemcpyAsync(d_o + 2 * off, origins + 2 * off, 8 * cnt, cudaMemcpyHostToDevice, cs));
        CKM(cudaMemcpyAsync(d_d + 2 * off, dirs + 2 * off, 8 * cnt, cudaMemcpyHostToDevice, cs));
        if (max_tois) CKM(cudaMemcpyAsync(d_mt + off, max_tois + off, 4 * cnt, cudaMemcpyHostToDevice, cs));
        rc = launch_ray_cast2(m, pose, (uint32_t)cnt, d_o + 2 * off, d_d + 2 * off, max_toi, max_tois ? d_mt + off : nullptr, d_toi + off,
                              d_feat + off, normal ? d_n + 2 * off : nullptr, cs);
        if (rc) break;
        CKM(cudaMemcpyAsync(toi + off, d_toi + off, 4 * cnt, cudaMemcpyDeviceToHost, cs));
        CKM(cudaMemcpyAsync(feature + off, d_feat + off, 4 * cnt, cudaMemcpyDeviceToHost, cs));
        if (normal) CKM(cudaMemcpyAsync(normal + 2 * off, d_n + 2 * off, 8 * cnt, cudaMemcpyDeviceToHost, cs));
    }
    for (int j = 0; j < used; ++j) cudaStreamSynchronize(m->chunk_stream[j]);
    if (rc) return rc;
    CKM(cudaGetLastError());
    return NCB_OK;
}

}  // extern "C"
#else
}  // namespace ncb (host shim: only slab_toi / ray_triangle above)
#endif  // NCB_HOST_SHIM
