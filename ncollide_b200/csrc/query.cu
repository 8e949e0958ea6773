// World-level ray queries on the broad phase's LBVH.
//
// Replaces (reference, file:line): pipeline/glue/query.rs:13-77 (interferences_with_ray) and :183-224
// (first_interference_with_ray) over BroadPhase::interferences_with_ray / first_interference_with_ray
// (dbvt_broad_phase.rs:401-461); per-shape RayCast::toi_and_normal_with_ray(position, ray, max_toi, solid = true):
// query/ray/ray_ball.rs:77-142, ray_cuboid.rs:7-27 + ray_aabb.rs:52-75,183-300, ray_plane.rs:44-79,
// ray_support_map.rs:15-35,139-160 + query/algorithms/gjk.rs:180-365 (ConvexHull through gjk::cast_ray).
//
// One thread per ray walks the LBVH with the reference's slab test (monotone under box inclusion, so the candidate set
// equals the one the two DBVTs produce), filters candidates by the query's collision groups, and runs the shape's ray cast.
// "all" mode appends (ray, handle, toi, normal, feature) rows; "first" mode keeps the smallest toi (ties: smallest handle).
#ifndef NCB_HOST_SHIM  // tests/host_shim compiles the per-shape functions of this file for the host (test infrastructure)
#include <cub/cub.cuh>
#endif
#include "bp_internal.h"
#include "shapes.cuh"

using namespace ncb;

namespace {

struct RayHit {
    bool hit;
    float toi;
    V3 normal;
    uint32_t feature;
};

__device__ __forceinline__ RayHit ray_cast_ball(V3 center, float radius, V3 o, V3 d, float max_toi) {
    RayHit h;
    h.hit = false;
    V3 dcenter = o - center;
    float a = norm_squared(d), b = dot(dcenter, d), c = norm_squared(dcenter) - radius * radius;
    bool inside = false;
    float t = 0.f;
    if (a == 0.f) {
        if (c > 0.f) return h;
        inside = true;
    } else if (c > 0.f && b > 0.f) {
        return h;
    } else {
        float delta = b * b - a * c;
        if (delta < 0.f) return h;
        t = (-b - sqrtf(delta)) / a;
        if (t <= 0.f) inside = true, t = 0.f;  // solid
    }
    if (!(t <= max_toi)) return h;
    V3 pos = o + d * t - center;
    V3 normal = normalize(pos);
    h.hit = true, h.toi = t, h.normal = inside ? -normal : normal, h.feature = FACE0;
    return h;
}

__device__ __forceinline__ RayHit ray_cast_cuboid(V3 he, const Iso& m, V3 o_w, V3 d_w, float max_toi) {
    RayHit h;
    h.hit = false;
    V3 o = iso_inv_point(m, o_w), d = iso_inv_vec(m, d_w);
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, mn[3] = {-he.x, -he.y, -he.z}, mx[3] = {he.x, he.y, he.z};
    float tmax = NCB_FMAX, tmin = -NCB_FMAX;
    int near_side = 0, far_side = 0;
    bool near_diag = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (dd[i] == 0.f) {
            if (oo[i] < mn[i] || oo[i] > mx[i]) return h;
        } else {
            float denom = 1.0f / dd[i];
            float tn = (mn[i] - oo[i]) * denom, tf = (mx[i] - oo[i]) * denom;
            bool flip = false;
            if (tn > tf) {
                flip = true;
                float t = tn;
                tn = tf, tf = t;
            }
            if (tn > tmin) {
                tmin = tn;
                near_side = flip ? -(i + 1) : (i + 1);
                near_diag = false;
            } else if (tn == tmin) {
                near_diag = true;
            }
            if (tf < tmax) {
                tmax = tf;
                far_side = !flip ? -(i + 1) : (i + 1);
            }
            if (tmax < 0.f || tmin > tmax) return h;
        }
    }
    float t;
    V3 n = v3(0.f, 0.f, 0.f);
    int side;
    if (tmin < 0.f) {
        t = 0.f, side = far_side;
    } else if (tmin <= max_toi) {
        t = tmin, side = near_side;
        if (near_diag)
            n = -normalize(d);
        else if (near_side != 0) {
            float v = near_side < 0 ? 1.f : -1.f;
            int ax = (near_side < 0 ? -near_side : near_side) - 1;
            n = v3(ax == 0 ? v : 0.f, ax == 1 ? v : 0.f, ax == 2 ? v : 0.f);
        }
    } else {
        return h;
    }
    h.hit = true, h.toi = t, h.normal = iso_mul_vec(m, n);
    h.feature = FID(NCB_FEATURE_FACE, side < 0 ? (-side - 1 + 3) : (side - 1));
    return h;
}

__device__ __forceinline__ RayHit ray_cast_plane(V3 pn, const Iso& m, V3 o_w, V3 d_w, float max_toi) {
    RayHit h;
    h.hit = false;
    V3 o = iso_inv_point(m, o_w), d = iso_inv_vec(m, d_w);
    float dot_normal_dpos = dot(pn, -o);
    if (dot_normal_dpos > 0.f) {
        h.hit = true, h.toi = 0.f, h.normal = v3(0.f, 0.f, 0.f), h.feature = FACE0;
        return h;
    }
    float t = dot_normal_dpos / dot(pn, d);
    if (t >= 0.f && t <= max_toi) h.hit = true, h.toi = t, h.normal = iso_mul_vec(m, pn), h.feature = FACE0;
    return h;
}

__device__ __forceinline__ bool ray_toi_with_plane(V3 center, V3 normal, V3 origin, V3 dir, float& t_out) {
    V3 dpos = center - origin;
    float denom = dot(normal, dir);
    if (relative_eq(denom, 0.f)) return false;
    float t = dot(normal, dpos) / denom;
    if (t >= 0.f) {
        t_out = t;
        return true;
    }
    return false;
}

// minkowski_ray_cast (gjk.rs:228-365) for (hull, identity) - (ConstantOrigin, identity), ray in the hull's local frame
__device__ __noinline__ RayHit ray_cast_hull(const HullView& H, const Iso& m, V3 o_w, V3 d_w, float max_toi) {
    RayHit h;
    h.hit = false;
    h.feature = FID_UNKNOWN;
    V3 ray_origin = iso_inv_point(m, o_w), ray_dir = iso_inv_vec(m, d_w);
    Support g;
    g.kind = 1;
    g.he = v3(0.f, 0.f, 0.f);
    g.hull = H;
    const float eps_tol = NCB_EPS * 10.0f;
    const float eps_rel = sqrtf(eps_tol);
    float ray_length = norm(ray_dir);
    if (relative_eq(ray_length, 0.f)) return h;
    float ltoi = 0.f;
    V3 curr_origin = ray_origin, curr_dir = ray_dir / ray_length;
    V3 ldir = -curr_dir;
    Simplex s;
    {
        CSOPoint sp;
        sp.orig1 = local_support_point(g, ldir);  // identity isometry: support_point == local_support_point
        sp.orig2 = v3(0.f, 0.f, 0.f);
        sp.point = sp.orig1 - sp.orig2;
        sp.point = sp.point + (-curr_origin);
        simplex_init(s, sp);
    }
    V3 proj = simplex_project_origin_and_reduce(s);
    float max_bound = NCB_FMAX;
    V3 dir;
    int niter = 0;
    bool last_chance = false;
    for (;;) {
        float old_max_bound = max_bound;
        float dist;
        if (unit_try_new_and_get(-proj, eps_tol, dir, dist))
            max_bound = dist;
        else {
            h.hit = true, h.toi = ltoi / ray_length, h.normal = iso_mul_vec(m, ldir);
            return h;
        }
        CSOPoint support_point;
        if (max_bound >= old_max_bound) {
            last_chance = true;
            V3 p = proj + curr_origin;
            support_point.point = p, support_point.orig1 = p, support_point.orig2 = v3(0.f, 0.f, 0.f);
        } else {
            support_point.orig1 = local_support_point(g, dir);
            support_point.orig2 = v3(0.f, 0.f, 0.f);
            support_point.point = support_point.orig1 - support_point.orig2;
        }
        if (last_chance && ltoi > 0.f) {
            h.hit = true, h.toi = ltoi / ray_length, h.normal = iso_mul_vec(m, ldir);
            return h;
        }
        float t;
        if (ray_toi_with_plane(support_point.point, dir, curr_origin, curr_dir, t)) {
            if (dot(dir, curr_dir) < 0.f && t > 0.f) {
                ldir = dir;
                ltoi += t;
                if (ltoi / ray_length > max_toi) return h;
                V3 shift = curr_dir * t;
                curr_origin = curr_origin + shift;
                max_bound = NCB_FMAX;
                for (int i = 0; i < s.dim + 1; ++i) s.v[i].point = s.v[i].point + (-shift);
                last_chance = false;
            }
        } else if (dot(dir, curr_dir) > eps_tol) {
            return h;
        }
        if (last_chance) return h;
        float min_bound = -dot(dir, support_point.point - curr_origin);
        if (max_bound - min_bound <= eps_rel * max_bound) return h;
        CSOPoint tp = support_point;
        tp.point = tp.point + (-curr_origin);
        (void)simplex_add_point(s, tp);
        proj = simplex_project_origin_and_reduce(s);
        if (s.dim == 3) {
            if (min_bound >= eps_tol) return h;
            h.hit = true, h.toi = ltoi / ray_length, h.normal = iso_mul_vec(m, ldir);
            return h;
        }
        niter += 1;
        if (niter == 10000) return h;
    }
}

#ifndef NCB_HOST_SHIM  // traversal kernels and host entry points: CUDA only
__device__ __forceinline__ bool slab_hit(const float* q, const float* inv, float4 lo, float4 hi, float& tmin) {  // ray_aabb.rs:13-50
    tmin = 0.f;
    float tmax = q[6];
    const float mn[3] = {lo.x, lo.y, lo.z}, mx[3] = {hi.x, hi.y, hi.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (q[3 + i] == 0.f) {
            if (q[i] < mn[i] || q[i] > mx[i]) return false;
        } else {
            float tn = (mn[i] - q[i]) * inv[i], tf = (mx[i] - q[i]) * inv[i];
            if (tn > tf) {
                float t = tn;
                tn = tf, tf = t;
            }
            tmin = fmaxf(tmin, tn);
            tmax = fminf(tmax, tf);
            if (tmin > tmax) return false;
        }
    }
    return true;
}

struct QueryArgs {
    const float* rays;  // 7 per ray
    uint32_t n_rays;
    const float4 *llo, *lhi, *nodes;
    uint32_t n, nout;
    const uint32_t* d_attached;
    DevObjects o;
    DevHulls H;
    uint32_t qg[3];
    int use_groups;
    // outputs
    unsigned long long* keys;  // ray << 32 | handle
    float4* vals;              // toi, normal
    uint32_t* feats;
    uint32_t cap;
    uint32_t* counter;
    uint32_t* trav_overflow;
};

template <bool FIRST>
__device__ __forceinline__ void visit_leaf(const QueryArgs& A, uint32_t ri, const float* q, uint32_t handle, RayHit& best, uint32_t& best_h) {
    if (A.d_attached[handle] != ST_ATTACHED) return;
    if (A.use_groups && A.o.groups) {  // CollisionGroups::can_interact_with_groups (collision_groups.rs:353-359)
        uint32_t m1 = __ldg(&A.o.groups[3 * handle]), w1 = __ldg(&A.o.groups[3 * handle + 1]), b1 = __ldg(&A.o.groups[3 * handle + 2]);
        if (!((m1 & A.qg[2]) == 0 && (A.qg[0] & b1) == 0 && (m1 & A.qg[1]) != 0 && (A.qg[0] & w1) != 0)) return;
    }
    uint32_t type = __ldg(&A.o.type[handle]);
    Shape sh = load_shape(A.o, A.H, handle, type);
    Iso m = load_iso(A.o, handle);
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
    if (!h.hit) return;
    if (FIRST) {
        if (!best.hit || h.toi < best.toi || (h.toi == best.toi && handle < best_h)) best = h, best_h = handle;
    } else {
        uint32_t k = atomicAdd(A.counter, 1u);
        if (k < A.cap) {
            A.keys[k] = ((unsigned long long)ri << 32) | handle;
            A.vals[k] = make_float4(h.toi, h.normal.x, h.normal.y, h.normal.z);
            A.feats[k] = h.feature;
        }
    }
}

// FIRST: once a hit is known, boxes entered later than it cannot hold a closer hit (a stored box contains its shape), so the
// slab test's upper bound shrinks to the best toi (plus a rounding guard) — the pruning the reference's best-first search does.
#define SHRINK()                                                                  \
    if (FIRST && best.hit) q[6] = fminf(q[6], best.toi * 1.00001f + 1e-6f)
template <bool FIRST>
__global__ void __launch_bounds__(128) k_world_ray_cast(QueryArgs A) {
    uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= A.n_rays) return;
    float q[7], inv[3];
    for (int k = 0; k < 7; ++k) q[k] = A.rays[7 * (size_t)ri + k];
    for (int k = 0; k < 3; ++k) inv[k] = 1.0f / q[3 + k];
    RayHit best;
    best.hit = false;
    uint32_t best_h = 0;
    uint32_t m = A.n - A.nout;
    float te;
    if (m >= 2) {
        uint32_t stack[64];
        int sp = 0;
        uint32_t node = 0;
        for (;;) {
            const float4* rec = A.nodes + 4 * (size_t)node;
            float4 Llo = __ldg(rec + 0), Lhi = __ldg(rec + 1), Rlo = __ldg(rec + 2), Rhi = __ldg(rec + 3);
            uint32_t left = __float_as_uint(Llo.w), right = __float_as_uint(Lhi.w);
            float tl, tr;
            bool goL = slab_hit(q, inv, Llo, Lhi, tl), goR = slab_hit(q, inv, Rlo, Rhi, tr);
            if (goL && (left & LEAF_BIT)) {
                visit_leaf<FIRST>(A, ri, q, __float_as_uint(__ldg(&A.llo[left & ~LEAF_BIT].w)), best, best_h);
                SHRINK();
                goL = false;
                if (FIRST && goR) goR = tr <= q[6];
            }
            if (goR && (right & LEAF_BIT)) {
                visit_leaf<FIRST>(A, ri, q, __float_as_uint(__ldg(&A.llo[right & ~LEAF_BIT].w)), best, best_h);
                SHRINK();
                goR = false;
            }
            if (goL && goR) {
                // FIRST: the child entered sooner is walked first, so that its hits prune the other one
                bool right_first = FIRST && tr < tl;
                if (sp < 64)
                    stack[sp++] = right_first ? left : right;
                else
                    atomicAdd(A.trav_overflow, 1u);
                node = right_first ? right : left;
            } else if (goL) {
                node = left;
            } else if (goR) {
                node = right;
            } else {
                if (sp == 0) break;
                node = stack[--sp];
            }
        }
    } else if (m == 1) {
        float4 lo = __ldg(&A.llo[0]), hi = __ldg(&A.lhi[0]);
        if (slab_hit(q, inv, lo, hi, te)) visit_leaf<FIRST>(A, ri, q, __float_as_uint(lo.w), best, best_h);
    }
    for (uint32_t o = m; o < A.n; ++o) {
        float4 lo = __ldg(&A.llo[o]), hi = __ldg(&A.lhi[o]);
        if (slab_hit(q, inv, lo, hi, te)) {
            visit_leaf<FIRST>(A, ri, q, __float_as_uint(lo.w), best, best_h);
            SHRINK();
        }
    }
    if (FIRST && best.hit) {
        uint32_t k = atomicAdd(A.counter, 1u);
        if (k < A.cap) {
            A.keys[k] = ((unsigned long long)ri << 32) | best_h;
            A.vals[k] = make_float4(best.toi, best.normal.x, best.normal.y, best.normal.z);
            A.feats[k] = best.feature;
        }
    }
}

// ---- glue::interferences_with_point (KIND 2) / interferences_with_aabb (KIND 0) (glue/query.rs:79-181) ------------------
template <int KIND>
__device__ __forceinline__ bool box_test(const float* q, float4 lo, float4 hi) {
    if (KIND == 0)  // AABB::intersects
        return lo.x <= q[3] && lo.y <= q[4] && lo.z <= q[5] && hi.x >= q[0] && hi.y >= q[1] && hi.z >= q[2];
    if (q[0] < lo.x || q[0] > hi.x) return false;  // AABB::contains_local_point
    if (q[1] < lo.y || q[1] > hi.y) return false;
    if (q[2] < lo.z || q[2] > hi.z) return false;
    return true;
}
template <int KIND>
__device__ __forceinline__ void visit_leaf_q(const QueryArgs& A, uint32_t qi, const float* q, uint32_t handle) {
    if (A.d_attached[handle] != ST_ATTACHED) return;
    if (A.use_groups && A.o.groups) {
        uint32_t m1 = __ldg(&A.o.groups[3 * handle]), w1 = __ldg(&A.o.groups[3 * handle + 1]), b1 = __ldg(&A.o.groups[3 * handle + 2]);
        if (!((m1 & A.qg[2]) == 0 && (A.qg[0] & b1) == 0 && (m1 & A.qg[1]) != 0 && (A.qg[0] & w1) != 0)) return;
    }
    if (KIND == 2) {  // PointQuery::contains_point of the shape
        uint32_t type = __ldg(&A.o.type[handle]);
        Shape sh = load_shape(A.o, A.H, handle, type);
        Iso m = load_iso(A.o, handle);
        V3 pt = v3(q[0], q[1], q[2]);
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
        if (!inside) return;
    }
    uint32_t k = atomicAdd(A.counter, 1u);
    if (k < A.cap) A.keys[k] = ((unsigned long long)qi << 32) | handle;
}
template <int KIND>
__global__ void __launch_bounds__(128) k_world_query(QueryArgs A) {
    const int W = KIND == 0 ? 6 : 3;
    uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= A.n_rays) return;
    float q[6];
    for (int k = 0; k < W; ++k) q[k] = A.rays[(size_t)W * qi + k];
    uint32_t m = A.n - A.nout;
    if (m >= 2) {
        uint32_t stack[64];
        int sp = 0;
        uint32_t node = 0;
        for (;;) {
            const float4* rec = A.nodes + 4 * (size_t)node;
            float4 Llo = __ldg(rec + 0), Lhi = __ldg(rec + 1), Rlo = __ldg(rec + 2), Rhi = __ldg(rec + 3);
            uint32_t left = __float_as_uint(Llo.w), right = __float_as_uint(Lhi.w);
            bool goL = box_test<KIND>(q, Llo, Lhi), goR = box_test<KIND>(q, Rlo, Rhi);
            if (goL && (left & LEAF_BIT)) {
                visit_leaf_q<KIND>(A, qi, q, __float_as_uint(__ldg(&A.llo[left & ~LEAF_BIT].w)));
                goL = false;
            }
            if (goR && (right & LEAF_BIT)) {
                visit_leaf_q<KIND>(A, qi, q, __float_as_uint(__ldg(&A.llo[right & ~LEAF_BIT].w)));
                goR = false;
            }
            if (goL) {
                if (goR) {
                    if (sp < 64)
                        stack[sp++] = right;
                    else
                        atomicAdd(A.trav_overflow, 1u);
                }
                node = left;
            } else if (goR) {
                node = right;
            } else {
                if (sp == 0) break;
                node = stack[--sp];
            }
        }
    } else if (m == 1) {
        float4 lo = __ldg(&A.llo[0]), hi = __ldg(&A.lhi[0]);
        if (box_test<KIND>(q, lo, hi)) visit_leaf_q<KIND>(A, qi, q, __float_as_uint(lo.w));
    }
    for (uint32_t o = m; o < A.n; ++o) {
        float4 lo = __ldg(&A.llo[o]), hi = __ldg(&A.lhi[o]);
        if (box_test<KIND>(q, lo, hi)) visit_leaf_q<KIND>(A, qi, q, __float_as_uint(lo.w));
    }
}
__global__ void k_unpack_keys(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[2 * i] = (uint32_t)(keys[i] >> 32), out[2 * i + 1] = (uint32_t)keys[i];
}

__global__ void k_iota(uint32_t* p, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
__global__ void k_gather_rows(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ order, const float4* __restrict__ vals,
                              const uint32_t* __restrict__ feats, uint32_t n, uint32_t* idx_out, float4* val_out, uint32_t* feat_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = keys[i];
    uint32_t src = order[i];
    idx_out[2 * i] = (uint32_t)(k >> 32), idx_out[2 * i + 1] = (uint32_t)k;
    val_out[i] = vals[src];
    feat_out[i] = feats[src];
}

}  // namespace

#define CKQ(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            char b__[512];                                                                                \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            ctx->err = b__;                                                                               \
            return NCB_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

// rows sorted by (ray, handle).  Host output arrays: idx[2 * cap], val[4 * cap], feat[cap].
int world_ray_cast(ncb_ctx* ctx, ncb_bp* bp, WorldQueryBufs& B, uint32_t n_rays, const float* rays, const uint32_t* groups, int first_only,
                   uint32_t* idx, float* val, uint32_t* feat, uint32_t cap, uint32_t* n_out) {
    CKQ(cudaSetDevice(ctx->device));
    if (n_out) *n_out = 0;
    if (n_rays == 0 || bp->tree_n == 0) return NCB_OK;
    if (ctx->has_capsules) {
        ctx->err = "world ray queries: the capsule ray cast is not on the device (worlds with capsules support updates only)";
        return NCB_ERR_UNSUPPORTED;
    }
    cudaStream_t s = ctx->stream;
    ncb_ctx* w = bp->work;
    CKQ(B.rays.reserve(7 * (size_t)n_rays));
    CKQ(cudaMemcpyAsync(B.rays.p, rays, 28 * (size_t)n_rays, cudaMemcpyHostToDevice, s));
    CKQ(B.counter.reserve(4));
    size_t want = B.keys.cap ? B.keys.cap : (first_only ? (size_t)n_rays + 64 : (size_t)8 * n_rays + 1024);
    if (first_only && want < n_rays) want = n_rays;
    uint32_t found = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        CKQ(B.keys.reserve(want));
        CKQ(B.vals.reserve(want));
        CKQ(B.feats.reserve(want));
        uint32_t c = (uint32_t)std::min(B.keys.cap, std::min(B.vals.cap, B.feats.cap));
        CKQ(cudaMemsetAsync(B.counter.p, 0, 16, s));
        QueryArgs A;
        A.rays = B.rays.p, A.n_rays = n_rays;
        A.llo = w->leaf_lo.p, A.lhi = w->leaf_hi.p, A.nodes = w->nodes.p;
        A.n = bp->tree_n, A.nout = bp->tree_outliers;
        A.d_attached = bp->d_attached.p;
        A.o = dev_objects(ctx), A.H = ctx->hulls;
        A.use_groups = groups != nullptr;
        for (int k = 0; k < 3; ++k) A.qg[k] = groups ? groups[k] : 0;
        A.keys = B.keys.p, A.vals = B.vals.p, A.feats = B.feats.p, A.cap = c, A.counter = B.counter.p, A.trav_overflow = trav_overflow_counter(ctx);
        unsigned g = (n_rays + 127) / 128;
        if (first_only)
            k_world_ray_cast<true><<<g, 128, 0, s>>>(A);
        else
            k_world_ray_cast<false><<<g, 128, 0, s>>>(A);
        CKQ(cudaGetLastError());
        CKQ(cudaMemcpyAsync(&found, B.counter.p, 4, cudaMemcpyDeviceToHost, s));
        CKQ(cudaStreamSynchronize(s));
        if (found <= c) break;
        want = (size_t)found + 1024;
    }
    if (n_out) *n_out = found;
    if (found == 0) return NCB_OK;
    // deterministic row order: sort by (ray, handle), carry the payload through an index
    CKQ(B.order_in.reserve(found));
    CKQ(B.order_out.reserve(found));
    CKQ(B.keys_sorted.reserve(found));
    k_iota<<<(found + 255) / 256, 256, 0, s>>>(B.order_in.p, found);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, B.keys.p, B.keys_sorted.p, B.order_in.p, B.order_out.p, (int)found, 0, 64);
    CKQ(B.cub_tmp.reserve(bytes + 256));
    bytes = B.cub_tmp.cap;
    CKQ(cub::DeviceRadixSort::SortPairs(B.cub_tmp.p, bytes, B.keys.p, B.keys_sorted.p, B.order_in.p, B.order_out.p, (int)found, 0, 64, s));
    CKQ(B.out_idx.reserve(2 * (size_t)found));
    CKQ(B.out_val.reserve(found));
    CKQ(B.out_feat.reserve(found));
    k_gather_rows<<<(found + 255) / 256, 256, 0, s>>>(B.keys_sorted.p, B.order_out.p, B.vals.p, B.feats.p, found, B.out_idx.p, B.out_val.p, B.out_feat.p);
    CKQ(cudaGetLastError());
    uint32_t wr = found < cap ? found : cap;
    if (idx && wr) CKQ(cudaMemcpyAsync(idx, B.out_idx.p, 8 * (size_t)wr, cudaMemcpyDeviceToHost, s));
    if (val && wr) CKQ(cudaMemcpyAsync(val, B.out_val.p, 16 * (size_t)wr, cudaMemcpyDeviceToHost, s));
    if (feat && wr) CKQ(cudaMemcpyAsync(feat, B.out_feat.p, 4 * (size_t)wr, cudaMemcpyDeviceToHost, s));
    CKQ(cudaStreamSynchronize(s));
    return found > cap ? 1 : NCB_OK;
}

// kind 0: interferences_with_aabb (6 floats per query), kind 2: interferences_with_point (3 floats).  Rows (query, handle) sorted.
int world_query(ncb_ctx* ctx, ncb_bp* bp, WorldQueryBufs& B, int kind, uint32_t n_q, const float* q, const uint32_t* groups, uint32_t* idx,
                uint32_t cap, uint32_t* n_out) {
    CKQ(cudaSetDevice(ctx->device));
    if (n_out) *n_out = 0;
    if (n_q == 0 || bp->tree_n == 0) return NCB_OK;
    if (kind == 2 && ctx->has_capsules) {
        ctx->err = "world point queries: capsule point containment is not on the device (worlds with capsules support updates only)";
        return NCB_ERR_UNSUPPORTED;
    }
    cudaStream_t s = ctx->stream;
    ncb_ctx* w = bp->work;
    const int W = kind == 0 ? 6 : 3;
    CKQ(B.rays.reserve((size_t)W * n_q));
    CKQ(cudaMemcpyAsync(B.rays.p, q, 4 * (size_t)W * n_q, cudaMemcpyHostToDevice, s));
    CKQ(B.counter.reserve(4));
    size_t want = B.keys.cap ? B.keys.cap : (size_t)8 * n_q + 1024;
    uint32_t found = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        CKQ(B.keys.reserve(want));
        CKQ(cudaMemsetAsync(B.counter.p, 0, 16, s));
        QueryArgs A;
        A.rays = B.rays.p, A.n_rays = n_q;
        A.llo = w->leaf_lo.p, A.lhi = w->leaf_hi.p, A.nodes = w->nodes.p;
        A.n = bp->tree_n, A.nout = bp->tree_outliers;
        A.d_attached = bp->d_attached.p;
        A.o = dev_objects(ctx), A.H = ctx->hulls;
        A.use_groups = groups != nullptr;
        for (int k = 0; k < 3; ++k) A.qg[k] = groups ? groups[k] : 0;
        A.keys = B.keys.p, A.vals = nullptr, A.feats = nullptr, A.cap = (uint32_t)B.keys.cap, A.counter = B.counter.p, A.trav_overflow = trav_overflow_counter(ctx);
        unsigned g = (n_q + 127) / 128;
        if (kind == 0)
            k_world_query<0><<<g, 128, 0, s>>>(A);
        else
            k_world_query<2><<<g, 128, 0, s>>>(A);
        CKQ(cudaGetLastError());
        CKQ(cudaMemcpyAsync(&found, B.counter.p, 4, cudaMemcpyDeviceToHost, s));
        CKQ(cudaStreamSynchronize(s));
        if (found <= B.keys.cap) break;
        want = (size_t)found + 1024;
    }
    if (n_out) *n_out = found;
    if (found == 0) return NCB_OK;
    CKQ(B.keys_sorted.reserve(found));
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, B.keys.p, B.keys_sorted.p, (int)found, 0, 64);
    CKQ(B.cub_tmp.reserve(bytes + 256));
    bytes = B.cub_tmp.cap;
    CKQ(cub::DeviceRadixSort::SortKeys(B.cub_tmp.p, bytes, B.keys.p, B.keys_sorted.p, (int)found, 0, 64, s));
    CKQ(B.out_idx.reserve(2 * (size_t)found));
    k_unpack_keys<<<(found + 255) / 256, 256, 0, s>>>(B.keys_sorted.p, found, B.out_idx.p);
    CKQ(cudaGetLastError());
    uint32_t wr = found < cap ? found : cap;
    if (idx && wr) CKQ(cudaMemcpyAsync(idx, B.out_idx.p, 8 * (size_t)wr, cudaMemcpyDeviceToHost, s));
    CKQ(cudaStreamSynchronize(s));
    return found > cap ? 1 : NCB_OK;
}
#else
}  // namespace (host shim: only the per-shape ray casts above)
#endif  // NCB_HOST_SHIM
