// Broad phase on the device: per-object AABBs, LBVH (Morton codes -> radix sort -> Karras build -> bottom-up refit),
// overlap pair search with warp-aggregated emission, and a counting sort of the pairs by shape-type key.
//
// Replaces (reference, file:line):
//   k_aabb          pipeline/object/collision_object.rs:89-93, bounding_volume/aabb_{ball,cuboid,convex,plane}.rs,
//                   bounding_volume/aabb_utils.rs:59-79, bounding_volume/aabb.rs:180-199, dbvt_broad_phase.rs:341
//   LBVH            partitioning/dbvt.rs:158-255 (the incremental tree is replaced, not ported)
//   k_pair_search   pipeline/broad_phase/dbvt_broad_phase.rs:218-253, partitioning/bvh.rs:24-45,
//                   query/visitors/bounding_volume_interferences_collector.rs:41-51,
//                   pipeline/object/collision_groups.rs:353-359
//   pair keys       pipeline/narrow_phase/contact_generator/default_contact_dispatcher.rs:27-97
// The fresh-world pair set is tree independent: {(i, j) : j < i, fat_i ∩ fat_j != ∅, allowed(i, j)}.
#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include "ncb_internal.h"
#include "vec.cuh"

namespace ncb {

#define LEAF_BIT 0x80000000u
static const float OUTLIER_ABS = 1.0e30f;  // planes have +-f32::MAX/2 boxes (aabb_plane.rs:16-20): kept out of the tree

__device__ __forceinline__ int f2o(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ------------------------------------------------------------------------------------------------------------
// K1: AABBs.  One thread per object; 128-bit loads of the quaternion / shape record, float4 SoA stores.
// mode 0: bounding_volume::aabb(shape, pos); 1: + loosen(query_limit) (compute_aabb); 2: + loosened(margin).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_aabb(DevObjects o, DevHulls H, float margin, int mode, uint32_t begin, uint32_t end,
                                              float4* __restrict__ lo, float4* __restrict__ hi) {
    uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    uint32_t type = o.type[i];
    float4 q4 = __ldg(&o.rot[i]);
    float4 p4 = __ldg(&o.param[i]);
    V3 t = v3(o.pos[3 * i], o.pos[3 * i + 1], o.pos[3 * i + 2]);
    Quat q = Quat{q4.x, q4.y, q4.z, q4.w};
    V3 mins, maxs;
    if (type == NCB_SHAPE_BALL) {
        float r = p4.x;
        mins = t + v3(-r, -r, -r);
        maxs = t + v3(r, r, r);
    } else if (type == NCB_SHAPE_CUBOID) {
        V3 he = absolute_transform_vector(q, v3(p4.x, p4.y, p4.z));
        mins = t - he;
        maxs = t + he;
    } else if (type == NCB_SHAPE_CONVEX_HULL) {
        uint32_t h = (uint32_t)p4.x;
        uint32_t v0 = H.vert_off[h], v1 = H.vert_off[h + 1];
        Iso m = Iso{t, q};
        const float* P = H.points + 3 * (size_t)v0;
        V3 w0 = iso_mul_point(m, v3(__ldg(P), __ldg(P + 1), __ldg(P + 2)));
        mins = w0;
        maxs = w0;
        for (uint32_t k = 1; k < v1 - v0; ++k) {
            V3 w = iso_mul_point(m, v3(__ldg(P + 3 * k), __ldg(P + 3 * k + 1), __ldg(P + 3 * k + 2)));
            mins = vmin(mins, w);
            maxs = vmax(maxs, w);
        }
    } else if (type == NCB_SHAPE_CAPSULE) {
        capsule_aabb(Iso{t, q}, p4.x, p4.y, mins, maxs);
    } else {
        float mx = NCB_FMAX * 0.5f;
        mins = v3(-mx, -mx, -mx);
        maxs = v3(mx, mx, mx);
    }
    if (mode >= 1) {
        float ql = o.qlimit[i];
        mins = mins + v3(-ql, -ql, -ql);
        maxs = maxs + v3(ql, ql, ql);
    }
    if (mode >= 2) {
        mins = mins + v3(-margin, -margin, -margin);
        maxs = maxs + v3(margin, margin, margin);
    }
    lo[i] = make_float4(mins.x, mins.y, mins.z, 0.0f);
    hi[i] = make_float4(maxs.x, maxs.y, maxs.z, __uint_as_float(type));
}

cudaError_t launch_aabbs(ncb_ctx* c, const DevObjects& o, float margin, int mode, uint32_t begin, uint32_t end) {
    if (end <= begin) return cudaSuccess;
    uint32_t n = end - begin;
    k_aabb<<<(n + 255) / 256, 256, 0, c->stream>>>(o, c->hulls, margin, mode, begin, end, c->aabb_lo.p, c->aabb_hi.p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// K2: scene bounds of the AABB centres (outliers excluded) + Morton codes.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_outlier(float4 lo, float4 hi) {
    float m = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(lo.y)), fabsf(lo.z)), fmaxf(fmaxf(fabsf(hi.x), fabsf(hi.y)), fabsf(hi.z)));
    return !(m < OUTLIER_ABS);  // also true for NaN
}

__global__ void __launch_bounds__(256) k_bounds(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n,
                                                DevCounters* cnt) {
    float mn[3] = {NCB_FMAX, NCB_FMAX, NCB_FMAX}, mx[3] = {-NCB_FMAX, -NCB_FMAX, -NCB_FMAX};
    uint32_t nout = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 a = __ldg(&lo[i]), b = __ldg(&hi[i]);
        if (is_outlier(a, b)) {
            nout++;
            continue;
        }
        float cx = (a.x + b.x) * 0.5f, cy = (a.y + b.y) * 0.5f, cz = (a.z + b.z) * 0.5f;
        mn[0] = fminf(mn[0], cx), mn[1] = fminf(mn[1], cy), mn[2] = fminf(mn[2], cz);
        mx[0] = fmaxf(mx[0], cx), mx[1] = fmaxf(mx[1], cy), mx[2] = fmaxf(mx[2], cz);
    }
    for (int off = 16; off; off >>= 1) {
        for (int k = 0; k < 3; ++k) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], off));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], off));
        }
        nout += __shfl_xor_sync(0xffffffffu, nout, off);
    }
    // block-level combine in shared memory, then 7 global atomics per CTA (not per warp)
    __shared__ float s_mn[8][3], s_mx[8][3];
    __shared__ uint32_t s_out[8];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        for (int k = 0; k < 3; ++k) s_mn[warp][k] = mn[k], s_mx[warp][k] = mx[k];
        s_out[warp] = nout;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int nw = blockDim.x >> 5;
        for (int w = 1; w < nw; ++w) {
            for (int k = 0; k < 3; ++k) mn[k] = fminf(mn[k], s_mn[w][k]), mx[k] = fmaxf(mx[k], s_mx[w][k]);
            nout += s_out[w];
        }
        for (int k = 0; k < 3; ++k) {
            atomicMin(&cnt->bounds[k], f2o(mn[k]));
            atomicMax(&cnt->bounds[3 + k], f2o(mx[k]));
        }
        if (nout) atomicAdd(&cnt->n_outliers, nout);
    }
}

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void __launch_bounds__(256) k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n,
                                                const DevCounters* __restrict__ cnt, uint32_t* __restrict__ keys,
                                                uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = __ldg(&lo[i]), b = __ldg(&hi[i]);
    uint32_t key;
    if (is_outlier(a, b)) {
        key = 0x40000000u;  // after every 30-bit code
    } else {
        float bx = o2f(cnt->bounds[0]), by = o2f(cnt->bounds[1]), bz = o2f(cnt->bounds[2]);
        float ex = o2f(cnt->bounds[3]) - bx, ey = o2f(cnt->bounds[4]) - by, ez = o2f(cnt->bounds[5]) - bz;
        float e = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-20f));
        float s = 1023.0f / e;
        float cx = ((a.x + b.x) * 0.5f - bx) * s, cy = ((a.y + b.y) * 0.5f - by) * s, cz = ((a.z + b.z) * 0.5f - bz) * s;
        uint32_t ux = (uint32_t)fminf(fmaxf(cx, 0.0f), 1023.0f);
        uint32_t uy = (uint32_t)fminf(fmaxf(cy, 0.0f), 1023.0f);
        uint32_t uz = (uint32_t)fminf(fmaxf(cz, 0.0f), 1023.0f);
        key = (expand_bits10(ux) << 2) | (expand_bits10(uy) << 1) | expand_bits10(uz);
    }
    keys[i] = key;
    idx[i] = i;
}

// Gather the boxes into Morton order; lo.w <- handle, hi.w keeps the shape type.
__global__ void __launch_bounds__(256) k_gather_leaves(const float4* __restrict__ lo, const float4* __restrict__ hi,
                                                       const uint32_t* __restrict__ idx, const uint32_t* __restrict__ handle_map,
                                                       uint32_t n, float4* __restrict__ llo, float4* __restrict__ lhi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = __ldg(&idx[i]);
    float4 a = __ldg(&lo[h]), b = __ldg(&hi[h]);
    a.w = __uint_as_float(handle_map ? __ldg(&handle_map[h]) : h);  // the id pairs are reported with
    llo[i] = a;
    lhi[i] = b;
}

// ------------------------------------------------------------------------------------------------------------
// K4: Karras (2012) radix tree over the sorted codes (ties broken by position) + bottom-up refit.
// Node layout (4 x float4 per internal node, one 64 B record):
//   [0] left  box mins, w = left child  (LEAF_BIT | leaf position, or internal index)
//   [1] left  box maxs, w = right child
//   [2] right box mins, w = split  (last sorted position covered by the left child)
//   [3] right box maxs, w = last   (last sorted position covered by the node)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int delta_fn(const uint32_t* __restrict__ keys, int m, int i, int j) {
    if (j < 0 || j >= m) return -1;
    uint32_t a = __ldg(&keys[i]), b = __ldg(&keys[j]);
    if (a == b) return 32 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clz(a ^ b);
}

__global__ void __launch_bounds__(256) k_karras(const uint32_t* __restrict__ keys, uint32_t n, const DevCounters* __restrict__ cnt,
                                                float4* __restrict__ nodes, uint32_t* __restrict__ parent) {
    int m = (int)(n - cnt->n_outliers);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m - 1) return;
    int d = (delta_fn(keys, m, i, i + 1) - delta_fn(keys, m, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta_fn(keys, m, i, i - d);
    int lmax = 2;
    while (delta_fn(keys, m, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta_fn(keys, m, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta_fn(keys, m, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta_fn(keys, m, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    uint32_t left = (first == gamma) ? (LEAF_BIT | (uint32_t)gamma) : (uint32_t)gamma;
    uint32_t right = (last == gamma + 1) ? (LEAF_BIT | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    // only the .w lanes are written here; the boxes come from k_refit
    reinterpret_cast<uint32_t*>(&nodes[4 * (size_t)i + 0])[3] = left;
    reinterpret_cast<uint32_t*>(&nodes[4 * (size_t)i + 1])[3] = right;
    reinterpret_cast<uint32_t*>(&nodes[4 * (size_t)i + 2])[3] = (uint32_t)gamma;
    reinterpret_cast<uint32_t*>(&nodes[4 * (size_t)i + 3])[3] = (uint32_t)last;
    // parent links: bit 31 = "I am the right child"
    if (left & LEAF_BIT)
        parent[n + gamma] = (uint32_t)i;
    else
        parent[gamma] = (uint32_t)i;
    if (right & LEAF_BIT)
        parent[n + gamma + 1] = (uint32_t)i | LEAF_BIT;
    else
        parent[gamma + 1] = (uint32_t)i | LEAF_BIT;
    if (i == 0) parent[0] = 0xffffffffu;
}

__global__ void __launch_bounds__(256) k_refit(const float4* __restrict__ llo, const float4* __restrict__ lhi, uint32_t n,
                                               const DevCounters* __restrict__ cnt, float4* nodes, const uint32_t* __restrict__ parent,
                                               uint32_t* flags) {
    uint32_t m = n - cnt->n_outliers;
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m || m < 2) return;
    float4 a = __ldg(&llo[j]), b = __ldg(&lhi[j]);
    float3 mn = make_float3(a.x, a.y, a.z), mx = make_float3(b.x, b.y, b.z);
    uint32_t p = parent[n + j];
    for (;;) {
        uint32_t node = p & ~LEAF_BIT;
        bool right = (p & LEAF_BIT) != 0;
        float* rec = reinterpret_cast<float*>(&nodes[4 * (size_t)node + (right ? 2 : 0)]);
        // xyz only: the .w lanes hold topology
        __stcg(rec + 0, mn.x), __stcg(rec + 1, mn.y), __stcg(rec + 2, mn.z);
        __stcg(rec + 4, mx.x), __stcg(rec + 5, mx.y), __stcg(rec + 6, mx.z);
        __threadfence();
        if (atomicAdd(&flags[node], 1u) == 0) return;  // the sibling subtree is not finished: its thread continues
        const float* other = reinterpret_cast<const float*>(&nodes[4 * (size_t)node + (right ? 0 : 2)]);
        mn.x = fminf(mn.x, __ldcg(other + 0)), mn.y = fminf(mn.y, __ldcg(other + 1)), mn.z = fminf(mn.z, __ldcg(other + 2));
        mx.x = fmaxf(mx.x, __ldcg(other + 4)), mx.y = fmaxf(mx.y, __ldcg(other + 5)), mx.z = fmaxf(mx.z, __ldcg(other + 6));
        if (node == 0) return;
        p = parent[node];
    }
}

size_t lbvh_temp_bytes(uint32_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n, 0, 31);
    return bytes;
}

// Builds the LBVH over c->aabb_lo/hi[0..n).  Needs c->counters zeroed (bounds initialised) by the caller.
cudaError_t launch_lbvh_build(ncb_ctx* c, uint32_t n, const uint32_t* handle_map) {
    cudaStream_t s = c->stream;
    if (n == 0) return cudaSuccess;
    int gs = c->sm_count * 4;
    uint32_t nb = (n + 255) / 256;
    k_bounds<<<min((uint32_t)gs, nb), 256, 0, s>>>(c->aabb_lo.p, c->aabb_hi.p, n, c->counters.p);
    k_morton<<<nb, 256, 0, s>>>(c->aabb_lo.p, c->aabb_hi.p, n, c->counters.p, c->keys_a.p, c->idx_a.p);
    size_t bytes = c->cub_tmp.cap;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, bytes, c->keys_a.p, c->keys_b.p, c->idx_a.p, c->idx_b.p, (int)n, 0, 31, s);
    if (e != cudaSuccess) return e;
    timer_mark(c, "morton_sort", 6);
    k_gather_leaves<<<nb, 256, 0, s>>>(c->aabb_lo.p, c->aabb_hi.p, c->idx_b.p, handle_map, n, c->leaf_lo.p, c->leaf_hi.p);
    e = cudaMemsetAsync(c->flags.p, 0, (size_t)n * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    k_karras<<<nb, 256, 0, s>>>(c->keys_b.p, n, c->counters.p, c->nodes.p, c->parent.p);
    k_refit<<<nb, 256, 0, s>>>(c->leaf_lo.p, c->leaf_hi.p, n, c->counters.p, c->nodes.p, c->parent.p, c->flags.p);
    timer_mark(c, "lbvh_build", 4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Spatial sharding over several GPUs (after the all-gather of the fat AABBs every rank holds all N boxes):
// a rank OWNS the objects whose Morton code falls in its range of SHARD_BINS top-bit bins (ranges hold equal object counts)
// and takes as GHOSTS the other objects whose box meets the union box of its owned objects; its LBVH and pair search
// then run on owned + ghost objects only (about N / ranks + a surface layer) instead of all N.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t morton_of(float4 a, float4 b, const DevCounters* __restrict__ cnt) {
    if (is_outlier(a, b)) return 0x40000000u;
    float bx = o2f(cnt->bounds[0]), by = o2f(cnt->bounds[1]), bz = o2f(cnt->bounds[2]);
    float ex = o2f(cnt->bounds[3]) - bx, ey = o2f(cnt->bounds[4]) - by, ez = o2f(cnt->bounds[5]) - bz;
    float e = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-20f));
    float s = 1023.0f / e;
    float cx = ((a.x + b.x) * 0.5f - bx) * s, cy = ((a.y + b.y) * 0.5f - by) * s, cz = ((a.z + b.z) * 0.5f - bz) * s;
    uint32_t ux = (uint32_t)fminf(fmaxf(cx, 0.0f), 1023.0f);
    uint32_t uy = (uint32_t)fminf(fmaxf(cy, 0.0f), 1023.0f);
    uint32_t uz = (uint32_t)fminf(fmaxf(cz, 0.0f), 1023.0f);
    return (expand_bits10(ux) << 2) | (expand_bits10(uy) << 1) | expand_bits10(uz);
}
__global__ void __launch_bounds__(256) k_shard_hist(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n,
                                                    const DevCounters* __restrict__ cnt, uint32_t* __restrict__ bins, ShardScratch* sh) {
    __shared__ uint32_t h[SHARD_BINS];
    for (int k = threadIdx.x; k < SHARD_BINS; k += blockDim.x) h[k] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t key = morton_of(__ldg(&lo[i]), __ldg(&hi[i]), cnt);
        uint32_t bin = key >= 0x40000000u ? SHARD_BINS : (key >> 20);  // top 10 of the 30 code bits
        bins[i] = bin;
        if (bin < SHARD_BINS) atomicAdd(&h[bin], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < SHARD_BINS; k += blockDim.x)
        if (h[k]) atomicAdd(&sh->hist[k], h[k]);
}
// rank r owns the bins [split[r], split[r + 1]); outliers (bin SHARD_BINS) belong to the last rank
__global__ void k_shard_split(ShardScratch* sh, int world) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t total = 0;
    for (int k = 0; k < SHARD_BINS; ++k) total += sh->hist[k];
    uint32_t acc = 0;
    int r = 1;
    sh->split[0] = 0;
    for (int k = 0; k < SHARD_BINS && r < world; ++k) {
        acc += sh->hist[k];
        while (r < world && (unsigned long long)acc * world >= (unsigned long long)total * r) sh->split[r++] = k + 1;
    }
    for (; r < world; ++r) sh->split[r] = SHARD_BINS;
    sh->split[world] = SHARD_BINS + 1;
}
__device__ __forceinline__ int shard_owner(const ShardScratch* __restrict__ sh, int world, uint32_t bin) {
    int r = 0;
    while (r + 1 < world && bin >= sh->split[r + 1]) ++r;
    return r;
}
__global__ void __launch_bounds__(256) k_shard_region(const float4* __restrict__ lo, const float4* __restrict__ hi, const uint32_t* __restrict__ bins,
                                                      uint32_t n, ShardScratch* sh, int rank) {
    float mn[3] = {NCB_FMAX, NCB_FMAX, NCB_FMAX}, mx[3] = {-NCB_FMAX, -NCB_FMAX, -NCB_FMAX};
    uint32_t b0 = sh->split[rank], b1 = sh->split[rank + 1];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t bin = __ldg(&bins[i]);
        if (bin < b0 || bin >= b1 || bin >= SHARD_BINS) continue;  // infinite boxes do not shape the region
        float4 a = __ldg(&lo[i]), b = __ldg(&hi[i]);
        mn[0] = fminf(mn[0], a.x), mn[1] = fminf(mn[1], a.y), mn[2] = fminf(mn[2], a.z);
        mx[0] = fmaxf(mx[0], b.x), mx[1] = fmaxf(mx[1], b.y), mx[2] = fmaxf(mx[2], b.z);
    }
    for (int off = 16; off; off >>= 1)
        for (int k = 0; k < 3; ++k) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], off));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], off));
        }
    __shared__ float s_mn[8][3], s_mx[8][3];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 3; ++k) s_mn[warp][k] = mn[k], s_mx[warp][k] = mx[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            for (int k = 0; k < 3; ++k) mn[k] = fminf(mn[k], s_mn[w][k]), mx[k] = fmaxf(mx[k], s_mx[w][k]);
        if (mn[0] <= mx[0])
            for (int k = 0; k < 3; ++k) {
                atomicMin(&sh->region[k], f2o(mn[k]));
                atomicMax(&sh->region[3 + k], f2o(mx[k]));
            }
    }
}
__global__ void __launch_bounds__(256) k_shard_select(const float4* __restrict__ lo, const float4* __restrict__ hi, const uint32_t* __restrict__ bins,
                                                      uint32_t n, ShardScratch* sh, int rank, int world, uint32_t cap, uint32_t* __restrict__ sel,
                                                      float4* __restrict__ loc_lo, float4* __restrict__ loc_hi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    float4 a, b;
    int owner = 0;
    if (i < n) {
        a = __ldg(&lo[i]), b = __ldg(&hi[i]);
        uint32_t bin = __ldg(&bins[i]);
        owner = shard_owner(sh, world, bin);
        if (owner == rank) {
            take = true;
        } else {
            float rl[3] = {o2f(sh->region[0]), o2f(sh->region[1]), o2f(sh->region[2])};
            float rh[3] = {o2f(sh->region[3]), o2f(sh->region[4]), o2f(sh->region[5])};
            // inclusive like AABB::intersects: a ghost must be present whenever a pair with an owned object can exist
            take = a.x <= rh[0] && a.y <= rh[1] && a.z <= rh[2] && b.x >= rl[0] && b.y >= rl[1] && b.z >= rl[2];
        }
    }
    // one allocation per CTA: warp counts -> shared prefix -> a single atomicAdd
    __shared__ uint32_t s_cnt[8], s_own[8], s_base;
    unsigned mask = __ballot_sync(0xffffffffu, take);
    unsigned own = __ballot_sync(0xffffffffu, take && owner == rank);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_cnt[warp] = __popc(mask), s_own[warp] = __popc(own);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0, town = 0;
        for (int w = 0; w < 8; ++w) {
            uint32_t c = s_cnt[w];
            s_cnt[w] = tot;
            tot += c;
            town += s_own[w];
        }
        s_base = tot ? atomicAdd(&sh->m, tot) : 0;
        if (town) atomicAdd(&sh->n_owned, town);
    }
    __syncthreads();
    if (!take) return;
    uint32_t k = s_base + s_cnt[warp] + __popc(mask & ((1u << lane) - 1));
    if (k >= cap) return;
    sel[k] = i;
    loc_lo[k] = a;
    b.w = __uint_as_float((__float_as_uint(b.w) & 0xffu) | ((uint32_t)owner << 8));
    loc_hi[k] = b;
}

// Needs c->counters reset by the caller; leaves the global bounds in c->counters (reset again before the local build).
cudaError_t launch_shard_select(ncb_ctx* c, uint32_t n, int rank, int world, ShardScratch* sh, uint32_t* bins, uint32_t cap, uint32_t* sel,
                                float4* loc_lo, float4* loc_hi) {
    cudaStream_t s = c->stream;
    int gs = c->sm_count * 4;
    uint32_t nb = (n + 255) / 256;
    ShardScratch z;
    memset(&z, 0, sizeof z);
    for (int k = 0; k < 3; ++k) z.region[k] = 0x7f7fffff, z.region[3 + k] = (int)0x80800000;
    cudaError_t e = cudaMemcpyAsync(sh, &z, sizeof z, cudaMemcpyHostToDevice, s);  // pageable source: copied before the call returns
    if (e != cudaSuccess) return e;
    k_bounds<<<min((uint32_t)gs, nb), 256, 0, s>>>(c->aabb_lo.p, c->aabb_hi.p, n, c->counters.p);
    k_shard_hist<<<min((uint32_t)c->sm_count * 2, nb), 256, 0, s>>>(c->aabb_lo.p, c->aabb_hi.p, n, c->counters.p, bins, sh);
    k_shard_split<<<1, 32, 0, s>>>(sh, world);
    k_shard_region<<<min((uint32_t)gs, nb), 256, 0, s>>>(c->aabb_lo.p, c->aabb_hi.p, bins, n, sh, rank);
    k_shard_select<<<nb, 256, 0, s>>>(c->aabb_lo.p, c->aabb_hi.p, bins, n, sh, rank, world, cap, sel, loc_lo, loc_hi);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Routed sharding (several GPUs, second design): instead of all-gathering every fat AABB and letting every rank scan
// all N of them, each rank ROUTES the objects of its own block to the ranks that need them.  Per step and rank:
//   0  bounds of the own block's AABB centres                         -> all-reduce (max of [-min, max], 6 floats)
//   1  Morton bin per own object + 1024-bin histogram                 -> all-reduce (sum)
//   2  equal-count bin ranges (owner per bin); own objects appended to the send bucket of their owner;
//      union box of what goes to each owner                            -> all-to-all (owned records), all-reduce (regions)
//   3  own objects whose box meets the region of another rank          -> all-to-all (ghost records)
//   4  received records unpacked into the local box arrays (+ poses scattered into the replicated pose arrays)
// then the local LBVH build / pair search / narrow phase of the first design (ownership rule unchanged: a pair is reported
// by the rank that owns both objects, or owns one and has the lower rank of the two owners).  Every rank receives about
// N / ranks + ghosts records instead of N; nothing is scanned N times.  Bucket capacities are fixed per call (equal splits,
// so the collectives need no host-side counts); slot 0 of a bucket carries its record count.
// Record: float4 (lo.xyz, handle) (hi.xyz, type | owner << 8) [ (pos.xyz, rot.i) (rot.j, rot.k, rot.w, -) with poses ].
// ------------------------------------------------------------------------------------------------------------
__global__ void k_route_pack_bounds(const DevCounters* cnt, float* out) {
    if (threadIdx.x < 3) {
        out[threadIdx.x] = -o2f(cnt->bounds[threadIdx.x]);
        out[3 + threadIdx.x] = o2f(cnt->bounds[3 + threadIdx.x]);
    }
}
__device__ __forceinline__ uint32_t route_bin(float4 a, float4 b, const float* __restrict__ gb) {
    if (is_outlier(a, b)) return SHARD_BINS;
    float bx = -gb[0], by = -gb[1], bz = -gb[2];
    float ex = gb[3] - bx, ey = gb[4] - by, ez = gb[5] - bz;
    float e = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-20f));
    float s = 1023.0f / e;
    float cx = ((a.x + b.x) * 0.5f - bx) * s, cy = ((a.y + b.y) * 0.5f - by) * s, cz = ((a.z + b.z) * 0.5f - bz) * s;
    uint32_t ux = (uint32_t)fminf(fmaxf(cx, 0.0f), 1023.0f);
    uint32_t uy = (uint32_t)fminf(fmaxf(cy, 0.0f), 1023.0f);
    uint32_t uz = (uint32_t)fminf(fmaxf(cz, 0.0f), 1023.0f);
    return ((expand_bits10(ux) << 2) | (expand_bits10(uy) << 1) | expand_bits10(uz)) >> 20;  // top 10 of the 30 code bits
}
__global__ void __launch_bounds__(256) k_route_hist(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t begin, uint32_t end,
                                                    const float* __restrict__ gb, uint32_t* __restrict__ bins, int* __restrict__ hist) {
    __shared__ uint32_t h[SHARD_BINS];
    for (int k = threadIdx.x; k < SHARD_BINS; k += blockDim.x) h[k] = 0;
    __syncthreads();
    for (uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
        uint32_t bin = route_bin(__ldg(&lo[i]), __ldg(&hi[i]), gb);
        bins[i - begin] = bin;
        if (bin < SHARD_BINS) atomicAdd(&h[bin], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < SHARD_BINS; k += blockDim.x)
        if (h[k]) atomicAdd(&hist[k], (int)h[k]);
}
__global__ void k_route_split(const int* __restrict__ hist, int world, uint32_t* __restrict__ split) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long total = 0;
    for (int k = 0; k < SHARD_BINS; ++k) total += (uint32_t)hist[k];
    unsigned long long acc = 0;
    int r = 1;
    split[0] = 0;
    for (int k = 0; k < SHARD_BINS && r < world; ++k) {
        acc += (uint32_t)hist[k];
        while (r < world && acc * world >= total * r) split[r++] = k + 1;
    }
    for (; r < world; ++r) split[r] = SHARD_BINS;
    split[world] = SHARD_BINS + 1;  // outliers (bin SHARD_BINS) belong to the last rank
}
__device__ __forceinline__ int route_owner(const uint32_t* __restrict__ split, int world, uint32_t bin) {
    int r = 0;
    while (r + 1 < world && bin >= split[r + 1]) ++r;
    return r;
}
// stage 2: own objects -> the bucket of their owner; union box of the (finite) boxes per owner
__global__ void __launch_bounds__(256) k_route_owned(const float4* __restrict__ lo, const float4* __restrict__ hi, const float* __restrict__ pos,
                                                     const float4* __restrict__ rot, uint32_t begin, uint32_t end, const uint32_t* __restrict__ bins,
                                                     const uint32_t* __restrict__ split, int world, uint32_t cap, int recw, RouteDst dst,
                                                     uint32_t* __restrict__ counts, int* __restrict__ region) {
    // A rank's region is kept as SHARD_SUBS boxes, one per top-level Morton octant of its bin range, not as one union box: an
    // equal-count boundary that falls one bin short of an octant boundary would otherwise stretch the union box across the
    // scene and double that rank's ghosts (measured at 4 ranks: one rank's broad phase took 2x, profiles/r2_multi_gpu.txt).
    __shared__ int s_reg[SHARD_MAX_RANKS * SHARD_SUBS][6];
    for (int k = threadIdx.x; k < SHARD_MAX_RANKS * SHARD_SUBS * 6; k += blockDim.x) (&s_reg[0][0])[k] = (k % 6) < 3 ? 0x7f7fffff : (int)0x80800000;
    __syncthreads();
    uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end) {
        float4 a = __ldg(&lo[i]), b = __ldg(&hi[i]);
        uint32_t bin = __ldg(&bins[i - begin]);
        int owner = route_owner(split, world, bin);
        unsigned peers = __match_any_sync(__activemask(), owner);
        int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&counts[owner], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        uint32_t k = base + __popc(peers & ((1u << lane) - 1));
        if (k + 1 < cap) {  // slot 0 is the header
            float4* r = dst.p[owner] + (size_t)(1 + k) * recw;
            a.w = __uint_as_float(i);
            b.w = __uint_as_float((__float_as_uint(b.w) & 0xffu) | ((uint32_t)owner << 8));
            r[0] = a, r[1] = b;
            if (recw == 4) {
                float4 q = __ldg(&rot[i]);
                r[2] = make_float4(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2], q.x);
                r[3] = make_float4(q.y, q.z, q.w, 0.f);
            }
        }
        if (bin < SHARD_BINS) {  // infinite boxes do not shape a region
            int* r6 = s_reg[owner * SHARD_SUBS + (int)(bin / (SHARD_BINS / SHARD_SUBS))];
            atomicMin(&r6[0], f2o(a.x)), atomicMin(&r6[1], f2o(a.y)), atomicMin(&r6[2], f2o(a.z));
            atomicMax(&r6[3], f2o(b.x)), atomicMax(&r6[4], f2o(b.y)), atomicMax(&r6[5], f2o(b.z));
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < world * SHARD_SUBS * 6; k += blockDim.x) {
        int v = (&s_reg[0][0])[k];
        if ((k % 6) < 3) {
            if (v != 0x7f7fffff) atomicMin(&region[k], v);
        } else if (v != (int)0x80800000) {
            atomicMax(&region[k], v);
        }
    }
}
// bucket headers (slot 0 = record count, may exceed the capacity: the receiver clamps and flags) + regions as [-min, max] floats
__global__ void k_route_finish(const uint32_t* __restrict__ counts, int world, RouteDst dst, const int* __restrict__ region,
                               float* __restrict__ region_f) {
    int t = threadIdx.x;
    // header.y: the largest bucket of this sender.  Every receiver sees it from every sender, so all ranks derive the same
    // capacity requirement (the collectives use one capacity for all buckets of all ranks) without another collective.
    uint32_t mx = 0;
    for (int q = 0; q < world; ++q) mx = max(mx, counts[q]);
    if (t < world) dst.p[t][0] = make_float4(__uint_as_float(counts[t]), __uint_as_float(mx), 0.f, 0.f);
    if (region_f)
        for (int k = t; k < world * SHARD_SUBS * 6; k += blockDim.x) region_f[k] = (k % 6) < 3 ? -o2f(region[k]) : o2f(region[k]);
}
// stage 3: ghosts = own objects whose box meets the region of a rank that does not own them (inclusive test, like AABB::intersects)
__global__ void __launch_bounds__(256) k_route_ghosts(const float4* __restrict__ lo, const float4* __restrict__ hi, const float* __restrict__ pos,
                                                      const float4* __restrict__ rot, uint32_t begin, uint32_t end, const uint32_t* __restrict__ bins,
                                                      const uint32_t* __restrict__ split, const float* __restrict__ region_f, int world, uint32_t cap,
                                                      int recw, RouteDst dst, uint32_t* __restrict__ counts) {
    __shared__ float s_r[SHARD_MAX_RANKS * SHARD_SUBS][6];
    for (int k = threadIdx.x; k < world * SHARD_SUBS * 6; k += blockDim.x) (&s_r[0][0])[k] = (k % 6) < 3 ? -region_f[k] : region_f[k];
    __syncthreads();
    uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    float4 a = __ldg(&lo[i]), b = __ldg(&hi[i]);
    int owner = route_owner(split, world, __ldg(&bins[i - begin]));
    for (int q = 0; q < world; ++q) {
        if (q == owner) continue;
        bool take = false;
        for (int sub = 0; sub < SHARD_SUBS && !take; ++sub) {  // an empty sub-box has min = +MAX, max = -MAX: never met
            const float* r6 = s_r[q * SHARD_SUBS + sub];
            take = a.x <= r6[3] && a.y <= r6[4] && a.z <= r6[5] && b.x >= r6[0] && b.y >= r6[1] && b.z >= r6[2];
        }
        if (!take) continue;
        uint32_t k = atomicAdd(&counts[q], 1u);
        if (k + 1 < cap) {
            float4* r = dst.p[q] + (size_t)(1 + k) * recw;
            float4 a2 = a, b2 = b;
            a2.w = __uint_as_float(i);
            b2.w = __uint_as_float((__float_as_uint(b.w) & 0xffu) | ((uint32_t)owner << 8));
            r[0] = a2, r[1] = b2;
            if (recw == 4) {
                float4 qq = __ldg(&rot[i]);
                r[2] = make_float4(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2], qq.x);
                r[3] = make_float4(qq.y, qq.z, qq.w, 0.f);
            }
        }
    }
}
// stage 4: the 2 * world received buckets (owned, then ghosts) -> compact local arrays; sh->m = records, sh->n_owned = owned ones.
// A bucket whose header exceeds its capacity was truncated by its sender: flagged in sh->split[0] (overflow), the step is repeated.
__global__ void __launch_bounds__(256) k_route_unpack(const float4* __restrict__ recv_o, const float4* __restrict__ recv_g, int world, uint32_t cap_o,
                                                      uint32_t cap_g, int recw, uint32_t cap_local, uint32_t* __restrict__ sel,
                                                      float4* __restrict__ loc_lo, float4* __restrict__ loc_hi, float* __restrict__ pos,
                                                      float4* __restrict__ rot, ShardScratch* sh) {
    __shared__ uint32_t s_off[2 * SHARD_MAX_RANKS + 1], s_cnt[2 * SHARD_MAX_RANKS];
    if (threadIdx.x == 0) {
        uint32_t acc = 0, owned = 0, need_o = 0, need_g = 0;
        for (int bk = 0; bk < 2 * world; ++bk) {
            bool ghost = bk >= world;
            const float4* hdr = ghost ? recv_g + (size_t)(bk - world) * cap_g * recw : recv_o + (size_t)bk * cap_o * recw;
            float4 hv = __ldcg(hdr);
            uint32_t c = __float_as_uint(hv.x), big = __float_as_uint(hv.y), cap = (ghost ? cap_g : cap_o) - 1;
            if (ghost) need_g = max(need_g, big + 1); else need_o = max(need_o, big + 1);
            c = min(c, cap);
            s_off[bk] = acc, s_cnt[bk] = c;
            acc += c;
            if (!ghost) owned += c;
        }
        s_off[2 * world] = acc;
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            sh->m = acc, sh->n_owned = owned;
            sh->split[0] = need_o, sh->split[1] = need_g;  // capacities this step would have needed (host: grow + repeat when larger)
        }
    }
    __syncthreads();
    int bk = blockIdx.y;
    bool ghost = bk >= world;
    uint32_t cap = ghost ? cap_g : cap_o;
    const float4* bucket = ghost ? recv_g + (size_t)(bk - world) * cap_g * recw : recv_o + (size_t)bk * cap_o * recw;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < s_cnt[bk]; k += gridDim.x * blockDim.x) {
        const float4* r = bucket + (size_t)(1 + k) * recw;
        uint32_t dst = s_off[bk] + k;
        if (dst >= cap_local) continue;
        float4 a = __ldcg(&r[0]), b = __ldcg(&r[1]);  // L2 reads: a peer may have written the bucket over NVLink
        uint32_t handle = __float_as_uint(a.w);
        sel[dst] = handle;
        loc_lo[dst] = a, loc_hi[dst] = b;
        if (recw == 4) {
            float4 p = __ldcg(&r[2]), q = __ldcg(&r[3]);
            pos[3 * (size_t)handle] = p.x, pos[3 * (size_t)handle + 1] = p.y, pos[3 * (size_t)handle + 2] = p.z;
            rot[handle] = make_float4(p.w, q.x, q.y, q.z);
        }
    }
    (void)cap;
}

static RouteDst route_dst(const RouteBufs& R, int world, bool ghosts) {
    RouteDst d;
    for (int q = 0; q < SHARD_MAX_RANKS; ++q) d.p[q] = nullptr;
    for (int q = 0; q < world; ++q) {
        if (R.p2p)  // my bucket inside peer q's receive buffer
            d.p[q] = (ghosts ? R.peer_recv_g[q] : R.peer_recv_o[q]) + (size_t)R.p2p_rank * R.p2p_cap * R.recw;
        else
            d.p[q] = (ghosts ? R.send_g.p + (size_t)q * R.cap_g * R.recw : R.send_o.p + (size_t)q * R.cap_o * R.recw);
    }
    return d;
}

cudaError_t launch_route_stage(ncb_ctx* c, int stage, int rank, int world, uint32_t begin, uint32_t end, RouteBufs& R) {
    (void)rank;
    cudaStream_t s = c->stream;
    uint32_t n_own = end - begin, nb = (n_own + 255) / 256;
    int gs = c->sm_count * 4;
    uint32_t cap_o = R.p2p ? R.p2p_cap : R.cap_o, cap_g = R.p2p ? R.p2p_cap : R.cap_g;
    switch (stage) {
        case 0: {  // bounds of the own block (needs counters reset by the caller)
            if (n_own) k_bounds<<<min((uint32_t)gs, nb), 256, 0, s>>>(c->aabb_lo.p + begin, c->aabb_hi.p + begin, n_own, c->counters.p);
            k_route_pack_bounds<<<1, 32, 0, s>>>(c->counters.p, R.bounds.p);
            break;
        }
        case 1: {
            cudaMemsetAsync(R.hist.p, 0, SHARD_BINS * sizeof(int), s);
            if (n_own) k_route_hist<<<min((uint32_t)c->sm_count * 2, nb), 256, 0, s>>>(c->aabb_lo.p, c->aabb_hi.p, begin, end, R.bounds.p, R.bins.p, R.hist.p);
            break;
        }
        case 2: {
            int z[SHARD_MAX_RANKS * SHARD_SUBS * 6];
            for (int k = 0; k < SHARD_MAX_RANKS * SHARD_SUBS * 6; ++k) z[k] = (k % 6) < 3 ? 0x7f7fffff : (int)0x80800000;
            cudaMemcpyAsync(R.region_i.p, z, sizeof z, cudaMemcpyHostToDevice, s);
            cudaMemsetAsync(R.counts.p, 0, 2 * SHARD_MAX_RANKS * sizeof(uint32_t), s);
            k_route_split<<<1, 32, 0, s>>>(R.hist.p, world, R.split.p);
            RouteDst d = route_dst(R, world, false);
            if (n_own)
                k_route_owned<<<nb, 256, 0, s>>>(c->aabb_lo.p, c->aabb_hi.p, c->pos.p, c->rot.p, begin, end, R.bins.p, R.split.p, world, cap_o, R.recw, d,
                                                 R.counts.p, R.region_i.p);
            k_route_finish<<<1, 128, 0, s>>>(R.counts.p, world, d, R.region_i.p, R.region_f.p);
            break;
        }
        case 3: {
            RouteDst d = route_dst(R, world, true);
            if (n_own)
                k_route_ghosts<<<nb, 256, 0, s>>>(c->aabb_lo.p, c->aabb_hi.p, c->pos.p, c->rot.p, begin, end, R.bins.p, R.split.p, R.region_f.p, world,
                                                  cap_g, R.recw, d, R.counts.p + SHARD_MAX_RANKS);
            k_route_finish<<<1, 128, 0, s>>>(R.counts.p + SHARD_MAX_RANKS, world, d, nullptr, nullptr);
            break;
        }
        default: break;
    }
    return cudaGetLastError();
}

// ---- peer-memory rounds: push a small array into every peer's meta slot and raise my flag there; wait for all flags; reduce ----
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// One CTA per peer.  Everything this rank stored for the round (records written by the kernels before this one on the stream, the
// words copied here) is ordered before the flag by the system-scope fence + release store.
__global__ void __launch_bounds__(256) k_p2p_push(const uint32_t* __restrict__ src, uint32_t words, RoutePeers peers, int slot, int rank,
                                                  uint32_t value) {
    uint32_t* m = peers.meta[blockIdx.x];
    uint32_t* dst = m + ((size_t)slot * SHARD_MAX_RANKS + rank) * P2P_SLOT_WORDS;
    for (uint32_t t = threadIdx.x; t < words; t += blockDim.x) dst[t] = src[t];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(m + P2P_FLAGS_OFF + rank, value);
}
// Waits until every sender has raised its flag to `value` (flags only grow), then reduces the senders' slots: op 0 = float max,
// 1 = int sum, < 0 = nothing to reduce.  A peer that does not arrive within ~2 s sets the error word instead of hanging the GPU.
__global__ void __launch_bounds__(256) k_p2p_wait_reduce(uint32_t* __restrict__ meta, int world, uint32_t value, int slot, uint32_t words, int op,
                                                         uint32_t* __restrict__ out) {
    if (threadIdx.x < (unsigned)world) {
        const uint32_t* f = meta + P2P_FLAGS_OFF + threadIdx.x;
        long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - value) < 0) {
            if (clock64() - t0 > 4000000000ll) {
                atomicExch(meta + P2P_ERR_OFF, 1u);
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (op < 0) return;
    for (uint32_t t = threadIdx.x; t < words; t += blockDim.x) {
        if (op == 0) {
            float v = -NCB_FMAX;
            for (int q = 0; q < world; ++q) v = fmaxf(v, __uint_as_float(__ldcg(meta + ((size_t)slot * SHARD_MAX_RANKS + q) * P2P_SLOT_WORDS + t)));
            out[t] = __float_as_uint(v);
        } else {
            uint32_t v = 0;
            for (int q = 0; q < world; ++q) v += __ldcg(meta + ((size_t)slot * SHARD_MAX_RANKS + q) * P2P_SLOT_WORDS + t);
            out[t] = v;
        }
    }
}
cudaError_t launch_p2p_push(ncb_ctx* c, RouteBufs& R, const void* src, uint32_t words, int slot, int round) {
    RoutePeers peers;
    for (int q = 0; q < SHARD_MAX_RANKS; ++q) peers.meta[q] = q < R.p2p_world ? R.peer_meta[q] : nullptr;
    k_p2p_push<<<R.p2p_world, 256, 0, c->stream>>>((const uint32_t*)src, words, peers, slot, R.p2p_rank, 4 * R.epoch + (uint32_t)round);
    return cudaGetLastError();
}
cudaError_t launch_p2p_wait_reduce(ncb_ctx* c, RouteBufs& R, int round, int slot, uint32_t words, int op, void* out) {
    k_p2p_wait_reduce<<<1, 256, 0, c->stream>>>(R.p2p_meta.p, R.p2p_world, 4 * R.epoch + (uint32_t)round, slot, words, op, (uint32_t*)out);
    return cudaGetLastError();
}
cudaError_t launch_route_unpack(ncb_ctx* c, int world, RouteBufs& R, uint32_t cap_local, ShardScratch* sh, uint32_t* sel, float4* loc_lo,
                                float4* loc_hi) {
    uint32_t cap_o = R.p2p ? R.p2p_cap : R.cap_o, cap_g = R.p2p ? R.p2p_cap : R.cap_g;
    const float4* ro = R.p2p ? R.p2p_recv_o.p : R.recv_o.p;
    const float4* rg = R.p2p ? R.p2p_recv_g.p : R.recv_g.p;
    uint32_t per = max(cap_o, cap_g);
    dim3 grid(min((per + 255) / 256, (uint32_t)c->sm_count * 2), 2 * world);
    k_route_unpack<<<grid, 256, 0, c->stream>>>(ro, rg, world, cap_o, cap_g, R.recw, cap_local, sel, loc_lo, loc_hi, c->pos.p, c->rot.p, sh);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// K5: pair search.  One thread per query leaf, in Morton order (neighbouring lanes traverse neighbouring boxes).
// A query at sorted position i reports only leaves at positions j > i, so each unordered pair is emitted once;
// it is oriented (larger handle, smaller handle) = the argument order of interference_started.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool groups_allow(const uint32_t* __restrict__ g, uint32_t a, uint32_t b) {
    if (!g) return true;
    uint32_t m1 = __ldg(&g[3 * a]), w1 = __ldg(&g[3 * a + 1]), b1 = __ldg(&g[3 * a + 2]);
    uint32_t m2 = __ldg(&g[3 * b]), w2 = __ldg(&g[3 * b + 1]), b2 = __ldg(&g[3 * b + 2]);
    return (m1 & b2) == 0 && (m2 & b1) == 0 && (m1 & w2) != 0 && (m2 & w1) != 0;
}

// Spatially sharded update (several GPUs): hi.w of a leaf carries `type | owner rank << 8`.  A rank reports a pair when it
// owns both objects, or owns one of them and has the lower rank of the two owners (both owners see the pair: the other
// object is a ghost there).  my_rank < 0: no filter.
__device__ __forceinline__ bool shard_allow(int my_rank, uint32_t tq, uint32_t tj) {
    if (my_rank < 0) return true;
    uint32_t me = (uint32_t)my_rank, oq = tq >> 8, oj = tj >> 8;
    if (oq != me && oj != me) return false;
    if (oq == oj) return true;
    return me < (oq == me ? oj : oq);
}

// Warp-aggregated append: one atomicAdd per converged group of emitting lanes.
__device__ __forceinline__ void emit_pair(uint32_t ha, uint32_t ta, uint32_t hb, uint32_t tb, uint2* __restrict__ pairs,
                                          uint8_t* __restrict__ keys, uint32_t cap, DevCounters* cnt) {
    cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(&cnt->n_pairs, (uint32_t)g.size());
    base = g.shfl(base, 0);
    uint32_t slot = base + g.thread_rank();
    if (slot < cap) {
        uint32_t h1, h2, t1, t2;
        if (ha > hb) {
            h1 = ha, t1 = ta, h2 = hb, t2 = tb;
        } else {
            h1 = hb, t1 = tb, h2 = ha, t2 = ta;
        }
        pairs[slot] = make_uint2(h1, h2);
        keys[slot] = (uint8_t)pair_key(t1, t2);
    }
}

__device__ __forceinline__ bool boxes_intersect(float4 alo, float4 ahi, float4 blo, float4 bhi) {
    // AABB::intersects (aabb.rs:156-158): inclusive; false on NaN
    return alo.x <= bhi.x && alo.y <= bhi.y && alo.z <= bhi.z && ahi.x >= blo.x && ahi.y >= blo.y && ahi.z >= blo.z;
}

// Candidate leaves are buffered per thread during the walk and emitted after it: the walk itself then contains no
// atomics, no group / handle gathers and no stores, and the emission runs converged (one atomicAdd per warp for all
// the pairs its 32 queries found).  A query with more than PAIR_BUF candidates spills through emit_pair() in-loop.
#define PAIR_BUF 12
__global__ void __launch_bounds__(128) k_pair_search(const float4* __restrict__ llo, const float4* __restrict__ lhi,
                                                     const float4* __restrict__ nodes, uint32_t n, const uint32_t* __restrict__ groups,
                                                     uint32_t q_begin, uint32_t q_end, uint2* __restrict__ pairs,
                                                     uint8_t* __restrict__ keys, uint32_t cap, DevCounters* cnt, int my_rank) {
    uint32_t m = n - cnt->n_outliers;
    uint32_t i = q_begin + blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < m && i < q_end && m >= 2;
    uint32_t buf[PAIR_BUF];
    int nb = 0;
    uint32_t hq = 0, tq = 0;
    if (valid) {
        float4 qlo = __ldg(&llo[i]), qhi = __ldg(&lhi[i]);
        hq = __float_as_uint(qlo.w), tq = __float_as_uint(qhi.w);
        uint32_t stack[64];
        int sp = 0;
        uint32_t node = 0;
        for (;;) {
            const float4* rec = nodes + 4 * (size_t)node;
            float4 Llo = __ldg(rec + 0), Lhi = __ldg(rec + 1), Rlo = __ldg(rec + 2), Rhi = __ldg(rec + 3);
            uint32_t left = __float_as_uint(Llo.w), right = __float_as_uint(Lhi.w);
            uint32_t split = __float_as_uint(Rlo.w), last = __float_as_uint(Rhi.w);
            bool goL = split > i && boxes_intersect(qlo, qhi, Llo, Lhi);
            bool goR = last > i && boxes_intersect(qlo, qhi, Rlo, Rhi);
            if (goL && (left & LEAF_BIT)) {
                uint32_t j = left & ~LEAF_BIT;
                if (nb < PAIR_BUF)
                    buf[nb++] = j;
                else {
                    uint32_t hj = __float_as_uint(__ldg(&llo[j].w)), tj = __float_as_uint(__ldg(&lhi[j].w));
                    if (shard_allow(my_rank, tq, tj) && groups_allow(groups, hq, hj)) emit_pair(hq, tq, hj, tj, pairs, keys, cap, cnt);
                }
                goL = false;
            }
            if (goR && (right & LEAF_BIT)) {
                uint32_t j = right & ~LEAF_BIT;
                if (nb < PAIR_BUF)
                    buf[nb++] = j;
                else {
                    uint32_t hj = __float_as_uint(__ldg(&llo[j].w)), tj = __float_as_uint(__ldg(&lhi[j].w));
                    if (shard_allow(my_rank, tq, tj) && groups_allow(groups, hq, hj)) emit_pair(hq, tq, hj, tj, pairs, keys, cap, cnt);
                }
                goR = false;
            }
            if (goL) {
                if (goR) {
                    if (sp < 64)
                        stack[sp++] = right;
                    else
                        atomicAdd(&cnt->stack_overflow, 1u);  // a subtree would be skipped: never silent (Karras depth <= 62 keeps it 0)
                }
                node = left;
            } else if (goR) {
                node = right;
            } else {
                if (sp == 0) break;
                node = stack[--sp];
            }
        }
    }
    // ---- converged emission: resolve handles / types / groups, compact, one allocation per warp ----
    uint32_t hjs[PAIR_BUF];
    uint8_t tjs[PAIR_BUF];
    int na = 0;
#pragma unroll
    for (int k = 0; k < PAIR_BUF; ++k) {
        if (k < nb) {
            uint32_t j = buf[k];
            uint32_t hj = __float_as_uint(__ldg(&llo[j].w)), tj = __float_as_uint(__ldg(&lhi[j].w));
            if (shard_allow(my_rank, tq, tj) && groups_allow(groups, hq, hj)) {
                hjs[na] = hj;
                tjs[na] = (uint8_t)tj;
                na++;
            }
        }
    }
    int lane = threadIdx.x & 31;
    uint32_t incl = (uint32_t)na;
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(&cnt->n_pairs, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    uint32_t slot = base + incl - (uint32_t)na;
    for (int k = 0; k < na; ++k, ++slot) {
        if (slot < cap) {
            uint32_t hb = hjs[k], tb = tjs[k];
            uint32_t h1, h2, t1, t2;
            if (hq > hb) {
                h1 = hq, t1 = tq, h2 = hb, t2 = tb;
            } else {
                h1 = hb, t1 = tb, h2 = hq, t2 = tq;
            }
            pairs[slot] = make_uint2(h1, h2);
            keys[slot] = (uint8_t)pair_key(t1, t2);
        }
    }
}

// Outliers (planes, boxes beyond 1e30) sit after the tree leaves in the sorted arrays and are tested against
// everything: their boxes overlap (almost) every object, so a tree would not prune anything.
__global__ void __launch_bounds__(256) k_pair_outliers(const float4* __restrict__ llo, const float4* __restrict__ lhi, uint32_t n,
                                                       const uint32_t* __restrict__ groups, uint32_t q_begin, uint32_t q_end,
                                                       uint2* __restrict__ pairs, uint8_t* __restrict__ keys, uint32_t cap,
                                                       DevCounters* cnt, int my_rank) {
    uint32_t nout = cnt->n_outliers;
    if (nout == 0) return;
    uint32_t m = n - nout;
    uint32_t i = q_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i >= q_end) return;
    float4 alo = __ldg(&llo[i]), ahi = __ldg(&lhi[i]);
    uint32_t ha = __float_as_uint(alo.w), ta = __float_as_uint(ahi.w);
    for (uint32_t o = max(m, i + 1); o < n; ++o) {
        float4 blo = __ldg(&llo[o]), bhi = __ldg(&lhi[o]);
        if (boxes_intersect(alo, ahi, blo, bhi)) {
            uint32_t hb = __float_as_uint(blo.w), tb = __float_as_uint(bhi.w);
            if (shard_allow(my_rank, ta, tb) && groups_allow(groups, ha, hb)) emit_pair(ha, ta, hb, tb, pairs, keys, cap, cnt);
        }
    }
}

cudaError_t launch_pair_search(ncb_ctx* c, uint32_t n, const uint32_t* groups, uint32_t q_begin, uint32_t q_end, uint32_t cap_pairs, int my_rank) {
    if (n == 0) return cudaSuccess;
    cudaStream_t s = c->stream;
    q_end = min(q_end, n);
    if (q_end <= q_begin) return cudaSuccess;
    uint32_t nq = q_end - q_begin;
    k_pair_search<<<(nq + 127) / 128, 128, 0, s>>>(c->leaf_lo.p, c->leaf_hi.p, c->nodes.p, n, groups, q_begin, q_end, c->pairs_raw.p,
                                                   c->keys_raw.p, cap_pairs, c->counters.p, my_rank);
    k_pair_outliers<<<(nq + 255) / 256, 256, 0, s>>>(c->leaf_lo.p, c->leaf_hi.p, n, groups, q_begin, q_end, c->pairs_raw.p,
                                                     c->keys_raw.p, cap_pairs, c->counters.p, my_rank);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// K6: counting sort of the pairs by type key (histogram -> scan -> scatter), all sized on the device:
// persistent grid-stride kernels read n_pairs from the counters, so the host never synchronises mid-update.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_key_hist(const uint8_t* __restrict__ keys, uint32_t cap, DevCounters* cnt) {
    __shared__ uint32_t h[K_MAX];
    if (threadIdx.x < K_MAX) h[threadIdx.x] = 0;
    __syncthreads();
    uint32_t np = min(cnt->n_pairs, cap);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) atomicAdd(&h[keys[p] & (K_MAX - 1)], 1u);
    __syncthreads();
    if (threadIdx.x < K_MAX && h[threadIdx.x]) atomicAdd(&cnt->key_hist[threadIdx.x], h[threadIdx.x]);
}
__global__ void k_key_scan(DevCounters* cnt) {
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int k = 0; k < K_MAX; ++k) {
            cnt->key_start[k] = acc;
            cnt->key_cursor[k] = acc;
            cnt->epa_cursor[k] = acc;
            cnt->cp_cursor[k] = acc;
            cnt->epa_fetch[k] = acc;
            cnt->gjk_fetch[k] = acc;
            acc += cnt->key_hist[k];
        }
    }
}
__global__ void __launch_bounds__(256) k_key_scatter(const uint2* __restrict__ pin, const uint8_t* __restrict__ kin, uint32_t cap,
                                                     uint2* __restrict__ pout, uint8_t* __restrict__ algo_out, uint32_t* __restrict__ index_out,
                                                     DevCounters* cnt, const uint32_t* __restrict__ local_of, uint2* __restrict__ pout_local) {
    uint32_t np = min(cnt->n_pairs, cap);
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < np; base += stride) {
        uint32_t p = base + threadIdx.x;
        bool valid = p < np;
        uint32_t key = valid ? (kin[p] & (K_MAX - 1)) : 0xffu;
        unsigned peers = __match_any_sync(0xffffffffu, key);
        if (valid) {
            int lane = threadIdx.x & 31;
            int leader = __ffs(peers) - 1;
            uint32_t b = 0;
            if (lane == leader) b = atomicAdd(&cnt->key_cursor[key], (uint32_t)__popc(peers));
            b = __shfl_sync(peers, b, leader);
            uint32_t dst = b + __popc(peers & ((1u << lane) - 1));
            uint2 pr = pin[p];
            pout[dst] = pr;
            // sharded updates: the narrow phase reads its operands from compact rank-local arrays (see k_gather_local_objects)
            if (local_of) pout_local[dst] = make_uint2(__ldg(&local_of[pr.x]), __ldg(&local_of[pr.y]));
            algo_out[dst] = (uint8_t)algo_of_key(key);
            if (index_out) index_out[dst] = p;
        }
    }
}

cudaError_t launch_pair_sort(ncb_ctx* c, uint32_t cap_pairs, uint32_t* index_out, const uint32_t* local_of, uint2* pairs_local) {
    cudaStream_t s = c->stream;
    int gs = c->sm_count * 4;
    k_key_hist<<<gs, 256, 0, s>>>(c->keys_raw.p, cap_pairs, c->counters.p);
    k_key_scan<<<1, 32, 0, s>>>(c->counters.p);
    k_key_scatter<<<gs, 256, 0, s>>>(c->pairs_raw.p, c->keys_raw.p, cap_pairs, c->pairs.p, c->pair_algo.p, index_out,
                                       c->counters.p, local_of, pairs_local);
    return cudaGetLastError();
}

// Sharded updates: a rank holds ~N / ranks + ghosts of the N objects, scattered over the replicated object arrays by global handle.
// At 8 M objects those arrays are 600 MB — not L2-resident the way the 76 MB of a 1 M-object world are — and every operand read of
// the narrow phase became a DRAM access (GJK + 0.07 ms, manifold + 0.18 ms at 8 GPUs).  One gather pass copies the attributes of the
// objects this rank holds into compact arrays indexed by LOCAL id (the position in the selected / unpacked list), and records
// local_of[global handle]; the sorted pair list gets a twin with local ids for the narrow-phase kernels.  Results are indexed by pair,
// so nothing else changes; the reported pairs keep their global handles.
__global__ void __launch_bounds__(256) k_gather_local_objects(const uint32_t* __restrict__ sel, uint32_t m, DevObjects g, float* __restrict__ lpos,
                                                              float4* __restrict__ lrot, uint32_t* __restrict__ ltype, float4* __restrict__ lparam,
                                                              float* __restrict__ lqlimit, float2* __restrict__ lang_cs, float* __restrict__ lcap,
                                                              uint32_t* __restrict__ local_of) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t h = __ldg(&sel[k]);
    local_of[h] = k;
    lpos[3 * (size_t)k] = __ldg(g.pos + 3 * (size_t)h), lpos[3 * (size_t)k + 1] = __ldg(g.pos + 3 * (size_t)h + 1);
    lpos[3 * (size_t)k + 2] = __ldg(g.pos + 3 * (size_t)h + 2);
    lrot[k] = __ldg(&g.rot[h]);
    ltype[k] = __ldg(&g.type[h]);
    lparam[k] = __ldg(&g.param[h]);
    lqlimit[k] = __ldg(&g.qlimit[h]);
    if (lang_cs) lang_cs[k] = __ldg(&g.ang_cs[h]);
    if (lcap)
        for (int d = 0; d < 6; ++d) lcap[6 * (size_t)k + d] = __ldg(g.cap_pts + 6 * (size_t)h + d);
}
cudaError_t launch_gather_local_objects(ncb_ctx* c, const uint32_t* sel, uint32_t m, const DevObjects& g, DevObjects* out) {
    cudaError_t e;
    if ((e = c->loc_pos.reserve(3 * (size_t)m + 3)) != cudaSuccess || (e = c->loc_rot.reserve(m)) != cudaSuccess ||
        (e = c->loc_type.reserve(m)) != cudaSuccess || (e = c->loc_param.reserve(m)) != cudaSuccess ||
        (e = c->loc_qlimit.reserve(m)) != cudaSuccess || (e = c->local_of.reserve(c->n)) != cudaSuccess)
        return e;
    if (g.ang_stride && (e = c->loc_ang_cs.reserve(m)) != cudaSuccess) return e;
    if (g.cap_pts && (e = c->loc_cap.reserve(6 * (size_t)m)) != cudaSuccess) return e;
    if (m)
        k_gather_local_objects<<<(m + 255) / 256, 256, 0, c->stream>>>(sel, m, g, c->loc_pos.p, c->loc_rot.p, c->loc_type.p, c->loc_param.p,
                                                                       c->loc_qlimit.p, g.ang_stride ? c->loc_ang_cs.p : nullptr,
                                                                       g.cap_pts ? c->loc_cap.p : nullptr, c->local_of.p);
    DevObjects o = g;
    o.n = m;
    o.pos = c->loc_pos.p, o.rot = c->loc_rot.p, o.type = c->loc_type.p, o.param = c->loc_param.p, o.qlimit = c->loc_qlimit.p;
    o.groups = nullptr;  // the narrow phase does not read groups
    o.ang = nullptr;     // nor the raw angles (ang_cs holds their cos / sin)
    if (g.ang_stride) o.ang_cs = c->loc_ang_cs.p;
    if (g.cap_pts) o.cap_pts = c->loc_cap.p;
    *out = o;
    return cudaGetLastError();
}

}  // namespace ncb
