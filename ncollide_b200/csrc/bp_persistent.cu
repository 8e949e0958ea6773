// Persistent (multi-step) broad phase behind the reference's BroadPhase trait surface:
// create_proxy / remove / deferred_set_bounding_volume / update with interference_started / interference_stopped.
//
// Replaces (reference, file:line): pipeline/broad_phase/dbvt_broad_phase.rs:101-147 (purge), :174-259 (update),
// :262-347 (proxy, create_proxy, remove, deferred_set_bounding_volume); slab handle reuse (LIFO) as the `slab` crate.
//
// Semantics used: after every update() the reference's pair set equals {(i, j): stored boxes intersect, pair allowed}
//   - a proxy's stored box only changes when the new box is NOT contained in it (then it becomes new.loosened(margin));
//   - re-inserted leaves query both trees, every other pair is re-validated by the purge (`updated` is never reset in
//     the reference, dbvt_broad_phase.rs:39,206), DBVT internal boxes are supersets, so nothing is missed or kept wrongly.
// So an update is: apply pending boxes -> LBVH over the attached proxies -> all pairs -> sorted 64-bit keys ->
// started = new \ old, stopped = old \ new.  interference_started(a, b) gets the re-inserted leaf first (the later
// one in update order when both moved); interference_stopped gets SortedPair order (smaller handle first).
#include <cub/cub.cuh>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "ncb_internal.h"

using namespace ncb;

#include "bp_internal.h"

namespace {

__global__ void k_bp_fill(uint32_t* p, uint32_t v, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// create_proxy: pending box = bv as given (no margin), dbvt_broad_phase.rs:275-280
__global__ void k_bp_stage_create(const uint32_t* __restrict__ handles, const float* __restrict__ mm, uint32_t n, uint32_t seq0, float4* pend_lo,
                                  float4* pend_hi, uint32_t* pend_seq, uint32_t* d_attached) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = handles[i];
    const float* s = mm + 6 * (size_t)i;
    pend_lo[h] = make_float4(s[0], s[1], s[2], 0.f);
    pend_hi[h] = make_float4(s[3], s[4], s[5], 0.f);
    // a recycled slot may still carry a queued entry of its previous owner (the reference's queue is keyed by handle):
    // the new leaf then takes that earlier place in the update order
    pend_seq[h] = min(pend_seq[h], seq0 + i);
    d_attached[h] = ST_DETACHED;
}
// deferred_set_bounding_volume (:325-347), pass 1: which entries push, who wins per handle, first sequence number
__global__ void k_bp_stage_set1(const uint32_t* __restrict__ handles, const float* __restrict__ mm, uint32_t n, uint32_t seq0,
                                const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, const uint32_t* __restrict__ d_attached,
                                uint32_t* win, uint32_t* pend_seq) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = handles[i];
    const float* s = mm + 6 * (size_t)i;
    bool needs = true;
    if (d_attached[h] == ST_ATTACHED) {
        float4 lo = box_lo[h], hi = box_hi[h];
        bool contains = lo.x <= s[0] && lo.y <= s[1] && lo.z <= s[2] && hi.x >= s[3] && hi.y >= s[4] && hi.z >= s[5];  // aabb.rs:161
        needs = !contains;
    }
    if (needs) {
        atomicMax(&win[h], i + 1);
        atomicMin(&pend_seq[h], seq0 + i);
    }
}
// pass 2: the last pushing entry of each handle writes bv.loosened(margin)
__global__ void k_bp_stage_set2(const uint32_t* __restrict__ handles, const float* __restrict__ mm, uint32_t n, float margin,
                                const uint32_t* __restrict__ win, float4* pend_lo, float4* pend_hi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = handles[i];
    if (win[h] != i + 1) return;
    const float* s = mm + 6 * (size_t)i;
    pend_lo[h] = make_float4(s[0] + (-margin), s[1] + (-margin), s[2] + (-margin), 0.f);
    pend_hi[h] = make_float4(s[3] + margin, s[4] + margin, s[5] + margin, 0.f);
}
// deferred_recompute_all_proximities_with (:349-363): an attached proxy is queued at the FRONT with its stored box;
// a box already queued for it stays the one applied (it sits later in the queue)
__global__ void k_bp_recompute_with(const uint32_t* __restrict__ handles, uint32_t n, uint32_t front0, const float4* __restrict__ box_lo,
                                    const float4* __restrict__ box_hi, const uint32_t* __restrict__ d_attached, float4* pend_lo, float4* pend_hi,
                                    uint32_t* pend_seq) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = handles ? handles[i] : i;
    if (d_attached[h] != ST_ATTACHED) return;
    if (pend_seq[h] == SEQ_NONE) {
        pend_lo[h] = box_lo[h];
        pend_hi[h] = box_hi[h];
    }
    atomicMin(&pend_seq[h], front0 - i);
}
__global__ void k_bp_clear_win(const uint32_t* __restrict__ handles, uint32_t n, uint32_t* win) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) win[handles[i]] = 0;
}
// update, step 1: apply pending boxes
__global__ void k_bp_apply(uint32_t slots, float4* box_lo, float4* box_hi, const float4* __restrict__ pend_lo, const float4* __restrict__ pend_hi,
                           uint32_t* pend_seq, uint32_t* upd_seq, uint32_t* d_attached) {
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= slots) return;
    uint32_t s = pend_seq[h];
    pend_seq[h] = SEQ_NONE;
    if (d_attached[h] == ST_VACANT) s = SEQ_NONE;  // queued entries of removed proxies are dropped (:180-182)
    upd_seq[h] = s;
    if (s != SEQ_NONE) {
        box_lo[h] = pend_lo[h];
        box_hi[h] = pend_hi[h];
        d_attached[h] = ST_ATTACHED;
    }
}
__global__ void k_bp_gather(const uint32_t* __restrict__ alive, uint32_t m, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi,
                            float4* lo, float4* hi) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t h = alive[k];
    lo[k] = box_lo[h];
    hi[k] = box_hi[h];
}
__global__ void k_bp_keys(const uint2* __restrict__ pairs, uint32_t np, unsigned long long* keys) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    uint2 pr = pairs[p];
    uint32_t lo = min(pr.x, pr.y), hi = max(pr.x, pr.y);
    keys[p] = ((unsigned long long)lo << 32) | hi;
}
__device__ bool bp_contains_key(const unsigned long long* __restrict__ keys, uint32_t n, unsigned long long k) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        unsigned long long v = keys[mid];
        if (v < k)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo < n && keys[lo] == k;
}
// out = a \ b (both sorted); mode 1 orients the pair like interference_started
__global__ void k_bp_diff(const unsigned long long* __restrict__ a, uint32_t na, const unsigned long long* __restrict__ b, uint32_t nb, int mode,
                          const uint32_t* __restrict__ upd_seq, unsigned long long* out, uint32_t* counter) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    unsigned long long k = a[i];
    if (bp_contains_key(b, nb, k)) return;
    uint32_t slot = atomicAdd(counter, 1u);
    uint32_t lo = (uint32_t)(k >> 32), hi = (uint32_t)k;
    uint32_t first = lo, second = hi;
    if (mode == 1) {
        uint32_t sl = upd_seq[lo], sh = upd_seq[hi];
        // the re-inserted leaf comes first; when both were re-inserted, the later one met the earlier one in the tree
        bool hi_first = (sh != SEQ_NONE) && (sl == SEQ_NONE || sh > sl);
        if (hi_first) first = hi, second = lo;
    }
    out[slot] = ((unsigned long long)first << 32) | second;
}
__global__ void k_bp_unpack(const unsigned long long* __restrict__ ev, uint32_t n, uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = ev[i];
    out[2 * i] = (uint32_t)(k >> 32);
    out[2 * i + 1] = (uint32_t)k;
}
// remove(): drop the keys that involve a removed handle
__global__ void k_bp_mark_removed(unsigned long long* keys, uint32_t n, const uint32_t* __restrict__ d_attached, unsigned long long* out,
                                  uint32_t* counter) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = keys[i];
    uint32_t lo = (uint32_t)(k >> 32), hi = (uint32_t)k;
    if (d_attached[lo] == ST_REMOVING || d_attached[hi] == ST_REMOVING) {
        out[atomicAdd(counter, 1u)] = k;
        keys[i] = ~0ull;
    }
}
__global__ void k_bp_set_flags(const uint32_t* __restrict__ handles, uint32_t n, uint32_t v, uint32_t* d_attached) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d_attached[handles[i]] = v;
}


// ---- interferences_with_bounding_volume / _ray / _point (:388-432) on the LBVH of the last update -------------------
struct BpQuery {
    float a[7];
};
template <int KIND>
__device__ __forceinline__ bool bp_query_test(const BpQuery& q, const float* inv, float4 lo, float4 hi) {
    if (KIND == 0)  // AABB::intersects (aabb.rs:156-158)
        return lo.x <= q.a[3] && lo.y <= q.a[4] && lo.z <= q.a[5] && hi.x >= q.a[0] && hi.y >= q.a[1] && hi.z >= q.a[2];
    if (KIND == 2) {  // AABB::contains_local_point (aabb.rs:138-146)
        if (q.a[0] < lo.x || q.a[0] > hi.x) return false;
        if (q.a[1] < lo.y || q.a[1] > hi.y) return false;
        if (q.a[2] < lo.z || q.a[2] > hi.z) return false;
        return true;
    }
    // AABB::toi_with_ray(.., solid).is_some() (ray_aabb.rs:13-50); inv[i] = 1 / dir[i] as the reference computes it per box
    float tmin = 0.f, tmax = q.a[6];
    const float mn[3] = {lo.x, lo.y, lo.z}, mx[3] = {hi.x, hi.y, hi.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (q.a[3 + i] == 0.f) {
            if (q.a[i] < mn[i] || q.a[i] > mx[i]) return false;
        } else {
            float near = (mn[i] - q.a[i]) * inv[i], far = (mx[i] - q.a[i]) * inv[i];
            if (near > far) {
                float t = near;
                near = far;
                far = t;
            }
            tmin = fmaxf(tmin, near);
            tmax = fminf(tmax, far);
            if (tmin > tmax) return false;
        }
    }
    return true;
}

__device__ __forceinline__ void bp_query_emit(uint32_t qi, uint32_t handle, const uint32_t* __restrict__ d_attached, unsigned long long* out,
                                              uint32_t cap, uint32_t* counter) {
    if (d_attached[handle] != ST_ATTACHED) return;  // removed since the tree was built
    uint32_t slot = atomicAdd(counter, 1u);
    if (slot < cap) out[slot] = ((unsigned long long)qi << 32) | handle;
}

template <int KIND>
__global__ void __launch_bounds__(128) k_bp_query(const float* __restrict__ qin, uint32_t nq, const float4* __restrict__ llo,
                                                  const float4* __restrict__ lhi, const float4* __restrict__ nodes, uint32_t n, uint32_t nout,
                                                  const uint32_t* __restrict__ d_attached, unsigned long long* out, uint32_t cap,
                                                  uint32_t* counter, uint32_t* trav_overflow) {
    const int W = KIND == 0 ? 6 : (KIND == 1 ? 7 : 3);
    uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    BpQuery q;
    for (int k = 0; k < W; ++k) q.a[k] = qin[(size_t)W * qi + k];
    float inv[3] = {0.f, 0.f, 0.f};
    if (KIND == 1)
        for (int k = 0; k < 3; ++k) inv[k] = 1.0f / q.a[3 + k];
    uint32_t m = n - nout;
    if (m >= 2) {
        uint32_t stack[64];
        int sp = 0;
        uint32_t node = 0;
        for (;;) {
            const float4* rec = nodes + 4 * (size_t)node;
            float4 Llo = __ldg(rec + 0), Lhi = __ldg(rec + 1), Rlo = __ldg(rec + 2), Rhi = __ldg(rec + 3);
            uint32_t left = __float_as_uint(Llo.w), right = __float_as_uint(Lhi.w);
            bool goL = bp_query_test<KIND>(q, inv, Llo, Lhi), goR = bp_query_test<KIND>(q, inv, Rlo, Rhi);
            if (goL && (left & LEAF_BIT)) {
                bp_query_emit(qi, __float_as_uint(__ldg(&llo[left & ~LEAF_BIT].w)), d_attached, out, cap, counter);
                goL = false;
            }
            if (goR && (right & LEAF_BIT)) {
                bp_query_emit(qi, __float_as_uint(__ldg(&llo[right & ~LEAF_BIT].w)), d_attached, out, cap, counter);
                goR = false;
            }
            if (goL) {
                if (goR) {
                    if (sp < 64)
                        stack[sp++] = right;
                    else
                        atomicAdd(trav_overflow, 1u);
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
        float4 lo = __ldg(&llo[0]), hi = __ldg(&lhi[0]);
        if (bp_query_test<KIND>(q, inv, lo, hi)) bp_query_emit(qi, __float_as_uint(lo.w), d_attached, out, cap, counter);
    }
    for (uint32_t o = m; o < n; ++o) {  // leaves kept out of the tree (planes, huge boxes)
        float4 lo = __ldg(&llo[o]), hi = __ldg(&lhi[o]);
        if (bp_query_test<KIND>(q, inv, lo, hi)) bp_query_emit(qi, __float_as_uint(lo.w), d_attached, out, cap, counter);
    }
}

}  // namespace

#define CKB(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            char b__[512];                                                                                \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            bp->err = b__;                                                                                \
            bp->owner->err = b__;                                                                         \
            return NCB_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

static int bp_grow(ncb_bp* bp, size_t slots) {
    if (slots <= bp->slots_cap) return NCB_OK;
    size_t want = slots + slots / 2 + 1024;
    cudaStream_t s = bp->owner->stream;
    auto grow4 = [&](DevBuf<float4>& b) -> cudaError_t {
        DevBuf<float4> nb;
        cudaError_t e = nb.reserve(want);
        if (e != cudaSuccess) return e;
        if (b.p && bp->slots_cap) e = cudaMemcpyAsync(nb.p, b.p, bp->slots_cap * sizeof(float4), cudaMemcpyDeviceToDevice, s);
        cudaStreamSynchronize(s);
        b.release();
        b = std::move(nb);
        return e;
    };
    auto growu = [&](DevBuf<uint32_t>& b, uint32_t fill) -> cudaError_t {
        DevBuf<uint32_t> nb;
        cudaError_t e = nb.reserve(want);
        if (e != cudaSuccess) return e;
        k_bp_fill<<<(unsigned)((nb.cap + 255) / 256), 256, 0, s>>>(nb.p, fill, nb.cap);
        if (b.p && bp->slots_cap) e = cudaMemcpyAsync(nb.p, b.p, bp->slots_cap * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
        cudaStreamSynchronize(s);
        b.release();
        b = std::move(nb);
        return e;
    };
    CKB(grow4(bp->box_lo));
    CKB(grow4(bp->box_hi));
    CKB(grow4(bp->pend_lo));
    CKB(grow4(bp->pend_hi));
    CKB(growu(bp->pend_seq, SEQ_NONE));
    CKB(growu(bp->upd_seq, SEQ_NONE));
    CKB(growu(bp->win, 0));
    CKB(growu(bp->d_attached, ST_VACANT));
    bp->slots_cap = bp->box_lo.cap;
    return NCB_OK;
}

extern "C" {

int ncb_bp_create(ncb_ctx* ctx, float margin, ncb_bp** out) {
    if (!ctx || !out) return NCB_ERR_ARG;
    if (!(margin >= 0.f)) {
        ctx->err = "The loosening margin must be positive.";  // aabb.rs:181-184
        return NCB_ERR_ARG;
    }
    ncb_bp* bp = new ncb_bp;
    bp->owner = ctx;
    bp->margin = margin;
    ncb_ctx* w = new ncb_ctx;
    w->device = ctx->device;
    w->stream = ctx->stream;
    w->sm_count = ctx->sm_count;
    bp->work = w;
    *out = bp;
    return NCB_OK;
}

void ncb_bp_destroy(ncb_bp* bp) {
    if (!bp) return;
    cudaSetDevice(bp->owner->device);
    cudaStreamSynchronize(bp->owner->stream);
    ncb_ctx* w = bp->work;
    w->aabb_lo.release(), w->aabb_hi.release(), w->keys_a.release(), w->keys_b.release(), w->idx_a.release(), w->idx_b.release();
    w->cub_tmp.release(), w->leaf_lo.release(), w->leaf_hi.release(), w->nodes.release(), w->parent.release(), w->flags.release();
    w->pairs_raw.release(), w->keys_raw.release(), w->counters.release();
    if (w->h_counters) cudaFreeHost(w->h_counters);
    delete w;
    bp->box_lo.release(), bp->box_hi.release(), bp->pend_lo.release(), bp->pend_hi.release();
    bp->pend_seq.release(), bp->upd_seq.release(), bp->win.release(), bp->d_attached.release();
    bp->stage_f.release(), bp->stage_u.release(), bp->alive.release(), bp->groups_dev.release();
    bp->keys_old.release(), bp->keys_new.release(), bp->keys_tmp.release(), bp->cub_tmp.release();
    bp->q_in.release(), bp->ev_a.release(), bp->ev_b.release(), bp->ev_sorted.release(), bp->ev_u32.release(), bp->counters.release();
    delete bp;
}

int ncb_bp_create_proxies(ncb_bp* bp, uint32_t n, const float* aabb_minmax, uint32_t* out_handles) {
    if (!bp || (n && (!aabb_minmax || !out_handles))) return NCB_ERR_ARG;
    CKB(cudaSetDevice(bp->owner->device));
    if (n == 0) return NCB_OK;
    for (uint32_t i = 0; i < n; ++i) {  // Slab::insert
        size_t key = bp->next;
        if (key == bp->next_free.size()) {
            bp->next_free.push_back(-2);
            bp->attached.push_back(0);
            bp->next = key + 1;
        } else {
            bp->next = (size_t)bp->next_free[key];
            bp->next_free[key] = -2;
            bp->attached[key] = 0;
        }
        bp->len++;
        out_handles[i] = (uint32_t)key;
    }
    bp->slab_dirty = true;
    int r = bp_grow(bp, bp->next_free.size());
    if (r) return r;
    cudaStream_t s = bp->owner->stream;
    CKB(bp->stage_f.reserve(6 * (size_t)n));
    CKB(bp->stage_u.reserve(n));
    CKB(cudaMemcpyAsync(bp->stage_f.p, aabb_minmax, 24 * (size_t)n, cudaMemcpyHostToDevice, s));
    CKB(cudaMemcpyAsync(bp->stage_u.p, out_handles, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    k_bp_stage_create<<<(n + 255) / 256, 256, 0, s>>>(bp->stage_u.p, bp->stage_f.p, n, SEQ_BACK0 + bp->seq, bp->pend_lo.p, bp->pend_hi.p, bp->pend_seq.p,
                                                      bp->d_attached.p);
    CKB(cudaGetLastError());
    CKB(cudaStreamSynchronize(s));  // the staging buffers are reused by the next call
    bp->seq += n;
    return NCB_OK;
}

int ncb_bp_set_bounding_volumes(ncb_bp* bp, uint32_t n, const uint32_t* handles, const float* aabb_minmax) {
    if (!bp || (n && (!handles || !aabb_minmax))) return NCB_ERR_ARG;
    CKB(cudaSetDevice(bp->owner->device));
    if (n == 0) return NCB_OK;
    for (uint32_t i = 0; i < n; ++i)
        if (handles[i] >= bp->next_free.size() || bp->next_free[handles[i]] != -2) {
            bp->err = bp->owner->err = "Attempting to set the bounding volume of an object that does not exist.";  // :345
            return NCB_ERR_ARG;
        }
    cudaStream_t s = bp->owner->stream;
    CKB(bp->stage_f.reserve(6 * (size_t)n));
    CKB(bp->stage_u.reserve(n));
    CKB(cudaMemcpyAsync(bp->stage_f.p, aabb_minmax, 24 * (size_t)n, cudaMemcpyHostToDevice, s));
    CKB(cudaMemcpyAsync(bp->stage_u.p, handles, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    unsigned g = (n + 255) / 256;
    k_bp_clear_win<<<g, 256, 0, s>>>(bp->stage_u.p, n, bp->win.p);
    k_bp_stage_set1<<<g, 256, 0, s>>>(bp->stage_u.p, bp->stage_f.p, n, SEQ_BACK0 + bp->seq, bp->box_lo.p, bp->box_hi.p, bp->d_attached.p, bp->win.p,
                                      bp->pend_seq.p);
    k_bp_stage_set2<<<g, 256, 0, s>>>(bp->stage_u.p, bp->stage_f.p, n, bp->margin, bp->win.p, bp->pend_lo.p, bp->pend_hi.p);
    CKB(cudaGetLastError());
    CKB(cudaStreamSynchronize(s));
    bp->seq += n;
    return NCB_OK;
}

// BroadPhase::deferred_recompute_all_proximities_with (:349-363) for n handles, in order (unknown or not yet
// attached handles are ignored like in the reference).
int ncb_bp_recompute_with(ncb_bp* bp, uint32_t n, const uint32_t* handles) {
    if (!bp || (n && !handles)) return NCB_ERR_ARG;
    CKB(cudaSetDevice(bp->owner->device));
    std::vector<uint32_t> ok;
    ok.reserve(n);
    for (uint32_t i = 0; i < n; ++i)
        if (handles[i] < bp->next_free.size() && bp->next_free[handles[i]] == -2 && bp->attached[handles[i]]) ok.push_back(handles[i]);
    if (ok.empty()) return NCB_OK;
    cudaStream_t s = bp->owner->stream;
    uint32_t m = (uint32_t)ok.size();
    CKB(bp->stage_u.reserve(m));
    CKB(cudaMemcpyAsync(bp->stage_u.p, ok.data(), 4 * (size_t)m, cudaMemcpyHostToDevice, s));
    k_bp_recompute_with<<<(m + 255) / 256, 256, 0, s>>>(bp->stage_u.p, m, SEQ_FRONT0 - bp->front, bp->box_lo.p, bp->box_hi.p, bp->d_attached.p,
                                                        bp->pend_lo.p, bp->pend_hi.p, bp->pend_seq.p);
    CKB(cudaGetLastError());
    CKB(cudaStreamSynchronize(s));
    bp->front += m;
    return NCB_OK;
}

// BroadPhase::deferred_recompute_all_proximities (:365-386): every attached proxy, in slab order.
int ncb_bp_recompute_all(ncb_bp* bp) {
    if (!bp) return NCB_ERR_ARG;
    CKB(cudaSetDevice(bp->owner->device));
    uint32_t slots = (uint32_t)bp->next_free.size();
    if (slots == 0) return NCB_OK;
    cudaStream_t s = bp->owner->stream;
    k_bp_recompute_with<<<(slots + 255) / 256, 256, 0, s>>>(nullptr, slots, SEQ_FRONT0 - bp->front, bp->box_lo.p, bp->box_hi.p, bp->d_attached.p,
                                                            bp->pend_lo.p, bp->pend_hi.p, bp->pend_seq.p);
    CKB(cudaGetLastError());
    bp->front += slots;
    return NCB_OK;
}

static int bp_sort_keys(ncb_bp* bp, unsigned long long* in, unsigned long long* out, uint32_t n) {
    if (n == 0) return NCB_OK;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, in, out, (int)n, 0, 64);
    CKB(bp->cub_tmp.reserve(bytes + 256));
    bytes = bp->cub_tmp.cap;
    CKB(cub::DeviceRadixSort::SortKeys(bp->cub_tmp.p, bytes, in, out, (int)n, 0, 64, bp->owner->stream));
    return NCB_OK;
}

// Sorts the n events of `ev` (deterministic output order) in place.
static int bp_sort_events(ncb_bp* bp, DevBuf<unsigned long long>& ev, uint32_t n) {
    if (n < 2) return NCB_OK;
    CKB(bp->ev_sorted.reserve(n));
    int r = bp_sort_keys(bp, ev.p, bp->ev_sorted.p, n);
    if (r) return r;
    CKB(cudaMemcpyAsync(ev.p, bp->ev_sorted.p, 8 * (size_t)n, cudaMemcpyDeviceToDevice, bp->owner->stream));  // keeps both capacities stable
    return NCB_OK;
}

// BroadPhase::remove (:284-323): every interference of a removed proxy is dropped and reported (the reference calls
// its removal handler with the two proxies' data; here: the pair, smaller handle first); events land in the
// "stopped" list of ncb_bp_events.
int ncb_bp_remove(ncb_bp* bp, uint32_t n, const uint32_t* handles, uint32_t* n_removed) {
    if (!bp || (n && !handles)) return NCB_ERR_ARG;
    CKB(cudaSetDevice(bp->owner->device));
    if (n_removed) *n_removed = 0;
    bp->n_started = bp->n_stopped = 0;
    if (n == 0) return NCB_OK;
    std::vector<uint8_t> seen(bp->next_free.size(), 0);
    for (uint32_t i = 0; i < n; ++i) {
        if (handles[i] >= bp->next_free.size() || bp->next_free[handles[i]] != -2 || seen[handles[i]]) {
            bp->err = bp->owner->err = "Attempting to remove an object that does not exist.";  // :297
            return NCB_ERR_ARG;
        }
        seen[handles[i]] = 1;
    }
    cudaStream_t s = bp->owner->stream;
    CKB(bp->stage_u.reserve(n));
    CKB(bp->counters.reserve(4));
    CKB(cudaMemcpyAsync(bp->stage_u.p, handles, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    k_bp_set_flags<<<(n + 255) / 256, 256, 0, s>>>(bp->stage_u.p, n, ST_REMOVING, bp->d_attached.p);
    uint32_t nr = 0;
    if (bp->n_old) {
        CKB(cudaMemsetAsync(bp->counters.p, 0, 16, s));
        CKB(bp->ev_b.reserve(bp->n_old));
        k_bp_mark_removed<<<(bp->n_old + 255) / 256, 256, 0, s>>>(bp->keys_old.p, bp->n_old, bp->d_attached.p, bp->ev_b.p, bp->counters.p);
        CKB(cudaGetLastError());
        CKB(cudaMemcpyAsync(&nr, bp->counters.p, 4, cudaMemcpyDeviceToHost, s));
        CKB(cudaStreamSynchronize(s));
        if (nr) {
            CKB(bp->keys_tmp.reserve(bp->n_old));
            int r = bp_sort_keys(bp, bp->keys_old.p, bp->keys_tmp.p, bp->n_old);  // dropped keys (~0) go last
            if (r) return r;
            std::swap(bp->keys_old, bp->keys_tmp);
            bp->n_old -= nr;
            r = bp_sort_events(bp, bp->ev_b, nr);
            if (r) return r;
        }
    }
    k_bp_set_flags<<<(n + 255) / 256, 256, 0, s>>>(bp->stage_u.p, n, ST_VACANT, bp->d_attached.p);
    CKB(cudaGetLastError());
    CKB(cudaStreamSynchronize(s));
    for (uint32_t i = 0; i < n; ++i) {  // Slab::remove, in argument order
        uint32_t h = handles[i];
        if (bp->attached[h]) bp->n_attached--;
        bp->attached[h] = 0;
        bp->next_free[h] = (int64_t)bp->next;
        bp->next = h;
        bp->len--;
    }
    bp->slab_dirty = true;
    bp->n_stopped = nr;
    if (n_removed) *n_removed = nr;
    return NCB_OK;
}

// BroadPhase::update (:174-259).  groups: 3 words per handle slot (membership, whitelist, blacklist) or NULL —
// the `allow_proximity` filter of the reference's handler, evaluated like CollisionGroups::can_interact_with_groups.
int ncb_bp_update(ncb_bp* bp, const uint32_t* groups, uint32_t n_group_slots, uint32_t* n_started, uint32_t* n_stopped) {
    if (!bp) return NCB_ERR_ARG;
    CKB(cudaSetDevice(bp->owner->device));
    uint32_t slots = (uint32_t)bp->next_free.size();
    const uint32_t* dgroups = nullptr;
    if (groups && slots) {
        if (n_group_slots < slots) {
            bp->err = bp->owner->err = "ncb_bp_update: groups must cover every handle slot";
            return NCB_ERR_ARG;
        }
        CKB(bp->groups_dev.reserve(3 * (size_t)slots));
        CKB(cudaMemcpyAsync(bp->groups_dev.p, groups, 12 * (size_t)slots, cudaMemcpyHostToDevice, bp->owner->stream));
        dgroups = bp->groups_dev.p;
    }
    return bp_update_impl(bp, dgroups, n_started, n_stopped, true);
}

}  // extern "C"

// d_groups: device pointer, 3 words per handle slot, or nullptr
int bp_update_impl(ncb_bp* bp, const uint32_t* d_groups, uint32_t* n_started, uint32_t* n_stopped, bool want_events) {
    CKB(cudaSetDevice(bp->owner->device));
    cudaStream_t s = bp->owner->stream;
    ncb_ctx* w = bp->work;
    w->stream = s;
    bp->n_started = bp->n_stopped = 0;
    if (n_started) *n_started = 0;
    if (n_stopped) *n_stopped = 0;
    uint32_t slots = (uint32_t)bp->next_free.size();
    if (slots == 0) return NCB_OK;
    static const bool prof = getenv("NCB_SIM_PROFILE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!prof) return;
        cudaStreamSynchronize(s);
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[bp]    %-12s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    bool any_pending = bp->seq != 0 || bp->front != 0;
    if (!any_pending) return NCB_OK;  // no leaf was updated: the reference neither queries nor purges
    // 1. apply pending boxes; every occupied slot is attached afterwards
    k_bp_apply<<<(slots + 255) / 256, 256, 0, s>>>(slots, bp->box_lo.p, bp->box_hi.p, bp->pend_lo.p, bp->pend_hi.p, bp->pend_seq.p, bp->upd_seq.p,
                                                   bp->d_attached.p);
    CKB(cudaGetLastError());
    if (bp->slab_dirty) {
        std::vector<uint32_t> alive;
        alive.reserve(bp->len);
        for (uint32_t h = 0; h < slots; ++h)
            if (bp->next_free[h] == -2) {
                bp->attached[h] = 1;
                alive.push_back(h);
            }
        bp->n_attached = (uint32_t)alive.size();
        CKB(bp->alive.reserve(alive.size() + 1));
        if (!alive.empty()) CKB(cudaMemcpyAsync(bp->alive.p, alive.data(), 4 * alive.size(), cudaMemcpyHostToDevice, s));
        CKB(cudaStreamSynchronize(s));  // `alive` is a local
        bp->slab_dirty = false;
    }
    mark("apply+alive");
    bp->seq = bp->front = 0;
    uint32_t m = bp->n_attached;
    uint32_t n_new = 0;
    bp->tree_n = 0;
    if (m == 1) {  // a single leaf: no internal node; keep it where the queries look for it
        CKB(w->leaf_lo.reserve(1));
        CKB(w->leaf_hi.reserve(1));
        k_bp_gather<<<1, 32, 0, s>>>(bp->alive.p, 1, bp->box_lo.p, bp->box_hi.p, w->leaf_lo.p, w->leaf_hi.p);
        CKB(cudaMemcpyAsync(reinterpret_cast<uint32_t*>(w->leaf_lo.p) + 3, bp->alive.p, 4, cudaMemcpyDeviceToDevice, s));
        CKB(cudaGetLastError());
        bp->tree_n = 1;
        bp->tree_outliers = 1;  // scanned linearly
    }
    if (m >= 2) {
        // 2. LBVH over the attached proxies, leaf ids = handles
        CKB(w->aabb_lo.reserve(m));
        CKB(w->aabb_hi.reserve(m));
        CKB(w->keys_a.reserve(m));
        CKB(w->keys_b.reserve(m));
        CKB(w->idx_a.reserve(m));
        CKB(w->idx_b.reserve(m));
        CKB(w->leaf_lo.reserve(m));
        CKB(w->leaf_hi.reserve(m));
        CKB(w->nodes.reserve(4 * (size_t)m));
        CKB(w->parent.reserve(2 * (size_t)m));
        CKB(w->flags.reserve(m));
        CKB(w->cub_tmp.reserve(lbvh_temp_bytes(m) + 256));
        CKB(w->counters.reserve(1));
        if (!w->h_counters) CKB(cudaMallocHost((void**)&w->h_counters, sizeof(DevCounters)));
        const uint32_t* dgroups = d_groups;
        k_bp_gather<<<(m + 255) / 256, 256, 0, s>>>(bp->alive.p, m, bp->box_lo.p, bp->box_hi.p, w->aabb_lo.p, w->aabb_hi.p);
        CKB(cudaGetLastError());
        size_t cap_pairs = w->pairs_raw.cap ? w->pairs_raw.cap : (size_t)8 * m + 1024;
        for (int attempt = 0; attempt < 3; ++attempt) {
            CKB(w->pairs_raw.reserve(cap_pairs));
            CKB(w->keys_raw.reserve(w->pairs_raw.cap));
            DevCounters z;
            memset(&z, 0, sizeof z);
            for (int k = 0; k < 3; ++k) z.bounds[k] = 0x7f7fffff, z.bounds[3 + k] = (int)0x80800000;
            *w->h_counters = z;
            CKB(cudaMemcpyAsync(w->counters.p, w->h_counters, sizeof z, cudaMemcpyHostToDevice, s));
            CKB(launch_lbvh_build(w, m, bp->alive.p));
            CKB(launch_pair_search(w, m, dgroups, 0, m, (uint32_t)w->pairs_raw.cap));
            CKB(cudaMemcpyAsync(w->h_counters, w->counters.p, sizeof(DevCounters), cudaMemcpyDeviceToHost, s));
            CKB(cudaStreamSynchronize(s));
            n_new = w->h_counters->n_pairs;
            bp->tree_n = m;
            bp->tree_outliers = w->h_counters->n_outliers;
            if (n_new <= w->pairs_raw.cap) break;
            cap_pairs = (size_t)n_new + n_new / 8 + 1024;
        }
        mark("lbvh+pairs");
        // 3. sorted 64-bit keys
        if (n_new) {
            CKB(bp->keys_tmp.reserve(n_new));
            CKB(bp->keys_new.reserve(n_new));
            k_bp_keys<<<(n_new + 255) / 256, 256, 0, s>>>(w->pairs_raw.p, n_new, bp->keys_tmp.p);
            CKB(cudaGetLastError());
            int r = bp_sort_keys(bp, bp->keys_tmp.p, bp->keys_new.p, n_new);
            if (r) return r;
        }
    }
    mark("sort_keys");
    if (!want_events) {  // the stepping world diffs the sorted key lists itself (sim.cu)
        std::swap(bp->keys_old, bp->keys_new);
        bp->n_old = n_new;
        return NCB_OK;
    }
    // 4. started = new \ old, stopped = old \ new
    CKB(bp->counters.reserve(4));
    CKB(cudaMemsetAsync(bp->counters.p, 0, 16, s));
    if (n_new) {
        CKB(bp->ev_a.reserve(n_new));
        k_bp_diff<<<(n_new + 255) / 256, 256, 0, s>>>(bp->keys_new.p, n_new, bp->keys_old.p, bp->n_old, 1, bp->upd_seq.p, bp->ev_a.p, bp->counters.p);
    }
    if (bp->n_old) {
        CKB(bp->ev_b.reserve(bp->n_old));
        k_bp_diff<<<(bp->n_old + 255) / 256, 256, 0, s>>>(bp->keys_old.p, bp->n_old, bp->keys_new.p, n_new, 0, bp->upd_seq.p, bp->ev_b.p,
                                                          bp->counters.p + 1);
    }
    CKB(cudaGetLastError());
    uint32_t cnt[2] = {0, 0};
    CKB(cudaMemcpyAsync(cnt, bp->counters.p, 8, cudaMemcpyDeviceToHost, s));
    CKB(cudaStreamSynchronize(s));
    int r = bp_sort_events(bp, bp->ev_a, cnt[0]);
    if (r) return r;
    r = bp_sort_events(bp, bp->ev_b, cnt[1]);
    if (r) return r;
    mark("diff+events");
    std::swap(bp->keys_old, bp->keys_new);
    bp->n_old = n_new;
    bp->n_started = cnt[0];
    bp->n_stopped = cnt[1];
    if (n_started) *n_started = cnt[0];
    if (n_stopped) *n_stopped = cnt[1];
    return NCB_OK;
}

// ---- device-side staging for the stepping world: boxes of ALL objects as float4 arrays indexed by handle ------------
namespace {
__global__ void k_bp_stage_create_f4(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, uint32_t seq0, float4* pend_lo,
                                     float4* pend_hi, uint32_t* pend_seq, uint32_t* d_attached) {
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    float4 a = lo[h], b = hi[h];
    pend_lo[h] = make_float4(a.x, a.y, a.z, 0.f);
    pend_hi[h] = make_float4(b.x, b.y, b.z, 0.f);
    pend_seq[h] = min(pend_seq[h], seq0 + h);
    d_attached[h] = ST_DETACHED;
}
// deferred_set_bounding_volume for every object with a set `moved` flag, in handle order (glue/update.rs:77-83)
__global__ void k_bp_stage_set_f4(const float4* __restrict__ lo, const float4* __restrict__ hi, const uint8_t* __restrict__ moved, uint32_t n,
                                  uint32_t seq0, float margin, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi,
                                  const uint32_t* __restrict__ d_attached, float4* pend_lo, float4* pend_hi, uint32_t* pend_seq) {
    uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n || (moved && !moved[h])) return;
    float4 a = lo[h], b = hi[h];
    if (d_attached[h] == ST_ATTACHED) {
        float4 slo = box_lo[h], shi = box_hi[h];
        bool contains = slo.x <= a.x && slo.y <= a.y && slo.z <= a.z && shi.x >= b.x && shi.y >= b.y && shi.z >= b.z;
        if (contains) return;
    }
    pend_lo[h] = make_float4(a.x + (-margin), a.y + (-margin), a.z + (-margin), 0.f);
    pend_hi[h] = make_float4(b.x + margin, b.y + margin, b.z + margin, 0.f);
    pend_seq[h] = min(pend_seq[h], seq0 + h);
}
}  // namespace

int bp_create_all_device(ncb_bp* bp, uint32_t n, const float4* lo, const float4* hi) {
    if (!bp->next_free.empty()) {
        bp->err = bp->owner->err = "bp_create_all_device: the broad phase must be empty";
        return NCB_ERR_STATE;
    }
    CKB(cudaSetDevice(bp->owner->device));
    bp->next_free.assign(n, -2);
    bp->attached.assign(n, 0);
    bp->next = n, bp->len = n;
    bp->slab_dirty = true;
    int r = bp_grow(bp, n);
    if (r) return r;
    if (n == 0) return NCB_OK;
    k_bp_stage_create_f4<<<(n + 255) / 256, 256, 0, bp->owner->stream>>>(lo, hi, n, SEQ_BACK0 + bp->seq, bp->pend_lo.p, bp->pend_hi.p, bp->pend_seq.p,
                                                                         bp->d_attached.p);
    CKB(cudaGetLastError());
    bp->seq += n;
    return NCB_OK;
}

namespace {
__global__ void k_bp_stage_create_listed(const uint32_t* __restrict__ handles, uint32_t m, const float4* __restrict__ lo, const float4* __restrict__ hi,
                                         uint32_t seq0, float4* pend_lo, float4* pend_hi, uint32_t* pend_seq, uint32_t* d_attached) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t h = handles[k];
    float4 a = lo[h], b = hi[h];
    pend_lo[h] = make_float4(a.x, a.y, a.z, 0.f);
    pend_hi[h] = make_float4(b.x, b.y, b.z, 0.f);
    pend_seq[h] = min(pend_seq[h], seq0 + k);
    d_attached[h] = ST_DETACHED;
}
}  // namespace

// create_proxy for m new objects whose handles the caller predicted from its own slab: the proxy slab must hand out the same
// handles (both recycle last-freed-first); boxes are read from the float4 arrays indexed by handle.
int bp_create_listed_device(ncb_bp* bp, uint32_t m, const uint32_t* handles_host, const uint32_t* handles_dev, const float4* lo, const float4* hi) {
    CKB(cudaSetDevice(bp->owner->device));
    for (uint32_t i = 0; i < m; ++i) {  // Slab::insert
        size_t key = bp->next;
        if (key == bp->next_free.size()) {
            bp->next_free.push_back(-2);
            bp->attached.push_back(0);
            bp->next = key + 1;
        } else {
            bp->next = (size_t)bp->next_free[key];
            bp->next_free[key] = -2;
            bp->attached[key] = 0;
        }
        bp->len++;
        if (key != handles_host[i]) {
            bp->err = bp->owner->err = "bp_create_listed_device: proxy slab and object slab are out of step";
            return NCB_ERR_STATE;
        }
    }
    bp->slab_dirty = true;
    int r = bp_grow(bp, bp->next_free.size());
    if (r) return r;
    k_bp_stage_create_listed<<<(m + 255) / 256, 256, 0, bp->owner->stream>>>(handles_dev, m, lo, hi, SEQ_BACK0 + bp->seq, bp->pend_lo.p, bp->pend_hi.p,
                                                                            bp->pend_seq.p, bp->d_attached.p);
    CKB(cudaGetLastError());
    bp->seq += m;
    return NCB_OK;
}

int bp_set_moved_device(ncb_bp* bp, uint32_t n, const float4* lo, const float4* hi, const uint8_t* moved) {
    CKB(cudaSetDevice(bp->owner->device));
    if (n == 0) return NCB_OK;
    k_bp_stage_set_f4<<<(n + 255) / 256, 256, 0, bp->owner->stream>>>(lo, hi, moved, n, SEQ_BACK0 + bp->seq, bp->margin, bp->box_lo.p, bp->box_hi.p,
                                                                      bp->d_attached.p, bp->pend_lo.p, bp->pend_hi.p, bp->pend_seq.p);
    CKB(cudaGetLastError());
    bp->seq += n;
    return NCB_OK;
}

extern "C" {

static int bp_fetch(ncb_bp* bp, const unsigned long long* ev, uint32_t n, uint32_t* out) {
    if (!n || !out) return NCB_OK;
    cudaStream_t s = bp->owner->stream;
    CKB(bp->ev_u32.reserve(2 * (size_t)n));
    k_bp_unpack<<<(n + 255) / 256, 256, 0, s>>>(ev, n, bp->ev_u32.p);
    CKB(cudaGetLastError());
    CKB(cudaMemcpyAsync(out, bp->ev_u32.p, 8 * (size_t)n, cudaMemcpyDeviceToHost, s));
    CKB(cudaStreamSynchronize(s));
    return NCB_OK;
}

// Events of the last update() / remove(): started[2 * n_started] in interference_started argument order,
// stopped[2 * n_stopped] in interference_stopped order.  Either pointer may be NULL.
int ncb_bp_events(ncb_bp* bp, uint32_t* started, uint32_t* stopped) {
    if (!bp) return NCB_ERR_ARG;
    CKB(cudaSetDevice(bp->owner->device));
    int r = bp_fetch(bp, bp->ev_a.p, bp->n_started, started);
    if (r) return r;
    return bp_fetch(bp, bp->ev_b.p, bp->n_stopped, stopped);
}

// Batched BroadPhase::interferences_with_bounding_volume (kind 0, 6 floats per query: mins, maxs), _with_ray (kind 1,
// 7 floats: origin, dir, max_toi) and _with_point (kind 2, 3 floats) (:388-432).  out[2 * k] = (query index, handle),
// sorted; cap in entries; *n_out = entries found; returns 1 when truncated.
int ncb_bp_query(ncb_bp* bp, int kind, uint32_t n_queries, const float* queries, uint32_t* out, uint32_t cap, uint32_t* n_out) {
    if (!bp || kind < 0 || kind > 2 || (n_queries && !queries)) return NCB_ERR_ARG;
    CKB(cudaSetDevice(bp->owner->device));
    if (n_out) *n_out = 0;
    if (n_queries == 0 || bp->tree_n == 0) return NCB_OK;
    cudaStream_t s = bp->owner->stream;
    ncb_ctx* w = bp->work;
    const int W = kind == 0 ? 6 : (kind == 1 ? 7 : 3);
    CKB(bp->q_in.reserve((size_t)W * n_queries));
    CKB(cudaMemcpyAsync(bp->q_in.p, queries, 4 * (size_t)W * n_queries, cudaMemcpyHostToDevice, s));
    CKB(bp->counters.reserve(4));
    size_t want = bp->ev_a.cap ? bp->ev_a.cap : (size_t)16 * n_queries + 1024;
    uint32_t found = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        CKB(bp->ev_a.reserve(want));
        CKB(cudaMemsetAsync(bp->counters.p, 0, 16, s));
        unsigned g = (n_queries + 127) / 128;
        uint32_t c = (uint32_t)bp->ev_a.cap;
        if (kind == 0)
            k_bp_query<0><<<g, 128, 0, s>>>(bp->q_in.p, n_queries, w->leaf_lo.p, w->leaf_hi.p, w->nodes.p, bp->tree_n, bp->tree_outliers,
                                            bp->d_attached.p, bp->ev_a.p, c, bp->counters.p, trav_overflow_counter(bp->owner));
        else if (kind == 1)
            k_bp_query<1><<<g, 128, 0, s>>>(bp->q_in.p, n_queries, w->leaf_lo.p, w->leaf_hi.p, w->nodes.p, bp->tree_n, bp->tree_outliers,
                                            bp->d_attached.p, bp->ev_a.p, c, bp->counters.p, trav_overflow_counter(bp->owner));
        else
            k_bp_query<2><<<g, 128, 0, s>>>(bp->q_in.p, n_queries, w->leaf_lo.p, w->leaf_hi.p, w->nodes.p, bp->tree_n, bp->tree_outliers,
                                            bp->d_attached.p, bp->ev_a.p, c, bp->counters.p, trav_overflow_counter(bp->owner));
        CKB(cudaGetLastError());
        CKB(cudaMemcpyAsync(&found, bp->counters.p, 4, cudaMemcpyDeviceToHost, s));
        CKB(cudaStreamSynchronize(s));
        if (found <= bp->ev_a.cap) break;
        want = (size_t)found + 1024;
    }
    bp->n_started = 0;  // the event buffer was reused
    int r = bp_sort_events(bp, bp->ev_a, found);
    if (r) return r;
    if (n_out) *n_out = found;
    uint32_t wr = found < cap ? found : cap;
    r = bp_fetch(bp, bp->ev_a.p, wr, out);
    if (r) return r;
    return (out && found > cap) ? 1 : NCB_OK;
}

int ncb_bp_num_interferences(ncb_bp* bp, uint32_t* n) {
    if (!bp || !n) return NCB_ERR_ARG;
    *n = bp->n_old;
    return NCB_OK;
}

// The current interference set, sorted, (smaller handle, larger handle) per pair; cap in pairs.
int ncb_bp_pairs(ncb_bp* bp, uint32_t* pairs, uint32_t cap, uint32_t* n) {
    if (!bp) return NCB_ERR_ARG;
    CKB(cudaSetDevice(bp->owner->device));
    if (n) *n = bp->n_old;
    uint32_t w = bp->n_old < cap ? bp->n_old : cap;
    int r = bp_fetch(bp, bp->keys_old.p, w, pairs);
    if (r) return r;
    return bp->n_old > cap && pairs ? 1 : NCB_OK;
}

// BroadPhase::proxy (:262-273): 1 and the stored box when the proxy is attached, 0 otherwise
int ncb_bp_proxy(ncb_bp* bp, uint32_t handle, float* minmax) {
    if (!bp || !minmax) return NCB_ERR_ARG;
    if (handle >= bp->next_free.size() || bp->next_free[handle] != -2 || !bp->attached[handle]) return 0;
    CKB(cudaSetDevice(bp->owner->device));
    float4 lo, hi;
    CKB(cudaMemcpyAsync(&lo, bp->box_lo.p + handle, 16, cudaMemcpyDeviceToHost, bp->owner->stream));
    CKB(cudaMemcpyAsync(&hi, bp->box_hi.p + handle, 16, cudaMemcpyDeviceToHost, bp->owner->stream));
    CKB(cudaStreamSynchronize(bp->owner->stream));
    minmax[0] = lo.x, minmax[1] = lo.y, minmax[2] = lo.z, minmax[3] = hi.x, minmax[4] = hi.y, minmax[5] = hi.z;
    return 1;
}

}  // extern "C"
