// Stepping world: CollisionWorld::update over several steps with the reference's temporal coherence.
//
// Replaces (reference, file:line): pipeline/world.rs:104-119 (update), pipeline/glue/update.rs:65-138 (perform_broad_phase /
// perform_narrow_phase), pipeline/narrow_phase/narrow_phase.rs:56-104 (update_contact: save_cache_and_clear, generate,
// ContactEvents), :168-197 (update: only pairs with a changed object), :200-278 (handle_interaction: edges created /
// removed by the broad phase callbacks, in the callback's argument order),
// contact_generator/convex_polyhedron_convex_polyhedron_manifold_generator.rs:25-33,98,106,139 (last_gjk_dir),
// query/contact/contact_manifold.rs:134-236 (manifold cache), pipeline/object/collision_object.rs:215-222 (set_position flags).
//
// State per interference pair lives in a slot (stable while the pair exists): orientation (h1, h2), dispatcher key,
// last_gjk_dir, persistent manifold.  The sorted pair list of the persistent broad phase (bp_persistent.cu) maps to slots.
#include <cub/cub.cuh>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "bp_internal.h"

using namespace ncb;

struct ncb_sim {
    ncb_ctx* ctx = nullptr;
    ncb_bp* bp = nullptr;
    float margin = 0.f;
    uint32_t n = 0;
    bool first = true;
    std::vector<uint8_t> alive;          // host mirror of the object slab
    std::vector<uint32_t> free_handles;  // vacant object handles, reused last-freed-first (CollisionObjectSlab)
    std::vector<uint32_t> groups_changed;  // objects with COLLISION_GROUPS_CHANGED since the last update (collision_object.rs:10-25)
    DevBuf<uint8_t> moved;
    DevBuf<uint8_t> sel_flags;
    DevBuf<unsigned long long> keys_tmp;
    DevBuf<uint32_t> slot_tmp, sel_count;
    DevBuf<uint32_t> stage_h;
    DevBuf<float> stage_p, stage_r;
    // pair table, sorted-key order
    DevBuf<unsigned long long> keys_prev;
    uint32_t n_prev = 0;
    DevBuf<uint32_t> slot_prev, slot_new, raw_slot;
    // slot-indexed state
    size_t slot_cap = 0;
    uint32_t next_slot_bound = 0;  // host copy of the bump allocator
    uint32_t n_free_host = 0;      // host copy of the free-list length
    DevBuf<uint2> slot_pair;
    DevBuf<uint8_t> slot_key;
    DevBuf<float4> slot_dir;             // last_gjk_dir of a contact pair / sep_axis of a proximity pair (xyz + valid flag)
    DevBuf<uint8_t> slot_prox;           // Proximity status of a proximity pair (Interaction::Proximity(_, status))
    DevBuf<uint32_t> pm_hdr;
    DevBuf<float4> pm_entry;
    DevBuf<uint32_t> free_slots;
    DevBuf<uint32_t> cnt;  // [0] n_free (signed) [1] next_slot [2] n_events [3] pm_overflow [4] n_prox_events
    DevBuf<uint4> prox_events;  // ProximityEvent rows (h1, h2, prev, new) of the last step
    DevBuf<uint8_t> exp_prox;
    uint32_t n_prox_events = 0;
    DevBuf<unsigned long long> events, events_sorted;
    DevBuf<uint8_t> cub_tmp;
    // export of the last step
    DevBuf<uint32_t> exp_count, exp_start, exp_ids, exp_events;
    DevBuf<ncb_contact> exp_contacts;
    DevBuf<uint2> exp_pairs;
    DevBuf<uint8_t> exp_algo, exp_mcount;
    uint32_t n_pairs = 0, n_contacts = 0, n_events = 0, n_active = 0, pm_overflow = 0;
    ncb_update_counts last = {};
    WorldQueryBufs qbufs;
};

#define CKS(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            char b__[512];                                                                                \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            sim->ctx->err = b__;                                                                          \
            return NCB_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

namespace {

__device__ __forceinline__ bool is_prox_key(uint8_t k) { return k >= K_PROX_BALL_BALL && k <= K_PROX_SM_HULL; }

__device__ int sim_find(const unsigned long long* __restrict__ keys, uint32_t n, unsigned long long k) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (keys[mid] < k)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (lo < n && keys[lo] == k) ? (int)lo : -1;
}
__device__ int pm_live_count(const uint32_t* __restrict__ pm_hdr, const float4* __restrict__ pm_entry, uint32_t slot) {
    int n0 = (int)(pm_hdr[(size_t)slot * PM_HDR_WORDS] & 0xffu), live = 0;
    const float4* e = pm_entry + (size_t)slot * PM_CAP * PM_ENTRY_F4;
    for (int i = 0; i < n0; ++i) live += (__float_as_uint(e[4 * i + 3].w) >> 8) & 1u;
    return live;
}

// CollisionObject::set_position for a batch (collision_object.rs:215-222)
__global__ void k_sim_scatter_poses(const uint32_t* __restrict__ handles, const float* __restrict__ pos, const float* __restrict__ rot, uint32_t m,
                                    float* dpos, float4* drot, uint8_t* moved) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t h = handles ? handles[k] : k;
    dpos[3 * (size_t)h] = pos[3 * (size_t)k], dpos[3 * (size_t)h + 1] = pos[3 * (size_t)k + 1], dpos[3 * (size_t)h + 2] = pos[3 * (size_t)k + 2];
    drot[h] = make_float4(rot[4 * (size_t)k], rot[4 * (size_t)k + 1], rot[4 * (size_t)k + 2], rot[4 * (size_t)k + 3]);
    moved[h] = 1;
}
// interference_stopped -> handle_interaction(.., false) (narrow_phase.rs:248-277): the edge goes, Stopped if it had contacts
__global__ void k_sim_release(const unsigned long long* __restrict__ prev, uint32_t n_prev, const unsigned long long* __restrict__ cur, uint32_t n_cur,
                              const uint32_t* __restrict__ slot_prev, const uint2* __restrict__ slot_pair, const uint32_t* __restrict__ pm_hdr,
                              const float4* __restrict__ pm_entry, uint32_t* free_slots, uint32_t* cnt, unsigned long long* events, uint32_t cap_events,
                              const uint8_t* __restrict__ slot_key, const uint8_t* __restrict__ slot_prox, uint4* prox_events) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_prev) return;
    if (sim_find(cur, n_cur, prev[j]) >= 0) return;
    uint32_t slot = slot_prev[j];
    if (is_prox_key(slot_key[slot])) {
        // a proximity edge goes: ProximityEvent(h1, h2, prev, Disjoint) unless it was Disjoint already (narrow_phase.rs:266-274)
        uint8_t st = slot_prox[slot];
        if (st != NCB_PROXIMITY_DISJOINT) {
            uint2 pr = slot_pair[slot];
            uint32_t k = atomicAdd(&cnt[4], 1u);
            if (k < cap_events) prox_events[k] = make_uint4(pr.x, pr.y, st, NCB_PROXIMITY_DISJOINT);
        }
    } else if (pm_live_count(pm_hdr, pm_entry, slot) > 0) {
        uint2 pr = slot_pair[slot];
        uint32_t k = atomicAdd(&cnt[2], 1u);
        if (k < cap_events) events[k] = ((unsigned long long)pr.x << 32) | pr.y;
    }
    free_slots[atomicAdd(&cnt[0], 1u)] = slot;
}
// interference_started -> handle_interaction(.., true) (:216-247): a new edge in the callback's argument order, fresh generator
__global__ void k_sim_assign(const unsigned long long* __restrict__ cur, uint32_t n_cur, const unsigned long long* __restrict__ prev, uint32_t n_prev,
                             const uint32_t* __restrict__ slot_prev, uint32_t* slot_new, const uint32_t* __restrict__ upd_seq,
                             const uint32_t* __restrict__ type, const uint32_t* __restrict__ free_slots, uint32_t* cnt, uint2* slot_pair,
                             uint8_t* slot_key, float4* slot_dir, uint32_t* pm_hdr, const uint8_t* __restrict__ qkind, uint8_t* slot_prox) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cur) return;
    unsigned long long k = cur[i];
    int j = sim_find(prev, n_prev, k);
    if (j >= 0) {
        slot_new[i] = slot_prev[j];
        return;
    }
    int idx = atomicAdd(reinterpret_cast<int*>(&cnt[0]), -1) - 1;
    uint32_t slot = idx >= 0 ? free_slots[idx] : atomicAdd(&cnt[1], 1u);
    uint32_t lo = (uint32_t)(k >> 32), hi = (uint32_t)k;
    uint32_t sl = upd_seq[lo], sh = upd_seq[hi];
    bool hi_first = (sh != SEQ_NONE) && (sl == SEQ_NONE || sh > sl);
    uint32_t h1 = hi_first ? hi : lo, h2 = hi_first ? lo : hi;
    slot_new[i] = slot;
    slot_pair[slot] = make_uint2(h1, h2);
    uint32_t t1 = type[h1] & NCB_TYPE_MASK, t2 = type[h2] & NCB_TYPE_MASK;
    uint8_t key = (uint8_t)pair_key(t1, t2);
    if (qkind && key != K_NONE && (qkind[h1] | qkind[h2])) {  // (_, Proximity) | (Proximity, _): a proximity detector (narrow_phase.rs:240-246)
        key = (t1 == NCB_SHAPE_BALL && t2 == NCB_SHAPE_BALL) ? K_PROX_BALL_BALL
              : (t1 == NCB_SHAPE_PLANE || t2 == NCB_SHAPE_PLANE) ? K_PROX_PLANE
              : (t1 == NCB_SHAPE_CONVEX_HULL || t2 == NCB_SHAPE_CONVEX_HULL) ? K_PROX_SM_HULL
                                                                             : K_PROX_SM;
        slot_prox[slot] = NCB_PROXIMITY_DISJOINT;
    }
    slot_key[slot] = key;
    slot_dir[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < PM_HDR_WORDS; ++w) pm_hdr[(size_t)slot * PM_HDR_WORDS + w] = 0;
}
__global__ void k_sim_fix_free(uint32_t* cnt) {
    if ((int)cnt[0] < 0) cnt[0] = 0;
}
// NarrowPhase::update (:168-197): only the edges with a changed endpoint are regenerated
__global__ void k_sim_active(const uint32_t* __restrict__ slot_new, uint32_t n_cur, const uint2* __restrict__ slot_pair, const uint8_t* __restrict__ slot_key,
                             const uint8_t* __restrict__ moved, uint2* pairs_raw, uint8_t* keys_raw, uint32_t* raw_slot, DevCounters* cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cur) return;
    uint32_t slot = slot_new[i];
    uint2 pr = slot_pair[slot];
    uint8_t key = slot_key[slot];
    if (key == K_NONE || !(moved[pr.x] | moved[pr.y])) return;
    uint32_t k = atomicAdd(&cnt->n_pairs, 1u);
    pairs_raw[k] = pr;
    keys_raw[k] = key;
    raw_slot[k] = slot;
}
__global__ void k_sim_compose(uint32_t* pair_index, const uint32_t* __restrict__ raw_slot, const DevCounters* __restrict__ cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt->n_pairs) return;
    pair_index[i] = raw_slot[pair_index[i]];
}
__global__ void k_sim_count(const uint32_t* __restrict__ slot_new, uint32_t n_cur, const uint32_t* __restrict__ pm_hdr, const float4* __restrict__ pm_entry,
                            uint32_t* count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cur) return;
    count[i] = (uint32_t)pm_live_count(pm_hdr, pm_entry, slot_new[i]);
}
// ContactManifold::contacts(): slab order, live entries (contact_manifold.rs:59-68)
__global__ void k_sim_export(const uint32_t* __restrict__ slot_new, uint32_t n_cur, const uint2* __restrict__ slot_pair, const uint8_t* __restrict__ slot_key,
                             const uint32_t* __restrict__ pm_hdr, const float4* __restrict__ pm_entry, const uint32_t* __restrict__ start,
                             uint2* out_pairs, uint8_t* out_algo, uint8_t* out_count, ncb_contact* contacts, uint32_t* ids) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cur) return;
    uint32_t slot = slot_new[i];
    out_pairs[i] = slot_pair[slot];
    out_algo[i] = (uint8_t)algo_of_key(slot_key[slot]);
    int n0 = (int)(pm_hdr[(size_t)slot * PM_HDR_WORDS] & 0xffu);
    const float4* e = pm_entry + (size_t)slot * PM_CAP * PM_ENTRY_F4;
    uint32_t dst = start[i], done = 0;
    int written = 0;
    for (;;) {  // selection by increasing slab slot
        int best = -1;
        uint32_t best_slot = 0xffffffffu;
        for (int k = 0; k < n0; ++k) {
            uint32_t meta = __float_as_uint(e[4 * k + 3].w);
            if (!((meta >> 8) & 1u) || ((done >> k) & 1u)) continue;
            if ((meta & 0xffu) < best_slot) best_slot = meta & 0xffu, best = k;
        }
        if (best < 0) break;
        done |= 1u << best;
        float4 a = e[4 * best], b = e[4 * best + 1], c = e[4 * best + 2];
        uint32_t meta = __float_as_uint(e[4 * best + 3].w);
        ncb_contact o;
        o.world1[0] = a.x, o.world1[1] = a.y, o.world1[2] = a.z;
        o.world2[0] = b.x, o.world2[1] = b.y, o.world2[2] = b.z;
        o.normal[0] = c.x, o.normal[1] = c.y, o.normal[2] = c.z;
        o.depth = a.w;
        o.f1 = __float_as_uint(b.w), o.f2 = __float_as_uint(c.w);
        o.pair = i;
        contacts[dst] = o;
        ids[dst] = ((meta >> 9) << 8) | (meta & 0xffu);
        dst++, written++;
    }
    out_count[i] = (uint8_t)written;
}
__global__ void k_sim_unpack_events(const unsigned long long* __restrict__ ev, uint32_t n, uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = ev[i];
    out[3 * i] = (uint32_t)(k >> 32) & 0x7fffffffu;
    out[3 * i + 1] = (uint32_t)k;
    out[3 * i + 2] = (uint32_t)(k >> 63);
}

__global__ void k_sim_export_prox(const uint32_t* __restrict__ slot_new, uint32_t n_cur, const uint8_t* __restrict__ slot_key,
                                  const uint8_t* __restrict__ slot_prox, uint8_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cur) return;
    uint32_t slot = slot_new[i];
    out[i] = is_prox_key(slot_key[slot]) ? slot_prox[slot] : (uint8_t)NCB_PROXIMITY_NONE;
}

// ---- CollisionWorld::remove / add between updates ------------------------------------------------------------------------
// pairs of a removed object leave the pair table silently (glue/setup.rs:56-58): flag them, release their slots
__global__ void k_sim_flag_removed(const unsigned long long* __restrict__ keys, uint32_t n, const uint32_t* __restrict__ d_attached,
                                   const uint32_t* __restrict__ slots, uint8_t* keep, uint32_t* free_slots, uint32_t* cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = keys[i];
    bool gone = d_attached[(uint32_t)(k >> 32)] == ST_VACANT || d_attached[(uint32_t)k] == ST_VACANT;
    keep[i] = gone ? 0 : 1;
    if (gone) free_slots[atomicAdd(&cnt[0], 1u)] = slots[i];
}
__global__ void k_sim_clear_moved(const uint32_t* __restrict__ handles, uint32_t n, uint8_t* moved, uint8_t v) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) moved[handles[i]] = v;
}
// new objects: scatter their records into the object arrays
__global__ void k_sim_scatter_objects(const uint32_t* __restrict__ handles, uint32_t m, const float* __restrict__ pos, const float* __restrict__ rot,
                                      const uint32_t* __restrict__ type, const float* __restrict__ param, const uint32_t* __restrict__ groups,
                                      const float* __restrict__ qlimit, const float2* __restrict__ ang_cs, float* dpos, float4* drot, uint32_t* dtype,
                                      float4* dparam, uint32_t* dgroups, float* dqlimit, float2* dang_cs, uint8_t* moved) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t h = handles[k];
    for (int d = 0; d < 3; ++d) dpos[3 * (size_t)h + d] = pos[3 * (size_t)k + d];
    drot[h] = make_float4(rot[4 * (size_t)k], rot[4 * (size_t)k + 1], rot[4 * (size_t)k + 2], rot[4 * (size_t)k + 3]);
    dtype[h] = type[k];
    dparam[h] = make_float4(param[4 * (size_t)k], param[4 * (size_t)k + 1], param[4 * (size_t)k + 2], param[4 * (size_t)k + 3]);
    if (dgroups)
        for (int d = 0; d < 3; ++d) dgroups[3 * (size_t)h + d] = groups ? groups[3 * (size_t)k + d] : (d < 2 ? 0x3FFFFFFFu : 0u);
    dqlimit[h] = qlimit[k];
    if (dang_cs) dang_cs[h] = ang_cs[k];
    moved[h] = 1;
}
__global__ void k_sim_scatter_kinds(const uint32_t* __restrict__ handles, uint32_t m, const uint8_t* __restrict__ kinds, uint8_t* qkind) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) qkind[handles[k]] = kinds ? kinds[k] : 0;
}
__global__ void k_sim_fill_groups(uint32_t* g, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[3 * (size_t)i] = g[3 * (size_t)i + 1] = 0x3FFFFFFFu, g[3 * (size_t)i + 2] = 0u;
}
__global__ void k_sim_fill_f2(float2* p, float2 v, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <typename T>
cudaError_t grow_keep(DevBuf<T>& b, size_t want, size_t keep, cudaStream_t s) {
    if (want <= b.cap) return cudaSuccess;
    DevBuf<T> nb;
    cudaError_t e = nb.reserve(want + want / 2);
    if (e != cudaSuccess) return e;
    if (b.p && keep) e = cudaMemcpyAsync(nb.p, b.p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s);
    cudaStreamSynchronize(s);
    b.release();
    b = std::move(nb);
    return e;
}

}  // namespace

static int sim_grow_slots(ncb_sim* sim, size_t want) {
    if (want <= sim->slot_cap) return NCB_OK;
    cudaStream_t s = sim->ctx->stream;
    size_t keep = sim->slot_cap;
    CKS(grow_keep(sim->slot_pair, want, keep, s));
    CKS(grow_keep(sim->slot_key, want, keep, s));
    CKS(grow_keep(sim->slot_dir, want, keep, s));
    CKS(grow_keep(sim->slot_prox, want, keep, s));
    CKS(grow_keep(sim->pm_hdr, want * PM_HDR_WORDS, keep * PM_HDR_WORDS, s));
    CKS(grow_keep(sim->pm_entry, want * PM_CAP * PM_ENTRY_F4, keep * PM_CAP * PM_ENTRY_F4, s));
    CKS(grow_keep(sim->free_slots, want, keep, s));
    size_t cap = sim->slot_pair.cap;
    cap = std::min(cap, (size_t)sim->slot_key.cap);
    cap = std::min(cap, (size_t)sim->slot_dir.cap);
    cap = std::min(cap, (size_t)sim->slot_prox.cap);
    cap = std::min(cap, sim->pm_hdr.cap / PM_HDR_WORDS);
    cap = std::min(cap, sim->pm_entry.cap / (PM_CAP * PM_ENTRY_F4));
    cap = std::min(cap, (size_t)sim->free_slots.cap);
    sim->slot_cap = cap;
    return NCB_OK;
}

extern "C" {

int ncb_sim_create(ncb_ctx* ctx, float margin, ncb_sim** out) {
    if (!ctx || !out) return NCB_ERR_ARG;
    if (ctx->n == 0) {
        ctx->err = "ncb_sim_create: call ncb_set_objects first";
        return NCB_ERR_STATE;
    }
    ncb_sim* sim = new ncb_sim;
    sim->ctx = ctx;
    sim->margin = margin;
    sim->n = ctx->n;
    int r = ncb_bp_create(ctx, margin, &sim->bp);
    if (r) {
        delete sim;
        return r;
    }
    CKS(cudaSetDevice(ctx->device));
    sim->alive.assign(sim->n, 1);
    CKS(sim->moved.reserve(sim->n));
    CKS(cudaMemsetAsync(sim->moved.p, 1, sim->n, ctx->stream));  // a new object has every update flag set
    CKS(sim->cnt.reserve(8));
    CKS(cudaMemsetAsync(sim->cnt.p, 0, 32, ctx->stream));
    *out = sim;
    return NCB_OK;
}

void ncb_sim_destroy(ncb_sim* sim) {
    if (!sim) return;
    cudaSetDevice(sim->ctx->device);
    cudaStreamSynchronize(sim->ctx->stream);
    ncb_bp_destroy(sim->bp);
    sim->moved.release(), sim->stage_h.release(), sim->stage_p.release(), sim->stage_r.release();
    sim->keys_prev.release(), sim->slot_prev.release(), sim->slot_new.release(), sim->raw_slot.release();
    sim->slot_pair.release(), sim->slot_key.release(), sim->slot_dir.release(), sim->pm_hdr.release(), sim->pm_entry.release();
    sim->slot_prox.release(), sim->prox_events.release(), sim->exp_prox.release();
    sim->sel_flags.release(), sim->keys_tmp.release(), sim->slot_tmp.release(), sim->sel_count.release();
    sim->free_slots.release(), sim->cnt.release(), sim->events.release(), sim->events_sorted.release(), sim->cub_tmp.release();
    sim->exp_count.release(), sim->exp_start.release(), sim->exp_ids.release(), sim->exp_events.release(), sim->exp_contacts.release();
    sim->exp_pairs.release(), sim->exp_algo.release(), sim->exp_mcount.release();
    sim->qbufs.release();
    delete sim;
}

// CollisionObject::set_position for m objects (handles == NULL: objects 0..m-1); marks them for the next step.
int ncb_sim_set_positions(ncb_sim* sim, uint32_t m, const uint32_t* handles, const float* pos, const float* rot) {
    if (!sim || (m && (!pos || !rot))) return NCB_ERR_ARG;
    CKS(cudaSetDevice(sim->ctx->device));
    if (m == 0) return NCB_OK;
    if (m > sim->n && !handles) return NCB_ERR_ARG;
    if (handles)
        for (uint32_t k = 0; k < m; ++k)
            if (handles[k] >= sim->n || !sim->alive[handles[k]]) {
                sim->ctx->err = "ncb_sim_set_positions: unknown object handle";
                return NCB_ERR_ARG;
            }
    cudaStream_t s = sim->ctx->stream;
    CKS(sim->stage_p.reserve(3 * (size_t)m));
    CKS(sim->stage_r.reserve(4 * (size_t)m));
    CKS(cudaMemcpyAsync(sim->stage_p.p, pos, 12 * (size_t)m, cudaMemcpyHostToDevice, s));
    CKS(cudaMemcpyAsync(sim->stage_r.p, rot, 16 * (size_t)m, cudaMemcpyHostToDevice, s));
    const uint32_t* dh = nullptr;
    if (handles) {
        CKS(sim->stage_h.reserve(m));
        CKS(cudaMemcpyAsync(sim->stage_h.p, handles, 4 * (size_t)m, cudaMemcpyHostToDevice, s));
        dh = sim->stage_h.p;
    }
    k_sim_scatter_poses<<<(m + 255) / 256, 256, 0, s>>>(dh, sim->stage_p.p, sim->stage_r.p, m, sim->ctx->pos.p, sim->ctx->rot.p, sim->moved.p);
    CKS(cudaGetLastError());
    CKS(cudaStreamSynchronize(s));  // staging buffers are reused
    return NCB_OK;
}

// CollisionObject::set_collision_groups (collision_object.rs:246-250) on a batch of live objects: sets COLLISION_GROUPS_CHANGED.  The
// next update then asks the broad phase to recompute all proximities of the object (needs_broad_phase_redispatch, glue/update.rs:83-86:
// pairs that are no longer allowed stop, newly allowed ones start) and the narrow phase to update the object's pairs
// (needs_narrow_phase_update); the bounding volume is not touched.  groups: 3 words per handle (membership, whitelist, blacklist).
__global__ void k_sim_scatter_groups(const uint32_t* __restrict__ handles, const uint32_t* __restrict__ g, uint32_t m, uint32_t* dgroups,
                                     uint8_t* moved) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t h = handles[k];
    dgroups[3 * (size_t)h] = g[3 * k], dgroups[3 * (size_t)h + 1] = g[3 * k + 1], dgroups[3 * (size_t)h + 2] = g[3 * k + 2];
    moved[h] = 1;  // the object's pairs take part in the next narrow-phase update; its stored box stays (the same box is re-submitted)
}
int ncb_sim_set_collision_groups(ncb_sim* sim, uint32_t m, const uint32_t* handles, const uint32_t* groups) {
    if (!sim || (m && (!handles || !groups))) return NCB_ERR_ARG;
    ncb_ctx* ctx = sim->ctx;
    CKS(cudaSetDevice(ctx->device));
    if (m == 0) return NCB_OK;
    for (uint32_t k = 0; k < m; ++k)
        if (handles[k] >= sim->n || !sim->alive[handles[k]]) {
            ctx->err = "ncb_sim_set_collision_groups: unknown object handle";
            return NCB_ERR_ARG;
        }
    cudaStream_t s = ctx->stream;
    if (!ctx->has_groups) {  // the world used default groups so far: materialise them
        CKS(ctx->groups.reserve(3 * (size_t)sim->n));
        k_sim_fill_groups<<<(sim->n + 255) / 256, 256, 0, s>>>(ctx->groups.p, sim->n);
        ctx->has_groups = true;
    }
    CKS(sim->stage_h.reserve(m));
    CKS(sim->stage_p.reserve(3 * (size_t)m));
    CKS(cudaMemcpyAsync(sim->stage_h.p, handles, 4 * (size_t)m, cudaMemcpyHostToDevice, s));
    CKS(cudaMemcpyAsync(sim->stage_p.p, groups, 12 * (size_t)m, cudaMemcpyHostToDevice, s));
    k_sim_scatter_groups<<<(m + 255) / 256, 256, 0, s>>>(sim->stage_h.p, reinterpret_cast<const uint32_t*>(sim->stage_p.p), m, ctx->groups.p,
                                                         sim->moved.p);
    CKS(cudaGetLastError());
    CKS(cudaStreamSynchronize(s));  // staging buffers are reused
    sim->groups_changed.insert(sim->groups_changed.end(), handles, handles + m);
    return NCB_OK;
}

// CollisionWorld::update (world.rs:104-119)
int ncb_sim_step(ncb_sim* sim, ncb_update_counts* counts) {
    if (!sim) return NCB_ERR_ARG;
    ncb_ctx* ctx = sim->ctx;
    CKS(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    uint32_t n = sim->n;
    if (ctx->n != n) {
        ctx->err = "ncb_sim_step: the object set changed since ncb_sim_create";
        return NCB_ERR_STATE;
    }
    // NCB_SIM_PROFILE=1: wall time per phase (with a stream synchronisation at every phase boundary) on stderr
    static const bool prof = getenv("NCB_SIM_PROFILE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!prof) return;
        cudaStreamSynchronize(s);
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[sim] %-14s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    // ---- perform_broad_phase (glue/update.rs:65-99): swept AABB = shape AABB loosened by the query limit
    int r = reserve_broad(ctx, n);
    if (r) return r;
    DevObjects objs = dev_objects(ctx);
    CKS(launch_aabbs(ctx, objs, sim->margin, 1, 0, n));
    if (sim->first) {
        r = bp_create_all_device(sim->bp, n, ctx->aabb_lo.p, ctx->aabb_hi.p);  // CollisionWorld::add -> create_proxies
        if (r) return r;
    }
    r = bp_set_moved_device(sim->bp, n, ctx->aabb_lo.p, ctx->aabb_hi.p, sim->moved.p);
    if (r) return r;
    if (!sim->groups_changed.empty()) {
        // objects.foreach visits the objects in handle order and pushes each redispatch to the FRONT of the update queue
        std::sort(sim->groups_changed.begin(), sim->groups_changed.end());
        sim->groups_changed.erase(std::unique(sim->groups_changed.begin(), sim->groups_changed.end()), sim->groups_changed.end());
        std::vector<uint32_t> live;
        for (uint32_t h : sim->groups_changed)
            if (h < sim->n && sim->alive[h]) live.push_back(h);
        sim->groups_changed.clear();
        if (!live.empty()) {
            r = ncb_bp_recompute_with(sim->bp, (uint32_t)live.size(), live.data());
            if (r) return r;
        }
    }
    mark("aabb+stage");
    // no event lists wanted: started / stopped pairs are found below by diffing the sorted key lists against the pair table
    r = bp_update_impl(sim->bp, ctx->has_groups ? ctx->groups.p : nullptr, nullptr, nullptr, false);
    if (r) return r;
    mark("bp_update");
    uint32_t n_cur = sim->bp->n_old;
    const unsigned long long* cur = sim->bp->keys_old.p;
    // ---- interaction edges follow the started / stopped callbacks
    {
        // started - stopped == n_cur - n_prev, and the stopped pairs release their slots before the started ones take theirs
        uint32_t net = n_cur > sim->n_prev ? n_cur - sim->n_prev : 0;
        r = sim_grow_slots(sim, (size_t)sim->next_slot_bound + (net > sim->n_free_host ? net - sim->n_free_host : 0) + 16);
        if (r) return r;
    }
    uint32_t cap_events = sim->n_prev + n_cur + 16;
    CKS(sim->events.reserve(cap_events));
    CKS(sim->slot_new.reserve(n_cur + 1));
    CKS(sim->raw_slot.reserve(n_cur + 1));
    CKS(sim->prox_events.reserve(cap_events));
    CKS(cudaMemsetAsync(sim->cnt.p + 2, 0, 12, s));
    if (sim->n_prev)
        k_sim_release<<<(sim->n_prev + 255) / 256, 256, 0, s>>>(sim->keys_prev.p, sim->n_prev, cur, n_cur, sim->slot_prev.p, sim->slot_pair.p, sim->pm_hdr.p,
                                                                sim->pm_entry.p, sim->free_slots.p, sim->cnt.p, sim->events.p, cap_events, sim->slot_key.p,
                                                                sim->slot_prox.p, sim->prox_events.p);
    if (n_cur)
        k_sim_assign<<<(n_cur + 255) / 256, 256, 0, s>>>(cur, n_cur, sim->keys_prev.p, sim->n_prev, sim->slot_prev.p, sim->slot_new.p, sim->bp->upd_seq.p,
                                                         ctx->type.p, sim->free_slots.p, sim->cnt.p, sim->slot_pair.p, sim->slot_key.p, sim->slot_dir.p,
                                                         sim->pm_hdr.p, ctx->has_prox ? ctx->qkind.p : nullptr, sim->slot_prox.p);
    k_sim_fix_free<<<1, 1, 0, s>>>(sim->cnt.p);
    CKS(cudaGetLastError());
    CKS(sim->keys_prev.reserve(n_cur + 1));
    if (n_cur) CKS(cudaMemcpyAsync(sim->keys_prev.p, cur, 8 * (size_t)n_cur, cudaMemcpyDeviceToDevice, s));
    std::swap(sim->slot_prev, sim->slot_new);  // slot_prev now describes the current pair list
    sim->n_prev = n_cur;
    mark("pair_table");
    // ---- perform_narrow_phase: regenerate the edges with a changed endpoint
    size_t cap_pairs = (size_t)n_cur + 1024;
    r = reserve_pairs(ctx, cap_pairs);
    if (r) return r;
    CKS(ctx->pair_index.reserve(cap_pairs));
    r = reset_counters(ctx);
    if (r) return r;
    if (n_cur) {
        k_sim_active<<<(n_cur + 255) / 256, 256, 0, s>>>(sim->slot_prev.p, n_cur, sim->slot_pair.p, sim->slot_key.p, sim->moved.p, ctx->pairs_raw.p,
                                                         ctx->keys_raw.p, sim->raw_slot.p, ctx->counters.p);
        CKS(launch_pair_sort(ctx, (uint32_t)cap_pairs, ctx->pair_index.p));
        k_sim_compose<<<(n_cur + 255) / 256, 256, 0, s>>>(ctx->pair_index.p, sim->raw_slot.p, ctx->counters.p);
        PersistArgs ps;
        ps.dir = sim->slot_dir.p;
        ps.pm_hdr = sim->pm_hdr.p;
        ps.pm_entry = sim->pm_entry.p;
        ps.events = sim->events.p;
        ps.n_events = sim->cnt.p + 2;
        ps.cap_events = cap_events;
        ps.pm_overflow = sim->cnt.p + 3;
        CKS(launch_narrow_phase_persistent(ctx, objs, ctx->pairs.p, ctx->pair_index.p, (uint32_t)cap_pairs, ps));
        if (ctx->has_prox)  // update_proximity for the sensor pairs with a changed endpoint (their four key segments)
            CKS(launch_proximity_persistent(ctx, objs, ctx->pairs.p, ctx->pair_index.p, sim->slot_dir.p, sim->slot_prox.p, sim->prox_events.p, sim->cnt.p + 4,
                                            cap_events));
    }
    mark("narrow");
    // ---- export: pairs in sorted order, live contacts in slab order
    CKS(sim->exp_count.reserve(n_cur + 1));
    CKS(sim->exp_start.reserve(n_cur + 1));
    CKS(sim->exp_pairs.reserve(n_cur + 1));
    CKS(sim->exp_algo.reserve(n_cur + 1));
    CKS(sim->exp_mcount.reserve(n_cur + 1));
    CKS(sim->exp_prox.reserve(n_cur + 1));
    uint32_t total = 0;
    if (n_cur) {
        k_sim_count<<<(n_cur + 255) / 256, 256, 0, s>>>(sim->slot_prev.p, n_cur, sim->pm_hdr.p, sim->pm_entry.p, sim->exp_count.p);
        size_t bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, sim->exp_count.p, sim->exp_start.p, (int)n_cur);
        CKS(sim->cub_tmp.reserve(bytes + 256));
        bytes = sim->cub_tmp.cap;
        CKS(cub::DeviceScan::ExclusiveSum(sim->cub_tmp.p, bytes, sim->exp_count.p, sim->exp_start.p, (int)n_cur, s));
        uint32_t last_start = 0, last_count = 0;
        CKS(cudaMemcpyAsync(&last_start, sim->exp_start.p + (n_cur - 1), 4, cudaMemcpyDeviceToHost, s));
        CKS(cudaMemcpyAsync(&last_count, sim->exp_count.p + (n_cur - 1), 4, cudaMemcpyDeviceToHost, s));
        CKS(cudaStreamSynchronize(s));
        total = last_start + last_count;
        CKS(sim->exp_contacts.reserve(total + 1));
        CKS(sim->exp_ids.reserve(total + 1));
        k_sim_export<<<(n_cur + 255) / 256, 256, 0, s>>>(sim->slot_prev.p, n_cur, sim->slot_pair.p, sim->slot_key.p, sim->pm_hdr.p, sim->pm_entry.p,
                                                         sim->exp_start.p, sim->exp_pairs.p, sim->exp_algo.p, sim->exp_mcount.p, sim->exp_contacts.p,
                                                         sim->exp_ids.p);
        k_sim_export_prox<<<(n_cur + 255) / 256, 256, 0, s>>>(sim->slot_prev.p, n_cur, sim->slot_key.p, sim->slot_prox.p, sim->exp_prox.p);
        CKS(cudaGetLastError());
    }
    mark("export");
    // ---- counters, events (sorted for a deterministic order), flags cleared (world.rs:115-118)
    uint32_t hc[5] = {0, 0, 0, 0, 0};
    CKS(cudaMemcpyAsync(hc, sim->cnt.p, 20, cudaMemcpyDeviceToHost, s));
    r = read_counters(ctx);
    if (r) return r;
    sim->n_free_host = hc[0];
    sim->next_slot_bound = hc[1];
    sim->n_events = hc[2] < cap_events ? hc[2] : cap_events;
    sim->pm_overflow = hc[3];
    sim->n_prox_events = hc[4] < cap_events ? hc[4] : cap_events;
    if (sim->n_events > 1) {
        CKS(sim->events_sorted.reserve(sim->n_events));
        size_t bytes = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, bytes, sim->events.p, sim->events_sorted.p, (int)sim->n_events, 0, 64);
        CKS(sim->cub_tmp.reserve(bytes + 256));
        bytes = sim->cub_tmp.cap;
        CKS(cub::DeviceRadixSort::SortKeys(sim->cub_tmp.p, bytes, sim->events.p, sim->events_sorted.p, (int)sim->n_events, 0, 64, s));
        CKS(cudaMemcpyAsync(sim->events.p, sim->events_sorted.p, 8 * (size_t)sim->n_events, cudaMemcpyDeviceToDevice, s));
    }
    CKS(cudaMemsetAsync(sim->moved.p, 0, n, s));
    CKS(cudaStreamSynchronize(s));
    mark("finish");
    sim->first = false;
    sim->n_pairs = n_cur;
    sim->n_contacts = total;
    sim->n_active = ctx->last_counters.n_pairs;
    if (counts) {
        memset(counts, 0, sizeof *counts);
        counts->n_pairs = n_cur;
        counts->n_contacts = total;
        counts->epa_overflow = ctx->last_counters.epa_overflow + sim->pm_overflow;
        counts->ref_panics = ctx->last_counters.ref_panics;
        counts->n_epa_pairs = ctx->last_counters.epa_cursor[K_CUBOID_CUBOID] - ctx->last_counters.key_start[K_CUBOID_CUBOID];
        counts->n_manifold_jobs = sim->n_active;  // pairs regenerated in this step
        const DevCounters& lc = ctx->last_counters;
        counts->n_proximity_pairs = lc.key_hist[K_PROX_BALL_BALL] + lc.key_hist[K_PROX_PLANE] + lc.key_hist[K_PROX_SM] + lc.key_hist[K_PROX_SM_HULL];  // updated
    }
    return NCB_OK;
}

// CollisionWorld::remove (world.rs:129-144) between updates: the objects leave the world, their proxies leave the broad
// phase, their pairs disappear without events (glue/setup.rs:50-62); the handles are recycled last-freed-first by ncb_sim_add.
int ncb_sim_remove(ncb_sim* sim, uint32_t m, const uint32_t* handles) {
    if (!sim || (m && !handles)) return NCB_ERR_ARG;
    if (m == 0) return NCB_OK;
    ncb_ctx* ctx = sim->ctx;
    CKS(cudaSetDevice(ctx->device));
    if (sim->first) {
        ctx->err = "ncb_sim_remove: call ncb_sim_step once before removing objects";
        return NCB_ERR_STATE;
    }
    std::vector<uint8_t> seen(sim->n, 0);
    for (uint32_t k = 0; k < m; ++k) {
        if (handles[k] >= sim->n || !sim->alive[handles[k]] || seen[handles[k]]) {
            ctx->err = "ncb_sim_remove: unknown object handle";
            return NCB_ERR_ARG;
        }
        seen[handles[k]] = 1;
    }
    cudaStream_t s = ctx->stream;
    uint32_t nr = 0;
    int r = ncb_bp_remove(sim->bp, m, handles, &nr);  // same order => the proxy slab recycles in lockstep with the object slab
    if (r < 0) return r;
    for (uint32_t k = 0; k < m; ++k) {
        sim->alive[handles[k]] = 0;
        sim->free_handles.push_back(handles[k]);
    }
    CKS(sim->stage_h.reserve(m));
    CKS(cudaMemcpyAsync(sim->stage_h.p, handles, 4 * (size_t)m, cudaMemcpyHostToDevice, s));
    k_sim_clear_moved<<<(m + 255) / 256, 256, 0, s>>>(sim->stage_h.p, m, sim->moved.p, 0);
    if (sim->n_prev) {
        // compact the pair table: entries of removed objects go, their state slots return to the free list
        uint32_t np = sim->n_prev;
        CKS(sim->sel_flags.reserve(np));
        CKS(sim->keys_tmp.reserve(np));
        CKS(sim->slot_tmp.reserve(np));
        CKS(sim->sel_count.reserve(4));
        k_sim_flag_removed<<<(np + 255) / 256, 256, 0, s>>>(sim->keys_prev.p, np, sim->bp->d_attached.p, sim->slot_prev.p, sim->sel_flags.p,
                                                            sim->free_slots.p, sim->cnt.p);
        size_t b1 = 0, b2 = 0;
        cub::DeviceSelect::Flagged(nullptr, b1, sim->keys_prev.p, sim->sel_flags.p, sim->keys_tmp.p, sim->sel_count.p, (int)np);
        cub::DeviceSelect::Flagged(nullptr, b2, sim->slot_prev.p, sim->sel_flags.p, sim->slot_tmp.p, sim->sel_count.p, (int)np);
        CKS(sim->cub_tmp.reserve(std::max(b1, b2) + 256));
        size_t bytes = sim->cub_tmp.cap;
        CKS(cub::DeviceSelect::Flagged(sim->cub_tmp.p, bytes, sim->keys_prev.p, sim->sel_flags.p, sim->keys_tmp.p, sim->sel_count.p, (int)np, s));
        bytes = sim->cub_tmp.cap;
        CKS(cub::DeviceSelect::Flagged(sim->cub_tmp.p, bytes, sim->slot_prev.p, sim->sel_flags.p, sim->slot_tmp.p, sim->sel_count.p, (int)np, s));
        uint32_t kept = 0, hc0 = 0;
        CKS(cudaMemcpyAsync(&kept, sim->sel_count.p, 4, cudaMemcpyDeviceToHost, s));
        CKS(cudaMemcpyAsync(&hc0, sim->cnt.p, 4, cudaMemcpyDeviceToHost, s));
        CKS(cudaStreamSynchronize(s));
        if (kept) {
            CKS(cudaMemcpyAsync(sim->keys_prev.p, sim->keys_tmp.p, 8 * (size_t)kept, cudaMemcpyDeviceToDevice, s));
            CKS(cudaMemcpyAsync(sim->slot_prev.p, sim->slot_tmp.p, 4 * (size_t)kept, cudaMemcpyDeviceToDevice, s));
        }
        sim->n_prev = kept;
        sim->n_free_host = hc0;
    }
    CKS(cudaGetLastError());
    CKS(cudaStreamSynchronize(s));
    return NCB_OK;
}

// CollisionWorld::add (world.rs:64-96) between updates for the objects of `objs` (hulls refer to the library already set):
// handles come from the object slab (last freed first, else appended) and are returned in out_handles; each proxy is created
// with the object's swept AABB; the object takes part in the next ncb_sim_step with every update flag set.
int ncb_sim_add(ncb_sim* sim, const ncb_objects* objs, uint32_t* out_handles) { return ncb_sim_add_with_query_types(sim, objs, nullptr, out_handles); }

// Same; kinds[k] = 0 Contacts / 1 Proximity per new object (NULL: all Contacts), see ncb_set_query_types.
int ncb_sim_add_with_query_types(ncb_sim* sim, const ncb_objects* objs, const uint8_t* kinds, uint32_t* out_handles) {
    if (!sim || !objs || !out_handles) return NCB_ERR_ARG;
    ncb_ctx* ctx = sim->ctx;
    CKS(cudaSetDevice(ctx->device));
    uint32_t m = objs->n;
    if (m == 0) return NCB_OK;
    if (!(objs->pos && objs->rot && objs->shape_type && objs->shape_param && objs->query_limit && objs->ang_pred)) return NCB_ERR_ARG;
    bool add_capsules = false;
    for (uint32_t k = 0; k < m; ++k) {
        uint32_t t = objs->shape_type[k];
        if (t > NCB_SHAPE_CAPSULE) {
            ctx->err = "ncb_sim_add: shape_type must be NCB_SHAPE_BALL / CUBOID / CONVEX_HULL / PLANE / CAPSULE";
            return NCB_ERR_UNSUPPORTED;
        }
        add_capsules = add_capsules || t == NCB_SHAPE_CAPSULE;
        if (t == NCB_SHAPE_CONVEX_HULL) {
            float h = objs->shape_param[4 * (size_t)k];
            if (!(h >= 0.f && h < (float)ctx->hulls.n_hulls && h == (float)(uint32_t)h)) {
                ctx->err = "ncb_sim_add: convex-hull object names a hull id that is not in the uploaded library (ncb_set_hulls)";
                return NCB_ERR_ARG;
            }
        }
    }
    {
        bool any_kind = ctx->has_prox;
        for (uint32_t k = 0; kinds && k < m; ++k) any_kind = any_kind || kinds[k] != 0;
        if (any_kind && (add_capsules || ctx->has_capsules)) {
            ctx->err = "ncb_sim_add: sensors in a world with capsules are not supported on the device";
            return NCB_ERR_UNSUPPORTED;
        }
    }
    for (uint32_t k = 0; kinds && k < m; ++k)
        if (kinds[k] > 1) {
            ctx->err = "ncb_sim_add_with_query_types: kind must be 0 (Contacts) or 1 (Proximity)";
            return NCB_ERR_ARG;
        }
    if (sim->first) {
        ctx->err = "ncb_sim_add: call ncb_sim_step once before adding objects (the initial set comes from ncb_set_objects)";
        return NCB_ERR_STATE;
    }
    cudaStream_t s = ctx->stream;
    // handles
    uint32_t old_n = sim->n, new_n = old_n;
    std::vector<uint32_t> fh = sim->free_handles;
    for (uint32_t k = 0; k < m; ++k) {
        if (!fh.empty()) {
            out_handles[k] = fh.back();
            fh.pop_back();
        } else {
            out_handles[k] = new_n++;
        }
    }
    // grow the object arrays (contents kept)
    if (new_n > old_n) {
        CKS(grow_keep(ctx->pos, 3 * (size_t)new_n, 3 * (size_t)old_n, s));
        CKS(grow_keep(ctx->rot, new_n, old_n, s));
        CKS(grow_keep(ctx->type, new_n, old_n, s));
        CKS(grow_keep(ctx->param, new_n, old_n, s));
        CKS(grow_keep(ctx->qlimit, new_n, old_n, s));
        if (ctx->has_groups) CKS(grow_keep(ctx->groups, 3 * (size_t)new_n, 3 * (size_t)old_n, s));
        if (ctx->ang_stride) CKS(grow_keep(ctx->ang_cs, new_n, old_n, s));
        CKS(grow_keep(sim->moved, new_n, old_n, s));
        CKS(cudaMemsetAsync(sim->moved.p + old_n, 0, new_n - old_n, s));
    }
    if (objs->groups && !ctx->has_groups) {  // the world used default groups so far: materialise them
        CKS(ctx->groups.reserve(3 * (size_t)new_n));
        k_sim_fill_groups<<<(new_n + 255) / 256, 256, 0, s>>>(ctx->groups.p, new_n);
        ctx->has_groups = true;
    }
    // angular prediction table: (cos, sin) from the host libm like ncb_set_objects; a uniform table is expanded on the first deviation
    std::vector<float2> cs(m);
    bool same = ctx->ang_stride == 0;
    for (uint32_t k = 0; k < m; ++k) {
        cs[k] = make_float2(cosf(objs->ang_pred[k]), sinf(objs->ang_pred[k]));
        if (ctx->ang_stride == 0 && !ctx->h_ang_cs.empty() && (cs[k].x != ctx->h_ang_cs[0].x || cs[k].y != ctx->h_ang_cs[0].y)) same = false;
    }
    if (ctx->ang_stride == 0 && !same) {
        CKS(ctx->ang_cs.reserve(new_n));
        k_sim_fill_f2<<<(new_n + 255) / 256, 256, 0, s>>>(ctx->ang_cs.p, ctx->h_ang_cs[0], new_n);
        ctx->ang_stride = 1;
    }
    // stage + scatter
    DevBuf<float> d_pos, d_rot, d_param, d_ql;
    DevBuf<uint32_t> d_type, d_groups, d_h;
    DevBuf<float2> d_cs;
    CKS(d_pos.reserve(3 * (size_t)m));
    CKS(d_rot.reserve(4 * (size_t)m));
    CKS(d_param.reserve(4 * (size_t)m));
    CKS(d_ql.reserve(m));
    CKS(d_type.reserve(m));
    CKS(d_h.reserve(m));
    CKS(d_cs.reserve(m));
    CKS(cudaMemcpyAsync(d_pos.p, objs->pos, 12 * (size_t)m, cudaMemcpyHostToDevice, s));
    CKS(cudaMemcpyAsync(d_rot.p, objs->rot, 16 * (size_t)m, cudaMemcpyHostToDevice, s));
    CKS(cudaMemcpyAsync(d_param.p, objs->shape_param, 16 * (size_t)m, cudaMemcpyHostToDevice, s));
    CKS(cudaMemcpyAsync(d_ql.p, objs->query_limit, 4 * (size_t)m, cudaMemcpyHostToDevice, s));
    CKS(cudaMemcpyAsync(d_type.p, objs->shape_type, 4 * (size_t)m, cudaMemcpyHostToDevice, s));
    CKS(cudaMemcpyAsync(d_h.p, out_handles, 4 * (size_t)m, cudaMemcpyHostToDevice, s));
    CKS(cudaMemcpyAsync(d_cs.p, cs.data(), 8 * (size_t)m, cudaMemcpyHostToDevice, s));
    if (objs->groups) {
        CKS(d_groups.reserve(3 * (size_t)m));
        CKS(cudaMemcpyAsync(d_groups.p, objs->groups, 12 * (size_t)m, cudaMemcpyHostToDevice, s));
    }
    k_sim_scatter_objects<<<(m + 255) / 256, 256, 0, s>>>(d_h.p, m, d_pos.p, d_rot.p, d_type.p, d_param.p, objs->groups ? d_groups.p : nullptr, d_ql.p, d_cs.p,
                                                          ctx->pos.p, ctx->rot.p, ctx->type.p, ctx->param.p, ctx->has_groups ? ctx->groups.p : nullptr,
                                                          ctx->qlimit.p, ctx->ang_stride ? ctx->ang_cs.p : nullptr, sim->moved.p);
    CKS(cudaGetLastError());
    {
        // query types of the new objects (a recycled handle must not inherit the kind of its previous owner)
        bool any_kind = false;
        for (uint32_t k = 0; kinds && k < m; ++k) any_kind = any_kind || kinds[k] != 0;
        if (any_kind && !ctx->has_prox) {  // first sensor of this world
            CKS(ctx->qkind.reserve(new_n));
            CKS(cudaMemsetAsync(ctx->qkind.p, 0, new_n, s));
            ctx->has_prox = true;
        } else if (ctx->has_prox && new_n > old_n) {
            CKS(grow_keep(ctx->qkind, new_n, old_n, s));
            CKS(cudaMemsetAsync(ctx->qkind.p + old_n, 0, new_n - old_n, s));
        }
        if (ctx->has_prox) {
            DevBuf<uint8_t> d_k;
            if (kinds) {
                CKS(d_k.reserve(m));
                CKS(cudaMemcpyAsync(d_k.p, kinds, m, cudaMemcpyHostToDevice, s));
            }
            k_sim_scatter_kinds<<<(m + 255) / 256, 256, 0, s>>>(d_h.p, m, kinds ? d_k.p : nullptr, ctx->qkind.p);
            CKS(cudaGetLastError());
            CKS(cudaStreamSynchronize(s));
            d_k.release();
        }
    }
    if (add_capsules || ctx->has_capsules) {  // capsule segments as 2-point hulls, for every object of the (grown) world
        CKS(ctx->cap_pts.reserve(6 * (size_t)new_n));
        CKS(launch_fill_cap_pts(ctx, new_n));
        ctx->has_capsules = true;
    }
    ctx->n = sim->n = new_n;
    sim->alive.resize(new_n, 0);
    for (uint32_t k = 0; k < m; ++k) sim->alive[out_handles[k]] = 1;
    sim->free_handles = fh;
    // proxies: swept AABB of the new objects at their current pose (glue/setup.rs:27-30)
    int r = reserve_broad(ctx, new_n);
    if (r) return r;
    CKS(launch_aabbs(ctx, dev_objects(ctx), sim->margin, 1, 0, new_n));
    r = bp_create_listed_device(sim->bp, m, out_handles, d_h.p, ctx->aabb_lo.p, ctx->aabb_hi.p);
    CKS(cudaStreamSynchronize(s));
    d_pos.release(), d_rot.release(), d_param.release(), d_ql.release(), d_type.release(), d_groups.release(), d_h.release(), d_cs.release();
    return r;
}

// Sizes of the last step: pairs, contacts, contact events.
int ncb_sim_sizes(ncb_sim* sim, uint32_t* n_pairs, uint32_t* n_contacts, uint32_t* n_events) {
    if (!sim) return NCB_ERR_ARG;
    if (n_pairs) *n_pairs = sim->n_pairs;
    if (n_contacts) *n_contacts = sim->n_contacts;
    if (n_events) *n_events = sim->n_events;
    return NCB_OK;
}

// Results of the last step; every array is sized from ncb_sim_sizes (NULL = not wanted).
//   pairs[2 * P]   (object1, object2) of every interference pair, sorted by (min handle, max handle); object order = the
//                  argument order of the interference_started callback that created the pair
//   algo[P], manifold_start[P], manifold_count[P], contacts[C] as ncb_world_fetch; contact_ids[C] = insertion counter << 8 |
//                  slab slot, stable per pair exactly as long as the reference's ContactId is
//   events[3 * E]  (object1, object2, 1 = ContactEvent::Started / 0 = Stopped), sorted
int ncb_sim_fetch(ncb_sim* sim, uint32_t* pairs, uint8_t* algo, uint32_t* manifold_start, uint8_t* manifold_count, ncb_contact* contacts,
                  uint32_t* contact_ids, uint32_t* events) {
    if (!sim) return NCB_ERR_ARG;
    CKS(cudaSetDevice(sim->ctx->device));
    cudaStream_t s = sim->ctx->stream;
    uint32_t P = sim->n_pairs, Cn = sim->n_contacts, E = sim->n_events;
    if (pairs && P) CKS(cudaMemcpyAsync(pairs, sim->exp_pairs.p, 8 * (size_t)P, cudaMemcpyDeviceToHost, s));
    if (algo && P) CKS(cudaMemcpyAsync(algo, sim->exp_algo.p, P, cudaMemcpyDeviceToHost, s));
    if (manifold_start && P) CKS(cudaMemcpyAsync(manifold_start, sim->exp_start.p, 4 * (size_t)P, cudaMemcpyDeviceToHost, s));
    if (manifold_count && P) CKS(cudaMemcpyAsync(manifold_count, sim->exp_mcount.p, P, cudaMemcpyDeviceToHost, s));
    if (contacts && Cn) CKS(cudaMemcpyAsync(contacts, sim->exp_contacts.p, sizeof(ncb_contact) * (size_t)Cn, cudaMemcpyDeviceToHost, s));
    if (contact_ids && Cn) CKS(cudaMemcpyAsync(contact_ids, sim->exp_ids.p, 4 * (size_t)Cn, cudaMemcpyDeviceToHost, s));
    if (events && E) {
        CKS(sim->exp_events.reserve(3 * (size_t)E));
        k_sim_unpack_events<<<(E + 255) / 256, 256, 0, s>>>(sim->events.p, E, sim->exp_events.p);
        CKS(cudaGetLastError());
        CKS(cudaMemcpyAsync(events, sim->exp_events.p, 12 * (size_t)E, cudaMemcpyDeviceToHost, s));
    }
    CKS(cudaStreamSynchronize(s));
    return NCB_OK;
}

// Proximity side of the last step: prox[P] = status per pair in the order of ncb_sim_fetch (NCB_PROXIMITY_NONE for contact pairs);
// events[4 * E] = ProximityEvent rows (collider1, collider2, prev_status, new_status), sorted; cap_events in rows; *n_events = rows
// that exist; returns 1 when events were truncated.  Any pointer may be NULL.
int ncb_sim_fetch_proximity(ncb_sim* sim, uint8_t* prox, uint32_t* events, uint32_t cap_events, uint32_t* n_events) {
    if (!sim) return NCB_ERR_ARG;
    CKS(cudaSetDevice(sim->ctx->device));
    cudaStream_t s = sim->ctx->stream;
    uint32_t P = sim->n_pairs, E = sim->n_prox_events;
    if (n_events) *n_events = E;
    if (prox && P) CKS(cudaMemcpyAsync(prox, sim->exp_prox.p, P, cudaMemcpyDeviceToHost, s));
    std::vector<uint4> ev;
    if (events && E) {
        ev.resize(E);
        CKS(cudaMemcpyAsync(ev.data(), sim->prox_events.p, sizeof(uint4) * (size_t)E, cudaMemcpyDeviceToHost, s));
    }
    CKS(cudaStreamSynchronize(s));
    if (events && E) {
        std::sort(ev.begin(), ev.end(), [](const uint4& a, const uint4& b) { return a.x != b.x ? a.x < b.x : (a.y != b.y ? a.y < b.y : a.z < b.z); });
        for (uint32_t k = 0; k < E && k < cap_events; ++k) events[4 * k] = ev[k].x, events[4 * k + 1] = ev[k].y, events[4 * k + 2] = ev[k].z, events[4 * k + 3] = ev[k].w;
    }
    return (events && E > cap_events) ? 1 : NCB_OK;
}

// glue::interferences_with_ray (first_only = 0) / first_interference_with_ray (first_only = 1) (glue/query.rs:13-77,183-224)
// against the boxes stored by the last ncb_sim_step: rays[7 * n] = origin, dir, max_toi; groups = the query's CollisionGroups
// (membership, whitelist, blacklist) or NULL.  Rows sorted by (ray, handle): idx[2 k] = (ray, handle), val[4 k] = (toi,
// normal), feat[k]; cap in rows; *n_out = rows found; returns 1 when truncated.  solid = true like the reference's queries.
int ncb_sim_ray_cast(ncb_sim* sim, uint32_t n_rays, const float* rays, const uint32_t* groups, int first_only, uint32_t* idx, float* val,
                     uint32_t* feat, uint32_t cap, uint32_t* n_out) {
    if (!sim || (n_rays && !rays)) return NCB_ERR_ARG;
    return world_ray_cast(sim->ctx, sim->bp, sim->qbufs, n_rays, rays, groups, first_only, idx, val, feat, cap, n_out);
}

// glue::interferences_with_aabb (kind 0; 6 floats per query: mins, maxs) / interferences_with_point (kind 2; 3 floats)
// (glue/query.rs:79-181): broad-phase candidates on the stored boxes, the query's collision groups, and for points the
// shape's PointQuery::contains_point.  idx[2 k] = (query, handle), sorted; cap in rows; returns 1 when truncated.
int ncb_sim_query(ncb_sim* sim, int kind, uint32_t n_queries, const float* queries, const uint32_t* groups, uint32_t* idx, uint32_t cap,
                  uint32_t* n_out) {
    if (!sim || (kind != 0 && kind != 2) || (n_queries && !queries)) return NCB_ERR_ARG;
    return world_query(sim->ctx, sim->bp, sim->qbufs, kind, n_queries, queries, groups, idx, cap, n_out);
}

}  // extern "C"
