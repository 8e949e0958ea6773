// Internal declarations shared by the translation units of libncb200.so (context, device buffers, launchers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/ncb200.h"

namespace ncb {

// Pair type keys (pair_search classifies; the narrow phase runs one persistent kernel per key segment).
enum PairKey : uint32_t {
    K_BALL_BALL = 0,
    K_PLANE_BALL = 1,
    K_PLANE_CUBOID = 2,
    K_PLANE_HULL = 3,
    K_BALL_CUBOID = 4,
    K_BALL_HULL = 5,
    K_CUBOID_CUBOID = 6,
    K_CUBOID_HULL = 7,
    K_HULL_HULL = 8,
    K_NONE = 9,
    // pairs with a GeometricQueryType::Proximity object (proximity.cu); adjacent
    K_PROX_BALL_BALL = 10,
    K_PROX_PLANE = 11,
    K_PROX_SM = 12,       // support map x support map without a hull operand (ball / cuboid): O(1) support functions
    K_PROX_SM_HULL = 13,  // ... with at least one convex hull (vertex scans)
    // pairs with a capsule (capsule.cuh; CapsuleCapsule / CapsuleShape generators); adjacent, one kernel runs the five segments
    K_CAPSULE_BALL = 14,
    K_CAPSULE_PLANE = 15,
    K_CAPSULE_CAPSULE = 16,
    K_CAPSULE_CUBOID = 17,
    K_CAPSULE_HULL = 18,
    K_COUNT = 19,
    K_MAX = 32  // size of the per-key counter arrays
};
#define NCB_TYPE_MASK 7u  // shape types travel in 3 bits (leaf records, key tables)

// The contact dispatcher as a table (default_contact_dispatcher.rs:27-97): key of a pair from its two shape types.
// Types: 0 ball, 1 cuboid, 2 convex hull, 3 plane, 4 capsule; anything else has no generator.
__host__ __device__ inline uint32_t pair_key(uint32_t t1, uint32_t t2) {
    t1 &= NCB_TYPE_MASK, t2 &= NCB_TYPE_MASK;
    if (t1 > 4 || t2 > 4) return K_NONE;
    if (t1 == 4 || t2 == 4) {  // CapsuleCapsule, then CapsuleShape on whatever the other shape is (dispatcher :64-78)
        uint32_t other = t1 == 4 ? t2 : t1;
        return other == 0 ? K_CAPSULE_BALL : other == 1 ? K_CAPSULE_CUBOID : other == 2 ? K_CAPSULE_HULL : other == 3 ? K_CAPSULE_PLANE : K_CAPSULE_CAPSULE;
    }
    uint32_t lo = t1 < t2 ? t1 : t2, hi = t1 < t2 ? t2 : t1;
    if (hi == 3) return lo == 0 ? K_PLANE_BALL : lo == 1 ? K_PLANE_CUBOID : lo == 2 ? K_PLANE_HULL : K_NONE;
    if (lo == 0) return hi == 0 ? K_BALL_BALL : hi == 1 ? K_BALL_CUBOID : K_BALL_HULL;
    if (lo == 1) return hi == 1 ? K_CUBOID_CUBOID : K_CUBOID_HULL;
    return K_HULL_HULL;
}
// NCB_ALGO_* of a key
__host__ __device__ inline uint32_t algo_of_key(uint32_t key) {
    switch (key) {
        case K_BALL_BALL: return NCB_ALGO_BALL_BALL;
        case K_PLANE_BALL: return NCB_ALGO_PLANE_BALL;
        case K_PLANE_CUBOID:
        case K_PLANE_HULL: return NCB_ALGO_PLANE_CONVEX;
        case K_BALL_CUBOID:
        case K_BALL_HULL: return NCB_ALGO_BALL_CONVEX;
        case K_CUBOID_CUBOID:
        case K_CUBOID_HULL:
        case K_HULL_HULL: return NCB_ALGO_CONVEX_CONVEX;
        case K_PROX_BALL_BALL:
        case K_PROX_PLANE:
        case K_PROX_SM:
        case K_PROX_SM_HULL: return NCB_ALGO_PROXIMITY;
        case K_CAPSULE_CAPSULE: return NCB_ALGO_CAPSULE_CAPSULE;
        case K_CAPSULE_BALL:
        case K_CAPSULE_PLANE:
        case K_CAPSULE_CUBOID:
        case K_CAPSULE_HULL: return NCB_ALGO_CAPSULE_SHAPE;
        default: return NCB_ALGO_NONE;
    }
}

struct DevHulls {
    uint32_t n_hulls;
    const uint32_t *vert_off, *face_off, *edge_off, *fadj_off, *vadj_off;
    const float* points;
    const float4* points4;  // the same vertices padded to 16 B: one load per vertex in the support scans of GJK / EPA
    const uint32_t *vert_first_adj, *vert_num_adj;
    const uint32_t *face_first, *face_num;
    const float* face_normal;
    const uint32_t *vaf, *eaf;
    const uint32_t *edge_vertices, *edge_faces;
    const float* edge_dir;
    const uint32_t *fav, *eav;
};

struct DevObjects {
    uint32_t n;
    const float* pos;        // 3 per object
    const float4* rot;       // i j k w
    const uint32_t* type;
    const float4* param;
    const uint32_t* groups;  // 3 per object or nullptr
    const float* qlimit;
    const float* ang;
    const float2* ang_cs;    // (cos, sin) of ang, evaluated on the host with libm like the reference does
    uint32_t ang_stride;     // 1: one entry per object; 0: every object has the same angular prediction (entry 0)
    const float* cap_pts;    // worlds with capsules: 6 floats per OBJECT, the capsule segment as the 2-point hull [b, a] (else nullptr)
};

// Counters living in one device allocation (zeroed per update with one memset).
struct DevCounters {
    uint32_t n_pairs;          // emitted by pair search (may exceed capacity)
    uint32_t n_contacts;       // allocated by the narrow phase (may exceed capacity)
    uint32_t n_contact_pairs;
    uint32_t epa_overflow;
    uint32_t ref_panics;
    uint32_t n_outliers;       // objects kept out of the LBVH (planes / non-finite boxes)
    uint32_t key_hist[K_MAX];     // pairs per PairKey
    uint32_t key_start[K_MAX];    // exclusive scan of key_hist
    uint32_t key_cursor[K_MAX];   // scatter cursors
    uint32_t epa_cursor[K_MAX];   // per key: end of the EPA work queue (starts at key_start[key])
    uint32_t cp_cursor[K_MAX];    // per key: end of the closest-points (manifold) work queue
    uint32_t epa_fetch[K_MAX];    // per key: next EPA queue entry to hand to an idle lane (dynamic fetch)
    uint32_t gjk_fetch[K_MAX];    // per key: next pair of the key segment to hand to an idle lane
    int bounds[6];             // ordered-int encoded min xyz / max xyz of AABB centres
    uint32_t epa_long_n;       // EPA pairs that outgrew the first-tier (shared-memory) polytope store: queue of the second tier
    uint32_t epa_long_fetch;
    uint32_t epa_defer_n;      // pairs beyond the second tier / segment simplices: queue of the last resort (k_cc_epa_big)
    uint32_t epa_defer_fetch;
    uint32_t stack_overflow;   // BVH traversals (pair search, ray casts, queries) that ran out of their 64-entry stack: must stay 0
    uint32_t prox_hist[4];     // proximity pairs per status (Intersecting, WithinMargin, Disjoint)
    uint32_t man_split;        // end of the manifold queue when the first EPA tier finished: what lies beyond comes from the later tiers
};

// Persistent narrow-phase state of a stepping world (sim.cu), indexed by state slot.
#define PM_CAP 24        // cache entries (live + stale) of one persistent manifold
#define PM_HDR_WORDS 8   // 2 + PM_CAP / 4
#define PM_ENTRY_F4 4    // float4 per entry
struct PersistArgs {
    float4* dir;        // last_gjk_dir (xyz) + valid flag (w) per slot
    uint32_t* pm_hdr;   // PM_HDR_WORDS per slot
    float4* pm_entry;   // PM_CAP * PM_ENTRY_F4 per slot
    unsigned long long* events;  // started << 63 | h1 << 32 | h2
    uint32_t* n_events;
    uint32_t cap_events;
    uint32_t* pm_overflow;
};

// Scratch of the spatial shard selection (broad.cu)
// The convex-convex manifold kernel runs in this many parts when results are shipped to the host while the update runs
// (ncb_world_update / ncb_world_fetch_early): the contacts of part k leave while part k + 1 computes.
#define NCB_MAN_PARTS 4

#define SHARD_BINS 1024
#define SHARD_MAX_RANKS 16
#define SHARD_SUBS 8  // sub-boxes of a rank's region in the routed sharding: one per top-level octant of the Morton bins
struct ShardScratch {
    uint32_t hist[SHARD_BINS];
    uint32_t split[SHARD_MAX_RANKS + 1];
    int region[6];  // ordered-int encoded union of the owned fat boxes
    uint32_t m, n_owned;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr, o.cap = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p, cap = o.cap;
            o.p = nullptr, o.cap = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }  // every buffer of a context / a temporary goes with its owner (error returns included)
};

// Buffers of the routed multi-GPU sharding (broad.cu): what the collectives of ncollide_b200/parallel.py read and write.
struct RouteBufs {
    DevBuf<float> bounds;      // 8 floats: [-min xyz, max xyz] of the own block's AABB centres (all-reduce max)
    DevBuf<int> hist;          // SHARD_BINS counts of the own block (all-reduce sum)
    DevBuf<uint32_t> split;    // world + 1 bin boundaries
    DevBuf<uint32_t> bins;     // bin per own object
    DevBuf<int> region_i;      // SHARD_MAX_RANKS x 6 ordered ints: union box of what this rank sends to each owner
    DevBuf<float> region_f;    // the same as [-min, max] floats (all-reduce max) = the owned region of every rank
    DevBuf<uint32_t> counts;   // [0, MAX_RANKS): owned records per destination, [MAX_RANKS, 2 MAX_RANKS): ghost records
    DevBuf<float4> send_o, recv_o, send_g, recv_g;  // world buckets of cap records of recw float4 (slot 0 = header)
    uint32_t cap_o = 0, cap_g = 0;
    int recw = 2;              // 2: boxes only; 4: boxes + poses
    // Peer-memory exchange (NVLink P2P, no NCCL in the step): every rank's recv buffers, meta slots and flags are mapped into all
    // peers (CUDA IPC between processes, raw pointers inside one process); the routing kernels store records straight into the
    // owner's bucket, small arrays (bounds, histogram, regions) go to a per-sender meta slot, and a system-scope flag per sender
    // closes each of the four rounds of a step.
    bool p2p = false;
    int p2p_rank = 0, p2p_world = 0;
    uint32_t p2p_cap = 0;              // records per bucket (incl. the header slot) = largest block + 1: cannot overflow
    uint32_t epoch = 0;                // steps so far; the flag value of round r (1..4) of a step is 4 * epoch + r
    DevBuf<float4> p2p_recv_o, p2p_recv_g;
    DevBuf<uint32_t> p2p_meta;
    float4* peer_recv_o[SHARD_MAX_RANKS] = {};
    float4* peer_recv_g[SHARD_MAX_RANKS] = {};
    uint32_t* peer_meta[SHARD_MAX_RANKS] = {};
    bool peer_opened[SHARD_MAX_RANKS][3] = {};  // mapped with cudaIpcOpenMemHandle (to be closed)
};
// meta buffer of the peer-memory exchange, in 32-bit words: three slots (bounds, histogram, regions) per sender, then flags
#define P2P_SLOT_WORDS 1024
#define P2P_FLAGS_OFF (3 * SHARD_MAX_RANKS * P2P_SLOT_WORDS)
#define P2P_ERR_OFF (P2P_FLAGS_OFF + 32)
#define P2P_META_WORDS (P2P_ERR_OFF + 32)
static_assert(SHARD_MAX_RANKS * SHARD_SUBS * 6 <= P2P_SLOT_WORDS, "a rank's region boxes must fit one meta slot");
struct RouteDst {
    float4* p[SHARD_MAX_RANKS];  // where the bucket for destination q starts (local send buffer, or the peer's recv buffer)
};
struct RoutePeers {
    uint32_t* meta[SHARD_MAX_RANKS];
};

struct StageTimer {
    static const int MAX = 24;
    bool enabled = false;
    int n = 0;
    const char* names[MAX];
    uint32_t launches[MAX];
    cudaEvent_t ev[MAX + 1];
    bool created = false;
};

}  // namespace ncb

struct ncb_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t side_stream = nullptr;  // second chain of the narrow phase (fork / join with events)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_tier1 = nullptr, ev_tier2 = nullptr;  // later EPA tiers run on the side stream beside the manifold kernel
    std::string err;
    int sm_count = 148;

    // objects
    uint32_t n = 0, n_planes = 0;
    bool has_groups = false;
    ncb::DevBuf<float> pos, qlimit, ang;
    ncb::DevBuf<float4> rot, param;
    ncb::DevBuf<float2> ang_cs;
    ncb::DevBuf<uint32_t> type, groups;
    ncb::DevBuf<uint32_t> trav_overflow;         // query / ray traversals that ran out of stack (cumulative; ncb_traversal_overflows)
    ncb::DevBuf<float> cap_pts;                  // capsule segments as 2-point hulls, 6 floats per object (only with capsules)
    bool has_capsules = false;
    std::vector<float2> h_ang_cs;
    uint32_t ang_stride = 1;
    // GeometricQueryType per object (ncb_set_query_types): 1 = Proximity(query_limit).  has_prox = at least one sensor.
    ncb::DevBuf<uint8_t> qkind;
    bool has_prox = false;
    ncb::DevBuf<uint8_t> prox;                   // Proximity status per sorted pair (255 for contact pairs)
    // hulls
    ncb::DevHulls hulls = {};
    std::vector<void*> hull_allocs;

    // broad phase
    ncb::DevBuf<float4> aabb_lo, aabb_hi;        // handle order
    ncb::DevBuf<uint32_t> keys_a, keys_b, idx_a, idx_b;
    ncb::DevBuf<uint8_t> cub_tmp;
    ncb::DevBuf<float4> leaf_lo, leaf_hi;        // Morton order; lo.w = handle bits, hi.w = shape type bits
    ncb::DevBuf<float4> nodes;                   // 4 float4 per internal node
    ncb::DevBuf<uint32_t> parent;                // [0,n) internal parents, [n,2n) leaf parents
    ncb::DevBuf<uint32_t> flags;
    ncb::DevBuf<uint2> pairs_raw, pairs;         // emission order / sorted by key
    ncb::DevBuf<uint8_t> keys_raw, pair_algo;
    ncb::DevBuf<ncb::DevCounters> counters;
    // narrow phase
    ncb::DevBuf<ncb_contact> contacts;
    ncb::DevBuf<ncb_kinematic> kinematics;       // aligned with contacts; only with want_kinematics (ncb_set_kinematics)
    bool want_kinematics = false, have_kinematics = false;
    ncb::DevBuf<uint32_t> manifold_start;
    ncb::DevBuf<uint8_t> manifold_count;
    ncb::DevBuf<uint32_t> pair_index;
    ncb::DevBuf<uint32_t> epa_long;              // EPA queue indices deferred to the second tier / to the last resort (2 x cap_pairs)
    ncb::DevBuf<uint32_t> epa_queue;             // 26 words per record
    ncb::DevBuf<uint32_t> cp_queue;              // 10 words per record

    ncb::DevCounters last_counters = {};
    uint32_t last_n_pairs = 0, last_n_contacts = 0;
    uint32_t cap_pairs_hint = 0, cap_contacts_hint = 0;

    ncb::StageTimer timer;
    bool timer_external = false;  // stage 0 (AABBs) already started the timer of this update
    ncb::DevCounters* h_counters = nullptr;  // pinned
    // overlapped result fetch of ncb_world_update (api.cu): copy stream, events, counter snapshots
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_pairs = nullptr, ev_copy = nullptr;
    // counter snapshots of an update that ships its results while it runs: [0] before the convex-convex EPA / manifold phases
    // (everything the other kernels wrote is final), [1 .. NCB_MAN_PARTS] after each part of the manifold kernel
    cudaEvent_t ev_snap[1 + NCB_MAN_PARTS] = {};
    ncb::DevBuf<uint32_t> snap;              // n_contacts per snapshot (device)
    ncb::DevCounters* h_snap = nullptr;      // pinned: [0] the counters after the pair sort; h_snap_n: the snapshots above
    uint32_t* h_snap_n = nullptr;
    struct EarlyFetch {
        bool active = false;  // armed for the next update
        bool valid = false;   // the last update filled pairs_done / contacts_done
        uint32_t* pairs = nullptr;
        uint8_t* algo = nullptr;
        ncb_contact* contacts = nullptr;
        uint32_t cap_pairs = 0, cap_contacts = 0;
        uint32_t pairs_done = 0, contacts_done = 0;  // rows already on their way to the host
    } early;
    // spatial sharding (several GPUs)
    ncb::DevBuf<ncb::ShardScratch> shard;
    ncb::DevBuf<uint32_t> shard_bins, shard_sel;
    ncb::DevBuf<float4> shard_lo, shard_hi;
    uint32_t shard_m = 0, shard_owned = 0;
    ncb::RouteBufs route;
    // sharded updates: compact copies of the attributes of the objects this rank holds, indexed by local id (broad.cu)
    ncb::DevBuf<float> loc_pos, loc_qlimit, loc_cap;
    ncb::DevBuf<float4> loc_rot, loc_param;
    ncb::DevBuf<uint32_t> loc_type, local_of;
    ncb::DevBuf<float2> loc_ang_cs;
    ncb::DevBuf<uint2> pairs_local;
    // ncollide2d slice (dim2.cu): device copies of the last 2-D world / pair batch and its results, kept between calls
    struct Dim2Bufs {
        ncb::DevBuf<float2> pos, rot;
        ncb::DevBuf<uint32_t> type, groups, start, feat, cnt;
        ncb::DevBuf<float4> param;
        ncb::DevBuf<float> ql, cang, sang, poly, nrm, contacts;
        ncb::DevBuf<uint8_t> count, qkind, prox;
        uint32_t last_pairs = 0;  // pairs of the last update (ncb2d_world_fetch_proximity)
        uint32_t last_n = 0;      // objects of the last update, whose boxes / tree are still in the context (ncb2d_world_ray_cast)
        bool last_groups = false;
        ncb::DevBuf<float> q_rays;
        ncb::DevBuf<unsigned long long> q_keys, q_keys2;
        ncb::DevBuf<float4> q_vals, q_vals2;
        ncb::DevBuf<uint32_t> q_feat, q_feat2, q_order, q_order2, q_cnt;
        ncb::DevBuf<uint8_t> q_tmp;
    } d2;
};

// api.cu helpers shared with sim.cu
ncb::DevObjects dev_objects(ncb_ctx* c);
int reserve_broad(ncb_ctx* ctx, uint32_t n);
int reserve_pairs(ncb_ctx* ctx, size_t cap);
int reset_counters(ncb_ctx* ctx);
int read_counters(ncb_ctx* ctx);
uint32_t* trav_overflow_counter(ncb_ctx* ctx);
cudaError_t launch_fill_cap_pts(ncb_ctx* ctx, uint32_t n);  // ctx->cap_pts from ctx->type / ctx->param (worlds with capsules)  // device counter of skipped subtrees in query / ray traversals (allocated on first use)

namespace ncb {

// ---- stage timers (CUDA events on the context's stream, only when ncb_profile_enable(ctx, 1)) ----------------
inline void timer_begin(ncb_ctx* c) {
    StageTimer& t = c->timer;
    t.n = 0;
    if (!t.enabled) return;
    if (!t.created) {
        for (int i = 0; i <= StageTimer::MAX; ++i) cudaEventCreate(&t.ev[i]);
        t.created = true;
    }
    cudaEventRecord(t.ev[0], c->stream);
}
inline void timer_mark(ncb_ctx* c, const char* name, uint32_t launches) {
    StageTimer& t = c->timer;
    if (!t.enabled || t.n >= StageTimer::MAX) return;
    t.names[t.n] = name;
    t.launches[t.n] = launches;
    t.n++;
    cudaEventRecord(t.ev[t.n], c->stream);
}

// broad.cu
cudaError_t launch_aabbs(ncb_ctx* c, const DevObjects& o, float margin, int fat, uint32_t begin, uint32_t end);
// handle_map (optional): leaf ids reported in pairs are handle_map[index] instead of the index into aabb_lo / aabb_hi
cudaError_t launch_lbvh_build(ncb_ctx* c, uint32_t n, const uint32_t* handle_map);
cudaError_t launch_pair_search(ncb_ctx* c, uint32_t n, const uint32_t* groups, uint32_t q_begin, uint32_t q_end, uint32_t cap_pairs, int my_rank = -1);
cudaError_t launch_shard_select(ncb_ctx* c, uint32_t n, int rank, int world, ShardScratch* sh, uint32_t* bins, uint32_t cap, uint32_t* sel,
                                float4* loc_lo, float4* loc_hi);
cudaError_t launch_pair_sort(ncb_ctx* c, uint32_t cap_pairs, uint32_t* index_out, const uint32_t* local_of = nullptr, uint2* pairs_local = nullptr);
cudaError_t launch_gather_local_objects(ncb_ctx* c, const uint32_t* sel, uint32_t m, const DevObjects& g, DevObjects* out);
cudaError_t launch_route_stage(ncb_ctx* c, int stage, int rank, int world, uint32_t begin, uint32_t end, RouteBufs& R);
cudaError_t launch_p2p_push(ncb_ctx* c, RouteBufs& R, const void* src, uint32_t words, int slot, int round);
cudaError_t launch_p2p_wait_reduce(ncb_ctx* c, RouteBufs& R, int round, int slot, uint32_t words, int op, void* out);
cudaError_t launch_route_unpack(ncb_ctx* c, int world, RouteBufs& R, uint32_t cap_local, ShardScratch* sh, uint32_t* sel, float4* loc_lo, float4* loc_hi);
size_t lbvh_temp_bytes(uint32_t n);
// narrow.cu
cudaError_t launch_narrow_phase(ncb_ctx* c, const DevObjects& o, const uint2* pairs, const uint32_t* pair_index, uint32_t cap_pairs,
                                uint32_t cap_contacts);
cudaError_t launch_classify_pairs(ncb_ctx* c, const uint2* pairs, uint32_t n);
// proximity.cu
cudaError_t launch_prox_rekey(ncb_ctx* c, uint32_t cap_pairs);
cudaError_t launch_proximity_segments(ncb_ctx* c, const DevObjects& o, const uint2* pairs, const uint32_t* pair_index, cudaStream_t s);
cudaError_t launch_proximity_persistent(ncb_ctx* c, const DevObjects& o, const uint2* pairs, const uint32_t* slot_of, float4* slot_dir,
                                        uint8_t* slot_prox, uint4* events, uint32_t* n_events, uint32_t cap_events);
cudaError_t launch_proximity_batch(ncb_ctx* c, const DevObjects& o, const uint2* pairs, uint32_t n, const float* margins, uint8_t* out);
cudaError_t launch_narrow_phase_persistent(ncb_ctx* c, const DevObjects& o, const uint2* pairs, const uint32_t* pair_index, uint32_t cap_pairs,
                                           const PersistArgs& ps);
}  // namespace ncb
