// Internal state of the persistent broad phase (bp_persistent.cu), shared with the stepping world (sim.cu).
#pragma once
#include <string>
#include <vector>
#include "ncb_internal.h"

using ncb::DevBuf;

#define SEQ_NONE 0xffffffffu
#define LEAF_BIT 0x80000000u  // child word of an LBVH node record (broad.cu)
#define SEQ_BACK0 0x80000000u   // entries pushed at the back of the reference's queue: SEQ_BACK0 + k
#define SEQ_FRONT0 0x7fffffffu  // entries pushed at its front (recompute_*): SEQ_FRONT0 - k
enum : uint32_t { ST_DETACHED = 0, ST_ATTACHED = 1, ST_REMOVING = 2, ST_VACANT = 3 };

struct ncb_bp {
    ncb_ctx* owner = nullptr;
    ncb_ctx* work = nullptr;  // private LBVH / pair buffers
    float margin = 0.f;
    // host slab (LIFO reuse like the `slab` crate)
    std::vector<int64_t> next_free;  // -2 occupied, else next vacant
    std::vector<uint8_t> attached;   // host mirror: 0 detached (pending), 1 attached
    size_t next = 0, len = 0;
    uint32_t n_attached = 0;
    bool slab_dirty = true;  // proxies were created / removed since `alive` was uploaded
    uint32_t seq = 0;    // back entries issued since the last update
    uint32_t front = 0;  // front entries issued since the last update
    // device state, indexed by handle slot
    DevBuf<float4> box_lo, box_hi, pend_lo, pend_hi;
    DevBuf<uint32_t> pend_seq, upd_seq, win, d_attached;
    size_t slots_cap = 0;
    // staging
    DevBuf<float> stage_f;
    DevBuf<uint32_t> stage_u, alive, groups_dev;
    DevBuf<unsigned long long> keys_old, keys_new, keys_tmp;
    DevBuf<uint8_t> cub_tmp;
    DevBuf<unsigned long long> ev_a, ev_b, ev_sorted;  // started / stopped events as (first << 32 | second)
    DevBuf<uint32_t> ev_u32, counters;
    uint32_t n_old = 0;
    uint32_t tree_n = 0, tree_outliers = 0;  // leaves of the LBVH built by the last update() (0: none yet)
    DevBuf<float> q_in;
    uint32_t n_started = 0, n_stopped = 0;  // events of the last update() / remove()
    std::string err;
};


// device-side entry points used by the stepping world: boxes already in HBM as two float4 arrays indexed by handle
int bp_create_all_device(ncb_bp* bp, uint32_t n, const float4* lo, const float4* hi);
int bp_create_listed_device(ncb_bp* bp, uint32_t m, const uint32_t* handles_host, const uint32_t* handles_dev, const float4* lo, const float4* hi);
int bp_set_moved_device(ncb_bp* bp, uint32_t n, const float4* lo, const float4* hi, const uint8_t* moved);
int bp_update_impl(ncb_bp* bp, const uint32_t* d_groups, uint32_t* n_started, uint32_t* n_stopped, bool want_events = true);

// world-level ray queries (query.cu)
struct WorldQueryBufs {
    DevBuf<float> rays;
    DevBuf<unsigned long long> keys, keys_sorted;
    DevBuf<float4> vals, out_val;
    DevBuf<uint32_t> feats, counter, order_in, order_out, out_idx, out_feat;
    DevBuf<uint8_t> cub_tmp;
    void release() {
        rays.release(), keys.release(), keys_sorted.release(), vals.release(), out_val.release(), feats.release(), counter.release();
        order_in.release(), order_out.release(), out_idx.release(), out_feat.release(), cub_tmp.release();
    }
};
int world_ray_cast(ncb_ctx* ctx, ncb_bp* bp, WorldQueryBufs& B, uint32_t n_rays, const float* rays, const uint32_t* groups, int first_only,
                   uint32_t* idx, float* val, uint32_t* feat, uint32_t cap, uint32_t* n_out);
int world_query(ncb_ctx* ctx, ncb_bp* bp, WorldQueryBufs& B, int kind, uint32_t n_q, const float* q, const uint32_t* groups, uint32_t* idx,
                uint32_t cap, uint32_t* n_out);
