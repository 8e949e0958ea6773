// ORACLE — TEST INFRASTRUCTURE ONLY (see na.hpp header).
// Reference-faithful multi-step DBVTBroadPhase, restated from
//   pipeline/broad_phase/dbvt_broad_phase.rs:22-73 (proxy / status), :101-147 (purge_some_contact_pairs),
//   :149-165 (update_activation_states), :174-259 (update), :262-347 (proxy / create_proxy / remove /
//   deferred_set_bounding_volume), partitioning/dbvt.rs:158-328 (insert / remove), partitioning/bvh.rs:24-45 (visit),
//   and the `slab` crate's LIFO slot reuse (handles are slab keys).
// Events are recorded instead of calling a handler: started (proxy1 = the re-inserted leaf, proxy2 = the leaf it met),
// stopped (SortedPair order), removed (pairs dropped by remove()).
#include <algorithm>
#include <deque>
#include <unordered_map>
#include <vector>
#include "na.hpp"
#include "oracle.h"

namespace orc {

struct BBox {
    V3 mins, maxs;
};
static inline bool bb_intersects(const BBox& a, const BBox& b) {
    return a.mins.x <= b.maxs.x && a.mins.y <= b.maxs.y && a.mins.z <= b.maxs.z && a.maxs.x >= b.mins.x && a.maxs.y >= b.mins.y &&
           a.maxs.z >= b.mins.z;
}
static inline bool bb_contains(const BBox& a, const BBox& b) {  // aabb.rs:161-163
    return a.mins.x <= b.mins.x && a.mins.y <= b.mins.y && a.mins.z <= b.mins.z && a.maxs.x >= b.maxs.x && a.maxs.y >= b.maxs.y &&
           a.maxs.z >= b.maxs.z;
}
static inline BBox bb_loosened(const BBox& a, real m) { return {a.mins + v3(-m, -m, -m), a.maxs + v3(m, m, m)}; }
static inline BBox bb_merged(const BBox& a, const BBox& b) { return {inf(a.mins, b.mins), sup(a.maxs, b.maxs)}; }
static inline V3 bb_center(const BBox& a) { return (a.mins + a.maxs) * real(0.5); }

// slab::Slab<T>: vacant entries form a LIFO free list
template <typename T>
struct Slab {
    std::vector<T> items;
    std::vector<int64_t> next_free;  // -2 = occupied, otherwise the next vacant index (or items.size() sentinel)
    size_t next = 0, len = 0;
    size_t insert(const T& v) {
        size_t key = next;
        if (key == items.size()) {
            items.push_back(v);
            next_free.push_back(-2);
            next = key + 1;
        } else {
            next = (size_t)next_free[key];
            items[key] = v;
            next_free[key] = -2;
        }
        len++;
        return key;
    }
    void remove(size_t key) {
        next_free[key] = (int64_t)next;
        next = key;
        len--;
    }
    bool contains(size_t key) const { return key < items.size() && next_free[key] == -2; }
    void clear() {
        items.clear();
        next_free.clear();
        next = 0;
        len = 0;
    }
    T& operator[](size_t k) { return items[k]; }
    const T& operator[](size_t k) const { return items[k]; }
};

enum ParentKind { P_ROOT, P_LEFT, P_RIGHT };
struct ParentRef {
    ParentKind kind = P_ROOT;
    size_t id = 0;
};
struct NodeId {
    bool leaf = true;
    size_t id = 0;
};

struct Dbvt {
    struct Leaf {
        BBox bv;
        V3 center;
        uint32_t data;
        ParentRef parent;
    };
    struct Internal {
        BBox bv;
        V3 center;
        NodeId left, right;
        ParentRef parent;
    };
    Slab<Leaf> leaves;
    Slab<Internal> internals;
    NodeId root;
    bool empty() const { return leaves.len == 0; }

    V3 center_of(NodeId n) const { return n.leaf ? leaves[n.id].center : internals[n.id].center; }

    size_t insert(const BBox& bv, uint32_t data) {  // dbvt.rs:158-255
        Leaf leaf{bv, bb_center(bv), data, ParentRef{}};
        if (empty()) {
            size_t id = leaves.insert(leaf);
            leaves[id].parent = ParentRef{P_ROOT, 0};
            root = NodeId{true, id};
            return id;
        }
        if (!root.leaf) {
            NodeId curr = root;
            for (;;) {
                if (!curr.leaf) {
                    Internal& n = internals[curr.id];
                    n.bv = bb_merged(n.bv, leaf.bv);
                    NodeId l = n.left, r = n.right;
                    real d1 = norm_squared(center_of(l) - leaf.center);
                    real d2 = norm_squared(center_of(r) - leaf.center);
                    curr = d1 < d2 ? l : r;
                } else {
                    size_t id = curr.id;
                    BBox pbv = bb_merged(leaves[id].bv, leaf.bv);
                    ParentRef gp = leaves[id].parent;
                    size_t new_id = leaves.insert(leaf);
                    Internal parent{pbv, bb_center(pbv), curr, NodeId{true, new_id}, gp};
                    size_t pid = internals.insert(parent);
                    leaves[id].parent = ParentRef{P_LEFT, pid};
                    leaves[new_id].parent = ParentRef{P_RIGHT, pid};
                    if (gp.kind == P_LEFT)
                        internals[gp.id].left = NodeId{false, pid};
                    else if (gp.kind == P_RIGHT)
                        internals[gp.id].right = NodeId{false, pid};
                    return new_id;
                }
            }
        }
        size_t id = root.id;
        size_t new_id = leaves.insert(leaf);
        BBox rbv = bb_merged(leaves[id].bv, leaves[new_id].bv);
        Internal r{rbv, bb_center(rbv), NodeId{true, id}, NodeId{true, new_id}, ParentRef{P_ROOT, 0}};
        size_t rid = internals.insert(r);
        leaves[id].parent = ParentRef{P_LEFT, rid};
        leaves[new_id].parent = ParentRef{P_RIGHT, rid};
        root = NodeId{false, rid};
        return new_id;
    }

    Leaf remove(size_t leaf_id) {  // dbvt.rs:260-328
        Leaf leaf = leaves[leaf_id];
        leaves.remove(leaf_id);
        if (leaf.parent.kind != P_ROOT) {
            size_t p = leaf.parent.id;
            NodeId other = leaf.parent.kind == P_RIGHT ? internals[p].left : internals[p].right;
            ParentRef pp = internals[p].parent;
            if (pp.kind == P_ROOT) {
                if (other.leaf)
                    leaves[other.id].parent = ParentRef{P_ROOT, 0};
                else
                    internals[other.id].parent = ParentRef{P_ROOT, 0};
                root = other;
            } else {
                if (other.leaf)
                    leaves[other.id].parent = pp;
                else
                    internals[other.id].parent = pp;
                if (pp.kind == P_RIGHT)
                    internals[pp.id].right = other;
                else
                    internals[pp.id].left = other;
            }
            internals.remove(p);
        } else {
            leaves.clear();
            internals.clear();
        }
        return leaf;
    }

    void visit(const BBox& q, std::vector<uint32_t>& collector, std::vector<NodeId>& stack) const {  // bvh.rs:24-45
        if (empty()) return;
        stack.clear();
        stack.push_back(root);
        while (!stack.empty()) {
            NodeId node = stack.back();
            stack.pop_back();
            if (!node.leaf) {
                const Internal& n = internals[node.id];
                if (bb_intersects(n.bv, q)) {
                    stack.push_back(n.left);
                    stack.push_back(n.right);
                }
            } else {
                const Leaf& l = leaves[node.id];
                if (bb_intersects(l.bv, q)) collector.push_back(l.data);
            }
        }
    }
    // the same walk with any bounding-volume predicate (ray / point collectors, query/visitors/*.rs)
    template <typename Pred>
    void visit_pred(const Pred& pred, std::vector<uint32_t>& collector, std::vector<NodeId>& stack) const {
        if (empty()) return;
        stack.clear();
        stack.push_back(root);
        while (!stack.empty()) {
            NodeId node = stack.back();
            stack.pop_back();
            if (!node.leaf) {
                const Internal& n = internals[node.id];
                if (pred(n.bv)) {
                    stack.push_back(n.left);
                    stack.push_back(n.right);
                }
            } else {
                const Leaf& l = leaves[node.id];
                if (pred(l.bv)) collector.push_back(l.data);
            }
        }
    }
};

// AABB::toi_with_ray(.., solid = true).is_some()  (query/ray/ray_aabb.rs:13-50, ray.rs:157-159)
static bool bb_intersects_ray(const BBox& b, V3 o, V3 d, real max_toi) {
    real tmin = 0, tmax = max_toi;
    const real oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    const real mn[3] = {b.mins.x, b.mins.y, b.mins.z}, mx[3] = {b.maxs.x, b.maxs.y, b.maxs.z};
    for (int i = 0; i < 3; ++i) {
        if (dd[i] == real(0)) {
            if (oo[i] < mn[i] || oo[i] > mx[i]) return false;
        } else {
            real denom = real(1) / dd[i];
            real near = (mn[i] - oo[i]) * denom, far = (mx[i] - oo[i]) * denom;
            if (near > far) std::swap(near, far);
            tmin = std::fmax(tmin, near);
            tmax = std::fmin(tmax, far);
            if (tmin > tmax) return false;
        }
    }
    return true;
}
// AABB::contains_local_point (bounding_volume/aabb.rs:138-146)
static bool bb_contains_point(const BBox& b, V3 p) {
    if (p.x < b.mins.x || p.x > b.maxs.x) return false;
    if (p.y < b.mins.y || p.y > b.maxs.y) return false;
    if (p.z < b.mins.z || p.z > b.maxs.z) return false;
    return true;
}

enum StatusKind { ON_STATIC, ON_DYNAMIC, DETACHED, DELETED };
struct Proxy {
    StatusKind status = DETACHED;
    size_t leaf = 0;       // tree leaf id (ON_*), or index in leaves_to_update (DETACHED with has_slot)
    size_t energy = 0;
    bool has_slot = false;
    bool updated = true;
};

struct PendingLeaf {
    BBox bv;
    uint32_t handle;
};

static const size_t DEACTIVATION_THRESHOLD = 100;

}  // namespace orc

using namespace orc;

struct orc_bp {
    real margin;
    Slab<Proxy> proxies;
    Dbvt tree, stree;
    std::unordered_map<uint64_t, bool> pairs;  // key = lo << 32 | hi
    bool purge_all = false;
    std::vector<PendingLeaf> leaves_to_update;
    std::deque<std::pair<uint32_t, BBox>> proxies_to_update;
    std::vector<uint32_t> collector;
    std::vector<NodeId> stack;
    const BBox& bv_of(const Proxy& p) const { return p.status == ON_STATIC ? stree.leaves[p.leaf].bv : tree.leaves[p.leaf].bv; }
};

static bool allowed(const uint32_t* g, uint32_t a, uint32_t b) {
    if (a == b) return false;
    if (!g) return true;
    uint32_t m1 = g[3 * a], w1 = g[3 * a + 1], b1 = g[3 * a + 2];
    uint32_t m2 = g[3 * b], w2 = g[3 * b + 1], b2 = g[3 * b + 2];
    return (m1 & b2) == 0 && (m2 & b1) == 0 && (m1 & w2) != 0 && (m2 & w1) != 0;
}

extern "C" {

orc_bp* orc_bp_create(real margin) {
    orc_bp* b = new orc_bp;
    b->margin = margin;
    return b;
}
void orc_bp_destroy(orc_bp* b) { delete b; }

uint32_t orc_bp_create_proxy(orc_bp* b, const real* mm) {  // :275-280
    uint32_t h = (uint32_t)b->proxies.insert(Proxy{});
    b->proxies_to_update.push_back({h, BBox{{mm[0], mm[1], mm[2]}, {mm[3], mm[4], mm[5]}}});
    return h;
}

// returns -1 on "Attempting to set the bounding volume of an object that does not exist." (a panic in the reference)
int orc_bp_set_bounding_volume(orc_bp* b, uint32_t h, const real* mm) {  // :325-347
    if (!b->proxies.contains(h)) return -1;
    const Proxy& p = b->proxies[h];
    BBox bv{{mm[0], mm[1], mm[2]}, {mm[3], mm[4], mm[5]}};
    bool needs_update = true;
    if (p.status == ON_STATIC || p.status == ON_DYNAMIC) needs_update = !bb_contains(b->bv_of(p), bv);
    if (needs_update) b->proxies_to_update.push_back({h, bb_loosened(bv, b->margin)});
    return 0;
}

// :282-323.  removed_pairs (optional): (pair.0, pair.1) of every pair dropped, in SortedPair order.
int orc_bp_remove(orc_bp* b, uint32_t n, const uint32_t* handles, uint32_t* removed_pairs, uint64_t cap, uint64_t* n_removed) {
    for (uint32_t i = 0; i < n; ++i) {
        if (!b->proxies.contains(handles[i])) return -1;  // "Attempting to remove an object that does not exist."
        Proxy& p = b->proxies[handles[i]];
        if (p.status == ON_STATIC)
            b->stree.remove(p.leaf);
        else if (p.status == ON_DYNAMIC)
            b->tree.remove(p.leaf);
        p.status = DELETED;
    }
    uint64_t nr = 0;
    for (auto it = b->pairs.begin(); it != b->pairs.end();) {
        uint32_t lo = (uint32_t)(it->first >> 32), hi = (uint32_t)it->first;
        if (b->proxies[lo].status == DELETED || b->proxies[hi].status == DELETED) {
            if (removed_pairs && nr < cap) removed_pairs[2 * nr] = lo, removed_pairs[2 * nr + 1] = hi;
            nr++;
            it = b->pairs.erase(it);
        } else
            ++it;
    }
    if (n_removed) *n_removed = nr;
    for (uint32_t i = 0; i < n; ++i) b->proxies.remove(handles[i]);
    return 0;
}

// :174-259.  groups: 3 u32 per handle (indexed by handle) or NULL.
void orc_bp_update(orc_bp* b, const uint32_t* groups, uint32_t* started, uint64_t cap_s, uint64_t* n_started, uint32_t* stopped, uint64_t cap_p,
                   uint64_t* n_stopped) {
    uint64_t ns = 0, np = 0;
    while (!b->proxies_to_update.empty()) {
        auto item = b->proxies_to_update.front();
        b->proxies_to_update.pop_front();
        uint32_t h = item.first;
        if (!b->proxies.contains(h)) continue;
        Proxy& p = b->proxies[h];
        bool set_status = true;
        if (p.status == ON_STATIC) {
            b->stree.remove(p.leaf);
            b->leaves_to_update.push_back({item.second, h});
        } else if (p.status == ON_DYNAMIC) {
            b->tree.remove(p.leaf);
            b->leaves_to_update.push_back({item.second, h});
        } else if (p.status == DETACHED && !p.has_slot) {
            b->leaves_to_update.push_back({item.second, h});
        } else if (p.status == DETACHED) {
            b->leaves_to_update[p.leaf] = {item.second, h};
            set_status = false;
        }
        p.updated = true;
        if (set_status) {
            p.status = DETACHED;
            p.has_slot = true;
            p.leaf = b->leaves_to_update.size() - 1;
        }
    }
    bool some_leaves_updated = !b->leaves_to_update.empty();
    for (const PendingLeaf& leaf : b->leaves_to_update) {
        b->collector.clear();
        b->tree.visit(leaf.bv, b->collector, b->stack);
        b->stree.visit(leaf.bv, b->collector, b->stack);
        for (uint32_t k2 : b->collector) {
            if (allowed(groups, leaf.handle, k2)) {
                uint32_t lo = std::min(leaf.handle, k2), hi = std::max(leaf.handle, k2);
                uint64_t key = ((uint64_t)lo << 32) | hi;
                auto it = b->pairs.find(key);
                if (it != b->pairs.end())
                    it->second = true;
                else {
                    if (started && ns < cap_s) started[2 * ns] = leaf.handle, started[2 * ns + 1] = k2;
                    ns++;
                    b->pairs.emplace(key, true);
                }
            }
        }
        Proxy& p1 = b->proxies[leaf.handle];
        size_t id = b->tree.insert(leaf.bv, leaf.handle);
        p1.status = ON_DYNAMIC;
        p1.leaf = id;
        p1.energy = DEACTIVATION_THRESHOLD;
        p1.has_slot = false;
    }
    b->leaves_to_update.clear();
    if (some_leaves_updated) {  // purge_some_contact_pairs :101-147
        for (auto it = b->pairs.begin(); it != b->pairs.end();) {
            bool retain = true;
            if (b->purge_all || !it->second) {
                it->second = true;
                uint32_t lo = (uint32_t)(it->first >> 32), hi = (uint32_t)it->first;
                const Proxy &p1 = b->proxies[lo], &p2 = b->proxies[hi];
                if (b->purge_all || p1.updated || p2.updated) {
                    bool keep = allowed(groups, lo, hi) && bb_intersects(b->bv_of(p1), b->bv_of(p2));
                    if (!keep) {
                        if (stopped && np < cap_p) stopped[2 * np] = lo, stopped[2 * np + 1] = hi;
                        np++;
                        retain = false;
                    }
                }
            }
            it->second = false;
            if (retain)
                ++it;
            else
                it = b->pairs.erase(it);
        }
    }
    // update_activation_states :149-165
    for (size_t k = 0; k < b->proxies.items.size(); ++k) {
        if (!b->proxies.contains(k)) continue;
        Proxy& p = b->proxies[k];
        if (p.status == ON_DYNAMIC) {
            if (p.energy == 1) {
                Dbvt::Leaf old = b->tree.remove(p.leaf);
                p.leaf = b->stree.insert(old.bv, old.data);
                p.status = ON_STATIC;
            } else
                p.energy -= 1;
        }
    }
    if (n_started) *n_started = ns;
    if (n_stopped) *n_stopped = np;
}

// :349-363.  Attached proxies are re-queued at the FRONT of the pending queue with their stored box.
void orc_bp_recompute_with(orc_bp* b, uint32_t h) {
    if (!b->proxies.contains(h)) return;
    const Proxy& p = b->proxies[h];
    if (p.status != ON_STATIC && p.status != ON_DYNAMIC) return;
    b->proxies_to_update.push_front({h, b->bv_of(p)});
}

// :365-386.  Every attached proxy, in slab order, each pushed in front of the previous one; purge_all stays set.
void orc_bp_recompute_all(orc_bp* b) {
    for (size_t k = 0; k < b->proxies.items.size(); ++k) {
        if (!b->proxies.contains(k)) continue;
        const Proxy& p = b->proxies[k];
        if (p.status != ON_STATIC && p.status != ON_DYNAMIC) continue;
        b->proxies_to_update.push_front({(uint32_t)k, b->bv_of(p)});
    }
    b->purge_all = true;
}

// interferences_with_bounding_volume / _ray / _point (:388-432): dynamic tree first, then the static tree.
// kind 0: q = 6 reals (mins, maxs); 1: q = 7 reals (origin, dir, max_toi); 2: q = 3 reals.  Returns the count.
uint64_t orc_bp_query(orc_bp* b, int kind, const real* q, uint32_t* out, uint64_t cap) {
    b->collector.clear();
    if (kind == 0) {
        BBox bv{{q[0], q[1], q[2]}, {q[3], q[4], q[5]}};
        auto pred = [&](const BBox& x) { return bb_intersects(x, bv); };
        b->tree.visit_pred(pred, b->collector, b->stack);
        b->stree.visit_pred(pred, b->collector, b->stack);
    } else if (kind == 1) {
        V3 o = v3(q[0], q[1], q[2]), d = v3(q[3], q[4], q[5]);
        real max_toi = q[6];
        auto pred = [&](const BBox& x) { return bb_intersects_ray(x, o, d, max_toi); };
        b->tree.visit_pred(pred, b->collector, b->stack);
        b->stree.visit_pred(pred, b->collector, b->stack);
    } else {
        V3 p = v3(q[0], q[1], q[2]);
        auto pred = [&](const BBox& x) { return bb_contains_point(x, p); };
        b->tree.visit_pred(pred, b->collector, b->stack);
        b->stree.visit_pred(pred, b->collector, b->stack);
    }
    for (size_t k = 0; k < b->collector.size() && k < cap; ++k) out[k] = b->collector[k];
    return b->collector.size();
}

uint64_t orc_bp_num_interferences(const orc_bp* b) { return b->pairs.size(); }

// BroadPhase::proxy (:262-273): 1 + box when the proxy is attached to a tree, else 0
int orc_bp_proxy(const orc_bp* b, uint32_t h, real* mm) {
    if (!b->proxies.contains(h)) return 0;
    const Proxy& p = b->proxies[h];
    if (p.status != ON_STATIC && p.status != ON_DYNAMIC) return 0;
    const BBox& bv = b->bv_of(p);
    mm[0] = bv.mins.x, mm[1] = bv.mins.y, mm[2] = bv.mins.z, mm[3] = bv.maxs.x, mm[4] = bv.maxs.y, mm[5] = bv.maxs.z;
    return 1;
}

uint64_t orc_bp_pairs(const orc_bp* b, uint32_t* out, uint64_t cap) {
    uint64_t n = 0;
    for (auto& kv : b->pairs) {
        if (out && n < cap) out[2 * n] = (uint32_t)(kv.first >> 32), out[2 * n + 1] = (uint32_t)kv.first;
        n++;
    }
    return n;
}

}  // extern "C"
