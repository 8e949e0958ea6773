// ORACLE — TEST INFRASTRUCTURE ONLY (see na.hpp header).
// AABB computation and broad phase, restated from the reference:
//   bounding_volume/aabb_ball.rs:8-13, aabb_cuboid.rs:9-14, aabb_convex.rs:9-11 + aabb_utils.rs:59-79,
//   aabb_plane.rs:13-21, aabb.rs:156-199 (intersects / loosen), pipeline/object/collision_object.rs:89-93,
//   pipeline/glue/setup.rs:27-30, pipeline/broad_phase/dbvt_broad_phase.rs:174-259,325-347,
//   partitioning/dbvt.rs:158-255, partitioning/bvh.rs:24-45,
//   query/visitors/bounding_volume_interferences_collector.rs:41-51,
//   pipeline/object/collision_groups.rs:353-359,389-404.
#include <algorithm>
#include <unordered_map>
#include <vector>
#include "oracle.h"
#include "scene.hpp"

namespace orc {

struct AABB {
    V3 mins, maxs;
};

static inline bool aabb_intersects(const AABB& a, const AABB& b) {  // aabb.rs:156-158 (inclusive)
    return a.mins.x <= b.maxs.x && a.mins.y <= b.maxs.y && a.mins.z <= b.maxs.z && a.maxs.x >= b.mins.x &&
           a.maxs.y >= b.mins.y && a.maxs.z >= b.mins.z;
}
static inline void aabb_loosen(AABB& a, real amount) {  // aabb.rs:180-187
    V3 m = {-amount, -amount, -amount}, p = {amount, amount, amount};
    a.mins = a.mins + m;
    a.maxs = a.maxs + p;
}
static inline AABB aabb_merged(const AABB& a, const AABB& b) { return {inf(a.mins, b.mins), sup(a.maxs, b.maxs)}; }
static inline V3 aabb_center(const AABB& a) { return (a.mins + a.maxs) * real(0.5); }  // na::center

AABB shape_aabb(const Objects& o, uint32_t i) {
    Iso m = o.iso(i);
    switch (o.shape_type[i]) {
        case BALL: {
            real r = o.shape_param[4 * i];
            V3 c = m.t;
            return {c + v3(-r, -r, -r), c + v3(r, r, r)};
        }
        case CUBOID: {
            V3 he = absolute_transform_vector(m, o.p3(i));
            return {m.t - he, m.t + he};
        }
        case HULL: {
            Hull H = hull_view(o.hulls, o.hull_id(i));
            V3 w0 = iso_mul_point(m, H.pt(0));
            AABB a = {w0, w0};
            for (uint32_t k = 1; k < H.nv; ++k) {
                V3 w = iso_mul_point(m, H.pt(k));
                a.mins = inf(a.mins, w);
                a.maxs = sup(a.maxs, w);
            }
            return a;
        }
        case CAPSULE: {  // aabb_support_map.rs:35-45 -> aabb_utils.rs:9-31 (support points along +-x, +-y, +-z) with capsule.rs:72-85
            real hh = o.shape_param[4 * i], r = o.shape_param[4 * i + 1];
            auto support = [&](V3 dir) {
                V3 ld = iso_inv_vec(m, dir);
                V3 d = normalize(ld);
                return iso_mul_point(m, v3(0, std::copysign(hh, d.y), 0) + d * r);
            };
            AABB a;
            a.maxs = v3(support(v3(1, 0, 0)).x, support(v3(0, 1, 0)).y, support(v3(0, 0, 1)).z);
            a.mins = v3(support(v3(-1, 0, 0)).x, support(v3(0, -1, 0)).y, support(v3(0, 0, -1)).z);
            return a;
        }
        default: {  // PLANE
            real mx = FMAX * real(0.5);
            return {v3(-mx, -mx, -mx), v3(mx, mx, mx)};
        }
    }
}

// Fat AABB a fresh-world update ends up with: ((tight -+ query_limit) -+ margin)
// (collision_object.rs:89-93 then dbvt_broad_phase.rs:341).
AABB fat_aabb(const Objects& o, uint32_t i, real margin) {
    AABB a = shape_aabb(o, i);
    aabb_loosen(a, o.query_limit[i]);
    aabb_loosen(a, margin);
    return a;
}

static inline bool groups_can_interact(const uint32_t* g, uint32_t a, uint32_t b) {  // collision_groups.rs:353-359
    if (!g) return true;
    uint32_t m1 = g[3 * a], w1 = g[3 * a + 1], b1 = g[3 * a + 2];
    uint32_t m2 = g[3 * b], w2 = g[3 * b + 1], b2 = g[3 * b + 2];
    return (m1 & b2) == 0 && (m2 & b1) == 0 && (m1 & w2) != 0 && (m2 & w1) != 0;
}

// ---------------------------------------------------------------------------------------------
// Reference-faithful DBVT (insert + visit only: all a fresh-world update needs).
// ---------------------------------------------------------------------------------------------
struct DBVT {
    struct Leaf {
        AABB bv;
        V3 center;
        uint32_t data;
        int32_t parent;  // internal index, <0 = root
        bool right;
    };
    struct Internal {
        AABB bv;
        V3 center;
        int32_t left, right;  // >=0 internal, <0 => leaf ~id
        int32_t parent;
        bool is_right;
    };
    std::vector<Leaf> leaves;
    std::vector<Internal> internals;
    int32_t root = 0;  // same encoding
    bool empty() const { return leaves.empty(); }

    void insert(const AABB& bv, uint32_t data) {  // dbvt.rs:158-255
        V3 c = aabb_center(bv);
        if (leaves.empty()) {
            leaves.push_back({bv, c, data, -1, false});
            root = ~0;
            return;
        }
        if (root >= 0) {
            int32_t curr = root;
            for (;;) {
                if (curr >= 0) {
                    Internal& n = internals[curr];
                    n.bv = aabb_merged(n.bv, bv);
                    int32_t l = n.left, r = n.right;
                    V3 cl = l >= 0 ? internals[l].center : leaves[~l].center;
                    V3 cr = r >= 0 ? internals[r].center : leaves[~r].center;
                    real d1 = norm_squared(cl - c), d2 = norm_squared(cr - c);
                    curr = d1 < d2 ? l : r;
                } else {
                    int32_t id = ~curr;
                    AABB pbv = aabb_merged(leaves[id].bv, bv);
                    int32_t gp = leaves[id].parent;
                    bool gp_right = leaves[id].right;
                    int32_t new_id = (int32_t)leaves.size();
                    leaves.push_back({bv, c, data, 0, true});
                    int32_t pid = (int32_t)internals.size();
                    internals.push_back({pbv, aabb_center(pbv), curr, ~new_id, gp, gp_right});
                    leaves[id].parent = pid;
                    leaves[id].right = false;
                    leaves[new_id].parent = pid;
                    if (gp_right)
                        internals[gp].right = pid;
                    else
                        internals[gp].left = pid;
                    return;
                }
            }
        } else {
            int32_t id = ~root;
            int32_t new_id = (int32_t)leaves.size();
            leaves.push_back({bv, c, data, 0, true});
            AABB rbv = aabb_merged(leaves[id].bv, leaves[new_id].bv);
            int32_t rid = (int32_t)internals.size();
            internals.push_back({rbv, aabb_center(rbv), ~id, ~new_id, -1, false});
            leaves[id].parent = rid;
            leaves[id].right = false;
            leaves[new_id].parent = rid;
            root = rid;
        }
    }

    // bvh.rs:24-45 with BoundingVolumeInterferencesCollector (visitors/...collector.rs:41-51)
    void visit(const AABB& q, std::vector<uint32_t>& collector, std::vector<int32_t>& stack) const {
        if (leaves.empty()) return;
        stack.clear();
        stack.push_back(root);
        while (!stack.empty()) {
            int32_t node = stack.back();
            stack.pop_back();
            if (node >= 0) {
                const Internal& n = internals[node];
                if (aabb_intersects(n.bv, q)) {
                    stack.push_back(n.left);
                    stack.push_back(n.right);
                }
            } else {
                const Leaf& l = leaves[~node];
                if (aabb_intersects(l.bv, q)) collector.push_back(l.data);
            }
        }
    }
};

struct PairHash {
    size_t operator()(uint64_t k) const {
        k ^= k >> 33;
        k *= 0xff51afd7ed558ccdULL;
        k ^= k >> 33;
        k *= 0xc4ceb9fe1a85ec53ULL;
        k ^= k >> 33;
        return (size_t)k;
    }
};

// DBVTBroadPhase::update on a fresh world (dbvt_broad_phase.rs:174-259): leaves are inserted in handle
// order; each queries the dynamic tree first; pairs go through a hash set keyed by SortedPair; the
// started callback gets (later handle, earlier handle).
void broad_phase_dbvt(uint32_t n, const AABB* fat, const uint32_t* groups, std::vector<uint32_t>& pairs_out) {
    DBVT tree;
    tree.leaves.reserve(n);
    tree.internals.reserve(n);
    std::unordered_map<uint64_t, bool, PairHash> pairs;
    std::vector<uint32_t> collector;
    std::vector<int32_t> stack;
    for (uint32_t i = 0; i < n; ++i) {
        collector.clear();
        tree.visit(fat[i], collector, stack);
        for (uint32_t j : collector) {
            if (groups_can_interact(groups, i, j)) {
                uint32_t lo = std::min(i, j), hi = std::max(i, j);
                uint64_t key = ((uint64_t)lo << 32) | hi;
                auto it = pairs.find(key);
                if (it != pairs.end())
                    it->second = true;
                else {
                    pairs_out.push_back(i);
                    pairs_out.push_back(j);
                    pairs.emplace(key, true);
                }
            }
        }
        tree.insert(fat[i], i);
    }
}

// Tree-independent definition of the fresh-world pair set, by sweep along x.
void broad_phase_sweep(uint32_t n, const AABB* fat, const uint32_t* groups, std::vector<uint32_t>& pairs_out) {
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return fat[a].mins.x < fat[b].mins.x || (fat[a].mins.x == fat[b].mins.x && a < b);
    });
    for (uint32_t a = 0; a < n; ++a) {
        uint32_t i = order[a];
        for (uint32_t b = a + 1; b < n; ++b) {
            uint32_t j = order[b];
            if (fat[j].mins.x > fat[i].maxs.x) break;
            if (aabb_intersects(fat[i], fat[j]) && groups_can_interact(groups, i, j)) {
                pairs_out.push_back(std::max(i, j));
                pairs_out.push_back(std::min(i, j));
            }
        }
    }
}

}  // namespace orc

using namespace orc;

static Objects make_objects(const orc_objects* o) {
    Objects r;
    r.n = o->n;
    r.pos = o->pos;
    r.rot = o->rot;
    r.shape_type = o->shape_type;
    r.shape_param = o->shape_param;
    r.groups = o->groups;
    r.query_limit = o->query_limit;
    r.ang_pred = o->ang_pred;
    r.hulls = reinterpret_cast<const HullLibrary*>(o->hulls);
    r.query_kind = o->query_kind;
    return r;
}

extern "C" {

// mode 0: bounding_volume::aabb(shape, pos); 1: compute_aabb (+ query_limit); 2: fat (+ margin)
void orc_compute_aabbs(const orc_objects* objs, real margin, int mode, real* out_minmax) {
    Objects o = make_objects(objs);
    for (uint32_t i = 0; i < o.n; ++i) {
        AABB a = shape_aabb(o, i);
        if (mode >= 1) aabb_loosen(a, o.query_limit[i]);
        if (mode >= 2) aabb_loosen(a, margin);
        real* d = out_minmax + 6 * (size_t)i;
        d[0] = a.mins.x, d[1] = a.mins.y, d[2] = a.mins.z, d[3] = a.maxs.x, d[4] = a.maxs.y, d[5] = a.maxs.z;
    }
}

// mode 0: reference-faithful DBVT incremental update; mode 1: sweep (tree-independent definition);
// mode 2: O(N^2) brute force.  Pairs are written as (larger handle, smaller handle).
// Returns the number of pairs found (which may exceed cap; only the first cap are written).
uint64_t orc_broad_phase(uint32_t n, const real* aabb_minmax, const uint32_t* groups, int mode, uint32_t* out_pairs,
                         uint64_t cap) {
    std::vector<AABB> fat(n);
    for (uint32_t i = 0; i < n; ++i) {
        const real* s = aabb_minmax + 6 * (size_t)i;
        fat[i] = {{s[0], s[1], s[2]}, {s[3], s[4], s[5]}};
    }
    std::vector<uint32_t> pairs;
    if (mode == 0)
        broad_phase_dbvt(n, fat.data(), groups, pairs);
    else if (mode == 1)
        broad_phase_sweep(n, fat.data(), groups, pairs);
    else {
        for (uint32_t i = 0; i < n; ++i)
            for (uint32_t j = 0; j < i; ++j)
                if (aabb_intersects(fat[i], fat[j]) && groups_can_interact(groups, i, j)) {
                    pairs.push_back(i);
                    pairs.push_back(j);
                }
    }
    uint64_t np = pairs.size() / 2;
    uint64_t w = std::min(np, cap);
    if (out_pairs) std::memcpy(out_pairs, pairs.data(), (size_t)w * 8);
    return np;
}

}  // extern "C"
