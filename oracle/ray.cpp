// ORACLE — TEST INFRASTRUCTURE ONLY (see na.hpp header).
// TriMesh ray casting restated from the reference:
//   shape/trimesh.rs:100-144 (per-triangle local AABB + BVT::new_balanced), bounding_volume/aabb_triangle.rs:27-41,
//   partitioning/bvt.rs:281-404 (median partitioning, post-order node layout), utils/median.rs:5-17,
//   partitioning/bvh.rs:101-160 (best_first_search, Rust BinaryHeap), query/ray/ray_trimesh.rs:22-50,150-190,
//   query/ray/ray_aabb.rs:13-50, query/ray/ray_triangle.rs:9-25,32-114, query/ray/ray.rs:36-41.
// PARITY UNPINNED: the reference has no test at all for TriMesh / triangle ray casting.
#include <algorithm>
#include <cstring>
#include <vector>
#include "na.hpp"
#include "oracle.h"

namespace orc {

struct Box {
    V3 mins, maxs;
};

// ray_aabb.rs:13-50 with m = identity
static bool aabb_toi_with_ray(const Box& b, V3 origin, V3 dir, real max_toi, bool solid, real* toi) {
    real tmin = 0, tmax = max_toi;
    for (int i = 0; i < 3; ++i) {
        if (dir[i] == 0) {
            if (origin[i] < b.mins[i] || origin[i] > b.maxs[i]) return false;
        } else {
            real denom = real(1) / dir[i];
            real near = (b.mins[i] - origin[i]) * denom;
            real far = (b.maxs[i] - origin[i]) * denom;
            if (near > far) std::swap(near, far);
            tmin = std::fmax(tmin, near);
            tmax = std::fmin(tmax, far);
            if (tmin > tmax) return false;
        }
    }
    *toi = (tmin == 0 && !solid) ? tmax : tmin;
    return true;
}

// ray_triangle.rs:32-114.  fid: 0 front, 1 back.
static bool ray_triangle(V3 a, V3 b, V3 c, V3 origin, V3 dir, real* toi, V3* normal, int* fid, V3* bary = nullptr) {
    V3 ab = b - a, ac = c - a;
    V3 n = cross(ab, ac);
    real d = dot(n, dir);
    if (d == 0) return false;
    V3 ap = origin - a;
    real t = dot(ap, n);
    if ((t < 0 && d < 0) || (t > 0 && d > 0)) return false;
    *fid = d < 0 ? 0 : 1;
    d = std::fabs(d);
    V3 e = -cross(dir, ap);
    real v, w;
    if (t < 0) {
        v = -dot(ac, e);
        if (v < 0 || v > d) return false;
        w = dot(ab, e);
        if (w < 0 || v + w > d) return false;
        real invd = real(1) / d;
        *toi = -t * invd;
        *normal = -normalize(n);
        v = v * invd;
        w = w * invd;
    } else {
        v = dot(ac, e);
        if (v < 0 || v > d) return false;
        w = -dot(ab, e);
        if (w < 0 || v + w > d) return false;
        real invd = real(1) / d;
        *toi = t * invd;
        *normal = normalize(n);
        v = v * invd;
        w = w * invd;
    }
    if (bary) *bary = v3(-v - w + real(1), v, w);  // ray_triangle.rs:113
    return true;
}

struct BVT {
    // node ids: >= 0 internal index, < 0 leaf ~index
    struct Internal {
        Box bv;
        int32_t left, right;
    };
    struct Leaf {
        Box bv;
        uint32_t data;
    };
    std::vector<Internal> internals;
    std::vector<Leaf> leaves;
    int32_t root = -1;
    bool has_root = false;
};

typedef std::pair<uint32_t, Box> Elt;

// bvt.rs:290-350 + :364-404
static int32_t bvt_build(int depth, std::vector<Elt>& elts, BVT& out) {
    if (elts.size() == 1) {
        out.leaves.push_back({elts[0].second, elts[0].first});
        return ~(int32_t)(out.leaves.size() - 1);
    }
    int sep_axis = depth % 3;
    std::vector<real> med;
    med.reserve(elts.size());
    for (auto& l : elts) med.push_back(((l.second.mins + l.second.maxs) * real(0.5))[sep_axis]);
    std::sort(med.begin(), med.end());  // values only: stability is irrelevant
    size_t n = med.size();
    real median = (n % 2 == 0) ? (med[n / 2 - 1] + med[n / 2]) / real(2) : med[n / 2];
    std::vector<Elt> left, right;
    Box bb = elts[0].second;
    bool insert_left = false;
    for (auto& l : elts) {
        bb.mins = inf(bb.mins, l.second.mins);
        bb.maxs = sup(bb.maxs, l.second.maxs);
        real pos = ((l.second.mins + l.second.maxs) * real(0.5))[sep_axis];
        if (pos < median || (pos == median && insert_left)) {
            left.push_back(l);
            insert_left = false;
        } else {
            right.push_back(l);
            insert_left = true;
        }
    }
    if (left.empty()) {
        left.push_back(right.back());
        right.pop_back();
    } else if (right.empty()) {
        right.push_back(left.back());
        left.pop_back();
    }
    std::vector<Elt>().swap(elts);
    int32_t l = bvt_build(depth + 1, left, out);
    int32_t r = bvt_build(depth + 1, right, out);
    out.internals.push_back({bb, l, r});
    return (int32_t)(out.internals.size() - 1);
}

struct Weighted {
    int32_t node;
    real cost;  // stored as -cost in the reference; here `cost` IS the heap key (i.e. -toi)
};
static inline bool w_le(const Weighted& a, const Weighted& b) { return a.cost <= b.cost; }
struct Heap {  // Rust BinaryHeap<WeightedValue>
    std::vector<Weighted> data;
    void sift_up(size_t start, size_t pos) {
        Weighted elt = data[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (w_le(elt, data[parent])) break;
            data[pos] = data[parent];
            pos = parent;
        }
        data[pos] = elt;
    }
    void push(Weighted w) {
        data.push_back(w);
        sift_up(0, data.size() - 1);
    }
    bool pop(Weighted* out) {
        if (data.empty()) return false;
        Weighted item = data.back();
        data.pop_back();
        if (!data.empty()) {
            std::swap(item, data[0]);
            size_t end = data.size(), pos = 0, child = 1;
            Weighted elt = data[0];
            while (end >= 2 && child <= end - 2) {
                if (w_le(data[child], data[child + 1])) child += 1;
                data[pos] = data[child];
                pos = child;
                child = 2 * pos + 1;
            }
            if (child == end - 1) {
                data[pos] = data[child];
                pos = child;
            }
            data[pos] = elt;
            sift_up(0, pos);
        }
        *out = item;
        return true;
    }
};

}  // namespace orc

using namespace orc;

struct orc_trimesh {
    std::vector<V3> verts;
    std::vector<uint32_t> idx;
    std::vector<Box> tri_box;
    BVT bvt;
    uint32_t ntris;
};

struct Hit {
    bool some = false;
    uint32_t tri = 0;
    real toi = 0;
    V3 normal = {0, 0, 0};
    int fid = 0;
    V3 bary = {0, 0, 0};  // TriMeshRayToiAndNormalAndUVsVisitor's third result (ray_trimesh.rs:199-240)
};

// TriMeshRayToiAndNormalVisitor::visit (ray_trimesh.rs:156-190)
static int visit(const orc_trimesh* m, real best, const Box& bv, const uint32_t* data, V3 o, V3 d, real max_toi, real* cost, Hit* result) {
    real toi;
    if (!aabb_toi_with_ray(bv, o, d, max_toi, true, &toi)) return 0;  // Stop
    *cost = toi;
    result->some = false;
    if (data && toi < best) {
        uint32_t t = *data;
        V3 a = m->verts[m->idx[3 * t]], b = m->verts[m->idx[3 * t + 1]], c = m->verts[m->idx[3 * t + 2]];
        real ttoi;
        V3 n;
        int f;
        V3 bary;
        if (ray_triangle(a, b, c, o, d, &ttoi, &n, &f, &bary) && ttoi <= max_toi) {
            *cost = ttoi;
            result->some = true;
            result->tri = t;
            result->toi = ttoi;
            result->normal = n;
            result->fid = f;
            result->bary = bary;
        }
    }
    return 1;  // Continue
}

// BVH::best_first_search (bvh.rs:101-160)
static Hit best_first(const orc_trimesh* m, V3 o, V3 d, real max_toi, Heap& queue) {
    Hit best_result;
    const BVT& t = m->bvt;
    if (!t.has_root) return best_result;
    queue.data.clear();
    real best_cost = FMAX;
    auto content = [&](int32_t node, const Box** bv, const uint32_t** data) {
        if (node >= 0) {
            *bv = &t.internals[node].bv;
            *data = nullptr;
        } else {
            *bv = &t.leaves[~node].bv;
            *data = &t.leaves[~node].data;
        }
    };
    const Box* bv;
    const uint32_t* data;
    content(t.root, &bv, &data);
    real cost;
    Hit res;
    if (!visit(m, best_cost, *bv, data, o, d, max_toi, &cost, &res)) return best_result;
    if (res.some) {
        best_cost = cost;
        best_result = res;
    }
    queue.push({t.root, -cost});
    Weighted entry;
    while (queue.pop(&entry)) {
        if (-entry.cost >= best_cost) break;
        if (entry.node < 0) continue;  // leaves have no children
        for (int i = 0; i < 2; ++i) {
            int32_t child = i == 0 ? t.internals[entry.node].left : t.internals[entry.node].right;
            content(child, &bv, &data);
            if (visit(m, best_cost, *bv, data, o, d, max_toi, &cost, &res)) {
                if (cost < best_cost) {
                    if (res.some) {
                        best_cost = cost;
                        best_result = res;
                    }
                    queue.push({child, -cost});
                }
            }
        }
    }
    return best_result;
}

extern "C" {

orc_trimesh* orc_trimesh_create(uint32_t n_verts, const real* xyz, uint32_t n_tris, const uint32_t* idx) {
    orc_trimesh* m = new orc_trimesh;
    m->verts.resize(n_verts);
    for (uint32_t i = 0; i < n_verts; ++i) m->verts[i] = v3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    m->idx.assign(idx, idx + 3 * (size_t)n_tris);
    m->ntris = n_tris;
    std::vector<Elt> leaves;
    leaves.reserve(n_tris);
    m->tri_box.resize(n_tris);
    for (uint32_t t = 0; t < n_tris; ++t) {
        V3 a = m->verts[idx[3 * t]], b = m->verts[idx[3 * t + 1]], c = m->verts[idx[3 * t + 2]];
        Box bx;
        for (int k = 0; k < 3; ++k) {  // aabb_triangle.rs:27-41
            bx.mins[k] = std::fmin(std::fmin(a[k], b[k]), c[k]);
            bx.maxs[k] = std::fmax(std::fmax(a[k], b[k]), c[k]);
        }
        m->tri_box[t] = bx;
        leaves.push_back({t, bx});
    }
    if (n_tris) {
        m->bvt.internals.reserve(n_tris);
        m->bvt.leaves.reserve(n_tris);
        m->bvt.root = bvt_build(0, leaves, m->bvt);
        m->bvt.has_root = true;
    }
    return m;
}
void orc_trimesh_destroy(orc_trimesh* m) { delete m; }

// toi_and_normal_and_uv_with_ray (ray_trimesh.rs:52-94) for a batch; uvs == NULL: toi_and_normal_with_ray (the reference's own
// fall-back, :59-61).  max_tois: one max_toi per ray (NULL: max_toi).  uv_out: 2 reals per ray.
void orc_trimesh_ray_cast_uv(const orc_trimesh* m, const real* pose, uint64_t n_rays, const real* origins, const real* dirs, real max_toi_all,
                             const real* max_tois, const real* uvs, int mode, real* toi, uint32_t* face, real* normal, real* uv_out) {
    Iso iso = iso_identity();
    if (pose) iso = Iso{{pose[0], pose[1], pose[2]}, {pose[3], pose[4], pose[5], pose[6]}};
    Heap queue;
    for (uint64_t r = 0; r < n_rays; ++r) {
        V3 o = v3(origins[3 * r], origins[3 * r + 1], origins[3 * r + 2]);
        V3 d = v3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
        // ray.inverse_transform_by(m) (ray.rs:36-41)
        V3 lo = iso_inv_point(iso, o), ld = iso_inv_vec(iso, d);
        real max_toi = max_tois ? max_tois[r] : max_toi_all;
        Hit h;
        if (mode == 0) {
            h = best_first(m, lo, ld, max_toi, queue);
        } else {
            for (uint32_t t = 0; t < m->ntris; ++t) {
                real bt;
                if (!aabb_toi_with_ray(m->tri_box[t], lo, ld, max_toi, true, &bt)) continue;
                V3 a = m->verts[m->idx[3 * t]], b = m->verts[m->idx[3 * t + 1]], c = m->verts[m->idx[3 * t + 2]];
                real tt;
                V3 n;
                int f;
                V3 bary;
                if (ray_triangle(a, b, c, lo, ld, &tt, &n, &f, &bary) && tt <= max_toi) {
                    if (!h.some || tt < h.toi) {
                        h.some = true;
                        h.tri = t;
                        h.toi = tt;
                        h.normal = n;
                        h.fid = f;
                        h.bary = bary;
                    }
                }
            }
        }
        if (h.some) {
            toi[r] = h.toi;
            face[r] = h.fid == 1 ? h.tri + m->ntris : h.tri;  // ray_trimesh.rs:41-45
            V3 wn = iso_mul_vec(iso, h.normal);                // :47
            if (normal) normal[3 * r] = wn.x, normal[3 * r + 1] = wn.y, normal[3 * r + 2] = wn.z;
            if (uv_out) {
                real ux = 0, uy = 0;
                if (uvs) {  // ray_trimesh.rs:76-84
                    uint32_t i0 = m->idx[3 * h.tri], i1 = m->idx[3 * h.tri + 1], i2 = m->idx[3 * h.tri + 2];
                    ux = uvs[2 * i0] * h.bary.x + uvs[2 * i1] * h.bary.y + uvs[2 * i2] * h.bary.z;
                    uy = uvs[2 * i0 + 1] * h.bary.x + uvs[2 * i1 + 1] * h.bary.y + uvs[2 * i2 + 1] * h.bary.z;
                }
                uv_out[2 * r] = ux, uv_out[2 * r + 1] = uy;
            }
        } else {
            toi[r] = -1;
            face[r] = 0xffffffffu;
            if (normal) normal[3 * r] = normal[3 * r + 1] = normal[3 * r + 2] = 0;
            if (uv_out) uv_out[2 * r] = uv_out[2 * r + 1] = 0;
        }
    }
}
void orc_trimesh_ray_cast(const orc_trimesh* m, const real* pose, uint64_t n_rays, const real* origins, const real* dirs, real max_toi,
                          int mode, real* toi, uint32_t* face, real* normal) {
    orc_trimesh_ray_cast_uv(m, pose, n_rays, origins, dirs, max_toi, nullptr, nullptr, mode, toi, face, normal, nullptr);
}

void orc_aabb_toi_with_ray(const real* mm, const real* origin, const real* dir, real max_toi, int solid, real* toi) {
    Box b = {{mm[0], mm[1], mm[2]}, {mm[3], mm[4], mm[5]}};
    real t;
    if (aabb_toi_with_ray(b, v3(origin[0], origin[1], origin[2]), v3(dir[0], dir[1], dir[2]), max_toi, solid != 0, &t))
        *toi = t;
    else
        *toi = -1;
}

}  // extern "C"
