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
static bool aabb_toi_with_ray(const Box& b, V3 origin, V3 dir, real max_toi, bool solid, real* toi, int dim = 3) {
    real tmin = 0, tmax = max_toi;
    for (int i = 0; i < dim; ++i) {
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
static int32_t bvt_build(int depth, std::vector<Elt>& elts, BVT& out, int dim = 3) {
    if (elts.size() == 1) {
        out.leaves.push_back({elts[0].second, elts[0].first});
        return ~(int32_t)(out.leaves.size() - 1);
    }
    int sep_axis = depth % dim;  // bvt.rs: depth % DIM
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
    int32_t l = bvt_build(depth + 1, left, out, dim);
    int32_t r = bvt_build(depth + 1, right, out, dim);
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

// BVH::best_first_search (bvh.rs:101-160); visit(best, bv, data, &cost, &res) -> 0 Stop / 1 Continue
template <class Visit>
static Hit best_first_search(const BVT& t, Heap& queue, Visit visit) {
    Hit best_result;
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
    if (!visit(best_cost, *bv, data, &cost, &res)) return best_result;
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
            if (visit(best_cost, *bv, data, &cost, &res)) {
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
static Hit best_first(const orc_trimesh* m, V3 o, V3 d, real max_toi, Heap& queue) {
    return best_first_search(m->bvt, queue, [&](real best, const Box& bv, const uint32_t* data, real* cost, Hit* res) {
        return visit(m, best, bv, data, o, d, max_toi, cost, res);
    });
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

// ---- ncollide2d: Polyline ray casting ---------------------------------------------------------------------------------------------
//   shape/polyline.rs:57-120 (Polyline::new: one leaf per edge, Segment::local_aabb = local_support_map_aabb, BVT::new_balanced),
//   bounding_volume/aabb_utils.rs:34-56, shape/segment.rs:81-125,182-191, partitioning/bvt.rs (depth % DIM with DIM = 2),
//   query/ray/ray_polyline.rs:8-147 (best-first search; Face(1) -> Face(edge + edges.len())), query/ray/ray_support_map.rs:219-293
//   (RayCast for Segment, dim2: line / line parameters; the scaled, NOT normalised, normal; max_toi is never applied to the segment hit),
//   query/closest_points/closest_points_line_line.rs:27-70, query/ray/ray.rs:36-41.
// PINNED on build/ncollide2d/tests/geometry/ray_cast.rs (the ten Segment known-answer tests) — tests/test_rays2d.py.
namespace orc {

struct Seg2Hit {
    bool some = false;
    real toi = 0;
    real nx = 0, ny = 0;
    int kind = 0, id = 0;  // FeatureId: kind 1 Face, 2 Vertex
};

// RayCast for Segment::toi_and_normal_with_ray (dim2) with the segment already in the ray's frame
static Seg2Hit segment_ray(real ax, real ay, real bx, real by, real ox, real oy, real dx, real dy) {
    Seg2Hit h;
    real sdx = bx - ax, sdy = by - ay;  // seg.scaled_direction()
    // closest_points_line_line_parameters_eps(ray.origin, ray.dir, seg.a, seg_dir, eps)
    real rx = ox - ax, ry = oy - ay;
    real a = dx * dx + dy * dy, e = sdx * sdx + sdy * sdy, f = sdx * rx + sdy * ry;
    real s, t;
    bool parallel = false;
    if (a <= EPS && e <= EPS) {
        s = 0, t = 0;
    } else if (a <= EPS) {
        s = 0, t = f / e;
    } else {
        real c = dx * rx + dy * ry;
        if (e <= EPS) {
            s = -c / a, t = 0;
        } else {
            real b = dx * sdx + dy * sdy;
            real ae = a * e, bb = b * b, denom = ae - bb;
            parallel = denom <= EPS || ulps_eq(ae, bb);
            s = !parallel ? (b * f - c * e) / denom : real(0);
            t = (b * s + f) / e;
        }
    }
    real nx = sdy, ny = -sdx;  // seg.scaled_normal()
    if (parallel) {
        real px = ax - ox, py = ay - oy;  // dpos
        if (std::fabs(px * nx + py * ny) < EPS) {  // collinear
            real dist1 = px * dx + py * dy;
            real dist2 = dist1 + (sdx * dx + sdy * dy);
            bool p1 = dist1 >= 0, p2 = dist2 >= 0;
            if (p1 && p2) {
                h.some = true, h.nx = nx, h.ny = ny, h.kind = 2;
                if (dist1 <= dist2)
                    h.toi = dist1 / (dx * dx + dy * dy), h.id = 0;
                else
                    h.toi = dist2 / (dx * dx + dy * dy), h.id = 1;
            } else if (p1 || p2) {  // the ray origin lies on the segment
                h.some = true, h.toi = 0, h.nx = nx, h.ny = ny, h.kind = 1, h.id = 0;
            }
        }
    } else if (s >= 0 && t >= 0 && t <= 1) {
        h.some = true, h.toi = s, h.kind = 1;
        if (nx * dx + ny * dy > 0)
            h.nx = -nx, h.ny = -ny, h.id = 1;
        else
            h.nx = nx, h.ny = ny, h.id = 0;
    }
    return h;
}

}  // namespace orc

struct orc2_polyline {
    std::vector<real> pts;
    std::vector<uint32_t> idx;
    std::vector<Box> seg_box;
    BVT bvt;
    uint32_t nedges = 0;
};

extern "C" {

orc2_polyline* orc2_polyline_create(uint32_t n_points, const real* xy, uint32_t n_edges, const uint32_t* idx) {
    orc2_polyline* m = new orc2_polyline;
    m->pts.assign(xy, xy + 2 * (size_t)n_points);
    if (idx) {
        m->idx.assign(idx, idx + 2 * (size_t)n_edges);
    } else {  // polyline.rs:58-63: the line strip
        n_edges = n_points ? n_points - 1 : 0;
        for (uint32_t i = 0; i < n_edges; ++i) m->idx.push_back(i), m->idx.push_back(i + 1);
    }
    m->nedges = n_edges;
    std::vector<Elt> leaves;
    for (uint32_t i = 0; i < n_edges; ++i) {
        const real* a = &m->pts[2 * m->idx[2 * i]];
        const real* b = &m->pts[2 * m->idx[2 * i + 1]];
        Box bx;
        bx.mins = v3(0, 0, 0), bx.maxs = v3(0, 0, 0);
        for (int k = 0; k < 2; ++k) {  // local_support_map_aabb over Segment::local_support_point: a if a . dir > b . dir else b
            real da = a[k] * real(1) + a[1 - k] * real(0), db = b[k] * real(1) + b[1 - k] * real(0);
            bx.maxs[k] = da > db ? a[k] : b[k];
            real na = a[k] * real(-1) + a[1 - k] * real(0), nb = b[k] * real(-1) + b[1 - k] * real(0);
            bx.mins[k] = na > nb ? a[k] : b[k];
        }
        m->seg_box.push_back(bx);
        leaves.push_back({i, bx});
    }
    if (n_edges) {
        m->bvt.internals.reserve(n_edges);
        m->bvt.leaves.reserve(n_edges);
        m->bvt.root = bvt_build(0, leaves, m->bvt, 2);
        m->bvt.has_root = true;
    }
    return m;
}
void orc2_polyline_destroy(orc2_polyline* m) { delete m; }

// RayCast for Polyline::toi_and_normal_with_ray for a batch.  pose: x y re im (NULL: identity).  mode 0: the reference's best-first
// search; mode 1: the definition the device follows (minimum toi over the edges whose AABB and segment are hit, ties -> smallest edge).
// feature: edge (Face(0) / Vertex) or edge + n_edges (Face(1)); 0xffffffff = None.
void orc2_polyline_ray_cast(const orc2_polyline* m, const real* pose, uint64_t n_rays, const real* origins, const real* dirs, real max_toi_all,
                            const real* max_tois, int mode, real* toi, uint32_t* feature, real* normal) {
    real tx = 0, ty = 0, re = 1, im = 0;
    if (pose) tx = pose[0], ty = pose[1], re = pose[2], im = pose[3];
    Heap queue;
    for (uint64_t r = 0; r < n_rays; ++r) {
        real ox = origins[2 * r], oy = origins[2 * r + 1], dx = dirs[2 * r], dy = dirs[2 * r + 1];
        if (pose) {  // ray.inverse_transform_by(m): conjugate rotation of (origin - translation) and of dir
            real px = ox - tx, py = oy - ty;
            ox = re * px + im * py, oy = -im * px + re * py;
            real qx = dx, qy = dy;
            dx = re * qx + im * qy, dy = -im * qx + re * qy;
        }
        real max_toi = max_tois ? max_tois[r] : max_toi_all;
        V3 o = v3(ox, oy, 0), d = v3(dx, dy, 0);
        auto seg = [&](uint32_t e) {
            const real* a = &m->pts[2 * m->idx[2 * e]];
            const real* b = &m->pts[2 * m->idx[2 * e + 1]];
            // segment_at(e).transformed(identity): rotation by 1 + 0i and a zero translation change nothing but the sign of a zero
            return segment_ray(a[0], a[1], b[0], b[1], ox, oy, dx, dy);
        };
        Hit h;
        if (mode == 0) {
            h = best_first_search(m->bvt, queue, [&](real best, const Box& bv, const uint32_t* data, real* cost, Hit* res) {
                real bt;  // PolylineRayToiAndNormalVisitor::visit (ray_polyline.rs:113-146)
                if (!aabb_toi_with_ray(bv, o, d, max_toi, true, &bt, 2)) return 0;
                *cost = bt;
                res->some = false;
                if (data && bt < best) {
                    Seg2Hit sh = seg(*data);
                    if (sh.some) {
                        *cost = sh.toi;
                        res->some = true, res->tri = *data, res->toi = sh.toi, res->normal = v3(sh.nx, sh.ny, 0);
                        res->fid = (sh.kind == 1 && sh.id == 1) ? 1 : 0;
                    }
                }
                return 1;
            });
        } else {
            for (uint32_t e = 0; e < m->nedges; ++e) {
                real bt;
                if (!aabb_toi_with_ray(m->seg_box[e], o, d, max_toi, true, &bt, 2)) continue;
                Seg2Hit sh = seg(e);
                if (sh.some && (!h.some || sh.toi < h.toi)) {
                    h.some = true, h.tri = e, h.toi = sh.toi, h.normal = v3(sh.nx, sh.ny, 0);
                    h.fid = (sh.kind == 1 && sh.id == 1) ? 1 : 0;
                }
            }
        }
        if (h.some) {
            toi[r] = h.toi;
            feature[r] = h.fid == 1 ? h.tri + m->nedges : h.tri;  // ray_polyline.rs:41-45
            if (normal) {  // m * res.normal
                normal[2 * r] = pose ? re * h.normal.x - im * h.normal.y : h.normal.x;
                normal[2 * r + 1] = pose ? im * h.normal.x + re * h.normal.y : h.normal.y;
            }
        } else {
            toi[r] = -1;
            feature[r] = 0xffffffffu;
            if (normal) normal[2 * r] = normal[2 * r + 1] = 0;
        }
    }
}

// Segment::toi_and_normal_with_ray(m, ray, max_toi, solid) (dim2) for one segment: returns 1 for Some.  feature: kind << 30 | id.
int orc2_segment_ray_cast(const real* ab, const real* pose, const real* origin, const real* dir, real* toi, real* normal, uint32_t* feature) {
    real ax = ab[0], ay = ab[1], bx = ab[2], by = ab[3];
    if (pose) {  // self.transformed(m)
        real tx = pose[0], ty = pose[1], re = pose[2], im = pose[3];
        real x = ax, y = ay;
        ax = (re * x - im * y) + tx, ay = (im * x + re * y) + ty;
        x = bx, y = by;
        bx = (re * x - im * y) + tx, by = (im * x + re * y) + ty;
    }
    Seg2Hit h = segment_ray(ax, ay, bx, by, origin[0], origin[1], dir[0], dir[1]);
    if (!h.some) return 0;
    *toi = h.toi;
    normal[0] = h.nx, normal[1] = h.ny;
    *feature = ((uint32_t)h.kind << 30) | (uint32_t)h.id;
    return 1;
}

}  // extern "C"
