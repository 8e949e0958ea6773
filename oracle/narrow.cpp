// ORACLE — TEST INFRASTRUCTURE ONLY (see na.hpp header).
// Narrow phase restated from the reference:
//   pipeline/narrow_phase/contact_generator/default_contact_dispatcher.rs:27-97 (dispatch priority)
//   .../ball_ball_manifold_generator.rs:28-67 + query/contact/contact_ball_ball.rs:8-38
//   .../plane_ball_manifold_generator.rs:30-83
//   .../plane_convex_polyhedron_manifold_generator.rs:29-81
//   .../ball_convex_polyhedron_manifold_generator.rs:28-122 + query/point/point_aabb.rs:9-135,
//        query/point/point_support_map.rs:15-53,90-117
//   .../convex_polyhedron_convex_polyhedron_manifold_generator.rs:83-167
//   shape/cuboid.rs:185-405, shape/convex.rs:388-540, shape/convex_polygonal_feature3.rs:217-401,
//   utils/point_in_poly2d.rs:5-26, query/ray/ray_plane.rs:9-23,
//   query/closest_points/closest_points_segment_segment.rs:62-141,
//   query/contact/contact_manifold.rs:165-236 (distance-based tracking, 0.02),
//   pipeline/narrow_phase/narrow_phase.rs:56-104,168-197, pipeline/object/query_type.rs:39-51.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>
#include "gjk_epa.hpp"
#include "oracle.h"
#include "scene.hpp"

namespace orc {

struct AABB {
    V3 mins, maxs;
};
AABB fat_aabb(const Objects& o, uint32_t i, real margin);
void broad_phase_dbvt(uint32_t n, const AABB* fat, const uint32_t* groups, std::vector<uint32_t>& pairs_out);

// ---------------------------------------------------------------------------------------------
// ConvexPolygonalFeature (convex_polygonal_feature3.rs:39-54)
// ---------------------------------------------------------------------------------------------
struct Feature {
    std::vector<V3> vertices;
    std::vector<V3> edge_normals;
    bool has_normal = false;
    V3 normal = {0, 0, 0};
    uint32_t feature_id = FID_UNKNOWN;
    std::vector<uint32_t> vertices_id, edges_id;
    void clear() {
        vertices.clear();
        edge_normals.clear();
        vertices_id.clear();
        edges_id.clear();
        has_normal = false;
        feature_id = FID_UNKNOWN;
    }
    void push(V3 p, uint32_t id) {
        vertices.push_back(p);
        vertices_id.push_back(id);
    }
    void transform_by(const Iso& m) {
        for (auto& p : vertices) p = iso_mul_point(m, p);
        for (auto& n : edge_normals) n = iso_mul_vec(m, n);
        if (has_normal) normal = iso_mul_vec(m, normal);
    }
    void recompute_edge_normals() {  // :157-170
        edge_normals.clear();
        for (size_t i1 = 0; i1 < vertices.size(); ++i1) {
            size_t i2 = (i1 + 1) % vertices.size();
            V3 dpt = vertices[i2] - vertices[i1];
            V3 sn = cross(dpt, normal), nn;
            if (try_normalize(sn, EPS, &nn))
                edge_normals.push_back(nn);
            else
                edge_normals.push_back(v3(0, 0, 0));
        }
    }
    size_t nedges() const {
        size_t l = vertices.size();
        return l == 1 ? 0 : (l == 2 ? 1 : l);
    }
    bool edge(uint32_t edge_id, V3* a, V3* b) const {  // :134-143
        for (size_t i1 = 0; i1 < vertices.size(); ++i1) {
            if (i1 < edges_id.size() && edges_id[i1] == edge_id) {
                size_t i2 = (i1 + 1) % vertices.size();
                *a = vertices[i1];
                *b = vertices[i2];
                return true;
            }
        }
        return false;
    }
};

// ---------------------------------------------------------------------------------------------
// Cuboid (cuboid.rs)
// ---------------------------------------------------------------------------------------------
static void cuboid_face(V3 he, uint32_t i, Feature& out) {  // cuboid.rs:185-277 (dim3)
    out.clear();
    uint32_t i1;
    real sign;
    if (i < 3) {
        i1 = i;
        sign = 1;
    } else {
        i1 = i - 3;
        sign = -1;
    }
    uint32_t i2 = (i1 + 1) % 3, i3 = (i1 + 2) % 3;
    uint32_t edge_i2, edge_i3;
    if (sign > 0) {
        edge_i2 = i2;
        edge_i3 = i3;
    } else {
        edge_i2 = i3;
        edge_i3 = i2;
    }
    uint32_t mask_i2 = ~(1u << edge_i2), mask_i3 = ~(1u << edge_i3);
    V3 vertex = he;
    vertex[i1] *= sign;
    uint32_t sbit = sign < 0 ? 1 : 0, msbit = sign < 0 ? 0 : 1;
    uint32_t vertex_id = sbit << i1;
    out.push(vertex, fid(F_VERTEX, vertex_id));
    out.edges_id.push_back(fid(F_EDGE, edge_i2 | ((vertex_id & mask_i2) << 2)));

    vertex[i2] = -sign * he[i2];
    vertex[i3] = sign * he[i3];
    vertex_id |= (msbit << i2) | (sbit << i3);
    out.push(vertex, fid(F_VERTEX, vertex_id));
    out.edges_id.push_back(fid(F_EDGE, edge_i3 | ((vertex_id & mask_i3) << 2)));

    vertex[i2] = -he[i2];
    vertex[i3] = -he[i3];
    vertex_id |= (1u << i2) | (1u << i3);
    out.push(vertex, fid(F_VERTEX, vertex_id));
    out.edges_id.push_back(fid(F_EDGE, edge_i2 | ((vertex_id & mask_i2) << 2)));

    vertex[i2] = sign * he[i2];
    vertex[i3] = -sign * he[i3];
    vertex_id = (sbit << i1) | (sbit << i2) | (msbit << i3);
    out.push(vertex, fid(F_VERTEX, vertex_id));
    out.edges_id.push_back(fid(F_EDGE, edge_i3 | ((vertex_id & mask_i3) << 2)));

    V3 normal = v3(0, 0, 0);
    normal[i1] = sign;
    out.normal = normal;
    out.has_normal = true;
    out.feature_id = sign > 0 ? fid(F_FACE, i1) : fid(F_FACE, i1 + 3);
    out.recompute_edge_normals();
}

static void cuboid_support_face_toward(V3 he, const Iso& m, V3 dir, Feature& out) {  // cuboid.rs:279-307
    out.clear();
    V3 ld = iso_inv_vec(m, dir);
    uint32_t iamax = 0;
    real amax = std::fabs(ld[0]);
    for (uint32_t i = 1; i < 3; ++i) {
        real c = std::fabs(ld[i]);
        if (c > amax) {
            amax = c;
            iamax = i;
        }
    }
    if (ld[iamax] > 0)
        cuboid_face(he, iamax, out);
    else
        cuboid_face(he, iamax + 3, out);
    out.transform_by(m);
}

static void cuboid_support_feature_toward(V3 he, const Iso& m, V3 dir, real angle, Feature& out) {  // cuboid.rs:309-405
    V3 ld = iso_inv_vec(m, dir);
    real cang = std::cos(angle);
    V3 sp = he;
    out.clear();
    real sang = std::sin(angle);
    uint32_t sp_id = 0;
    for (uint32_t i1 = 0; i1 < 3; ++i1) {
        real sign = signum(ld[i1]);
        if (sign * ld[i1] >= cang) {
            if (sign > 0)
                cuboid_face(he, i1, out);
            else
                cuboid_face(he, i1 + 3, out);
            out.transform_by(m);
            return;
        } else if (sign < 0) {
            sp[i1] *= sign;
            sp_id |= 1u << i1;
        }
    }
    for (uint32_t i = 0; i < 3; ++i) {
        real sign = signum(ld[i]);
        if (sign * ld[i] <= sang) {
            sp[i] = -he[i];
            V3 p1 = sp;
            sp[i] = he[i];
            V3 p2 = sp;
            uint32_t p2_id = sp_id & ~(1u << i);
            out.push(iso_mul_point(m, p1), fid(F_VERTEX, sp_id | (1u << i)));
            out.push(iso_mul_point(m, p2), fid(F_VERTEX, p2_id));
            uint32_t edge_id = fid(F_EDGE, i | (p2_id << 2));
            out.edges_id.push_back(edge_id);
            out.feature_id = edge_id;
            return;
        }
    }
    out.push(iso_mul_point(m, sp), fid(F_VERTEX, sp_id));
    out.feature_id = fid(F_VERTEX, sp_id);
}

static V3 cuboid_feature_normal(uint32_t f) {  // cuboid.rs:502-563 (dim3)
    uint32_t id = fid_id(f);
    V3 dir = v3(0, 0, 0);
    switch (fid_kind(f)) {
        case F_FACE:
            if (id < 3)
                dir[id] = 1;
            else
                dir[id - 3] = -1;
            return dir;
        case F_EDGE: {
            uint32_t edge = id & 3u, face1 = (edge + 1) % 3, face2 = (edge + 2) % 3, signs = id >> 2;
            dir[face1] = (signs & (1u << face1)) ? -1 : 1;
            dir[face2] = (signs & (1u << face2)) ? -1 : 1;
            return normalize(dir);
        }
        default:
            for (uint32_t i = 0; i < 3; ++i) dir[i] = (id & (1u << i)) ? -1 : 1;
            return normalize(dir);
    }
}

// AABB::project_point_with_feature for a cuboid (point_aabb.rs:9-135, point_cuboid.rs:17-27)
static void cuboid_project_point_with_feature(V3 he, const Iso& m, V3 pt, bool* inside_out, V3* proj_out, uint32_t* feature) {
    V3 mins = v3(0, 0, 0) + (-he), maxs = v3(0, 0, 0) + he;
    V3 ls_pt = iso_inv_point(m, pt);
    V3 mins_pt = mins - ls_pt, pt_maxs = ls_pt - maxs;
    V3 zero = v3(0, 0, 0);
    V3 shift = sup(mins_pt, zero) - sup(pt_maxs, zero);
    bool inside = shift.x == 0 && shift.y == 0 && shift.z == 0;
    V3 ls_proj;
    if (!inside) {
        ls_proj = ls_pt + shift;
    } else {
        real best = -FMAX;
        bool is_mins = false;
        int best_id = 0;
        for (int i = 0; i < 3; ++i) {
            real mins_pt_i = mins_pt[i], pt_maxs_i = pt_maxs[i];
            if (mins_pt_i < pt_maxs_i) {
                if (pt_maxs[i] > best) {
                    best_id = i;
                    is_mins = false;
                    best = pt_maxs_i;
                }
            } else if (mins_pt_i > best) {
                best_id = i;
                is_mins = true;
                best = mins_pt_i;
            }
        }
        shift = v3(0, 0, 0);
        if (is_mins)
            shift[best_id] = best;
        else
            shift[best_id] = -best;
        ls_proj = ls_pt + shift;
    }
    *inside_out = inside;
    *proj_out = iso_mul_point(m, ls_proj);
    int nzero_shifts = 0, last_zero_shift = 0, last_not_zero_shift = 0;
    for (int i = 0; i < 3; ++i) {
        if (shift[i] == 0) {
            nzero_shifts++;
            last_zero_shift = i;
        } else
            last_not_zero_shift = i;
    }
    V3 center = (mins + maxs) * real(0.5);
    if (nzero_shifts == 3) {
        for (int i = 0; i < 3; ++i) {
            if (ls_proj[i] > maxs[i] - EPS) {
                *feature = fid(F_FACE, i);
                return;
            }
            if (ls_proj[i] <= mins[i] + EPS) {
                *feature = fid(F_FACE, i + 3);
                return;
            }
        }
        *feature = FID_UNKNOWN;
    } else if (nzero_shifts == 2) {
        if (ls_proj[last_not_zero_shift] < center[last_not_zero_shift])
            *feature = fid(F_FACE, last_not_zero_shift + 3);
        else
            *feature = fid(F_FACE, last_not_zero_shift);
    } else {
        uint32_t id = 0;
        for (int i = 0; i < 3; ++i)
            if (ls_proj[i] < center[i]) id |= 1u << i;
        if (nzero_shifts == 0)
            *feature = fid(F_VERTEX, id);
        else
            *feature = fid(F_EDGE, (id << 2) | (uint32_t)last_zero_shift);
    }
}

// ---------------------------------------------------------------------------------------------
// ConvexHull (convex.rs)
// ---------------------------------------------------------------------------------------------
static void hull_face(const Hull& H, uint32_t id, Feature& out) {  // convex.rs:443-461
    out.clear();
    uint32_t first = H.face_first[id], last = first + H.face_num[id];
    for (uint32_t i = first; i < last; ++i) {
        uint32_t vid = H.vaf[i], eid = H.eaf[i];
        out.push(H.pt(vid), fid(F_VERTEX, vid));
        out.edges_id.push_back(fid(F_EDGE, eid));
    }
    out.normal = H.fnormal(id);
    out.has_normal = true;
    out.feature_id = fid(F_FACE, id);
    out.recompute_edge_normals();
}
static void hull_support_face_toward(const Hull& H, const Iso& m, V3 dir, Feature& out) {  // convex.rs:487-509
    V3 ls_dir = iso_inv_vec(m, dir);
    uint32_t best = 0;
    real max_dot = dot(H.fnormal(0), ls_dir);
    for (uint32_t i = 1; i < H.nf; ++i) {
        real d = dot(H.fnormal(i), ls_dir);
        if (d > max_dot) {
            max_dot = d;
            best = i;
        }
    }
    hull_face(H, best, out);
    out.transform_by(m);
}
static uint32_t hull_support_feature_id_toward_eps(const Hull& H, V3 local_dir, real eps) {  // convex.rs:388-415
    real seps = std::sin(eps), ceps = std::cos(eps);
    uint32_t sp = 0;
    real best_dot = dot(H.pt(0), local_dir);
    for (uint32_t i = 1; i < H.nv; ++i) {
        real d = dot(H.pt(i), local_dir);
        if (d > best_dot) {
            best_dot = d;
            sp = i;
        }
    }
    uint32_t first = H.vert_first_adj[sp], num = H.vert_num_adj[sp];
    for (uint32_t i = 0; i < num; ++i) {
        uint32_t face_id = H.fav[first + i];
        if (dot(H.fnormal(face_id), local_dir) >= ceps) return fid(F_FACE, face_id);
    }
    for (uint32_t i = 0; i < num; ++i) {
        uint32_t edge_id = H.eav[first + i];
        if (std::fabs(dot(H.edir(edge_id), local_dir)) <= seps) return fid(F_EDGE, edge_id);
    }
    return fid(F_VERTEX, sp);
}
static void hull_support_feature_toward(const Hull& H, const Iso& m, V3 dir, real angle, Feature& out) {  // convex.rs:511-540
    out.clear();
    V3 local_dir = iso_inv_vec(m, dir);
    uint32_t f = hull_support_feature_id_toward_eps(H, local_dir, angle);
    switch (fid_kind(f)) {
        case F_VERTEX:
            out.push(H.pt(fid_id(f)), f);
            out.feature_id = f;
            break;
        case F_EDGE: {
            uint32_t e = fid_id(f), v1 = H.edge_vertices[2 * e], v2 = H.edge_vertices[2 * e + 1];
            out.push(H.pt(v1), fid(F_VERTEX, v1));
            out.push(H.pt(v2), fid(F_VERTEX, v2));
            out.feature_id = f;
            out.edges_id.push_back(f);
            break;
        }
        default:
            hull_face(H, fid_id(f), out);
            break;
    }
    out.transform_by(m);
}
static V3 hull_feature_normal(const Hull& H, uint32_t f) {  // convex.rs:463-485
    uint32_t id = fid_id(f);
    switch (fid_kind(f)) {
        case F_FACE:
            return H.fnormal(id);
        case F_EDGE:
            return normalize(H.fnormal(H.edge_faces[2 * id]) + H.fnormal(H.edge_faces[2 * id + 1]));
        default: {
            uint32_t first = H.vert_first_adj[id], last = first + H.vert_num_adj[id];
            V3 n = v3(0, 0, 0);
            for (uint32_t i = first; i < last; ++i) n = n + H.fnormal(H.fav[i]);
            return normalize(n);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Segment as a ConvexPolyhedron (segment.rs:152-375, dim3), for the segment of a Capsule: a = (0, -hh, 0), b = (0, hh, 0)
// (capsule.rs:53-61).  Groundwork for SURVEY §8f N3 (capsule generators); oracle only so far.
// ---------------------------------------------------------------------------------------------
static void segment_support_face_toward(real hh, const Iso& m, Feature& out) {  // segment.rs:286-299
    out.clear();
    out.push(v3(0, -hh, 0), fid(F_VERTEX, 0));
    out.push(v3(0, hh, 0), fid(F_VERTEX, 1));
    out.edges_id.push_back(fid(F_EDGE, 0));
    out.feature_id = fid(F_EDGE, 0);
    out.transform_by(m);
}
static void segment_support_feature_toward(real hh, const Iso& m, V3 dir, real eps, Feature& out) {  // segment.rs:301-341
    out.clear();
    V3 a = iso_mul_point(m, v3(0, -hh, 0)), b = iso_mul_point(m, v3(0, hh, 0));  // self.transformed(transform)
    real ceps = std::sin(eps);
    V3 seg_dir;
    if (!unit_try_new(b - a, EPS, &seg_dir)) return;
    real cang = dot(dir, seg_dir);
    if (cang > ceps) {
        out.feature_id = fid(F_VERTEX, 1);
        out.push(b, fid(F_VERTEX, 1));
    } else if (cang < -ceps) {
        out.feature_id = fid(F_VERTEX, 0);
        out.push(a, fid(F_VERTEX, 0));
    } else {
        out.push(a, fid(F_VERTEX, 0));
        out.push(b, fid(F_VERTEX, 1));
        out.edges_id.push_back(fid(F_EDGE, 0));
        out.feature_id = fid(F_EDGE, 0);
    }
}
static V3 segment_feature_normal(real hh, uint32_t f) {  // segment.rs:237-284
    V3 direction;
    if (!unit_try_new(v3(0, hh, 0) - v3(0, -hh, 0), EPS, &direction)) return v3(0, 1, 0);
    switch (fid_kind(f)) {
        case F_VERTEX:
            return fid_id(f) == 0 ? direction : -direction;
        case F_EDGE: {
            int iamin = 0;  // first component of smallest absolute value
            for (int k = 1; k < 3; ++k)
                if (std::fabs(direction[k]) < std::fabs(direction[iamin])) iamin = k;
            V3 normal = v3(0, 0, 0);
            normal[iamin] = 1;
            normal = normal - direction * direction[iamin];
            return normalize(normal);
        }
        default: {  // Face(id): the 2-D formula, z left at 0
            V3 dir = fid_id(f) == 0 ? v3(direction.y, -direction.x, 0) : v3(-direction.y, direction.x, 0);
            return dir;
        }
    }
}
// PointQuery::project_point_with_feature for Segment (point_segment.rs:14-40,50-91)
static void segment_project_point_with_feature(real hh, const Iso& m, V3 pt, bool* inside, V3* proj_out, uint32_t* feature) {
    V3 a = v3(0, -hh, 0), b = v3(0, hh, 0);
    V3 ls_pt = iso_inv_point(m, pt);
    V3 ab = b - a, ap = ls_pt - a;
    real ab_ap = dot(ab, ap), sqnab = norm_squared(ab);
    V3 proj;
    if (ab_ap <= 0) {
        *feature = fid(F_VERTEX, 0);
        proj = iso_mul_point(m, a);
    } else if (ab_ap >= sqnab) {
        *feature = fid(F_VERTEX, 1);
        proj = iso_mul_point(m, b);
    } else {
        real u = ab_ap / sqnab;
        *feature = fid(F_EDGE, 0);
        proj = iso_mul_point(m, a + ab * u);
    }
    *inside = relative_eq_v3(proj, pt);
    *proj_out = proj;
}

// ContactPreprocessor of a capsule (capsule.rs:87-132): the sub-detector works on the capsule's segment with the prediction enlarged
// by the radius; every contact it produces is moved out to the capsule's surface and its segment feature is renamed.
struct Preproc {
    bool active = false;
    real radius = 0;
    // returns false when the contact must be ignored
    bool process(Contact& c, uint32_t& f1, uint32_t& f2, bool is_first) const {
        if (!active) return true;
        uint32_t f = is_first ? f1 : f2, actual;
        switch (fid_kind(f)) {
            case F_VERTEX: actual = fid(F_FACE, fid_id(f)); break;
            case F_EDGE: actual = fid(F_FACE, 2); break;
            case F_FACE: actual = fid(F_FACE, 2 + fid_id(f)); break;
            default: return false;
        }
        if (is_first) {
            f1 = actual;
            c.k.dil1 = radius;  // kinematic.set_dilation1
            c.world1 = c.world1 + c.normal * radius;
            c.depth += radius;
        } else {
            f2 = actual;
            c.k.dil2 = radius;
            c.world2 = c.world2 - c.normal * radius;
            c.depth += radius;
        }
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// Shapes
// ---------------------------------------------------------------------------------------------
struct Shape {
    uint32_t type;
    real radius;
    V3 he;      // cuboid half extents / plane normal / (half_height, radius, 0) of a capsule
    Hull hull;
    real hh = 0;  // capsule / segment half height
};
static Shape get_shape(const Objects& o, uint32_t i) {
    Shape s;
    s.type = o.shape_type[i];
    s.radius = o.shape_param[4 * i];
    s.he = o.p3(i);
    if (s.type == HULL) s.hull = hull_view(o.hulls, o.hull_id(i));
    if (s.type == CAPSULE) s.hh = o.shape_param[4 * i], s.radius = o.shape_param[4 * i + 1];
    return s;
}
static Support as_support(const Shape& s) {
    Support g;
    g.he = s.he;
    g.radius = s.radius;
    if (s.type == CUBOID)
        g.kind = Support::S_CUBOID;
    else if (s.type == HULL) {
        g.kind = Support::S_HULL;
        g.hull = s.hull;
    } else if (s.type == SEGMENT) {
        g.kind = Support::S_SEGMENT;
        g.he = v3(s.hh, 0, 0);
    } else if (s.type == CAPSULE) {
        g.kind = Support::S_CAPSULE;
        g.he = v3(s.hh, 0, 0);
    } else
        g.kind = Support::S_BALL;
    return g;
}
static bool is_convex_polyhedron(const Shape& s) { return s.type == CUBOID || s.type == HULL || s.type == SEGMENT; }
static void support_face_toward(const Shape& s, const Iso& m, V3 dir, Feature& out) {
    if (s.type == CUBOID)
        cuboid_support_face_toward(s.he, m, dir, out);
    else if (s.type == SEGMENT)
        segment_support_face_toward(s.hh, m, out);
    else
        hull_support_face_toward(s.hull, m, dir, out);
}
static void support_feature_toward(const Shape& s, const Iso& m, V3 dir, real angle, Feature& out) {
    if (s.type == CUBOID)
        cuboid_support_feature_toward(s.he, m, dir, angle, out);
    else if (s.type == SEGMENT)
        segment_support_feature_toward(s.hh, m, dir, angle, out);
    else
        hull_support_feature_toward(s.hull, m, dir, angle, out);
}

// ---------------------------------------------------------------------------------------------
// ContactManifold (contact_manifold.rs:14-236): Slab<(TrackedContact, usize)> + DistanceBased(0.02) cache,
// persistence = 1.  A fresh manifold (no save_cache_and_clear before) behaves like the one-shot case.
// ---------------------------------------------------------------------------------------------
struct Tracked {
    Contact c;
    uint32_t f1, f2;
    uint32_t id = 0;  // insertion counter of this manifold: stable exactly as long as the reference's ContactId is
};
struct Manifold {
    struct Slot {
        Tracked t;
        size_t remaining = 0;    // the `usize` of the tuple
        bool occupied = false;
        int64_t next_free = -1;  // slab free-list link
    };
    // slab::Slab: vacant entries form a LIFO free list; `next` == entries.len() when the list is empty
    std::vector<Slot> slab;
    size_t slab_next = 0;
    std::vector<std::pair<V3, size_t>> cache;
    size_t deepest = 0, ncontacts = 0;
    static constexpr size_t persistence = 1;
    uint32_t next_id = 0;

    size_t slab_insert(const Tracked& t, size_t remaining) {
        size_t key = slab_next;
        if (key == slab.size()) {
            slab.push_back(Slot{});
            slab_next = key + 1;
        } else {
            slab_next = (size_t)slab[key].next_free;
        }
        slab[key].t = t;
        slab[key].remaining = remaining;
        slab[key].occupied = true;
        return key;
    }
    void slab_remove(size_t key) {
        slab[key].occupied = false;
        slab[key].next_free = (int64_t)slab_next;
        slab_next = key;
    }
    size_t len() const { return ncontacts; }
    // contacts(): slab order, entries with remaining == persistence (:59-68)
    template <typename F>
    void for_each_contact(F f) const {
        for (size_t i = 0; i < slab.size(); ++i)
            if (slab[i].occupied && slab[i].remaining == persistence) f(slab[i].t, i);
    }
    void save_cache_and_clear() {  // :134-156
        std::vector<std::pair<V3, size_t>> kept;
        for (auto& c : cache)
            if (slab[c.second].remaining != 0) kept.push_back(c);
        cache.swap(kept);
        deepest = 0;
        ncontacts = 0;
        for (size_t i = 0; i < slab.size(); ++i) {  // Slab::retain visits the keys in increasing order
            if (!slab[i].occupied) continue;
            if (slab[i].remaining == 0)
                slab_remove(i);
            else
                slab[i].remaining -= 1;
        }
    }
    // preprocessor1 / preprocessor2 of :171-181 (only the capsule's exists on the path); a rejected contact is dropped
    void push(Contact c, uint32_t f1, uint32_t f2, V3 tracking_pt, const Preproc* pp1, const Preproc* pp2) {
        if (pp1 && !pp1->process(c, f1, f2, true)) return;
        if (pp2 && !pp2->process(c, f1, f2, false)) return;
        push(c, f1, f2, tracking_pt);
    }
    void push(const Contact& c, uint32_t f1, uint32_t f2, V3 tracking_pt) {  // :165-236
        const real threshold = real(0.02);
        bool is_deepest = ncontacts == 0 || c.depth > slab[deepest].t.c.depth;
        size_t closest = cache.size();
        real closest_dist = threshold * threshold;
        for (size_t i = 0; i < cache.size(); ++i) {
            real d = norm_squared(tracking_pt - cache[i].first);
            if (d < closest_dist) {
                closest_dist = d;
                closest = i;
            }
        }
        if (closest == cache.size()) {
            Tracked t{c, f1, f2, 0};
            t.id = ++next_id;  // narrow_phase.rs:85-89 hands a fresh id to every contact whose id is null
            size_t i = slab_insert(t, persistence);
            cache.push_back({tracking_pt, i});
            ncontacts += 1;
            if (is_deepest) deepest = i;
        } else {
            size_t ci = cache[closest].second;
            if (is_deepest) deepest = ci;
            Slot& sl = slab[ci];
            if (sl.remaining == persistence) {
                if (c.depth <= sl.t.c.depth) return;  // keep the contact already in cache because it is deeper
            } else {
                ncontacts += 1;
                sl.remaining = persistence;
            }
            sl.t.c = c;  // the TrackedContact keeps its id
            sl.t.f1 = f1, sl.t.f2 = f2;
            cache[closest].first = tracking_pt;
        }
    }
};

static Contact contact_new_wo_depth(V3 w1, V3 w2, V3 n) { return {w1, w2, n, -dot(n, w2 - w1)}; }

// ---------------------------------------------------------------------------------------------
// Generators
// ---------------------------------------------------------------------------------------------
static const uint32_t FACE0 = (F_FACE << 30);

static void gen_ball_ball(const Iso& ma, real r1, const Iso& mb, real r2, real prediction, Manifold& mf) {
    V3 c1 = ma.t, c2 = mb.t;
    V3 delta = c2 - c1;
    real d2 = norm_squared(delta);
    real sum_radius = r1 + r2;
    real sre = sum_radius + prediction;
    if (d2 < sre * sre) {
        V3 normal = d2 != 0 ? normalize(delta) : v3(1, 0, 0);
        Contact c = {c1 + normal * r1, c2 + normal * (-r2), normal, sum_radius - std::sqrt(d2)};
        c.k.dil1 = r1, c.k.dil2 = r2;  // both approximations: (Face(0), origin, Point) (ball_ball_manifold_generator.rs:46-58)
        mf.push(c, FACE0, FACE0, v3(0, 0, 0));
    }
}

// plane_ball_manifold_generator.rs:30-83.  (m1, plane) (m2, ball)
static void gen_plane_ball(const Iso& m1, V3 plane_n, const Iso& m2, real radius, real prediction, bool flip, Manifold& mf) {
    V3 n = iso_mul_vec(m1, plane_n);
    V3 pc = m1.t, bc = m2.t;
    real dist = dot(bc - pc, n);
    real depth = -dist + radius;
    if (depth > -prediction) {
        V3 world1 = bc + n * (-dist);
        V3 world2 = bc + n * (-radius);
        V3 local1 = iso_inv_point(m1, world1);  // plane side: (Face(0), local1, Plane(plane.normal)); ball side: (Face(0), origin, Point)
        if (!flip) {
            Contact c = {world1, world2, n, depth};
            c.k.local1 = local1, c.k.g1 = G_PLANE, c.k.dir1 = plane_n, c.k.dil2 = radius;
            mf.push(c, FACE0, FACE0, v3(0, 0, 0));
        } else {
            Contact c = {world2, world1, -n, depth};
            c.k.local2 = local1, c.k.g2 = G_PLANE, c.k.dir2 = plane_n, c.k.dil1 = radius;
            mf.push(c, FACE0, FACE0, v3(0, 0, 0));
        }
    }
}

// plane_convex_polyhedron_manifold_generator.rs:29-81
// pp1 / pp2: the preprocessors of (m1, plane) / (m2, cp) as do_update_to receives them
static void gen_plane_convex(const Iso& m1, V3 plane_n, const Iso& m2, const Shape& cp, real prediction, bool flip, Manifold& mf,
                             const Preproc* pp1 = nullptr, const Preproc* pp2 = nullptr) {
    V3 n = iso_mul_vec(m1, plane_n);
    V3 pc = m1.t;
    Feature feat;
    support_face_toward(cp, m2, -n, feat);
    for (size_t i = 0; i < feat.vertices.size(); ++i) {
        V3 world2 = feat.vertices[i];
        V3 dpt = world2 - pc;
        real dist = dot(dpt, n);
        if (dist <= prediction) {
            V3 world1 = world2 + (-n * dist);
            V3 local1 = iso_inv_point(m1, world1);
            V3 local2 = iso_inv_point(m2, world2);
            uint32_t f2 = feat.vertices_id[i];
            if (!flip) {
                Contact c = {world1, world2, n, -dist};
                c.k.local1 = local1, c.k.g1 = G_PLANE, c.k.dir1 = plane_n, c.k.local2 = local2;  // approx2 = Point
                mf.push(c, FACE0, f2, local2, pp1, pp2);
            } else {
                Contact c = {world2, world1, -n, -dist};
                c.k.local1 = local2, c.k.local2 = local1, c.k.g2 = G_PLANE, c.k.dir2 = plane_n;
                mf.push(c, f2, FACE0, local2, pp2, pp1);
            }
        }
    }
}

// point_support_map.rs:15-53 (solid = false) for a hull
static void hull_project_point(const Hull& H, const Iso& m_in, V3 point, bool* inside, V3* proj, GJKStats* st) {
    Iso id = iso_identity();
    Iso m = m_in;
    m.t = (-point) + m_in.t;  // Translation::from(-point.coords) * m
    Support shape;
    shape.kind = Support::S_HULL;
    shape.hull = H;
    Support origin;
    origin.kind = Support::S_ORIGIN;
    V3 dir;
    if (!unit_try_new(-m.t, EPS, &dir)) dir = v3(1, 0, 0);
    VoronoiSimplex simplex;
    simplex.reset(cso_from_shapes(m, shape, id, origin, dir));
    GJKResult r = gjk_closest_points(m, shape, id, origin, FMAX, simplex, st);  // gjk::project_origin
    if (r.kind == GJK_CLOSEST_POINTS) {
        *inside = false;
        *proj = r.p1 + point;
        return;
    }
    // GJK_INTERSECTION (anything else is unreachable!() in the reference)
    EPA epa;
    if (st) st->epa_calls++;
    V3 p1, p2, n;
    *inside = true;
    if (epa.closest_points(m, shape, id, origin, simplex, &p1, &p2, &n, st))
        *proj = p1 + point;
    else {
        if (st) st->epa_fail++;
        *proj = point;
    }
}
// point_support_map.rs:97-116
static void hull_project_point_with_feature(const Hull& H, const Iso& m, V3 point, bool* inside, V3* proj, uint32_t* feature, GJKStats* st) {
    hull_project_point(H, m, point, inside, proj, st);
    V3 dpt = point - *proj;
    V3 local_dir = *inside ? iso_inv_vec(m, -dpt) : iso_inv_vec(m, dpt);
    V3 u;
    if (unit_try_new(local_dir, EPS, &u))
        *feature = hull_support_feature_id_toward_eps(H, u, real(3.14159265358979323846 / 180.0));
    else
        *feature = FID_UNKNOWN;
}

// ConvexPolyhedron::edge(id) in LOCAL coordinates: cuboid.rs:163-183, convex.rs:430-441, segment.rs:203-205
static void shape_edge(const Shape& cp, uint32_t f, V3* p1, V3* p2) {
    uint32_t eid = fid_id(f);
    if (cp.type == CUBOID) {
        real res[3] = {cp.he.x, cp.he.y, cp.he.z};
        uint32_t edge_i = eid & 3u, vertex_i = eid >> 2;
        for (uint32_t i = 0; i < 3; ++i)
            if (i != edge_i && (vertex_i & (1u << i))) res[i] = -res[i];
        *p1 = v3(res[0], res[1], res[2]);
        res[edge_i] = -res[edge_i];
        *p2 = v3(res[0], res[1], res[2]);
    } else if (cp.type == SEGMENT) {
        *p1 = v3(0, -cp.hh, 0), *p2 = v3(0, cp.hh, 0);
    } else {
        *p1 = cp.hull.pt(cp.hull.edge_vertices[2 * eid]), *p2 = cp.hull.pt(cp.hull.edge_vertices[2 * eid + 1]);
    }
}

// ball_convex_polyhedron_manifold_generator.rs:28-122.  (m1, ball) (m2, convex polyhedron)
static void gen_ball_convex(const Iso& m1, real radius, const Iso& m2, const Shape& cp, real prediction, bool flip, Manifold& mf,
                            GJKStats* st, const Preproc* pp1 = nullptr, const Preproc* pp2 = nullptr) {
    V3 ball_center = m1.t;
    bool inside;
    V3 world2;
    uint32_t f2;
    if (cp.type == CUBOID)
        cuboid_project_point_with_feature(cp.he, m2, ball_center, &inside, &world2, &f2);
    else if (cp.type == SEGMENT)
        segment_project_point_with_feature(cp.hh, m2, ball_center, &inside, &world2, &f2);
    else
        hull_project_point_with_feature(cp.hull, m2, ball_center, &inside, &world2, &f2, st);
    V3 dpt = world2 - ball_center;
    real depth;
    V3 normal, dir;
    real dist;
    if (unit_try_new_and_get(dpt, EPS, &dir, &dist)) {
        if (inside) {
            depth = dist + radius;
            normal = -dir;
        } else {
            depth = -dist + radius;
            normal = dir;
        }
    } else {
        if (f2 == FID_UNKNOWN) return;
        depth = radius;
        normal = -(cp.type == CUBOID ? cuboid_feature_normal(f2) : (cp.type == SEGMENT ? segment_feature_normal(cp.hh, f2) : hull_feature_normal(cp.hull, f2)));
    }
    if (depth >= -prediction) {
        V3 world1 = ball_center + normal * radius;
        // ball side: (Face(0), origin, Point) + dilation; polyhedron side: local2 and the geometry of f2 (:80-114)
        V3 local2 = iso_inv_point(m2, world2);
        uint32_t g2 = G_POINT;
        V3 d2 = v3(0, 0, 0);
        if (fid_kind(f2) == F_FACE) {
            g2 = G_PLANE, d2 = iso_inv_vec(m2, -normal);
        } else if (fid_kind(f2) == F_EDGE) {
            V3 e0, e1;
            shape_edge(cp, f2, &e0, &e1);
            g2 = G_LINE, d2 = normalize(e1 - e0);
        }
        if (!flip) {
            Contact c = {world1, world2, normal, depth};
            c.k.dil1 = radius, c.k.local2 = local2, c.k.g2 = g2, c.k.dir2 = d2;
            mf.push(c, FACE0, f2, v3(0, 0, 0), pp1, pp2);
        } else {
            Contact c = {world2, world1, -normal, depth};
            c.k.dil2 = radius, c.k.local1 = local2, c.k.g1 = g2, c.k.dir1 = d2;
            mf.push(c, f2, FACE0, v3(0, 0, 0), pp2, pp1);
        }
    }
}

// utils/point_in_poly2d.rs:5-26
static bool point_in_poly2d(V2 pt, const std::vector<V2>& poly) {
    if (poly.empty()) return false;
    real sign = 0;
    for (size_t i1 = 0; i1 < poly.size(); ++i1) {
        size_t i2 = (i1 + 1) % poly.size();
        V2 seg_dir = sub2(poly[i2], poly[i1]);
        V2 dpt = sub2(pt, poly[i1]);
        real perp = perp2(dpt, seg_dir);
        if (sign == 0)
            sign = perp;
        else if (sign * perp < 0)
            return false;
    }
    return true;
}
// ray_plane.rs:9-23
static bool line_toi_with_plane(V3 plane_center, V3 plane_normal, V3 line_origin, V3 line_dir, real* toi) {
    V3 dpos = plane_center - line_origin;
    real denom = dot(plane_normal, line_dir);
    if (relative_eq(denom, real(0))) return false;
    *toi = dot(plane_normal, dpos) / denom;
    return true;
}
// closest_points_segment_segment.rs:62-141 for D = 2.  Returns true iff both locations are OnEdge.
static bool seg_seg_2d(V2 a1, V2 b1, V2 a2, V2 b2, real* s_out, real* t_out) {
    const real eps = EPS, _0 = 0, _1 = 1;
    V2 d1 = sub2(b1, a1), d2 = sub2(b2, a2), r = sub2(a1, a2);
    real a = dot2(d1, d1), e = dot2(d2, d2), f = dot2(d2, r);
    real s, t;
    if (a <= eps && e <= eps) {
        s = _0;
        t = _0;
    } else if (a <= eps) {
        s = _0;
        t = clampf(f / e, _0, _1);
    } else {
        real c = dot2(d1, r);
        if (e <= eps) {
            t = _0;
            s = clampf(-c / a, _0, _1);
        } else {
            real b = dot2(d1, d2);
            real ae = a * e, bb = b * b, denom = ae - bb;
            bool parallel = denom <= eps || ulps_eq(ae, bb);
            if (!parallel)
                s = clampf((b * f - c * e) / denom, _0, _1);
            else
                s = _0;
            t = (b * s + f) / e;
            if (t < _0) {
                t = _0;
                s = clampf(-c / a, _0, _1);
            } else if (t > _1) {
                t = _1;
                s = clampf((b - c) / a, _0, _1);
            }
        }
    }
    *s_out = s;
    *t_out = t;
    return s != _0 && s != _1 && t != _0 && t != _1;
}

struct NewContact {
    Contact c;
    uint32_t f1, f2;
};

// convex_polygonal_feature3.rs:217-338
static void clip(const Feature& self, const Feature& other, V3 normal, real prediction, std::vector<NewContact>& out) {
    if (self.vertices.size() <= 2 && other.vertices.size() <= 2) return;
    V3 basis[2];
    orthonormal_basis(normal, &basis[0], &basis[1]);
    V3 ref_pt = self.vertices[0];
    std::vector<V2> poly1, poly2;
    for (auto& pt : self.vertices) {
        V3 dpt = pt - ref_pt;
        poly1.push_back({dot(basis[0], dpt), dot(basis[1], dpt)});
    }
    for (auto& pt : other.vertices) {
        V3 dpt = pt - ref_pt;
        poly2.push_back({dot(basis[0], dpt), dot(basis[1], dpt)});
    }
    if (poly2.size() > 2) {
        for (size_t i = 0; i < poly1.size(); ++i) {
            V2 pt = poly1[i];
            if (point_in_poly2d(pt, poly2)) {
                V3 origin = ref_pt + basis[0] * pt.x + basis[1] * pt.y;
                real toi2;
                if (line_toi_with_plane(other.vertices[0], other.normal, origin, normal, &toi2)) {
                    V3 world2 = origin + normal * toi2;
                    V3 world1 = self.vertices[i];
                    Contact c = contact_new_wo_depth(world1, world2, normal);
                    if (-c.depth <= prediction) out.push_back({c, self.vertices_id[i], other.feature_id});
                }
            }
        }
    }
    if (poly1.size() > 2) {
        for (size_t i = 0; i < poly2.size(); ++i) {
            V2 pt = poly2[i];
            if (point_in_poly2d(pt, poly1)) {
                V3 origin = ref_pt + basis[0] * pt.x + basis[1] * pt.y;
                real toi1;
                if (line_toi_with_plane(self.vertices[0], self.normal, origin, normal, &toi1)) {
                    V3 world1 = origin + normal * toi1;
                    V3 world2 = other.vertices[i];
                    Contact c = contact_new_wo_depth(world1, world2, normal);
                    if (-c.depth <= prediction) out.push_back({c, self.feature_id, other.vertices_id[i]});
                }
            }
        }
    }
    size_t nedges1 = self.nedges(), nedges2 = other.nedges();
    for (size_t i1 = 0; i1 < nedges1; ++i1) {
        size_t j1 = (i1 + 1) % poly1.size();
        for (size_t i2 = 0; i2 < nedges2; ++i2) {
            size_t j2 = (i2 + 1) % poly2.size();
            real s, t;
            if (seg_seg_2d(poly1[i1], poly1[j1], poly2[i2], poly2[j2], &s, &t)) {
                // Segment::point_at(OnEdge([1 - s, s])) = a * bcoords[0] + b.coords * bcoords[1]
                V3 world1 = self.vertices[i1] * (real(1) - s) + self.vertices[j1] * s;
                V3 world2 = other.vertices[i2] * (real(1) - t) + other.vertices[j2] * t;
                Contact c = contact_new_wo_depth(world1, world2, normal);
                if (-c.depth <= prediction) out.push_back({c, self.edges_id[i1], other.edges_id[i2]});
            }
        }
    }
}

// convex_polygonal_feature3.rs:341-401: returns false when the reference drops the contact
static bool feature_ok_for_manifold(const Feature& ft, uint32_t f) {
    switch (fid_kind(f)) {
        case F_FACE:
        case F_VERTEX:
            return true;
        case F_EDGE: {
            V3 a, b, d;
            if (!ft.edge(f, &a, &b)) return false;  // .expect("Invalid edge id.") would panic
            return unit_try_new(b - a, EPS, &d);
        }
        default:
            return false;
    }
}

// NeighborhoodGeometry of feature f of the polygonal feature ft, in the local frame of m (add_contact_to_manifold,
// convex_polygonal_feature3.rs:356-397); f passed feature_ok_for_manifold
static void feature_geometry(const Feature& ft, uint32_t f, const Iso& m, uint32_t* g, V3* dir) {
    *g = G_POINT, *dir = v3(0, 0, 0);
    if (fid_kind(f) == F_FACE) {
        *g = G_PLANE, *dir = iso_inv_vec(m, ft.normal);  // inverse_transform_unit_vector(self.normal.unwrap())
    } else if (fid_kind(f) == F_EDGE) {
        V3 a = v3(0, 0, 0), b = a, d = a;
        ft.edge(f, &a, &b);
        unit_try_new(b - a, EPS, &d);  // Segment::direction
        *g = G_LINE, *dir = iso_inv_vec(m, d);
    }
}

// convex_polyhedron_convex_polyhedron_manifold_generator.rs:83-167.  last_gjk_dir: the generator's persistent
// direction (nullptr / !*has_dir = fresh generator, None); updated like :106 and :139.
static void gen_convex_convex(const Iso& ma, const Shape& a, const Iso& mb, const Shape& b, real pred_linear, real ang1, real ang2,
                              Manifold& mf, GJKStats* st, V3* last_gjk_dir = nullptr, bool* has_dir = nullptr, const Preproc* pp1 = nullptr,
                              const Preproc* pp2 = nullptr) {
    Support ga = as_support(a), gb = as_support(b);
    VoronoiSimplex simplex;
    const V3* init = (last_gjk_dir && has_dir && *has_dir) ? last_gjk_dir : nullptr;
    V3 init_copy;
    if (init) {
        init_copy = *init;
        init = &init_copy;
    }
    GJKResult r = contact_support_map_support_map_with_params(ma, ga, mb, gb, pred_linear, simplex, init, st);
    if (last_gjk_dir && has_dir && (r.kind == GJK_CLOSEST_POINTS || r.kind == GJK_NO_INTERSECTION)) {
        *last_gjk_dir = r.dir;
        *has_dir = true;
    }
    std::vector<NewContact> new_contacts;
    Feature m1, m2;
    if (r.kind == GJK_CLOSEST_POINTS) {
        Contact contact = contact_new_wo_depth(r.p1, r.p2, r.dir);
        if (contact.depth > 0) {
            support_face_toward(a, ma, contact.normal, m1);
            support_face_toward(b, mb, -contact.normal, m2);
            clip(m1, m2, contact.normal, pred_linear, new_contacts);
        } else {
            support_feature_toward(a, ma, contact.normal, ang1, m1);
            support_feature_toward(b, mb, -contact.normal, ang2, m2);
            clip(m1, m2, contact.normal, pred_linear, new_contacts);
        }
        if (new_contacts.empty()) new_contacts.push_back({contact, m1.feature_id, m2.feature_id});
    }
    for (auto& nc : new_contacts) {
        if (!feature_ok_for_manifold(m1, nc.f1)) continue;
        if (!feature_ok_for_manifold(m2, nc.f2)) continue;
        V3 local1 = iso_inv_point(ma, nc.c.world1);
        // ConvexPolygonalFeature::add_contact_to_manifold (convex_polygonal_feature3.rs:340-401)
        nc.c.k.local1 = local1;
        nc.c.k.local2 = iso_inv_point(mb, nc.c.world2);
        feature_geometry(m1, nc.f1, ma, &nc.c.k.g1, &nc.c.k.dir1);
        feature_geometry(m2, nc.f2, mb, &nc.c.k.g2, &nc.c.k.dir2);
        mf.push(nc.c, nc.f1, nc.f2, local1, pp1, pp2);
    }
}

// default_contact_dispatcher.rs:27-97
enum Algo : uint8_t { A_NONE = 0, A_BALL_BALL, A_PLANE_BALL, A_PLANE_CONVEX, A_BALL_CONVEX, A_CONVEX_CONVEX, A_PROXIMITY_RESERVED,
                      A_CAPSULE_CAPSULE = 7, A_CAPSULE_SHAPE = 8 };

static uint8_t generate_contacts(const Objects& o, uint32_t i1, uint32_t i2, Manifold& mf, GJKStats* st, V3* last_gjk_dir = nullptr,
                                 bool* has_dir = nullptr) {
    Shape a = get_shape(o, i1), b = get_shape(o, i2);
    Iso ma = o.iso(i1), mb = o.iso(i2);
    // query_type.rs:39-51
    real linear = o.query_limit[i1] + o.query_limit[i2];
    real ang1 = o.ang_pred[i1], ang2 = o.ang_pred[i2];
    // CapsuleCapsuleManifoldGenerator (capsule_capsule_manifold_generator.rs:24-55) / CapsuleShapeManifoldGenerator
    // (capsule_shape_manifold_generator.rs:23-75): the capsule is replaced by its segment, the linear prediction grows by the radius,
    // the capsule's preprocessor is attached on that side, and the sub-detector the dispatcher picks for (segment, other) /
    // (other, segment) runs with the shapes in their original order (`flip` only restores that order).
    Preproc ppa, ppb;
    uint8_t capsule_algo = A_NONE;
    if (a.type == CAPSULE || b.type == CAPSULE) {
        capsule_algo = (a.type == CAPSULE && b.type == CAPSULE) ? A_CAPSULE_CAPSULE : A_CAPSULE_SHAPE;
        if (a.type == CAPSULE) {
            ppa.active = true, ppa.radius = a.radius;
            linear = linear + a.radius;
            a.type = SEGMENT;
        }
        if (b.type == CAPSULE) {
            ppb.active = true, ppb.radius = b.radius;
            linear = linear + b.radius;
            b.type = SEGMENT;
        }
    }
    const Preproc* pa = ppa.active ? &ppa : nullptr;
    const Preproc* pb = ppb.active ? &ppb : nullptr;
    bool a_ball = a.type == BALL, b_ball = b.type == BALL, a_plane = a.type == PLANE, b_plane = b.type == PLANE;
    bool a_sm = a.type != PLANE, b_sm = b.type != PLANE;  // is_support_map
    uint8_t algo = A_NONE;
    if (a_ball && b_ball) {
        gen_ball_ball(ma, a.radius, mb, b.radius, linear, mf);
        algo = A_BALL_BALL;
    } else if (a_plane && b_ball) {
        gen_plane_ball(ma, a.he, mb, b.radius, linear, false, mf);
        algo = A_PLANE_BALL;
    } else if (a_ball && b_plane) {
        gen_plane_ball(mb, b.he, ma, a.radius, linear, true, mf);
        algo = A_PLANE_BALL;
    } else if (a_plane && b_sm) {
        gen_plane_convex(ma, a.he, mb, b, linear, false, mf, pa, pb);
        algo = A_PLANE_CONVEX;
    } else if (b_plane && a_sm) {
        gen_plane_convex(mb, b.he, ma, a, linear, true, mf, pb, pa);
        algo = A_PLANE_CONVEX;
    } else if (a_ball && is_convex_polyhedron(b)) {
        gen_ball_convex(ma, a.radius, mb, b, linear, false, mf, st, pa, pb);
        algo = A_BALL_CONVEX;
    } else if (b_ball && is_convex_polyhedron(a)) {
        gen_ball_convex(mb, b.radius, ma, a, linear, true, mf, st, pb, pa);
        algo = A_BALL_CONVEX;
    } else if (is_convex_polyhedron(a) && is_convex_polyhedron(b)) {
        gen_convex_convex(ma, a, mb, b, linear, ang1, ang2, mf, st, last_gjk_dir, has_dir, pa, pb);
        algo = A_CONVEX_CONVEX;
    }
    return capsule_algo != A_NONE ? capsule_algo : algo;  // A_NONE e.g. plane x plane: pair kept by the broad phase, no interaction edge
}

// ---------------------------------------------------------------------------------------------
// Proximity detectors: default_proximity_dispatcher.rs:19-47, ball_ball_proximity_detector.rs:24-43 +
// proximity_ball_ball.rs:8-36, plane_support_map_proximity_detector.rs:37-67 + proximity_plane_support_map.rs:9-47,
// support_map_support_map_proximity_detector.rs:31-61 + proximity_support_map_support_map.rs:36-75 (GJK, exact_dist = false).
// Proximity as a byte: Intersecting = 0, WithinMargin = 1, Disjoint = 2 (proximity.rs:4-12); 255 = no detector.
// ---------------------------------------------------------------------------------------------
enum : uint8_t { PROX_INTERSECTING = 0, PROX_WITHIN_MARGIN = 1, PROX_DISJOINT = 2, PROX_NONE = 255 };
static const uint8_t A_PROXIMITY = 6;

static uint8_t proximity_ball_ball(V3 c1, real r1, V3 c2, real r2, real margin) {
    V3 delta_pos = c2 - c1;
    real distance_squared = norm_squared(delta_pos);
    real sum_radius = r1 + r2;
    real sum_radius_with_error = sum_radius + margin;
    if (distance_squared <= sum_radius_with_error * sum_radius_with_error)
        return distance_squared <= sum_radius * sum_radius ? PROX_INTERSECTING : PROX_WITHIN_MARGIN;
    return PROX_DISJOINT;
}
static uint8_t proximity_plane_support_map(const Iso& mplane, V3 plane_n, const Iso& mother, const Support& other, real margin) {
    V3 plane_normal = iso_mul_vec(mplane, plane_n);
    V3 plane_center = mplane.t;
    V3 deepest = other.support_point_toward(mother, -plane_normal);
    real distance = dot(plane_normal, plane_center - deepest);
    if (distance >= -margin) return distance >= real(0) ? PROX_INTERSECTING : PROX_WITHIN_MARGIN;
    return PROX_DISJOINT;
}
// sep_axis / has_axis: SupportMapSupportMapProximityDetector::sep_axis (None on a fresh detector and after Intersecting)
static uint8_t proximity_support_map_support_map(const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, real margin, V3* sep_axis,
                                                 bool* has_axis) {
    V3 dir;
    if (has_axis && *has_axis)
        dir = *sep_axis;
    else if (!unit_try_new(m2.t - m1.t, EPS, &dir))
        dir = v3(1, 0, 0);
    VoronoiSimplex simplex;
    simplex.reset(cso_from_shapes(m1, g1, m2, g2, dir));
    GJKResult r = gjk_closest_points(m1, g1, m2, g2, margin, simplex, nullptr, false);
    uint8_t res;
    V3 out_dir = dir;
    if (r.kind == GJK_INTERSECTION)
        res = PROX_INTERSECTING;
    else if (r.kind == GJK_PROXIMITY)
        res = PROX_WITHIN_MARGIN, out_dir = r.dir;
    else
        res = PROX_DISJOINT, out_dir = r.dir;
    if (has_axis) {
        *has_axis = res != PROX_INTERSECTING;
        if (*has_axis) *sep_axis = out_dir;
    }
    return res;
}
// ProximityDetector::update of the detector the default dispatcher picks for (a, b).
static uint8_t proximity_update(const Objects& o, uint32_t i1, uint32_t i2, real margin, V3* sep_axis = nullptr, bool* has_axis = nullptr) {
    Shape a = get_shape(o, i1), b = get_shape(o, i2);
    Iso ma = o.iso(i1), mb = o.iso(i2);
    if (a.type == BALL && b.type == BALL) return proximity_ball_ball(ma.t, a.radius, mb.t, b.radius, margin);
    if (a.type == PLANE && b.type != PLANE) return proximity_plane_support_map(ma, a.he, mb, as_support(b), margin);
    if (b.type == PLANE && a.type != PLANE) return proximity_plane_support_map(mb, b.he, ma, as_support(a), margin);
    if (a.type != PLANE && b.type != PLANE) return proximity_support_map_support_map(ma, as_support(a), mb, as_support(b), margin, sep_axis, has_axis);
    return PROX_NONE;
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

static void write_contact(orc_contact* d, const Tracked& t) {
    for (int k = 0; k < 3; ++k) {
        d->world1[k] = t.c.world1[k];
        d->world2[k] = t.c.world2[k];
        d->normal[k] = t.c.normal[k];
    }
    d->depth = t.c.depth;
    d->f1 = t.f1;
    d->f2 = t.f2;
}

static void write_kinematic(orc_kinematic* d, const Tracked& t) {
    const Kin& k = t.c.k;
    for (int i = 0; i < 3; ++i) d->local1[i] = k.local1[i], d->local2[i] = k.local2[i], d->dir1[i] = k.dir1[i], d->dir2[i] = k.dir2[i];
    d->dil1 = k.dil1, d->dil2 = k.dil2;
    d->g1 = k.g1, d->g2 = k.g2;
}

extern "C" {

// kin_out (optional, same capacity as out): the ContactKinematic of every contact (contact_kinematic.rs:57-66)
uint64_t orc_narrow_phase_kin(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, orc_contact* out, orc_kinematic* kin_out,
                              uint64_t cap, uint32_t* manifold_off, uint8_t* algo_out, uint32_t* stats) {
    Objects o = make_objects(objs);
    uint64_t nc = 0;
    GJKStats st;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        Manifold mf;
        uint8_t algo = generate_contacts(o, pairs[2 * p], pairs[2 * p + 1], mf, &st);
        if (algo_out) algo_out[p] = algo;
        if (manifold_off) manifold_off[p] = (uint32_t)nc;
        mf.for_each_contact([&](const Tracked& t, size_t) {
            if (nc < cap && out) write_contact(&out[nc], t);
            if (nc < cap && kin_out) write_kinematic(&kin_out[nc], t);
            nc++;
        });
    }
    if (manifold_off) manifold_off[n_pairs] = (uint32_t)nc;
    if (stats) {
        stats[0] = st.gjk_iters, stats[1] = st.epa_iters, stats[2] = st.epa_max_verts, stats[3] = st.epa_max_faces;
        stats[4] = st.epa_max_heap, stats[5] = st.epa_calls, stats[6] = st.epa_fail;
    }
    return nc;
}
uint64_t orc_narrow_phase(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, orc_contact* out, uint64_t cap,
                          uint32_t* manifold_off, uint8_t* algo_out, uint32_t* stats) {
    return orc_narrow_phase_kin(objs, n_pairs, pairs, out, nullptr, cap, manifold_off, algo_out, stats);
}

void orc_world_update_timed(const orc_objects* objs, real margin, double* times, uint64_t* counts) {
    using clk = std::chrono::steady_clock;
    Objects o = make_objects(objs);
    auto t0 = clk::now();
    std::vector<AABB> fat(o.n);
    for (uint32_t i = 0; i < o.n; ++i) fat[i] = fat_aabb(o, i, margin);
    auto t1 = clk::now();
    std::vector<uint32_t> pairs;
    broad_phase_dbvt(o.n, fat.data(), o.groups, pairs);
    auto t2 = clk::now();
    uint64_t nc = 0, np = pairs.size() / 2, n_with = 0;
    for (uint64_t p = 0; p < np; ++p) {
        Manifold mf;
        generate_contacts(o, pairs[2 * p], pairs[2 * p + 1], mf, nullptr);
        nc += mf.len();
        n_with += mf.len() != 0;
    }
    auto t3 = clk::now();
    times[0] = std::chrono::duration<double>(t1 - t0).count();
    times[1] = std::chrono::duration<double>(t2 - t1).count();
    times[2] = std::chrono::duration<double>(t3 - t2).count();
    counts[0] = np;
    counts[1] = nc;
    counts[2] = n_with;
}

// query::contact (contact_shape_shape.rs:10-48) between objects 0 and 1.
int orc_query_contact(const orc_objects* objs, real prediction, orc_contact* out) {
    Objects o = make_objects(objs);
    Shape a = get_shape(o, 0), b = get_shape(o, 1);
    Iso m1 = o.iso(0), m2 = o.iso(1);
    Manifold mf;
    bool have = false;
    Contact c;
    if (a.type == BALL && b.type == BALL) {
        gen_ball_ball(m1, a.radius, m2, b.radius, prediction, mf);
    } else if (a.type == PLANE && b.type != PLANE) {
        // contact_plane_support_map.rs:7-27
        V3 n = iso_mul_vec(m1, a.he);
        Support g = as_support(b);
        V3 deepest = g.support_point_toward(m2, -n);
        real distance = dot(n, m1.t - deepest);
        if (distance > -prediction) {
            c = {deepest + n * distance, deepest, n, distance};
            have = true;
        }
    } else if (b.type == PLANE && a.type != PLANE) {
        V3 n = iso_mul_vec(m2, b.he);
        Support g = as_support(a);
        V3 deepest = g.support_point_toward(m1, -n);
        real distance = dot(n, m2.t - deepest);
        if (distance > -prediction) {
            c = {deepest, deepest + n * distance, -n, distance};
            have = true;
        }
    } else if (a.type == BALL && is_convex_polyhedron(b)) {
        gen_ball_convex(m1, a.radius, m2, b, prediction, false, mf, nullptr);
    } else if (b.type == BALL && is_convex_polyhedron(a)) {
        gen_ball_convex(m2, b.radius, m1, a, prediction, true, mf, nullptr);
    } else if (a.type != PLANE && b.type != PLANE) {
        Support ga = as_support(a), gb = as_support(b);
        VoronoiSimplex simplex;
        GJKResult r = contact_support_map_support_map_with_params(m1, ga, m2, gb, prediction, simplex, nullptr, nullptr);
        if (r.kind == GJK_CLOSEST_POINTS) {
            c = contact_new_wo_depth(r.p1, r.p2, r.dir);
            have = true;
        }
    }
    if (!have && mf.len() != 0) {
        mf.for_each_contact([&](const Tracked& t, size_t) {
            if (!have) c = t.c;
            have = true;
        });
    }
    if (have) {
        Tracked t = {c, FID_UNKNOWN, FID_UNKNOWN};
        write_contact(out, t);
    }
    return have ? 1 : 0;
}

// contact_support_map_support_map (contact_support_map_support_map.rs:12-79) with a fresh simplex for a batch of cuboid / hull
// pairs: out[10 p] = p1, p2, normal, kind (0 = no contact within the prediction, 1 = closest points / penetration).
void orc_contact_sm_sm(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, const real* predictions, real* out, uint32_t* stats) {
    Objects o = make_objects(objs);
    GJKStats st;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        Shape a = get_shape(o, i1), b = get_shape(o, i2);
        // a capsule takes part through its segment, as in the capsule generators (capsule_capsule_manifold_generator.rs:35-36)
        if (a.type == CAPSULE) a.type = SEGMENT;
        if (b.type == CAPSULE) b.type = SEGMENT;
        real prediction = predictions ? predictions[p] : o.query_limit[i1] + o.query_limit[i2];
        VoronoiSimplex simplex;
        GJKResult r = contact_support_map_support_map_with_params(o.iso(i1), as_support(a), o.iso(i2), as_support(b), prediction, simplex, nullptr, &st);
        real* d = out + 10 * p;
        for (int k = 0; k < 10; ++k) d[k] = 0;
        if (r.kind == GJK_CLOSEST_POINTS) {
            for (int k = 0; k < 3; ++k) d[k] = r.p1[k], d[3 + k] = r.p2[k], d[6 + k] = r.dir[k];
            d[9] = 1;
        }
    }
    if (stats) stats[0] = st.gjk_iters, stats[1] = st.epa_iters, stats[2] = st.epa_calls, stats[3] = st.epa_fail;
}

// The reference's issue-#157 test (build/ncollide3d/tests/geometry/cylinder_cuboid_contact.rs): Cylinder(half_height, radius) at t1
// vs Cuboid(he) at t2, identity rotations.  out[0] = distance_support_map_support_map (distance_support_map_support_map.rs:8-53),
// out[1] = proximity_support_map_support_map(margin) as ORC_PROX_*, out[2] = contact_support_map_support_map(prediction).is_some().
void orc_kat_cylinder_cuboid(real half_height, real radius, const real* t1, const real* he, const real* t2, real margin, real prediction,
                             real* out) {
    Iso m1{{t1[0], t1[1], t1[2]}, {0, 0, 0, 1}}, m2{{t2[0], t2[1], t2[2]}, {0, 0, 0, 1}};
    Support g1, g2;
    g1.kind = Support::S_CYLINDER, g1.he = v3(half_height, 0, 0), g1.radius = radius;
    g2.kind = Support::S_CUBOID, g2.he = v3(he[0], he[1], he[2]), g2.radius = 0;
    {
        V3 dir;
        if (!unit_try_new(m1.t - m2.t, EPS, &dir)) dir = v3(1, 0, 0);
        VoronoiSimplex simplex;
        simplex.reset(cso_from_shapes(m1, g1, m2, g2, dir));
        GJKResult r = gjk_closest_points(m1, g1, m2, g2, FMAX, simplex, nullptr, true);
        out[0] = r.kind == GJK_CLOSEST_POINTS ? norm(r.p1 - r.p2) : real(0);
    }
    out[1] = proximity_support_map_support_map(m1, g1, m2, g2, margin, nullptr, nullptr);
    {
        VoronoiSimplex simplex;
        GJKResult r = contact_support_map_support_map_with_params(m1, g1, m2, g2, prediction, simplex, nullptr, nullptr);
        out[2] = r.kind == GJK_CLOSEST_POINTS ? real(1) : real(0);
    }
}

void orc_proximity(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, const real* margins, uint8_t* out) {
    Objects o = make_objects(objs);
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        real margin = margins ? margins[p] : o.query_limit[i1] + o.query_limit[i2];
        out[p] = proximity_update(o, i1, i2, margin);
    }
}
void orc_proximity_warm(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, const real* margins, real* axis_io, uint8_t* out) {
    Objects o = make_objects(objs);
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        real margin = margins ? margins[p] : o.query_limit[i1] + o.query_limit[i2];
        V3 axis = v3(axis_io[4 * p], axis_io[4 * p + 1], axis_io[4 * p + 2]);
        bool has_axis = axis_io[4 * p + 3] != real(0);
        out[p] = proximity_update(o, i1, i2, margin, &axis, &has_axis);
        axis_io[4 * p] = axis.x, axis_io[4 * p + 1] = axis.y, axis_io[4 * p + 2] = axis.z, axis_io[4 * p + 3] = has_axis ? real(1) : real(0);
    }
}
int orc_query_proximity(const orc_objects* objs, real margin) {
    Objects o = make_objects(objs);
    return proximity_update(o, 0, 1, margin);
}
uint64_t orc_narrow_phase_kinds(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, orc_contact* out, uint64_t cap,
                                uint32_t* manifold_off, uint8_t* algo_out, uint8_t* prox_out) {
    Objects o = make_objects(objs);
    uint64_t nc = 0;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        if (manifold_off) manifold_off[p] = (uint32_t)nc;
        if (o.is_proximity(i1) || o.is_proximity(i2)) {  // narrow_phase.rs:226-247
            uint8_t st = proximity_update(o, i1, i2, o.query_limit[i1] + o.query_limit[i2]);
            if (prox_out) prox_out[p] = st;
            if (algo_out) algo_out[p] = st == PROX_NONE ? A_NONE : A_PROXIMITY;
            continue;
        }
        Manifold mf;
        uint8_t algo = generate_contacts(o, i1, i2, mf, nullptr);
        if (algo_out) algo_out[p] = algo;
        if (prox_out) prox_out[p] = PROX_NONE;
        mf.for_each_contact([&](const Tracked& t, size_t) {
            if (nc < cap && out) write_contact(&out[nc], t);
            nc++;
        });
    }
    if (manifold_off) manifold_off[n_pairs] = (uint32_t)nc;
    return nc;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Stepping world: CollisionWorld::update over several steps (pipeline/world.rs:104-119, glue/update.rs:65-138,
// narrow_phase.rs:56-104,168-278): persistent broad phase (bp_persistent.cpp), interaction edges created /
// removed by its started / stopped callbacks with the callback's argument order, per-edge generator state
// (last_gjk_dir) and ContactManifold cache, contact events, update flags.
// ---------------------------------------------------------------------------------------------
#include <map>

namespace orc {
AABB shape_aabb(const Objects& o, uint32_t i);
struct Edge {
    uint32_t h1, h2;
    uint8_t algo;
    Manifold manifold;
    V3 last_gjk_dir{0, 0, 0};
    bool has_dir = false;
    // Interaction::Proximity(detector, status) (interaction_graph.rs): status + the detector's sep_axis
    bool is_prox = false;
    uint8_t prox = 2;  // Proximity::Disjoint
    V3 sep_axis{0, 0, 0};
    bool has_axis = false;
};
}  // namespace orc

struct orc_sim {
    Objects o;
    std::vector<real> pos, rot, param, qlimit, ang;  // the sim owns the per-object arrays (objects can be added later)
    std::vector<uint32_t> type, groups;
    std::vector<uint8_t> qkind;  // empty = all Contacts
    std::vector<uint32_t> prox_events;  // (h1, h2, prev, new)
    std::vector<uint8_t> flags;  // 1 = POSITION_CHANGED (| everything, for a new object)
    std::vector<uint8_t> alive;
    std::vector<uint32_t> free_handles;  // CollisionObjectSlab: vacant keys, reused last-freed-first
    void rebind() {
        o.n = (uint32_t)alive.size();
        o.pos = pos.data(), o.rot = rot.data(), o.shape_type = type.data(), o.shape_param = param.data();
        o.groups = groups.empty() ? nullptr : groups.data();
        o.query_limit = qlimit.data(), o.ang_pred = ang.data();
        o.query_kind = qkind.empty() ? nullptr : qkind.data();
    }
    orc_bp* bp = nullptr;
    std::map<uint64_t, Edge> edges;  // key = min << 32 | max (iteration order is not observable: events are sorted)
    std::vector<uint32_t> events;    // (h1, h2, started)
    bool first = true;
};

static uint8_t dispatch_algo(const Objects& o, uint32_t i1, uint32_t i2) {  // default_contact_dispatcher.rs:27-97
    uint32_t a = o.shape_type[i1], b = o.shape_type[i2];
    if (a == CAPSULE && b == CAPSULE) return A_CAPSULE_CAPSULE;
    if (a == CAPSULE || b == CAPSULE) return A_CAPSULE_SHAPE;
    if (a == BALL && b == BALL) return A_BALL_BALL;
    if ((a == PLANE && b == BALL) || (a == BALL && b == PLANE)) return A_PLANE_BALL;
    if (a == PLANE && b == PLANE) return A_NONE;
    if (a == PLANE || b == PLANE) return A_PLANE_CONVEX;
    if (a == BALL || b == BALL) return A_BALL_CONVEX;
    return A_CONVEX_CONVEX;
}

extern "C" {

orc_sim* orc_sim_create(const orc_objects* objs, real margin) {
    orc_sim* s = new orc_sim;
    s->o = make_objects(objs);
    size_t n0 = objs->n;
    s->pos.assign(objs->pos, objs->pos + 3 * n0);
    s->rot.assign(objs->rot, objs->rot + 4 * n0);
    s->type.assign(objs->shape_type, objs->shape_type + n0);
    s->param.assign(objs->shape_param, objs->shape_param + 4 * n0);
    if (objs->groups) s->groups.assign(objs->groups, objs->groups + 3 * n0);
    s->qlimit.assign(objs->query_limit, objs->query_limit + n0);
    s->ang.assign(objs->ang_pred, objs->ang_pred + n0);
    if (objs->query_kind) s->qkind.assign(objs->query_kind, objs->query_kind + n0);
    s->flags.assign(n0, 1);
    s->alive.assign(n0, 1);
    s->rebind();
    s->bp = orc_bp_create(margin);
    // CollisionWorld::add -> glue::create_proxies (glue/setup.rs:20-36): proxy box = compute_swept_aabb()
    for (uint32_t i = 0; i < s->o.n; ++i) {
        AABB a = shape_aabb(s->o, i);
        real ql = s->o.query_limit[i];
        real mm[6] = {a.mins.x + (-ql), a.mins.y + (-ql), a.mins.z + (-ql), a.maxs.x + ql, a.maxs.y + ql, a.maxs.z + ql};
        uint32_t h = orc_bp_create_proxy(s->bp, mm);
        (void)h;
    }
    return s;
}
void orc_sim_destroy(orc_sim* s) {
    if (!s) return;
    orc_bp_destroy(s->bp);
    delete s;
}
// CollisionObject::set_position for n objects (handles == NULL: objects 0..n-1)
void orc_sim_set_positions(orc_sim* s, uint32_t n, const uint32_t* handles, const real* pos, const real* rot) {
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t h = handles ? handles[k] : k;
        for (int d = 0; d < 3; ++d) s->pos[3 * (size_t)h + d] = pos[3 * (size_t)k + d];
        for (int d = 0; d < 4; ++d) s->rot[4 * (size_t)h + d] = rot[4 * (size_t)k + d];
        if (h < s->alive.size() && s->alive[h] && !(s->flags[h] & 2)) s->flags[h] |= 1;
    }
}

void orc_sim_step(orc_sim* s) {
    const Objects& o = s->o;
    s->events.clear();  // narrow_phase.clear_events()
    s->prox_events.clear();
    // perform_broad_phase (glue/update.rs:65-99)
    for (uint32_t i = 0; i < o.n; ++i) {
        if (!s->alive[i] || !s->flags[i]) continue;
        // flags: 1 POSITION_CHANGED, 2 a new object (every flag), 4 COLLISION_GROUPS_CHANGED (collision_object.rs:10-59)
        if (s->flags[i] & 3) {  // needs_bounding_volume_update
            AABB a = shape_aabb(o, i);
            real ql = o.query_limit[i];
            real mm[6] = {a.mins.x + (-ql), a.mins.y + (-ql), a.mins.z + (-ql), a.maxs.x + ql, a.maxs.y + ql, a.maxs.z + ql};
            orc_bp_set_bounding_volume(s->bp, i, mm);
        }
        // needs_broad_phase_redispatch: a new object also has SHAPE_CHANGED etc. set (a no-op on a detached proxy); changed groups
        if ((s->flags[i] & 6) || s->first) orc_bp_recompute_with(s->bp, i);
    }
    uint64_t cap = 1 << 16, ns = 0, np = 0;
    std::vector<uint32_t> st, sp;
    // the call mutates state, so size the buffers from the upper bounds first
    cap = std::max<uint64_t>(cap, 64ull * o.n + 1024);
    cap = std::max<uint64_t>(cap, std::min<uint64_t>((uint64_t)o.n * o.n / 2 + 1024, 1ull << 26));  // dense small worlds: every pair can start at once
    st.resize(2 * cap), sp.resize(2 * cap);
    orc_bp_update(s->bp, o.groups, st.data(), cap, &ns, sp.data(), cap, &np);
    if (ns > cap || np > cap) abort();
    // interference_started -> handle_interaction(.., true) (narrow_phase.rs:216-247)
    for (uint64_t k = 0; k < ns; ++k) {
        uint32_t b1 = st[2 * k], b2 = st[2 * k + 1];
        uint64_t key = ((uint64_t)std::min(b1, b2) << 32) | std::max(b1, b2);
        if (s->edges.count(key)) continue;
        uint8_t algo = dispatch_algo(o, b1, b2);
        if (algo == A_NONE) continue;  // plane x plane: neither dispatcher has an algorithm
        Edge e;
        e.h1 = b1, e.h2 = b2, e.algo = algo;
        if (o.is_proximity(b1) || o.is_proximity(b2)) e.is_prox = true, e.algo = A_PROXIMITY;  // narrow_phase.rs:240-246
        s->edges.emplace(key, std::move(e));
    }
    // interference_stopped -> handle_interaction(.., false) (:248-277)
    for (uint64_t k = 0; k < np; ++k) {
        uint64_t key = ((uint64_t)sp[2 * k] << 32) | sp[2 * k + 1];
        auto it = s->edges.find(key);
        if (it == s->edges.end()) continue;
        if (it->second.is_prox) {  // narrow_phase.rs:266-274
            if (it->second.prox != PROX_DISJOINT)
                s->prox_events.insert(s->prox_events.end(), {it->second.h1, it->second.h2, (uint32_t)it->second.prox, (uint32_t)PROX_DISJOINT});
        } else if (it->second.manifold.len() != 0)
            s->events.insert(s->events.end(), {it->second.h1, it->second.h2, 0u});
        s->edges.erase(it);
    }
    // perform_narrow_phase -> NarrowPhase::update (:168-197) -> update_contact (:56-104)
    for (auto& kv : s->edges) {
        Edge& e = kv.second;
        if (!s->flags[e.h1] && !s->flags[e.h2]) continue;
        if (e.is_prox) {  // update_proximity (narrow_phase.rs:123-143)
            uint8_t np_ = proximity_update(o, e.h1, e.h2, o.query_limit[e.h1] + o.query_limit[e.h2], &e.sep_axis, &e.has_axis);
            if (np_ != e.prox) s->prox_events.insert(s->prox_events.end(), {e.h1, e.h2, (uint32_t)e.prox, (uint32_t)np_});
            e.prox = np_;
            continue;
        }
        bool had = e.manifold.len() != 0;
        e.manifold.save_cache_and_clear();
        generate_contacts(o, e.h1, e.h2, e.manifold, nullptr, &e.last_gjk_dir, &e.has_dir);
        bool has = e.manifold.len() != 0;
        if (!has && had) s->events.insert(s->events.end(), {e.h1, e.h2, 0u});
        if (has && !had) s->events.insert(s->events.end(), {e.h1, e.h2, 1u});
    }
    std::fill(s->flags.begin(), s->flags.end(), 0);
    s->first = false;
}

// CollisionObject::set_collision_groups (collision_object.rs:246-250)
int orc_sim_set_collision_groups(orc_sim* s, uint32_t n, const uint32_t* handles, const uint32_t* groups) {
    for (uint32_t k = 0; k < n; ++k)
        if (handles[k] >= s->alive.size() || !s->alive[handles[k]]) return -1;
    if (s->groups.empty()) {  // CollisionGroups::new() for everybody so far
        s->groups.resize(3 * s->alive.size());
        for (size_t i = 0; i < s->alive.size(); ++i) s->groups[3 * i] = s->groups[3 * i + 1] = 0x3FFFFFFFu, s->groups[3 * i + 2] = 0;
        s->rebind();
    }
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t h = handles[k];
        for (int d = 0; d < 3; ++d) s->groups[3 * (size_t)h + d] = groups[3 * (size_t)k + d];
        s->flags[h] |= 4;
    }
    return 0;
}

// CollisionWorld::remove (world.rs:129-144): per handle, objects.remove + glue::remove_proxies (glue/setup.rs:50-62): the
// broad phase drops the proxy and its pairs WITHOUT notifying anybody, the interaction-graph node goes with its edges.
int orc_sim_remove(orc_sim* s, uint32_t n, const uint32_t* handles) {
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t h = handles[k];
        if (h >= s->alive.size() || !s->alive[h]) return -1;
        s->alive[h] = 0;
        s->flags[h] = 0;
        s->free_handles.push_back(h);
        if (orc_bp_remove(s->bp, 1, &h, nullptr, 0, nullptr) != 0) return -1;
        for (auto it = s->edges.begin(); it != s->edges.end();) {
            if (it->second.h1 == h || it->second.h2 == h)
                it = s->edges.erase(it);
            else
                ++it;
        }
    }
    return 0;
}
// CollisionWorld::add (world.rs:64-96) for the n objects of `objs` (same hull library): handles come from the slab
// (last freed first, else appended); the proxy is created with the swept AABB; every update flag is set.
int orc_sim_add(orc_sim* s, const orc_objects* objs, uint32_t* out_handles) {
    for (uint32_t k = 0; k < objs->n; ++k) {
        uint32_t h;
        if (!s->free_handles.empty()) {
            h = s->free_handles.back();
            s->free_handles.pop_back();
        } else {
            h = (uint32_t)s->alive.size();
            s->alive.push_back(0), s->flags.push_back(0), s->type.push_back(0), s->qlimit.push_back(0), s->ang.push_back(0);
            if (!s->qkind.empty()) s->qkind.push_back(0);
            s->pos.resize(s->pos.size() + 3), s->rot.resize(s->rot.size() + 4), s->param.resize(s->param.size() + 4);
            if (!s->groups.empty() || objs->groups) {
                size_t old = s->groups.size() / 3;
                s->groups.resize(3 * s->alive.size());
                for (size_t i = old; i < s->alive.size(); ++i) s->groups[3 * i] = s->groups[3 * i + 1] = 0x3FFFFFFFu, s->groups[3 * i + 2] = 0;
            }
        }
        for (int d = 0; d < 3; ++d) s->pos[3 * (size_t)h + d] = objs->pos[3 * (size_t)k + d];
        for (int d = 0; d < 4; ++d) s->rot[4 * (size_t)h + d] = objs->rot[4 * (size_t)k + d];
        for (int d = 0; d < 4; ++d) s->param[4 * (size_t)h + d] = objs->shape_param[4 * (size_t)k + d];
        s->type[h] = objs->shape_type[k];
        s->qlimit[h] = objs->query_limit[k];
        s->ang[h] = objs->ang_pred[k];
        if (objs->query_kind && objs->query_kind[k] && s->qkind.empty()) s->qkind.assign(s->alive.size(), 0);
        if (!s->qkind.empty()) s->qkind[h] = objs->query_kind ? objs->query_kind[k] : 0;
        if (!s->groups.empty()) {
            for (int d = 0; d < 3; ++d) s->groups[3 * (size_t)h + d] = objs->groups ? objs->groups[3 * (size_t)k + d] : (d < 2 ? 0x3FFFFFFFu : 0u);
        }
        s->alive[h] = 1;
        s->flags[h] = 2;
        s->rebind();
        AABB a = shape_aabb(s->o, h);
        real ql = s->o.query_limit[h];
        real mm[6] = {a.mins.x + (-ql), a.mins.y + (-ql), a.mins.z + (-ql), a.maxs.x + ql, a.maxs.y + ql, a.maxs.z + ql};
        uint32_t ph = orc_bp_create_proxy(s->bp, mm);
        if (ph != h) return -1;  // object slab and proxy slab evolve in lockstep
        if (out_handles) out_handles[k] = h;
    }
    return 0;
}

uint64_t orc_sim_num_pairs(const orc_sim* s) { return s->edges.size(); }
uint64_t orc_sim_num_contacts(const orc_sim* s) {
    uint64_t n = 0;
    for (auto& kv : s->edges) n += kv.second.manifold.len();
    return n;
}
// Edges sorted by (min handle, max handle): pairs = (h1, h2) in the edge's own orientation; manifold_off has n + 1 entries;
// ids[k] = insertion counter << 8 | slab slot of contact k inside its manifold.
void orc_sim_fetch(const orc_sim* s, uint32_t* pairs, uint8_t* algo, uint32_t* manifold_off, orc_contact* contacts, uint32_t* ids) {
    uint64_t p = 0, nc = 0;
    for (auto& kv : s->edges) {
        const Edge& e = kv.second;
        pairs[2 * p] = e.h1, pairs[2 * p + 1] = e.h2;
        algo[p] = e.algo;
        manifold_off[p] = (uint32_t)nc;
        e.manifold.for_each_contact([&](const Tracked& t, size_t slot) {
            write_contact(&contacts[nc], t);
            ids[nc] = (t.id << 8) | (uint32_t)slot;
            nc++;
        });
        p++;
    }
    manifold_off[p] = (uint32_t)nc;
}
// ContactEvents of the last step as (h1, h2, 1 = Started / 0 = Stopped) triples, in emission order.
uint64_t orc_sim_events(const orc_sim* s, uint32_t* out, uint64_t cap) {
    uint64_t n = s->events.size() / 3;
    for (uint64_t k = 0; k < n && k < cap; ++k)
        for (int d = 0; d < 3; ++d) out[3 * k + d] = s->events[3 * k + d];
    return n;
}
// ---- a bare set of Interaction::Contact edges driven by the caller (no broad phase): NarrowPhase::update_contact
// (narrow_phase.rs:56-104) on the listed edges, with the generator state (last_gjk_dir) and the ContactManifold cache kept per edge.
struct orc_edges {
    std::vector<Edge> e;
};
orc_edges* orc_edges_create(uint64_t n, const uint32_t* pairs) {
    orc_edges* E = new orc_edges;
    E->e.resize(n);
    for (uint64_t i = 0; i < n; ++i) E->e[i].h1 = pairs[2 * i], E->e[i].h2 = pairs[2 * i + 1], E->e[i].algo = A_NONE;
    return E;
}
void orc_edges_destroy(orc_edges* E) { delete E; }
// events: (h1, h2, 1 Started / 0 Stopped) rows in emission order; returns their number
uint64_t orc_edges_update(orc_edges* E, const orc_objects* objs, uint64_t n_update, const uint32_t* which, uint32_t* events, uint64_t cap) {
    Objects o = make_objects(objs);
    uint64_t ne = 0;
    for (uint64_t k = 0; k < n_update; ++k) {
        Edge& e = E->e[which[k]];
        bool had = e.manifold.len() != 0;
        e.manifold.save_cache_and_clear();
        e.algo = generate_contacts(o, e.h1, e.h2, e.manifold, nullptr, &e.last_gjk_dir, &e.has_dir);
        bool has = e.manifold.len() != 0;
        if (has != had) {
            if (ne < cap) events[3 * ne] = e.h1, events[3 * ne + 1] = e.h2, events[3 * ne + 2] = has ? 1u : 0u;
            ne++;
        }
    }
    return ne;
}
// manifold_off[n + 1]; contacts in ContactManifold::contacts() order; ids = insertion counter << 8 | slab slot; dirs[4 i] = the
// generator's last_gjk_dir + a "Some" flag.  Returns the number of contacts (may exceed cap).
uint64_t orc_edges_fetch(const orc_edges* E, uint32_t* manifold_off, orc_contact* contacts, uint32_t* ids, uint64_t cap, real* dirs) {
    uint64_t nc = 0, n = E->e.size();
    for (uint64_t i = 0; i < n; ++i) {
        const Edge& e = E->e[i];
        manifold_off[i] = (uint32_t)nc;
        e.manifold.for_each_contact([&](const Tracked& t, size_t slot) {
            if (nc < cap) {
                write_contact(&contacts[nc], t);
                ids[nc] = (t.id << 8) | (uint32_t)slot;
            }
            nc++;
        });
        if (dirs) dirs[4 * i] = e.last_gjk_dir.x, dirs[4 * i + 1] = e.last_gjk_dir.y, dirs[4 * i + 2] = e.last_gjk_dir.z, dirs[4 * i + 3] = e.has_dir ? 1 : 0;
    }
    manifold_off[n] = (uint32_t)nc;
    return nc;
}

void orc_sim_fetch_proximity(const orc_sim* s, uint8_t* prox) {
    uint64_t p = 0;
    for (auto& kv : s->edges) prox[p++] = kv.second.is_prox ? kv.second.prox : (uint8_t)PROX_NONE;
}
uint64_t orc_sim_proximity_events(const orc_sim* s, uint32_t* out, uint64_t cap) {
    uint64_t n = s->prox_events.size() / 4;
    for (uint64_t k = 0; k < n && k < cap; ++k)
        for (int d = 0; d < 4; ++d) out[4 * k + d] = s->prox_events[4 * k + d];
    return n;
}
uint64_t orc_sim_bp_num_interferences(const orc_sim* s) { return orc_bp_num_interferences(s->bp); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// World ray queries (pipeline/glue/query.rs:13-77,183-224): broad-phase candidates (interferences_with_ray on the stored
// boxes) -> collision-groups test against the query's groups -> shape.toi_and_normal_with_ray(position, ray, max_toi, true).
// Shape ray casts: query/ray/ray_ball.rs:77-142, ray_cuboid.rs + ray_aabb.rs:52-75,183-300, ray_plane.rs:44-79,
// ray_support_map.rs:15-35,139-160 + gjk.rs:180-365 (ConvexHull).  solid = true only (what the world queries use).
// ---------------------------------------------------------------------------------------------
namespace orc {
struct RayHit {
    bool hit = false;
    real toi = 0;
    V3 normal{0, 0, 0};
    uint32_t feature = FID_UNKNOWN;
};
static RayHit ray_cast_ball(V3 center, real radius, V3 o, V3 d, real max_toi) {
    RayHit h;
    V3 dcenter = o - center;
    real a = norm_squared(d), b = dot(dcenter, d), c = norm_squared(dcenter) - radius * radius;
    bool inside = false, some = false;
    real t = 0;
    if (a == real(0)) {
        if (c > real(0)) return h;
        inside = true, some = true, t = 0;
    } else if (c > real(0) && b > real(0)) {
        return h;
    } else {
        real delta = b * b - a * c;
        if (delta < real(0)) return h;
        t = (-b - std::sqrt(delta)) / a;
        if (t <= real(0)) {
            inside = true, t = 0;  // solid
        }
        some = true;
    }
    if (!some || !(t <= max_toi)) return h;
    V3 pos = o + d * t - center;
    V3 normal = normalize(pos);
    h.hit = true, h.toi = t, h.normal = inside ? -normal : normal, h.feature = FACE0;
    return h;
}
static RayHit ray_cast_cuboid(V3 he, const Iso& m, V3 o_w, V3 d_w, real max_toi) {
    RayHit h;
    V3 o = iso_inv_point(m, o_w), d = iso_inv_vec(m, d_w);
    const real oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, mn[3] = {-he.x, -he.y, -he.z}, mx[3] = {he.x, he.y, he.z};
    // clip_line (ray_aabb.rs:183-279)
    real tmax = FMAX, tmin = -FMAX;
    int near_side = 0, far_side = 0;
    bool near_diag = false, far_diag = false;
    for (int i = 0; i < 3; ++i) {
        if (dd[i] == real(0)) {
            if (oo[i] < mn[i] || oo[i] > mx[i]) return h;
        } else {
            real denom = real(1) / dd[i];
            real near = (mn[i] - oo[i]) * denom, far = (mx[i] - oo[i]) * denom;
            bool flip = false;
            if (near > far) {
                flip = true;
                std::swap(near, far);
            }
            if (near > tmin) {
                tmin = near;
                near_side = flip ? -(i + 1) : (i + 1);
                near_diag = false;
            } else if (near == tmin) {
                near_diag = true;
            }
            if (far < tmax) {
                tmax = far;
                far_side = !flip ? -(i + 1) : (i + 1);
                far_diag = false;
            } else if (far == tmax) {
                far_diag = true;
            }
            if (tmax < real(0) || tmin > tmax) return h;
        }
    }
    V3 near_n{0, 0, 0};
    if (near_diag)
        near_n = -normalize(d);
    else if (near_side != 0) {
        real* c = &near_n.x;
        if (near_side < 0)
            c[-near_side - 1] = real(1);
        else
            c[near_side - 1] = -real(1);
    }
    // ray_aabb (:282-300), solid = true
    real t;
    V3 n;
    int side;
    if (tmin < real(0)) {
        t = 0, n = v3(0, 0, 0), side = far_side;
    } else if (tmin <= max_toi) {
        t = tmin, n = near_n, side = near_side;
    } else {
        return h;
    }
    h.hit = true, h.toi = t, h.normal = iso_mul_vec(m, n);
    // side == 0 (zero direction, origin inside): Face(0usize - 1) wraps in a release build of the reference; ids are 30 bits here
    h.feature = (F_FACE << 30) | ((uint32_t)(side < 0 ? (-side - 1 + 3) : (side - 1)) & 0x3fffffffu);
    return h;
}
static RayHit ray_cast_plane(V3 pn, const Iso& m, V3 o_w, V3 d_w, real max_toi) {
    RayHit h;
    V3 o = iso_inv_point(m, o_w), d = iso_inv_vec(m, d_w);
    V3 dpos = -o;
    real dot_normal_dpos = dot(pn, dpos);
    if (dot_normal_dpos > real(0)) {  // solid: inside the half-space
        h.hit = true, h.toi = 0, h.normal = v3(0, 0, 0), h.feature = FACE0;
        return h;
    }
    real t = dot_normal_dpos / dot(pn, d);
    if (t >= real(0) && t <= max_toi) {
        V3 n = dot_normal_dpos > real(0) ? -pn : pn;
        h.hit = true, h.toi = t, h.normal = iso_mul_vec(m, n), h.feature = FACE0;
    }
    return h;
}
static RayHit ray_cast_support_map(const Support& g, const Iso& m, V3 o_w, V3 d_w, real max_toi);
static RayHit ray_cast_hull(const Hull& H, const Iso& m, V3 o_w, V3 d_w, real max_toi) {
    Support g;
    g.kind = Support::S_HULL;
    g.hull = H;
    return ray_cast_support_map(g, m, o_w, d_w, max_toi);
}
// ray_support_map.rs:15-35,114-137 (ConvexHull and Capsule share it): gjk::cast_ray in the shape's local frame, solid = true
static RayHit ray_cast_support_map(const Support& g, const Iso& m, V3 o_w, V3 d_w, real max_toi) {
    RayHit h;
    V3 o = iso_inv_point(m, o_w), d = iso_inv_vec(m, d_w);
    Support origin;
    origin.kind = Support::S_ORIGIN;
    Iso id = iso_identity();
    VoronoiSimplex simplex;
    V3 supp = g.support_point(id, -d);
    V3 p = supp - o;
    simplex.reset(CSOPoint{p, p, v3(0, 0, 0)});  // overwritten by minkowski_ray_cast's own reset, like the reference
    real toi;
    V3 normal;
    if (!minkowski_ray_cast(id, g, id, origin, o, d, max_toi, simplex, &toi, &normal)) return h;
    h.hit = true, h.toi = toi, h.normal = iso_mul_vec(m, normal), h.feature = FID_UNKNOWN;
    return h;
}
static RayHit shape_ray_cast(const Objects& o, uint32_t i, V3 ro, V3 rd, real max_toi) {
    Shape s = get_shape(o, i);
    Iso m = o.iso(i);
    switch (s.type) {
        case BALL: return ray_cast_ball(m.t, s.radius, ro, rd, max_toi);
        case CUBOID: return ray_cast_cuboid(s.he, m, ro, rd, max_toi);
        case HULL: return ray_cast_hull(s.hull, m, ro, rd, max_toi);
        case CAPSULE: return ray_cast_support_map(as_support(s), m, ro, rd, max_toi);
        default: return ray_cast_plane(s.he, m, ro, rd, max_toi);
    }
}
}  // namespace orc

extern "C" {

// One shape ray cast (RayCast::toi_and_normal_with_ray, solid = true) against object i; out = toi, normal xyz.
int orc_shape_ray_cast(const orc_objects* objs, uint32_t i, const real* origin, const real* dir, real max_toi, real* out, uint32_t* feature) {
    Objects o = make_objects(objs);
    RayHit h = shape_ray_cast(o, i, v3(origin[0], origin[1], origin[2]), v3(dir[0], dir[1], dir[2]), max_toi);
    if (!h.hit) return 0;
    out[0] = h.toi, out[1] = h.normal.x, out[2] = h.normal.y, out[3] = h.normal.z;
    *feature = h.feature;
    return 1;
}

// The same for a batch: ray k (7 reals: origin, dir, max_toi) against object which[k]; out[4 k] = toi, normal; feat[k]; hit[k].
void orc_shape_ray_cast_batch(const orc_objects* objs, uint64_t n, const uint32_t* which, const real* rays, real* out, uint32_t* feat, uint8_t* hit) {
    Objects o = make_objects(objs);
    for (uint64_t k = 0; k < n; ++k) {
        const real* q = rays + 7 * k;
        RayHit h = shape_ray_cast(o, which[k], v3(q[0], q[1], q[2]), v3(q[3], q[4], q[5]), q[6]);
        hit[k] = h.hit ? 1 : 0;
        out[4 * k] = h.hit ? h.toi : 0, out[4 * k + 1] = h.hit ? h.normal.x : 0, out[4 * k + 2] = h.hit ? h.normal.y : 0, out[4 * k + 3] = h.hit ? h.normal.z : 0;
        feat[k] = h.hit ? h.feature : 0u;
    }
}
// PointQuery::contains_point(position, point) of object which[k] for point k (point_ball.rs:45-47, point_cuboid.rs, point_plane.rs:8-19,
// point_support_map.rs:15-53 for hulls).
static bool shape_contains_point(const Objects& o, uint32_t h, V3 pt) {
    Shape sh = get_shape(o, h);
    Iso m = o.iso(h);
    if (sh.type == BALL) return norm_squared(iso_inv_point(m, pt)) <= sh.radius * sh.radius;
    if (sh.type == CUBOID) {
        V3 l = iso_inv_point(m, pt);
        return !(l.x < -sh.he.x || l.x > sh.he.x || l.y < -sh.he.y || l.y > sh.he.y || l.z < -sh.he.z || l.z > sh.he.z);
    }
    if (sh.type == HULL) {
        bool inside;
        V3 proj;
        hull_project_point(sh.hull, m, pt, &inside, &proj, nullptr);
        return inside;
    }
    if (sh.type == CAPSULE) {  // point_capsule.rs:6-33 (contains_point = project_point(.., solid = true).is_inside)
        bool seg_inside;
        V3 proj;
        uint32_t f;
        segment_project_point_with_feature(sh.hh, m, pt, &seg_inside, &proj, &f);
        V3 dir;
        real dist;
        if (unit_try_new_and_get(pt - proj, EPS, &dir, &dist)) return dist <= sh.radius;
        return true;
    }
    return dot(sh.he, iso_inv_point(m, pt)) <= real(0);
}
void orc_shape_contains_point_batch(const orc_objects* objs, uint64_t n, const uint32_t* which, const real* pts, uint8_t* inside) {
    Objects o = make_objects(objs);
    for (uint64_t k = 0; k < n; ++k) inside[k] = shape_contains_point(o, which[k], v3(pts[3 * k], pts[3 * k + 1], pts[3 * k + 2])) ? 1 : 0;
}

// glue::interferences_with_ray (first_only = 0: every hit, rows sorted by (ray, handle)) / first_interference_with_ray
// (first_only = 1: smallest toi, ties -> smallest handle).  rays: 7 reals (origin, dir, max_toi); groups: the query's
// CollisionGroups (3 words) or NULL.  Rows: idx[2k] = (ray, handle), val[4k] = (toi, normal), feat[k].  Returns the row count.
uint64_t orc_sim_ray_cast(orc_sim* s, uint64_t n_rays, const real* rays, const uint32_t* groups, int first_only, uint32_t* idx, real* val,
                          uint32_t* feat, uint64_t cap) {
    const Objects& o = s->o;
    uint64_t rows = 0;
    std::vector<uint32_t> cand(o.n + 1);
    for (uint64_t r = 0; r < n_rays; ++r) {
        const real* q = rays + 7 * r;
        uint64_t nc = orc_bp_query(s->bp, 1, q, cand.data(), cand.size());
        std::sort(cand.begin(), cand.begin() + nc);
        V3 ro = v3(q[0], q[1], q[2]), rd = v3(q[3], q[4], q[5]);
        bool have = false;
        RayHit best;
        uint32_t best_h = 0;
        for (uint64_t k = 0; k < nc; ++k) {
            uint32_t h = cand[k];
            if (groups && o.groups) {
                uint32_t m1 = o.groups[3 * h], w1 = o.groups[3 * h + 1], b1 = o.groups[3 * h + 2];
                uint32_t m2 = groups[0], w2 = groups[1], b2 = groups[2];
                if (!((m1 & b2) == 0 && (m2 & b1) == 0 && (m1 & w2) != 0 && (m2 & w1) != 0)) continue;
            }
            RayHit hit = shape_ray_cast(o, h, ro, rd, q[6]);
            if (!hit.hit) continue;
            if (first_only) {
                if (!have || hit.toi < best.toi) best = hit, best_h = h, have = true;
            } else {
                if (rows < cap) {
                    idx[2 * rows] = (uint32_t)r, idx[2 * rows + 1] = h;
                    val[4 * rows] = hit.toi, val[4 * rows + 1] = hit.normal.x, val[4 * rows + 2] = hit.normal.y, val[4 * rows + 3] = hit.normal.z;
                    feat[rows] = hit.feature;
                }
                rows++;
            }
        }
        if (first_only && have) {
            if (rows < cap) {
                idx[2 * rows] = (uint32_t)r, idx[2 * rows + 1] = best_h;
                val[4 * rows] = best.toi, val[4 * rows + 1] = best.normal.x, val[4 * rows + 2] = best.normal.y, val[4 * rows + 3] = best.normal.z;
                feat[rows] = best.feature;
            }
            rows++;
        }
    }
    return rows;
}

// glue::interferences_with_point (kind 2: candidates whose stored box contains the point, collision-groups test, then
// shape.contains_point(position, point); glue/query.rs:79-131) and interferences_with_aabb (kind 0: stored box intersects,
// collision-groups test; :133-181).  q: 3 / 6 reals per query.  Rows (query, handle) sorted.  Returns the row count.
uint64_t orc_sim_query(orc_sim* s, int kind, uint64_t n, const real* q, const uint32_t* groups, uint32_t* idx, uint64_t cap) {
    const Objects& o = s->o;
    uint64_t rows = 0;
    std::vector<uint32_t> cand(o.n + 1);
    const int W = kind == 0 ? 6 : 3;
    for (uint64_t r = 0; r < n; ++r) {
        const real* qq = q + W * r;
        uint64_t nc = orc_bp_query(s->bp, kind, qq, cand.data(), cand.size());
        std::sort(cand.begin(), cand.begin() + nc);
        for (uint64_t k = 0; k < nc; ++k) {
            uint32_t h = cand[k];
            if (groups && o.groups) {
                uint32_t m1 = o.groups[3 * h], w1 = o.groups[3 * h + 1], b1 = o.groups[3 * h + 2];
                if (!((m1 & groups[2]) == 0 && (groups[0] & b1) == 0 && (m1 & groups[1]) != 0 && (groups[0] & w1) != 0)) continue;
            }
            if (kind == 2) {
                if (!shape_contains_point(o, h, v3(qq[0], qq[1], qq[2]))) continue;
            }
            if (rows < cap) idx[2 * rows] = (uint32_t)r, idx[2 * rows + 1] = h;
            rows++;
        }
    }
    return rows;
}

}  // extern "C"
