// ORACLE — TEST INFRASTRUCTURE ONLY (see na.hpp header).
// Scene description shared by the oracle translation units: SoA object arrays, the flat convex-hull
// table library (the tables ConvexHull::try_new builds, shape/convex.rs:109-335), feature ids and
// the contact record.  Layout mirrors include/ncb200.h but is declared independently.
#pragma once
#include <cstdint>
#include <vector>
#include "na.hpp"

namespace orc {

// CAPSULE (oracle only so far: groundwork for SURVEY §8f N3): shape_param = (half_height, radius), axis = local y (capsule.rs:10-17).
// SEGMENT never appears in a scene: it is the capsule's segment handed to the sub-detectors (capsule.rs:53-61).
enum ShapeType : uint32_t { BALL = 0, CUBOID = 1, HULL = 2, PLANE = 3, CAPSULE = 4, SEGMENT = 5 };

// FeatureId (shape/feature_id.rs): kind in the top 2 bits, id in the low 30.
enum : uint32_t { F_VERTEX = 0u, F_EDGE = 1u, F_FACE = 2u, F_UNKNOWN = 3u };
static inline uint32_t fid(uint32_t kind, uint32_t id) { return (kind << 30) | (id & 0x3fffffffu); }
static inline uint32_t fid_kind(uint32_t f) { return f >> 30; }
static inline uint32_t fid_id(uint32_t f) { return f & 0x3fffffffu; }
static const uint32_t FID_UNKNOWN = 0xc0000000u;

// Flat hull library.  All per-hull ids are LOCAL to the hull.
struct HullLibrary {
    uint32_t n_hulls;
    const uint32_t* vert_off;  // [n_hulls+1] into points / vert_*           (unit: vertices)
    const uint32_t* face_off;  // [n_hulls+1] into face_*                    (unit: faces)
    const uint32_t* edge_off;  // [n_hulls+1] into edge_*                    (unit: edges, deleted ones included)
    const uint32_t* fadj_off;  // [n_hulls+1] into vertices/edges_adj_to_face
    const uint32_t* vadj_off;  // [n_hulls+1] into faces/edges_adj_to_vertex
    const real* points;       // xyz per vertex
    const uint32_t* vert_first_adj;
    const uint32_t* vert_num_adj;
    const uint32_t* face_first;
    const uint32_t* face_num;
    const real* face_normal;  // xyz per face
    const uint32_t* vertices_adj_to_face;
    const uint32_t* edges_adj_to_face;
    const uint32_t* edge_vertices;  // 2 per edge
    const uint32_t* edge_faces;     // 2 per edge
    const real* edge_dir;          // xyz per edge
    const uint32_t* faces_adj_to_vertex;
    const uint32_t* edges_adj_to_vertex;
};

// One hull of the library (pointers pre-offset).
struct Hull {
    uint32_t nv, nf, ne;
    const real* points;
    const uint32_t *vert_first_adj, *vert_num_adj;
    const uint32_t *face_first, *face_num;
    const real* face_normal;
    const uint32_t *vaf, *eaf;
    const uint32_t *edge_vertices, *edge_faces;
    const real* edge_dir;
    const uint32_t *fav, *eav;
    V3 pt(uint32_t i) const { return {points[3 * i], points[3 * i + 1], points[3 * i + 2]}; }
    V3 fnormal(uint32_t i) const { return {face_normal[3 * i], face_normal[3 * i + 1], face_normal[3 * i + 2]}; }
    V3 edir(uint32_t i) const { return {edge_dir[3 * i], edge_dir[3 * i + 1], edge_dir[3 * i + 2]}; }
};

static inline Hull hull_view(const HullLibrary* L, uint32_t h) {
    Hull H;
    uint32_t v0 = L->vert_off[h], f0 = L->face_off[h], e0 = L->edge_off[h], fa0 = L->fadj_off[h], va0 = L->vadj_off[h];
    H.nv = L->vert_off[h + 1] - v0;
    H.nf = L->face_off[h + 1] - f0;
    H.ne = L->edge_off[h + 1] - e0;
    H.points = L->points + 3 * (size_t)v0;
    H.vert_first_adj = L->vert_first_adj + v0;
    H.vert_num_adj = L->vert_num_adj + v0;
    H.face_first = L->face_first + f0;
    H.face_num = L->face_num + f0;
    H.face_normal = L->face_normal + 3 * (size_t)f0;
    H.vaf = L->vertices_adj_to_face + fa0;
    H.eaf = L->edges_adj_to_face + fa0;
    H.edge_vertices = L->edge_vertices + 2 * (size_t)e0;
    H.edge_faces = L->edge_faces + 2 * (size_t)e0;
    H.edge_dir = L->edge_dir + 3 * (size_t)e0;
    H.fav = L->faces_adj_to_vertex + va0;
    H.eav = L->edges_adj_to_vertex + va0;
    return H;
}

struct Objects {
    uint32_t n;
    const real* pos;          // 3 per object
    const real* rot;          // 4 per object (i, j, k, w)
    const uint32_t* shape_type;
    const real* shape_param;  // 4 per object
    const uint32_t* groups;    // 3 per object: membership, whitelist, blacklist (may be null = defaults)
    const real* query_limit;  // GeometricQueryType::Contacts(linear, _)
    const real* ang_pred;     // GeometricQueryType::Contacts(_, angular)
    const HullLibrary* hulls;
    const uint8_t* query_kind = nullptr;  // 1 = GeometricQueryType::Proximity(query_limit); null = all Contacts
    bool is_proximity(uint32_t i) const { return query_kind && query_kind[i] != 0; }
    Iso iso(uint32_t i) const {
        return Iso{{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, {rot[4 * i], rot[4 * i + 1], rot[4 * i + 2], rot[4 * i + 3]}};
    }
    V3 p3(uint32_t i) const { return {shape_param[4 * i], shape_param[4 * i + 1], shape_param[4 * i + 2]}; }
    uint32_t hull_id(uint32_t i) const {
        return (uint32_t)shape_param[4 * i];  // hull id stored as a number (exact below 2^24)
    }
};

// query/contact/contact.rs:15-27 + the feature ids carried by ContactKinematic.
struct ContactOut {
    real world1[3], world2[3], normal[3], depth;
    uint32_t f1, f2;
};

// query/contact/contact_kinematic.rs:57-66: NeighborhoodGeometry (kind + direction / normal) and tracked local point per side,
// and the dilations (margin1 / margin2)
enum { G_POINT = 0, G_LINE = 1, G_PLANE = 2 };
struct Kin {
    V3 local1{0, 0, 0}, local2{0, 0, 0};
    V3 dir1{0, 0, 0}, dir2{0, 0, 0};
    real dil1 = 0, dil2 = 0;
    uint32_t g1 = G_POINT, g2 = G_POINT;
};
// what the C interface returns per contact (oracle.h: orc_kinematic)
struct KinOut {
    real local1[3], local2[3], dir1[3], dir2[3], dil1, dil2;
    uint32_t g1, g2;
};

struct Contact {
    V3 world1, world2, normal;
    real depth;
    Kin k;  // the ContactKinematic pushed with the contact (contact_manifold.rs:165-171)
};

}  // namespace orc
