/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * CPU restatement of the ncollide3d hot path (one CollisionWorld::update step + TriMesh ray casting)
 * used to check the CUDA path and as bench.py's cpu_baseline.  The product never links this.
 * The Rust reference cannot be built in this image (no rustc/cargo; nalgebra not vendored), so this
 * is a restatement ("port"), pinned against the reference's own known-answer tests (tests/test_oracle_kat.py).
 * PARITY UNPINNED for shape_type 4 (Capsule: half_height, radius in shape_param; groundwork for SURVEY.md §8f N3): the reference has
 * no test that involves a capsule; that part is checked against closed-form geometry only (tests/test_oracle_capsule.py).
 */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H
#include <stdint.h>
#ifndef ORC_REAL
#define ORC_REAL float
#endif
typedef ORC_REAL real; /* f32 on the path; f64 build only for the reference's f64 KATs */
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_hull_library {
    uint32_t n_hulls;
    const uint32_t *vert_off, *face_off, *edge_off, *fadj_off, *vadj_off;
    const real* points;
    const uint32_t *vert_first_adj, *vert_num_adj;
    const uint32_t *face_first, *face_num;
    const real* face_normal;
    const uint32_t *vertices_adj_to_face, *edges_adj_to_face;
    const uint32_t *edge_vertices, *edge_faces;
    const real* edge_dir;
    const uint32_t *faces_adj_to_vertex, *edges_adj_to_vertex;
} orc_hull_library;

typedef struct orc_objects {
    uint32_t n;
    const real* pos;         /* 3 per object */
    const real* rot;         /* 4 per object, (i,j,k,w) */
    const uint32_t* shape_type; /* 0 ball, 1 cuboid, 2 convex hull, 3 plane, 4 capsule (oracle only so far) */
    const real* shape_param; /* 4 per object: r | half extents | hull id bits | plane normal */
    const uint32_t* groups;   /* 3 per object or NULL */
    const real* query_limit;
    const real* ang_pred;
    const orc_hull_library* hulls;
    const uint8_t* query_kind; /* per object: 0 = GeometricQueryType::Contacts, 1 = Proximity(query_limit); NULL = all Contacts */
} orc_objects;

typedef struct orc_contact {
    real world1[3], world2[3], normal[3], depth;
    uint32_t f1, f2;
} orc_contact;

/* ContactKinematic of a contact (query/contact/contact_kinematic.rs:57-66): tracked local points, NeighborhoodGeometry per side
 * (g: 0 Point, 1 Line(dir), 2 Plane(dir)), dilations (margin1 / margin2).  The feature ids are in orc_contact. */
typedef struct orc_kinematic {
    real local1[3], local2[3], dir1[3], dir2[3], dil1, dil2;
    uint32_t g1, g2;
} orc_kinematic;

void orc_compute_aabbs(const orc_objects* objs, real margin, int mode, real* out_minmax);
uint64_t orc_broad_phase(uint32_t n, const real* aabb_minmax, const uint32_t* groups, int mode, uint32_t* out_pairs,
                         uint64_t cap);

/* Narrow phase for the given (object1, object2) pairs, object1 being the first argument of
 * generate_contacts.  manifold_off has n_pairs+1 entries.  algo_out (optional) gets the dispatched
 * algorithm per pair.  Returns the number of contacts (may exceed cap). */
uint64_t orc_narrow_phase_kin(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, orc_contact* out, orc_kinematic* kin_out,
                              uint64_t cap, uint32_t* manifold_off, uint8_t* algo_out, uint32_t* stats);
uint64_t orc_narrow_phase(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, orc_contact* out, uint64_t cap,
                          uint32_t* manifold_off, uint8_t* algo_out, uint32_t* stats);

/* contact_support_map_support_map (query/contact/contact_support_map_support_map.rs:12-79: GJK, then EPA when penetrating) with a
 * fresh simplex for a batch of cuboid / hull pairs; predictions == NULL: query_limit sums.  out[10 p] = p1, p2, normal, found flag;
 * stats[4] = GJK iterations, EPA iterations, EPA calls, EPA failures. */
void orc_contact_sm_sm(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, const real* predictions, real* out, uint32_t* stats);

/* The reference's cylinder / cuboid known-answer test (tests/geometry/cylinder_cuboid_contact.rs): out[3] = distance, proximity
 * status, contact.is_some(), through the same GJK / EPA restatement with a Cylinder support map (shape/cylinder.rs:47-62). */
void orc_kat_cylinder_cuboid(real half_height, real radius, const real* t1, const real* he, const real* t2, real margin, real prediction,
                             real* out);

/* Proximity (query/proximity/proximity.rs:4-12) as a byte; ORC_PROX_NONE = the dispatcher has no detector (plane x plane). */
#define ORC_PROX_INTERSECTING 0
#define ORC_PROX_WITHIN_MARGIN 1
#define ORC_PROX_DISJOINT 2
#define ORC_PROX_NONE 255
#define ORC_ALGO_PROXIMITY 6
/* ProximityDetector::update with FRESH detectors (default_proximity_dispatcher.rs:19-47 and the four detectors) for a batch
 * of (object1, object2) pairs; margins == NULL: query_limit[o1] + query_limit[o2] (narrow_phase.rs:131-139). */
void orc_proximity(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, const real* margins, uint8_t* out);
/* Same with the detector state carried by the caller: axis_io[4 p] = sep_axis xyz + a "Some" flag, read and written back
 * (SupportMapSupportMapProximityDetector::sep_axis, support_map_support_map_proximity_detector.rs:53-57). */
void orc_proximity_warm(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, const real* margins, real* axis_io, uint8_t* out);
/* query::proximity(m1, g1, m2, g2, margin) (proximity_shape_shape.rs:8-33) for objects 0 and 1. */
int orc_query_proximity(const orc_objects* objs, real margin);
/* orc_narrow_phase for a world whose objects carry query kinds: pairs with a Proximity object go to the proximity detectors
 * (narrow_phase.rs:226-247): algo = ORC_ALGO_PROXIMITY, no contacts, prox_out[p] = status; other pairs: prox_out[p] = 255. */
uint64_t orc_narrow_phase_kinds(const orc_objects* objs, uint64_t n_pairs, const uint32_t* pairs, orc_contact* out, uint64_t cap,
                                uint32_t* manifold_off, uint8_t* algo_out, uint8_t* prox_out);

/* One reference-faithful fresh-world CollisionWorld::update (DBVT broad phase + narrow phase); returns
 * seconds spent in [aabb, broad, narrow] through times[3]; n_pairs / n_contacts through counts[2]. */
void orc_world_update_timed(const orc_objects* objs, real margin, double* times, uint64_t* counts);

/* One-shot query::contact(m1, g1, m2, g2, prediction) (query/contact/contact_shape_shape.rs:10-48) for
 * objects 0 and 1 of objs; returns 1 and fills out when a contact exists. */
int orc_query_contact(const orc_objects* objs, real prediction, orc_contact* out);

/* Multi-step DBVTBroadPhase (reference-faithful: two DBVTs, slab handles, pair hash map, purge, activation states).
 * Events are returned instead of calling a BroadPhaseInterferenceHandler: started = (re-inserted leaf, leaf it met),
 * stopped = SortedPair order. groups: 3 u32 per handle (indexed by handle) or NULL (only a != b is required). */
typedef struct orc_bp orc_bp;
orc_bp* orc_bp_create(real margin);
void orc_bp_destroy(orc_bp*);
uint32_t orc_bp_create_proxy(orc_bp*, const real* minmax);
int orc_bp_set_bounding_volume(orc_bp*, uint32_t handle, const real* minmax);
int orc_bp_remove(orc_bp*, uint32_t n, const uint32_t* handles, uint32_t* removed_pairs, uint64_t cap, uint64_t* n_removed);
void orc_bp_update(orc_bp*, const uint32_t* groups, uint32_t* started, uint64_t cap_s, uint64_t* n_started, uint32_t* stopped,
                   uint64_t cap_p, uint64_t* n_stopped);
void orc_bp_recompute_with(orc_bp*, uint32_t handle);
void orc_bp_recompute_all(orc_bp*);
uint64_t orc_bp_query(orc_bp*, int kind, const real* q, uint32_t* out, uint64_t cap);
uint64_t orc_bp_num_interferences(const orc_bp*);
int orc_bp_proxy(const orc_bp*, uint32_t handle, real* minmax);
uint64_t orc_bp_pairs(const orc_bp*, uint32_t* out, uint64_t cap);

/* Stepping world: CollisionWorld::update over several steps (persistent broad phase, interaction edges in the callback
 * orientation, GJK warm start, ContactManifold cache, contact events).  The shape arrays / hull library of objs must
 * outlive the sim; positions are copied.  Results are listed by edges sorted on (min handle, max handle). */
typedef struct orc_sim orc_sim;
orc_sim* orc_sim_create(const orc_objects* objs, real margin);
void orc_sim_destroy(orc_sim*);
void orc_sim_set_positions(orc_sim*, uint32_t n, const uint32_t* handles, const real* pos, const real* rot);
void orc_sim_step(orc_sim*);
int orc_sim_set_collision_groups(orc_sim*, uint32_t n, const uint32_t* handles, const uint32_t* groups);
int orc_sim_remove(orc_sim*, uint32_t n, const uint32_t* handles);
int orc_sim_add(orc_sim*, const orc_objects* objs, uint32_t* out_handles);
uint64_t orc_sim_num_pairs(const orc_sim*);
uint64_t orc_sim_num_contacts(const orc_sim*);
void orc_sim_fetch(const orc_sim*, uint32_t* pairs, uint8_t* algo, uint32_t* manifold_off, orc_contact* contacts, uint32_t* ids);
uint64_t orc_sim_events(const orc_sim*, uint32_t* out, uint64_t cap);
/* A bare set of contact edges driven by the caller (no broad phase): NarrowPhase::update_contact on the listed edges, generator
 * state (last_gjk_dir) and ContactManifold cache kept per edge — what a state slot of the device's stepping world holds. */
typedef struct orc_edges orc_edges;
orc_edges* orc_edges_create(uint64_t n, const uint32_t* pairs);
void orc_edges_destroy(orc_edges*);
uint64_t orc_edges_update(orc_edges*, const orc_objects* objs, uint64_t n_update, const uint32_t* which, uint32_t* events, uint64_t cap);
uint64_t orc_edges_fetch(const orc_edges*, uint32_t* manifold_off, orc_contact* contacts, uint32_t* ids, uint64_t cap, real* dirs);
/* Proximity status per edge (same order as orc_sim_fetch; 255 for contact edges) and the ProximityEvents of the last step as
 * (h1, h2, prev, new) rows in emission order. */
void orc_sim_fetch_proximity(const orc_sim*, uint8_t* prox);
uint64_t orc_sim_proximity_events(const orc_sim*, uint32_t* out, uint64_t cap);
uint64_t orc_sim_bp_num_interferences(const orc_sim*);
uint64_t orc_sim_query(orc_sim*, int kind, uint64_t n, const real* q, const uint32_t* groups, uint32_t* idx, uint64_t cap);
int orc_shape_ray_cast(const orc_objects* objs, uint32_t i, const real* origin, const real* dir, real max_toi, real* out, uint32_t* feature);
void orc_shape_ray_cast_batch(const orc_objects* objs, uint64_t n, const uint32_t* which, const real* rays, real* out, uint32_t* feat, uint8_t* hit);
void orc_shape_contains_point_batch(const orc_objects* objs, uint64_t n, const uint32_t* which, const real* pts, uint8_t* inside);
uint64_t orc_sim_ray_cast(orc_sim*, uint64_t n_rays, const real* rays, const uint32_t* groups, int first_only, uint32_t* idx, real* val,
                          uint32_t* feat, uint64_t cap);

/* TriMesh ray casting. */
typedef struct orc_trimesh orc_trimesh;
orc_trimesh* orc_trimesh_create(uint32_t n_verts, const real* xyz, uint32_t n_tris, const uint32_t* idx);
void orc_trimesh_destroy(orc_trimesh*);
/* mode 0: reference-faithful BVT best-first search; mode 1: brute force, global min toi over accepted hits,
 * ties -> smallest face index.  pose = t(3) q(4) or NULL for identity.  toi < 0 => miss. face = i or i+T (back face). */
void orc_trimesh_ray_cast_uv(const orc_trimesh*, const real* pose, uint64_t n_rays, const real* origins, const real* dirs, real max_toi,
                             const real* max_tois, const real* uvs, int mode, real* toi, uint32_t* face, real* normal, real* uv_out);
void orc_trimesh_ray_cast(const orc_trimesh*, const real* pose, uint64_t n_rays, const real* origins, const real* dirs,
                          real max_toi, int mode, real* toi, uint32_t* face, real* normal);
void orc_aabb_toi_with_ray(const real* minmax, const real* origin, const real* dir, real max_toi, int solid, real* toi);

/* ncollide2d RayCast::toi_and_normal_with_ray(m, ray, max_toi, solid = true) of shape k for ray k (oracle/dim2.cpp).  rays: 5 reals
 * (origin, dir, max_toi); out: 3 reals (toi, normal); feature: kind << 30 | id or 0xffffffff. */
void orc2_ray_cast(uint64_t n, const uint32_t* type, const real* param, const real* pose, const real* poly_points, const real* rays,
                   uint8_t* found, real* out, uint32_t* feature);
/* ncollide2d Polyline ray casting (oracle/ray.cpp).  idx: 2 point indices per edge (NULL: the line strip 0-1, 1-2, ...).  pose = x y re im
 * or NULL.  mode as above.  feature = edge, or edge + n_edges for FeatureId::Face(1) of the segment; normal = the segment's scaled normal. */
typedef struct orc2_polyline orc2_polyline;
orc2_polyline* orc2_polyline_create(uint32_t n_points, const real* xy, uint32_t n_edges, const uint32_t* idx);
void orc2_polyline_destroy(orc2_polyline*);
void orc2_polyline_ray_cast(const orc2_polyline*, const real* pose, uint64_t n_rays, const real* origins, const real* dirs, real max_toi,
                            const real* max_tois, int mode, real* toi, uint32_t* feature, real* normal);
/* Segment::toi_and_normal_with_ray (dim2); ab = a.x a.y b.x b.y; returns 1 for Some; feature = kind << 30 | id (1 Face, 2 Vertex). */
int orc2_segment_ray_cast(const real* ab, const real* pose, const real* origin, const real* dir, real* toi, real* normal, uint32_t* feature);

/* ncollide2d query::contact for n pairs of 2-D shapes (oracle/dim2.cpp).  type: 0 ball, 1 cuboid, 2 convex polygon; param: 4 reals per
 * shape (radius | hx, hy | first point, count); pose: 4 reals (translation x y, UnitComplex re im); found: 1 Some, 0 None, 2 not restated;
 * out: 7 reals per pair (world1, world2, normal, depth). */
void orc2_contact(uint64_t n, const uint32_t* type1, const real* param1, const real* pose1, const uint32_t* type2, const real* param2,
                  const real* pose2, const real* poly_points, const real* poly_normals, real prediction, uint8_t* found, real* out,
                  uint32_t* panics);

/* 2-D world (oracle/dim2.cpp): fat AABBs as 6 reals with z = 0 (usable with orc_broad_phase), and the contact manifolds of given pairs. */
struct orc2_objects;
void orc2_compute_aabbs(const struct orc2_objects* o, real margin, real* out);
uint64_t orc2_narrow_phase(const struct orc2_objects* o, uint64_t n_pairs, const uint32_t* pairs, uint32_t* manifold_off, real* contacts,
                           uint32_t* feats, uint64_t cap, uint32_t* panics, uint8_t* prox);

/* glue::interferences_with_ray / first_interference_with_ray over a 2-D world by brute force (boxes: 6 reals per object as written by
 * orc2_compute_aabbs); rows (ray, handle) sorted; returns the number of rows (may exceed cap). */
uint64_t orc2_world_ray_cast(const struct orc2_objects* o, const real* boxes, const uint32_t* obj_groups, uint64_t n_rays, const real* rays,
                             const uint32_t* groups, int first_only, uint32_t* idx, real* val, uint32_t* feat, uint64_t cap);

/* PointQuery::contains_point of shape k for point k; interferences_with_aabb (kind 0) / interferences_with_point (kind 2) by brute force. */
void orc2_contains_point(uint64_t n, const uint32_t* type, const real* param, const real* pose, const real* poly_points, const real* pts,
                         uint8_t* out);
uint64_t orc2_world_query(const struct orc2_objects* o, const real* boxes, const uint32_t* obj_groups, int kind, uint64_t n_queries,
                          const real* queries, const uint32_t* groups, uint32_t* idx, uint64_t cap);

/* ncollide2d query::proximity for n pairs, one margin per pair; out: 0 Intersecting, 1 WithinMargin, 2 Disjoint, 255 plane x plane. */
void orc2_proximity(uint64_t n, const uint32_t* type1, const real* param1, const real* pose1, const uint32_t* type2, const real* param2,
                    const real* pose2, const real* poly_points, const real* margins, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
