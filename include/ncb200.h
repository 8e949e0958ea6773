/* ncb200 — C ABI of the B200-native collision hot path (one CollisionWorld::update step + TriMesh ray casting), widened to
 * the persistent BroadPhase (ncb_bp_*), the stepping CollisionWorld (ncb_sim_*), its world queries and
 * proximity-only interactions (GeometricQueryType::Proximity sensors, the 3-D half of N4).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ / torch types, no unwinding.
 * Every entry point names the reference (dimforge/ncollide, paths relative to the reference root) item it
 * replaces.  The reference itself has no FFI; INTEGRATION.md shows the Rust `extern "C"` block and the
 * BroadPhase / ContactManifoldGenerator / RayCast shims a maintainer would add on top of this header.
 *
 * Conventions
 *   - f32 everywhere on the path; ids are u32 on the wire (usize in the reference).
 *   - Return value: 0 = ok, < 0 = error (ncb_last_error gives the text), > 0 = an output capacity was too
 *     small: outputs were truncated, the n_* out-parameters hold the NEEDED counts; call again with larger buffers.
 *   - One context per GPU (one process per GPU); a context is not re-entrant (calls on it are serialised by the
 *     caller, as `&mut self` is in the reference).
 *   - There is no CPU fallback: every compute entry point fails with NCB_ERR_CUDA when no device is usable.
 */
#ifndef NCB200_H
#define NCB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NCB_OK 0
#define NCB_ERR_CUDA (-1)
#define NCB_ERR_ARG (-2)
#define NCB_ERR_STATE (-3)
#define NCB_ERR_UNSUPPORTED (-4)
#define NCB_ROUTE_REPEAT 2 /* ncb_world_update_routed stage 4: bucket capacities were raised, repeat from stage 2 */

/* shape_type values (shape/ball.rs, shape/cuboid.rs, shape/convex.rs, shape/plane.rs) */
#define NCB_SHAPE_BALL 0u
#define NCB_SHAPE_CUBOID 1u
#define NCB_SHAPE_CONVEX_HULL 2u
#define NCB_SHAPE_PLANE 3u
#define NCB_SHAPE_CAPSULE 4u /* shape/capsule.rs: shape_param = (half_height, radius, -, -), axis = local y */

/* FeatureId (shape/feature_id.rs): kind in bits 31..30 (0 vertex, 1 edge, 2 face, 3 unknown), id in bits 29..0. */
#define NCB_FEATURE_VERTEX 0u
#define NCB_FEATURE_EDGE 1u
#define NCB_FEATURE_FACE 2u
#define NCB_FEATURE_UNKNOWN 3u

/* Contact algorithm chosen per pair (DefaultContactDispatcher::get_contact_algorithm,
 * pipeline/narrow_phase/contact_generator/default_contact_dispatcher.rs:27-97). */
#define NCB_ALGO_NONE 0u
#define NCB_ALGO_BALL_BALL 1u
#define NCB_ALGO_PLANE_BALL 2u
#define NCB_ALGO_PLANE_CONVEX 3u
#define NCB_ALGO_BALL_CONVEX 4u
#define NCB_ALGO_CONVEX_CONVEX 5u
/* The pair involves a GeometricQueryType::Proximity object: a ProximityDetector ran instead of a contact generator
 * (DefaultProximityDispatcher::get_proximity_algorithm, proximity_detector/default_proximity_dispatcher.rs:19-47). */
#define NCB_ALGO_PROXIMITY 6u
#define NCB_ALGO_CAPSULE_CAPSULE 7u /* CapsuleCapsuleManifoldGenerator (capsule_capsule_manifold_generator.rs) */
#define NCB_ALGO_CAPSULE_SHAPE 8u   /* CapsuleShapeManifoldGenerator (capsule_shape_manifold_generator.rs) */

/* query::Proximity (query/proximity/proximity.rs:4-12) as a byte; NCB_PROXIMITY_NONE: not a proximity pair / no detector. */
#define NCB_PROXIMITY_INTERSECTING 0u
#define NCB_PROXIMITY_WITHIN_MARGIN 1u
#define NCB_PROXIMITY_DISJOINT 2u
#define NCB_PROXIMITY_NONE 255u

typedef struct ncb_ctx ncb_ctx;
typedef struct ncb_mesh ncb_mesh;

/* Tables of ConvexHull::try_new (shape/convex.rs:109-335) for a library of hulls, concatenated; ids are LOCAL to
 * each hull; *_off arrays have n_hulls + 1 entries.  Limits: <= 64 vertices per hull, <= 16 vertices per face. */
typedef struct ncb_hull_library {
    uint32_t n_hulls;
    const uint32_t* vert_off; /* -> points, vert_first_adj, vert_num_adj */
    const uint32_t* face_off; /* -> face_first, face_num, face_normal */
    const uint32_t* edge_off; /* -> edge_vertices, edge_faces, edge_dir */
    const uint32_t* fadj_off; /* -> vertices_adj_to_face, edges_adj_to_face */
    const uint32_t* vadj_off; /* -> faces_adj_to_vertex, edges_adj_to_vertex */
    const float* points;      /* xyz */
    const uint32_t* vert_first_adj;
    const uint32_t* vert_num_adj;
    const uint32_t* face_first;
    const uint32_t* face_num;
    const float* face_normal; /* xyz */
    const uint32_t* vertices_adj_to_face;
    const uint32_t* edges_adj_to_face;
    const uint32_t* edge_vertices; /* 2 per edge */
    const uint32_t* edge_faces;    /* 2 per edge */
    const float* edge_dir;         /* xyz */
    const uint32_t* faces_adj_to_vertex;
    const uint32_t* edges_adj_to_vertex;
} ncb_hull_library;

/* SoA collision objects = what CollisionObject holds on the path (pipeline/object/collision_object.rs:62-113):
 * position (Isometry3: translation + unit quaternion i,j,k,w), shape, CollisionGroups, GeometricQueryType (Contacts here;
 * ncb_set_query_types turns objects into Proximity sensors, query_limit then holds the proximity margin). */
typedef struct ncb_objects {
    uint32_t n;
    const float* pos;           /* 3 per object */
    const float* rot;           /* 4 per object */
    const uint32_t* shape_type; /* NCB_SHAPE_* */
    const float* shape_param;   /* 4 per object: radius | half extents | (float)hull id | plane normal */
    const uint32_t* groups;     /* 3 per object: membership, whitelist, blacklist; NULL = CollisionGroups::new() */
    const float* query_limit;   /* Contacts(linear, _) or Proximity(margin): GeometricQueryType::query_limit() */
    const float* ang_pred;      /* Contacts(_, angular) */
} ncb_objects;

/* query::Contact (query/contact/contact.rs:15-27) + the two feature ids of its ContactKinematic + owning pair. */
typedef struct ncb_contact {
    float world1[3];
    float world2[3];
    float normal[3];
    float depth;
    uint32_t f1, f2;
    uint32_t pair; /* index into the pair array of the same call */
} ncb_contact;

/* query::ContactKinematic (query/contact/contact_kinematic.rs:57-66) of a contact, besides the two feature ids that ncb_contact
 * carries: the tracked local points (local1 / local2, `approx.point`), the NeighborhoodGeometry of each side (geometry: 0 Point,
 * 1 Line(dir), 2 Plane(dir); dir in the object's local frame, zero for Point) and the dilations (margin1 / margin2).  Together with
 * ncb_contact this is what ContactManifold::push receives (contact_manifold.rs:165-171).  Produced on request (ncb_set_kinematics). */
typedef struct ncb_kinematic {
    float local1[3], local2[3];
    float dir1[3], dir2[3];
    float dilation1, dilation2;
    uint32_t geometry1, geometry2;
} ncb_kinematic;
#define NCB_GEOMETRY_POINT 0u
#define NCB_GEOMETRY_LINE 1u
#define NCB_GEOMETRY_PLANE 2u

typedef struct ncb_update_counts {
    uint32_t n_pairs;          /* broad-phase pairs (DBVTBroadPhase::num_interferences) */
    uint32_t n_contacts;       /* contacts over all manifolds */
    uint32_t n_contact_pairs;  /* pairs whose manifold is not empty */
    uint32_t n_algo[6];        /* pairs per NCB_ALGO_NONE .. NCB_ALGO_CONVEX_CONVEX (NCB_ALGO_PROXIMITY: n_proximity_pairs below) */
    uint32_t epa_overflow;     /* pairs that exceeded a fixed device capacity: EPA polytope (result = "no contact") or more than
                                * 32 distinct contacts in one manifold (extra contacts dropped); 0 expected, tests assert it */
    uint32_t ref_panics;       /* pairs on which the reference itself would have panicked (assert / unwrap) */
    uint32_t n_epa_pairs;      /* convex-convex pairs that needed EPA (GJK found the origin inside the CSO) */
    uint32_t n_manifold_jobs;  /* convex-convex pairs that reached feature clipping */
    uint32_t n_proximity_pairs; /* pairs handled by a proximity detector (NCB_ALGO_PROXIMITY) */
    uint32_t n_proximity[3];    /* of those: Intersecting, WithinMargin, Disjoint */
    uint32_t n_capsule_pairs[2]; /* pairs per NCB_ALGO_CAPSULE_CAPSULE, NCB_ALGO_CAPSULE_SHAPE */
    uint32_t stack_overflow;    /* BVH traversals that ran out of their fixed stack (a subtree was skipped): 0 expected, tests assert it */
    uint32_t n_epa_restarts;    /* EPA pairs beyond the shared-memory polytope capacities, restarted on the large store (not an error) */
} ncb_update_counts;

/* ---- context ------------------------------------------------------------------------------------------------ */
int ncb_create(int device, ncb_ctx** out);
void ncb_destroy(ncb_ctx* ctx);
const char* ncb_last_error(const ncb_ctx* ctx); /* NULL ctx: error of the last failed ncb_create on this thread */
/* Use the caller's CUDA stream (cudaStream_t as void*) instead of the context's own; NULL restores the context's own stream (to share
 * the legacy default stream pass cudaStreamLegacy, i.e. (void*)0x1).  Work already enqueued on the outgoing stream is waited for. */
int ncb_set_stream(ncb_ctx* ctx, void* cuda_stream);
void* ncb_get_stream(ncb_ctx* ctx);
int ncb_synchronize(ncb_ctx* ctx);

/* BVH walks of the query entry points (TriMesh ray casts, world / broad-phase queries) use a fixed 64-entry stack; a walk that runs
 * out of it skips a subtree and is counted here (cumulative over the context's life, plus the pair search of the last update /
 * ncb_broad_phase, which an update also reports in ncb_update_counts.stack_overflow).  The LBVH's depth bound (30 Morton bits + 32 tie-break bits) keeps both at 0; tests assert it. */
int ncb_traversal_overflows(ncb_ctx* ctx, uint32_t* out);

/* on != 0: fresh-world updates (ncb_world_update*, ncb_generate_contacts) also produce the ContactKinematic of every contact, in an
 * array aligned with the contacts (ncb_world_fetch_kinematics).  Off by default (64 more bytes per contact).  The stepping world
 * (ncb_sim_*) does not produce kinematics: the reference keeps the kinematic of a cached contact across updates, which would have to
 * be stored with every entry of the persistent manifolds. */
int ncb_set_kinematics(ncb_ctx* ctx, int on);
/* The kinematics of the last update's / ncb_generate_contacts' contacts, same order as the contacts.  NCB_ERR_STATE if they were not
 * requested before that call. */
int ncb_world_fetch_kinematics(ncb_ctx* ctx, ncb_kinematic* out, uint32_t cap_contacts);

/* ---- shapes and objects -------------------------------------------------------------------------------------- */
/* ConvexHull tables (host pointers), copied to the device once. */
int ncb_set_hulls(ncb_ctx* ctx, const ncb_hull_library* lib);
/* CollisionWorld::add for a whole world (pipeline/world.rs:66-96): host SoA -> device.  NCB_ERR_UNSUPPORTED when a shape_type is not
 * one of the NCB_SHAPE_* values. */
int ncb_set_objects(ncb_ctx* ctx, const ncb_objects* objs);
/* CollisionObject::set_position for all objects (pipeline/object/collision_object.rs:186-190). */
int ncb_set_positions(ncb_ctx* ctx, uint32_t n, const float* pos, const float* rot);
/* Same for the objects [begin, begin + count) only (multi-GPU: every rank uploads its own block; the blocks are then
 * all-gathered on the device buffers returned by ncb_device_ptr(ctx, 4 / 5)). pos / rot point at the block. */
int ncb_set_positions_range(ncb_ctx* ctx, uint32_t begin, uint32_t count, const float* pos, const float* rot);

/* GeometricQueryType per object (pipeline/object/query_type.rs:8-37): kinds[i] = 0 Contacts(query_limit, ang_pred) or
 * 1 Proximity(query_limit) — a sensor.  A pair with at least one sensor gets a Proximity status instead of a contact manifold
 * (NarrowPhase::handle_interaction, narrow_phase.rs:226-247).  kinds == NULL (or all 0): every object is Contacts, the state
 * after EVERY ncb_set_objects (ncb_world_update, which re-uploads the world, keeps the kinds when the object count is unchanged).
 * n must equal the object count.  NCB_ERR_UNSUPPORTED for a world with capsules (their proximity detectors are not on the device). */
int ncb_set_query_types(ncb_ctx* ctx, uint32_t n, const uint8_t* kinds);

/* ---- stage entry points (each mirrors one reference routine, host buffers in/out) ---------------------------- */
/* mode 0: bounding_volume::aabb(shape, position) (shape/shape.rs aabb, bounding_volume/aabb_*.rs);
 * mode 1: CollisionObjectRef::compute_aabb = mode 0 loosened by query_limit (collision_object.rs:89-93);
 * mode 2: additionally DBVTBroadPhase's loosened(margin) (dbvt_broad_phase.rs:341) = the box the broad phase stores.
 * out: 6 floats per object (mins, maxs). */
int ncb_compute_aabbs(ncb_ctx* ctx, float margin, int mode, float* out_minmax);
/* BroadPhase::update on a fresh proxy set (pipeline/broad_phase/broad_phase.rs:65, dbvt_broad_phase.rs:174-259):
 * proxies 0..n-1 with the given (already loosened) AABBs; out pairs = (larger handle, smaller handle), i.e. the
 * argument order of interference_started.  groups may be NULL. */
int ncb_broad_phase(ncb_ctx* ctx, uint32_t n, const float* aabb_minmax, const uint32_t* groups, uint32_t* out_pairs,
                    uint32_t cap_pairs, uint32_t* n_pairs);
/* ContactManifoldGenerator::generate_contacts for a batch of pairs over the objects set by ncb_set_objects
 * (contact_generator/contact_manifold_generator.rs:10-36); pairs[2p] is the first shape.  manifold_start/count
 * (optional, n_pairs entries) locate each pair's contacts inside out_contacts. */
int ncb_generate_contacts(ncb_ctx* ctx, uint32_t n_pairs, const uint32_t* pairs, ncb_contact* out_contacts,
                          uint32_t cap_contacts, uint32_t* n_contacts, uint32_t* manifold_start, uint8_t* manifold_count,
                          uint8_t* algo);

/* ProximityDetector::update with fresh detectors for a batch of (object1, object2) pairs over the objects set by
 * ncb_set_objects (proximity_detector/proximity_detector.rs:10-30; ball x ball, plane x support map, support map x support map
 * through GJK with exact_dist = false; balls are support maps here).  margins[p] or, when NULL, query_limit[o1] + query_limit[o2]
 * (narrow_phase.rs:138).  out[p] = NCB_PROXIMITY_* (NONE for plane x plane).  With one explicit margin per pair this is
 * query::proximity(m1, g1, m2, g2, margin) (query/proximity/proximity_shape_shape.rs:8-33). */
int ncb_proximity(ncb_ctx* ctx, uint32_t n_pairs, const uint32_t* pairs, const float* margins, uint8_t* out);

/* ---- fused hot path ------------------------------------------------------------------------------------------ */
/* One fresh-world CollisionWorld::update (pipeline/world.rs:104-119 -> glue/update.rs:117-135) over the objects
 * currently on the device; results stay on the device.  q_begin/q_end restrict the broad-phase QUERY leaves to a
 * slice of the Morton order (multi-GPU sharding, pass 0 / UINT32_MAX for everything). */
int ncb_world_update_device(ncb_ctx* ctx, float margin, uint32_t q_begin, uint32_t q_end, ncb_update_counts* counts);
/* Copy the results of the last update to host buffers (any pointer may be NULL). */
int ncb_world_fetch(ncb_ctx* ctx, uint32_t* pairs, uint32_t cap_pairs, uint8_t* pair_algo, uint32_t* manifold_start,
                    uint8_t* manifold_count, ncb_contact* contacts, uint32_t cap_contacts);
/* Proximity status of every pair of the last update, in the order of ncb_world_fetch's pairs: NCB_PROXIMITY_* for
 * NCB_ALGO_PROXIMITY pairs, NCB_PROXIMITY_NONE for the others.  On a fresh world every status other than Disjoint is what
 * the reference reports as ProximityEvent(o1, o2, Disjoint, status) (narrow_phase.rs:108-121).  All NONE without sensors. */
int ncb_world_fetch_proximity(ncb_ctx* ctx, uint8_t* prox, uint32_t cap_pairs);
/* Host-buffer convenience = ncb_set_objects + ncb_world_update_device + ncb_world_fetch (the end-to-end call). */
int ncb_world_update(ncb_ctx* ctx, const ncb_objects* objs, float margin, uint32_t* pairs, uint32_t cap_pairs,
                     uint8_t* pair_algo, uint32_t* manifold_start, uint8_t* manifold_count, ncb_contact* contacts,
                     uint32_t cap_contacts, ncb_update_counts* counts);
/* The end-to-end call of a world whose objects persist (pipeline/world.rs:64-119: `add` once, then per step
 * `CollisionObject::set_position` on the objects + `CollisionWorld::update`): uploads only the n poses (28 B per object, host
 * pointers, pinned memory recommended), updates, and returns the results like ncb_world_update.  Shapes, groups and query limits are
 * the ones of the last ncb_set_objects.  = ncb_set_positions + ncb_world_fetch_early + ncb_world_update_device + ncb_world_fetch. */
int ncb_world_update_poses(ncb_ctx* ctx, uint32_t n, const float* pos, const float* rot, float margin, uint32_t* pairs, uint32_t cap_pairs,
                           uint8_t* pair_algo, uint32_t* manifold_start, uint8_t* manifold_count, ncb_contact* contacts,
                           uint32_t cap_contacts, ncb_update_counts* counts);

/* Device pointers of internal buffers, for multi-GPU plumbing (NCCL all-gather of AABBs) and zero-copy consumers.
 * which: 0 aabb_lo (float4 per object: mins, w unused) 1 aabb_hi, 2 pairs (uint2), 3 contacts (ncb_contact),
 * 4 pos (3 floats), 5 rot (4 floats). */
void* ncb_device_ptr(ncb_ctx* ctx, int which);
/* Split update for multi-GPU: stage 0 = AABBs of objects [obj_begin, obj_end) only; stage 1 = everything after the
 * AABBs (LBVH + pair search for query slice + narrow phase). */
int ncb_world_update_stage(ncb_ctx* ctx, int stage, float margin, uint32_t begin, uint32_t end, ncb_update_counts* counts);
/* Multi-GPU stage 1 with spatial ownership: after stage 0 + the all-gather of the AABB arrays (ncb_device_ptr 0 / 1), rank
 * `rank` of `world` selects the objects it owns (equal-count Morton ranges) plus the ghosts around them, builds its LBVH
 * over those only and reports its share of the pairs + their contacts.  Every pair is reported by exactly one rank. */
int ncb_world_update_sharded(ncb_ctx* ctx, float margin, int rank, int world, ncb_update_counts* counts);
/* Multi-GPU update with ROUTED spatial ownership (the default of ncollide_b200/parallel.py): no all-gather of all boxes; every rank
 * sends each object of its own block [begin, end) to the rank that owns its Morton bin and, as a ghost, to the ranks whose region its
 * fat box meets (pipeline/broad_phase/dbvt_broad_phase.rs:174-259 sees the same pair set: every intersecting pair has both boxes on
 * the rank that reports it).  The caller runs one collective on the device buffers of ncb_route_buffer between the stages:
 *   stage 0: AABBs + centre bounds of the own block          -> all-reduce MAX of buffer 0
 *   stage 1: Morton bins + histogram                          -> all-reduce SUM of buffer 1
 *   stage 2: owner buckets + per-owner regions                -> all-to-all buffer 2 -> 3 (world equal parts), all-reduce MAX of buffer 4
 *   stage 3: ghost buckets                                    -> all-to-all buffer 5 -> 6
 *   stage 4: unpack, local LBVH, pair search, narrow phase; fills counts.  NCB_ROUTE_REPEAT: a bucket was too small on some rank;
 *            every rank has raised its capacities identically, repeat from stage 2.
 * with_poses != 0: records also carry the poses (a rank needs only the poses of its own block before stage 0). */
int ncb_world_update_routed(ncb_ctx* ctx, int stage, float margin, int rank, int world, uint32_t begin, uint32_t end, int with_poses,
                            ncb_update_counts* counts);
/* Device buffers of the routed update, for the caller's collectives.  which: 0 bounds (6 f32), 1 histogram (1024 i32), 2 / 3 owner
 * buckets send / recv, 4 regions (8 sub-boxes x 6 f32 per rank), 5 / 6 ghost buckets send / recv; *bytes = the extent the collective covers. */
void* ncb_route_buffer(ncb_ctx* ctx, int which, int world, uint64_t* bytes);
/* Peer-memory exchange for the routed update (NVLink P2P instead of NCCL inside the step).  ncb_route_p2p_alloc allocates this rank's
 * receive buffers (buckets of largest-block + 1 records: they cannot overflow) and exports them: `handles` receives 3 cudaIpcMemHandle_t
 * (3 x 64 bytes) for peers in other processes, `ptrs` 3 raw device pointers for peers inside this process (either may be NULL).  The
 * caller gathers every rank's handles (any transport; done once) and calls ncb_route_p2p_connect with world x 3 handles in rank order
 * (or handles_all NULL and world x 3 raw pointers when all ranks live in one process).  From then on the routing kernels store records
 * straight into the owner's bucket on the peer GPU, the small arrays (bounds, histogram, regions) go to per-sender slots, a
 * system-scope flag per sender closes each round, and ncb_world_update_routed(stage = -1) runs a whole step in one call with no
 * collective and no host work between the stages (stages 0..4 remain callable one by one: every rank must have been given stage s
 * before any rank is given stage s + 1 when the ranks share a stream).  A peer that does not arrive within ~2 s makes the step
 * return NCB_ERR_STATE instead of hanging. */
int ncb_route_p2p_alloc(ncb_ctx* ctx, int rank, int world, uint32_t n_total, void* handles, uint64_t* ptrs);
int ncb_route_p2p_connect(ncb_ctx* ctx, const void* handles_all, const uint64_t* ptrs_all);
int ncb_route_p2p_close(ncb_ctx* ctx);
/* Arms an overlapped fetch for the NEXT device update: while its narrow phase runs, the sorted pairs (+ algorithm) and the
 * contacts that are already final are copied into these host buffers (pinned memory recommended); ncb_world_fetch with the
 * same buffers then only copies the rest.  ncb_world_update does this by itself. */
int ncb_world_fetch_early(ncb_ctx* ctx, uint32_t* pairs, uint32_t cap_pairs, uint8_t* pair_algo, ncb_contact* contacts, uint32_t cap_contacts);

/* Per-stage device times of the last ncb_world_update_device, measured with CUDA events on the context's stream
 * when enabled.  names/ms are arrays of at least 16 entries; returns the number of stages filled. */
int ncb_profile_enable(ncb_ctx* ctx, int on);
int ncb_profile_get(ncb_ctx* ctx, const char** names, float* ms, uint32_t* launches);

/* ---- RayCast for TriMesh ------------------------------------------------------------------------------------- */
/* TriMesh::new (shape/trimesh.rs:100-197): uploads the mesh and builds the device BVH. */
int ncb_trimesh_create(ncb_ctx* ctx, uint32_t n_verts, const float* xyz, uint32_t n_tris, const uint32_t* idx, ncb_mesh** out);
void ncb_trimesh_destroy(ncb_mesh* mesh);
/* RayCast::toi_and_normal_with_ray for a batch (query/ray/ray_trimesh.rs:22-50): pose_tq = translation(3) +
 * quaternion ijkw(4) or NULL (identity); origins/dirs 3 floats per ray; toi < 0 = None; face = i, or i + n_tris
 * for a back-face hit (ray_trimesh.rs:41-45); normal (optional) in world space. */
int ncb_trimesh_ray_cast(ncb_mesh* mesh, const float* pose_tq, uint32_t n_rays, const float* origins, const float* dirs,
                         float max_toi, float* toi, uint32_t* face, float* normal);
/* TriMesh::uvs (shape/trimesh.rs, `uvs: Option<Vec<Point2<N>>>`): 2 floats per vertex, NULL clears them. */
int ncb_trimesh_set_uvs(ncb_mesh* mesh, const float* uvs);
/* RayCast::toi_and_normal_and_uv_with_ray for a batch (query/ray/ray_trimesh.rs:52-94; barycentric coordinates of
 * query/ray/ray_triangle.rs:92-114): like ncb_trimesh_ray_cast, plus uv (2 floats per ray, optional; zeros for a miss or when the mesh has
 * no uvs, where the reference falls back to toi_and_normal_with_ray) and max_tois (one limit per ray, optional; NULL = max_toi for all:
 * `Ray` casts of query/ray/ray.rs:125-160 each carry their own max_toi).  Host buffers; the batch is pipelined in chunks (upload | cast |
 * download on three streams), so pinned buffers make the call cost about max(copies, kernel). */
int ncb_trimesh_ray_cast_uv(ncb_mesh* mesh, const float* pose_tq, uint32_t n_rays, const float* origins, const float* dirs, float max_toi,
                            const float* max_tois, float* toi, uint32_t* face, float* normal, float* uv);
/* Same with device-resident rays / results (pointers are device pointers); asynchronous on the context's stream. */
int ncb_trimesh_ray_cast_device(ncb_mesh* mesh, const float* pose_tq_host, uint32_t n_rays, const float* d_origins,
                                const float* d_dirs, float max_toi, float* d_toi, uint32_t* d_face, float* d_normal);

/* ---- Persistent broad phase: the BroadPhase trait surface over several updates -------------- */
/* Replaces DBVTBroadPhase (pipeline/broad_phase/dbvt_broad_phase.rs) behind pipeline/broad_phase/broad_phase.rs:68-99.
 * Handles are slab keys handed out LIFO like the reference's (`slab` crate).  Boxes are 6 floats (mins, maxs).
 * After each ncb_bp_update the interference set equals the reference's: { (i, j) : stored boxes intersect (inclusive),
 * groups allow }, where a stored box changes only when a new box is not contained in it (then: new.loosened(margin)).
 * When the groups of a live handle change, call ncb_bp_recompute_with for it (as glue/update.rs:85 does). */
typedef struct ncb_bp ncb_bp;
/* DBVTBroadPhase::new(margin) (dbvt_broad_phase.rs:75-88) */
int ncb_bp_create(ncb_ctx* ctx, float margin, ncb_bp** out);
void ncb_bp_destroy(ncb_bp* bp);
/* BroadPhase::create_proxy (:275-280) for n boxes, in order; out_handles[n]. */
int ncb_bp_create_proxies(ncb_bp* bp, uint32_t n, const float* aabb_minmax, uint32_t* out_handles);
/* BroadPhase::deferred_set_bounding_volume (:325-347) for n (handle, box) entries, in order.  NCB_ERR_ARG when a
 * handle does not exist (the reference panics). */
int ncb_bp_set_bounding_volumes(ncb_bp* bp, uint32_t n, const uint32_t* handles, const float* aabb_minmax);
/* BroadPhase::deferred_recompute_all_proximities_with (:349-363) / deferred_recompute_all_proximities (:365-386):
 * the way the reference's world tells the broad phase that a proxy's groups (or the pair filter) changed. */
int ncb_bp_recompute_with(ncb_bp* bp, uint32_t n, const uint32_t* handles);
int ncb_bp_recompute_all(ncb_bp* bp);
/* BroadPhase::remove (:282-323).  The dropped interferences are readable as the "stopped" list of ncb_bp_events. */
int ncb_bp_remove(ncb_bp* bp, uint32_t n, const uint32_t* handles, uint32_t* n_removed);
/* BroadPhase::update (:174-259).  groups = 3 words per handle slot (membership, whitelist, blacklist) or NULL. */
int ncb_bp_update(ncb_bp* bp, const uint32_t* groups, uint32_t n_group_slots, uint32_t* n_started, uint32_t* n_stopped);
/* Events of the last ncb_bp_update / ncb_bp_remove, sorted: started[2 * n_started] = the arguments of
 * BroadPhaseInterferenceHandler::interference_started (re-inserted proxy first; when both moved, the one updated
 * later first), stopped[2 * n_stopped] = those of interference_stopped (smaller handle first). */
int ncb_bp_events(ncb_bp* bp, uint32_t* started, uint32_t* stopped);
/* BroadPhase::num_interferences (:349-351) */
int ncb_bp_num_interferences(ncb_bp* bp, uint32_t* n);
/* The current interference set, sorted, (smaller, larger) handle per pair; cap in pairs; returns 1 when truncated. */
int ncb_bp_pairs(ncb_bp* bp, uint32_t* pairs, uint32_t cap, uint32_t* n);
/* Batched BroadPhase::interferences_with_bounding_volume (kind 0; 6 floats per query: mins, maxs), interferences_with_ray
 * (kind 1; 7 floats: origin, dir, max_toi) and interferences_with_point (kind 2; 3 floats) (:388-432), against the boxes
 * stored by the last ncb_bp_update, removed proxies excluded.  out[2 * k] = (query index, handle), sorted; cap in
 * entries; *n_out = entries found; returns 1 when truncated.  Reuses the event buffer: read ncb_bp_events first. */
int ncb_bp_query(ncb_bp* bp, int kind, uint32_t n_queries, const float* queries, uint32_t* out, uint32_t cap, uint32_t* n_out);
/* BroadPhase::proxy (:262-273): returns 1 and the stored (loosened) box when the proxy is attached, else 0. */
int ncb_bp_proxy(ncb_bp* bp, uint32_t handle, float* minmax);

/* ---- Stepping world: CollisionWorld::update over several steps ------------------------------- */
/* Replaces world.rs:104-119 + glue/update.rs:65-138 + narrow_phase.rs:56-104,168-278 for a fixed object set (the one
 * given to ncb_set_objects; object handle = index): persistent broad phase, interaction pairs created / removed by its
 * started / stopped callbacks (object order = callback argument order), only pairs with a moved object regenerated,
 * GJK warm start (last_gjk_dir), ContactManifold cache with stable contact ids, ContactEvents. */
typedef struct ncb_sim ncb_sim;
int ncb_sim_create(ncb_ctx* ctx, float margin, ncb_sim** out);
void ncb_sim_destroy(ncb_sim* sim);
/* CollisionObject::set_position (collision_object.rs:215-222) for m objects; handles == NULL means objects 0..m-1. */
int ncb_sim_set_positions(ncb_sim* sim, uint32_t m, const uint32_t* handles, const float* pos, const float* rot);
/* CollisionWorld::remove (world.rs:129-144) / CollisionWorld::add (world.rs:64-96) between updates (after the first
 * ncb_sim_step).  Removed objects and their pairs disappear without events (glue/setup.rs:50-62); handles are recycled
 * last-freed-first like the reference's slabs, so ncb_sim_add returns the handles the reference would hand out.  New objects
 * refer to the hull library already set with ncb_set_hulls. */
int ncb_sim_remove(ncb_sim* sim, uint32_t m, const uint32_t* handles);
int ncb_sim_add(ncb_sim* sim, const ncb_objects* objs, uint32_t* out_handles);
/* ncb_sim_add with a GeometricQueryType per new object: kinds[k] = 0 Contacts / 1 Proximity(query_limit) (NULL = all Contacts). */
int ncb_sim_add_with_query_types(ncb_sim* sim, const ncb_objects* objs, const uint8_t* kinds, uint32_t* out_handles);
/* CollisionObject::set_collision_groups (pipeline/object/collision_object.rs:246-250) for a batch of live objects; groups = 3 words
 * per handle (membership, whitelist, blacklist; collision_groups.rs:26-37).  Sets COLLISION_GROUPS_CHANGED: the next ncb_sim_step
 * redispatches the objects in the broad phase (glue/update.rs:83-86: pairs that are no longer allowed stop — with a
 * ContactEvent::Stopped / ProximityEvent if they were touching —, newly allowed ones start) and updates their pairs in the narrow
 * phase (collision_object.rs:33-43). */
int ncb_sim_set_collision_groups(ncb_sim* sim, uint32_t n, const uint32_t* handles, const uint32_t* groups);
/* CollisionWorld::update.  counts: n_pairs, n_contacts, epa_overflow (+ manifold-cache overflows), ref_panics,
 * n_epa_pairs, n_manifold_jobs (= pairs regenerated in this step). */
int ncb_sim_step(ncb_sim* sim, ncb_update_counts* counts);
int ncb_sim_sizes(ncb_sim* sim, uint32_t* n_pairs, uint32_t* n_contacts, uint32_t* n_events);
/* Results of the last step, arrays sized from ncb_sim_sizes (NULL = not wanted): pairs[2 P] sorted by (min, max) handle in
 * the pair's own object order; algo / manifold_start / manifold_count [P]; contacts[C]; contact_ids[C] (stable per pair as
 * long as the reference's ContactId is); events[3 E] = (object1, object2, 1 Started | 0 Stopped), sorted. */
int ncb_sim_fetch(ncb_sim* sim, uint32_t* pairs, uint8_t* algo, uint32_t* manifold_start, uint8_t* manifold_count, ncb_contact* contacts,
                  uint32_t* contact_ids, uint32_t* events);

/* Proximity interactions of a stepping world: objects marked with ncb_set_query_types before ncb_sim_create
 * (or added with ncb_sim_add_with_query_types) are sensors; their pairs carry Interaction::Proximity(detector, status) — status
 * Disjoint on a new pair, the support-map detector's separating axis kept between updates — and emit ProximityEvents
 * (narrow_phase.rs:108-143,226-247,266-274).  prox[P] = status per pair in ncb_sim_fetch's order (NCB_PROXIMITY_NONE for contact
 * pairs); events[4 E] = (collider1, collider2, prev_status, new_status), sorted; cap_events in rows; *n_events = rows that exist;
 * returns 1 when the events were truncated.  Any pointer may be NULL. */
int ncb_sim_fetch_proximity(ncb_sim* sim, uint8_t* prox, uint32_t* events, uint32_t cap_events, uint32_t* n_events);

/* World ray queries: glue::interferences_with_ray (first_only = 0) / first_interference_with_ray
 * (first_only = 1) (pipeline/glue/query.rs:13-77,183-224) against the state of the last ncb_sim_step.  rays[7 n] = origin,
 * dir, max_toi; groups = the query's CollisionGroups (membership, whitelist, blacklist) or NULL.  Candidates come from the
 * broad phase's stored boxes, each is tested with its shape's RayCast::toi_and_normal_with_ray(position, ray, max_toi,
 * solid = true) (ball, cuboid, plane, convex hull through the GJK ray cast).  Rows sorted by (ray, handle): idx[2 k] =
 * (ray, handle), val[4 k] = (toi, normal), feat[k] = feature id; first_only keeps the smallest toi per ray (ties: smallest
 * handle).  cap in rows; *n_out = rows found; returns 1 when truncated. */
int ncb_sim_ray_cast(ncb_sim* sim, uint32_t n_rays, const float* rays, const uint32_t* groups, int first_only, uint32_t* idx, float* val,
                     uint32_t* feat, uint32_t cap, uint32_t* n_out);

/* glue::interferences_with_aabb (kind 0; 6 floats per query: mins, maxs) / interferences_with_point (kind 2; 3 floats)
 * (pipeline/glue/query.rs:79-181): candidates from the stored boxes, the query's collision groups, and for points the shape's
 * PointQuery::contains_point (ball, cuboid, plane, convex hull through gjk::project_origin).  idx[2 k] = (query, handle),
 * sorted; cap in rows; returns 1 when truncated. */
int ncb_sim_query(ncb_sim* sim, int kind, uint32_t n_queries, const float* queries, const uint32_t* groups, uint32_t* idx, uint32_t cap,
                  uint32_t* n_out);

const char* ncb_version(void);

/* ---- ncollide2d, first slice: batched query::contact between 2-D shapes ----------------------------------- */
/* ncollide2d::query::contact(m1, g1, m2, g2, prediction) (query/contact/contact_shape_shape.rs:15-60) for n_pairs pairs of 2-D balls,
 * cuboids and convex polygons: contact_ball_ball, contact_ball_convex_polyhedron (ball x cuboid / polygon, either order; the polygon
 * case projects the ball centre with GJK / EPA, point_support_map.rs:14-55) and contact_support_map_support_map = 2-D GJK
 * (gjk.rs:76-177 over voronoi_simplex2.rs) + 2-D EPA (epa2.rs).
 * Host buffers.  type: NCB2D_BALL / _CUBOID / _POLYGON; param: 4 floats per shape (radius | hx, hy | first point, point count into
 * poly_points; poly_normals = ConvexPolygon::normals aligned with poly_points, convex_polygon.rs:34-66, needed when a ball meets a
 * polygon); pose: 4 floats per shape = Isometry2 (translation x, y; UnitComplex re, im = cos, sin of the angle, evaluated by the
 * caller like Isometry2::new does).  found[p] = 1 for Some(contact), 0 for None; out: 7 floats per pair (world1 xy, world2 xy, normal
 * xy, depth).  ref_panics counts pairs on which the reference itself would panic (epa2.rs:279, peek on an empty heap), epa_overflow
 * pairs whose polytope outgrew the device capacity (72 vertices; they answer None). */
#define NCB2D_BALL 0u
#define NCB2D_CUBOID 1u
#define NCB2D_POLYGON 2u
#define NCB2D_PLANE 3u /* shape/plane.rs in 2-D (a half-space): shape_param = the unit normal (nx, ny) */
#define NCB2D_SEGMENT 4u /* shape/segment.rs: shape_param = (a.x, a.y, b.x, b.y), a != b; a ConvexPolyhedron with two vertices and two faces */
int ncb2d_contact(ncb_ctx* ctx, uint32_t n_pairs, const uint32_t* type1, const float* param1, const float* pose1, const uint32_t* type2,
                  const float* param2, const float* pose2, const float* poly_points, const float* poly_normals, uint32_t n_poly_points,
                  float prediction, uint8_t* found, float* out, uint32_t* ref_panics, uint32_t* epa_overflow);

/* ncollide2d::query::proximity(m1, g1, m2, g2, margin) (query/proximity/proximity_shape_shape.rs:8-33) for a batch, one margin per
 * pair: proximity_ball_ball, proximity_plane_support_map (either order), proximity_support_map_support_map (2-D GJK with
 * exact_dist = false; balls are support maps here).  out: NCB_PROXIMITY_INTERSECTING / _WITHIN_MARGIN / _DISJOINT per pair. */
int ncb2d_proximity(ncb_ctx* ctx, uint32_t n_pairs, const uint32_t* type1, const float* param1, const float* pose1, const uint32_t* type2,
                    const float* param2, const float* pose2, const float* poly_points, uint32_t n_poly_points, const float* margins,
                    uint8_t* out);
/* ncollide2d RayCast::toi_and_normal_with_ray(m, ray, max_toi, solid = true) of shape k for ray k, for a batch (query/ray/ray_ball.rs,
 * ray_cuboid.rs + ray_aabb.rs, ray_plane.rs, ray_support_map.rs:165-189 = ConvexPolygon through gjk::cast_ray, query/algorithms/gjk.rs:
 * 180-365).  type / param / pose as in ncb2d_contact; rays: 5 floats (origin x y, dir x y, max_toi).  found: 1 Some / 0 None; out: 3
 * floats (toi, normal x y); feature: 1 << 30 | face id (cuboid: ray_aabb.rs:62-66, far side + 3 as in the reference), 0xffffffff for
 * FeatureId::Unknown (polygons). */
int ncb2d_ray_cast(ncb_ctx* ctx, uint32_t n, const uint32_t* type, const float* param, const float* pose, const float* poly_points,
                   uint32_t n_poly_points, const float* rays, uint8_t* found, float* out, uint32_t* feature);
/* A fresh ncollide2d CollisionWorld of balls, cuboids and convex polygons (host SoA): pos = translation x y, rot = UnitComplex re im,
 * shape_param as in ncb2d_contact, groups = 3 words per object or NULL, query_limit / ang_pred = GeometricQueryType::Contacts(linear,
 * angular); poly_normals is required when the world holds polygons. */
typedef struct ncb2d_objects {
    uint32_t n;
    const float* pos;
    const float* rot;
    const uint32_t* shape_type;
    const float* shape_param;
    const uint32_t* groups;
    const float* query_limit;
    const float* ang_pred;
    const float* poly_points;
    const float* poly_normals;
    uint32_t n_poly_points;
    const uint8_t* query_kind; /* NULL: every object is GeometricQueryType::Contacts(query_limit, ang_pred); else per object 0 Contacts,
                                  1 Proximity(query_limit): a sensor */
} ncb2d_objects;
/* ncollide2d CollisionWorld::update for such a world (pipeline/world.rs:104-119): fat AABBs -> broad-phase pairs (object1 = larger handle,
 * like the 3-D path) -> contact manifolds from BallBall / BallConvexPolyhedron / ConvexPolyhedronConvexPolyhedron generators with 2-D
 * features and clipping (shape/convex_polygonal_feature2.rs).  Output: pairs (2 words each, grouped by generator), manifold_start / _count per
 * pair, contacts (7 floats: world1, world2, normal, depth) and features (2 words per contact: kind << 30 | id, kind 1 = face, 2 = vertex;
 * a ball's feature is face 0).  diag (optional, 4 words): reference panics, EPA capacity overflows, manifolds beyond 4 contacts,
 * traversal-stack overflows — all expected 0.  Returns 1 when an output was truncated (the needed counts are in n_pairs / n_contacts). */
int ncb2d_world_update(ncb_ctx* ctx, const ncb2d_objects* objs, float margin, uint32_t* pairs, uint32_t cap_pairs, uint32_t* manifold_start,
                       uint8_t* manifold_count, float* contacts, uint32_t* features, uint32_t cap_contacts, uint32_t* n_pairs,
                       uint32_t* n_contacts, uint32_t* diag);
/* Proximity status of every pair of the last ncb2d_world_update, in the order of its pairs: NCB_PROXIMITY_* for pairs with a sensor (the
 * narrow phase runs the proximity detector instead of a contact generator on them, narrow_phase.rs:138-167: BallBall / PlaneSupportMap /
 * SupportMapSupportMap detectors with margin = the two query limits added; their manifolds are empty), NCB_PROXIMITY_NONE for the others. */
int ncb2d_world_fetch_proximity(ncb_ctx* ctx, uint8_t* prox, uint32_t cap_pairs);
/* World ray queries of ncollide2d: glue::interferences_with_ray (first_only = 0) / first_interference_with_ray (first_only = 1)
 * (pipeline/glue/query.rs:13-77,183-224) against the world of the last ncb2d_world_update (its objects and broad-phase boxes stay on the
 * device).  rays[5 n] = origin x y, dir x y, max_toi; groups = the query's CollisionGroups (membership, whitelist, blacklist) or NULL.
 * A candidate is an object whose stored (fat) box the ray enters within max_toi; it is kept when the groups allow it and its shape's
 * RayCast::toi_and_normal_with_ray(position, ray, max_toi, solid = true) answers Some.  Rows sorted by (ray, handle): idx[2 k] = (ray,
 * handle), val[3 k] = (toi, normal), feat[k] as in ncb2d_ray_cast; first_only keeps the smallest toi per ray (ties: smallest handle).
 * cap in rows; *n_out = rows found; returns 1 when truncated.  The queries read the context's broad-phase buffers as the last
 * ncb2d_world_update left them: any other update on the same context (3-D world, stepping world) invalidates them — query before it,
 * or keep a context per world. */
int ncb2d_world_ray_cast(ncb_ctx* ctx, uint32_t n_rays, const float* rays, const uint32_t* groups, int first_only, uint32_t* idx, float* val,
                         uint32_t* feat, uint32_t cap, uint32_t* n_out);
/* glue::interferences_with_aabb (kind 0; 4 floats per query: mins x y, maxs x y) / interferences_with_point (kind 2; 2 floats)
 * (pipeline/glue/query.rs:79-181) against the world of the last ncb2d_world_update: candidates from the stored boxes, the query's collision
 * groups, and for points the shape's PointQuery::contains_point (ball, cuboid, plane; convex polygon through gjk::project_origin).
 * idx[2 k] = (query, handle), sorted; cap in rows; returns 1 when truncated. */
int ncb2d_world_query(ncb_ctx* ctx, int kind, uint32_t n_queries, const float* queries, const uint32_t* groups, uint32_t* idx, uint32_t cap,
                      uint32_t* n_out);

/* ---- ncollide2d: RayCast for Polyline (the 2-D counterpart of the TriMesh ray path) -------------------------------------------------- */
/* Polyline::new(points, indices) (shape/polyline.rs:57-120): 2 floats per point, 2 point indices per edge; edges == NULL builds the
 * line strip 0-1, 1-2, ... like `indices = None`.  One leaf per edge (Segment::local_aabb), device BVH.  The handle is a ncb_mesh. */
int ncb2d_polyline_create(ncb_ctx* ctx, uint32_t n_points, const float* xy, uint32_t n_edges, const uint32_t* edges, ncb_mesh** out);
void ncb2d_polyline_destroy(ncb_mesh* polyline);
/* RayCast::toi_and_normal_with_ray for a batch (query/ray/ray_polyline.rs:22-50,113-146; the segment test is RayCast for Segment, dim2,
 * query/ray/ray_support_map.rs:219-293).  pose = Isometry2 (x, y, re, im) or NULL; origins / dirs 2 floats per ray; max_tois optional (one
 * limit per ray).  toi < 0 = None; feature = edge, or edge + n_edges when the segment answered FeatureId::Face(1) (ray_polyline.rs:41-45);
 * normal (optional, 2 floats) = pose * the segment's scaled normal, NOT normalised, as the reference returns it.  As in the reference,
 * max_toi limits the AABB tests only: a segment hit beyond it is reported when its AABB was entered before it.  Host buffers, chunked
 * upload | cast | download like ncb_trimesh_ray_cast_uv. */
int ncb2d_polyline_ray_cast(ncb_mesh* polyline, const float* pose, uint32_t n_rays, const float* origins, const float* dirs, float max_toi,
                            const float* max_tois, float* toi, uint32_t* feature, float* normal);
/* Same with device-resident rays / results (8-byte aligned device pointers; pose on the host); asynchronous on the context's stream. */
int ncb2d_polyline_ray_cast_device(ncb_mesh* polyline, const float* pose, uint32_t n_rays, const float* d_origins, const float* d_dirs,
                                   float max_toi, const float* d_max_tois, float* d_toi, uint32_t* d_feature, float* d_normal);

#ifdef __cplusplus
}
#endif
#endif /* NCB200_H */
