// Proximity-only interactions on the device: objects whose GeometricQueryType is Proximity(margin)
// ("sensors") get a Proximity status per interfering pair instead of a contact manifold.
//
// Replaces (reference, file:line):
//   NarrowPhase::handle_interaction / update_proximity   pipeline/narrow_phase/narrow_phase.rs:123-143,226-247
//   DefaultProximityDispatcher                           proximity_detector/default_proximity_dispatcher.rs:19-47
//   BallBallProximityDetector                            proximity_detector/ball_ball_proximity_detector.rs:24-43,
//                                                        query/proximity/proximity_ball_ball.rs:8-36
//   PlaneSupportMap / SupportMapPlane detectors          proximity_detector/plane_support_map_proximity_detector.rs:37-67,
//                                                        query/proximity/proximity_plane_support_map.rs:9-47
//   SupportMapSupportMapProximityDetector                proximity_detector/support_map_support_map_proximity_detector.rs:31-61,
//                                                        query/proximity/proximity_support_map_support_map.rs:36-75,
//                                                        query/algorithms/gjk.rs:76-177 with exact_dist = false
//   Ball as a support map                                shape/ball.rs:29-48 (support_point ignores the rotation and
//                                                        normalises the direction; support_point_toward does not)
//   query::proximity                                     query/proximity/proximity_shape_shape.rs:8-33
//
// World path: after the pair search a small pass re-keys the pairs that involve a sensor (four extra key segments: ball x
// ball, plane x support map, support map x support map without / with a hull operand), the counting sort groups them, and ONE persistent kernel runs the
// four segments (contiguous in the sorted pair array) one pair per thread; the contact kernels never see those pairs.
// The kernels run only when ncb_set_query_types marked at least one sensor: a world without sensors takes the old path.
#include "gjk.cuh"
#include "ncb_internal.h"
#include "shapes.cuh"
#include "vec.cuh"

namespace ncb {

namespace {

// A support-mapped operand of the proximity GJK: the Support of gjk.cuh plus kind 3 = ball (radius in he.x).
NCB_HD Support prox_support(const Shape& s) {
    Support g;
    g.kind = s.type == NCB_SHAPE_CUBOID ? 0 : (s.type == NCB_SHAPE_CONVEX_HULL ? 1 : 3);
    g.he = s.type == NCB_SHAPE_BALL ? v3(s.radius, 0.f, 0.f) : s.he;
    g.hull = s.hull;
    return g;
}
// SupportMap::support_point (support_map.rs:26-29); Ball overrides it (ball.rs:31-33)
NCB_HD V3 prox_support_point(const Support& g, const Iso& m, V3 dir) {
    if (g.kind == 3) return m.t + normalize(dir) * g.he.x;
    return support_point(g, m, dir);
}
// SupportMap::support_point_toward (support_map.rs:32-35); Ball: ball.rs:36-38
NCB_HD V3 prox_support_point_toward(const Support& g, const Iso& m, V3 unit_dir) {
    if (g.kind == 3) return m.t + unit_dir * g.he.x;
    return support_point(g, m, unit_dir);
}
// CSOPoint::from_shapes (cso_point.rs:70-85).  As in gjk.cuh, the two independent support evaluations are issued in a canonical
// order (the vertex-scanning hull operand second) so that lanes holding (x, hull) and (hull, x) pairs run the scan together;
// each operand still sees its own isometry and direction, the values are unchanged.
NCB_HD CSOPoint prox_cso(const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, V3 dir) {
    CSOPoint c;
    bool swap = g1.kind == 1 && g2.kind != 1;
    const Support& ga = swap ? g2 : g1;
    const Support& gb = swap ? g1 : g2;
    const Iso& ia = swap ? m2 : m1;
    const Iso& ib = swap ? m1 : m2;
    V3 da = swap ? -dir : dir;
    V3 sa = prox_support_point(ga, ia, da);
    V3 sb = prox_support_point(gb, ib, -da);
    c.orig1 = swap ? sb : sa;
    c.orig2 = swap ? sa : sb;
    c.point = c.orig1 - c.orig2;
    return c;
}

NCB_HD uint8_t proximity_ball_ball(V3 c1, float r1, V3 c2, float r2, float margin) {
    V3 delta_pos = c2 - c1;
    float distance_squared = norm_squared(delta_pos);
    float sum_radius = r1 + r2;
    float sum_radius_with_error = sum_radius + margin;
    if (distance_squared <= sum_radius_with_error * sum_radius_with_error)
        return distance_squared <= sum_radius * sum_radius ? NCB_PROXIMITY_INTERSECTING : NCB_PROXIMITY_WITHIN_MARGIN;
    return NCB_PROXIMITY_DISJOINT;
}

NCB_HD uint8_t proximity_plane_support_map(const Iso& mplane, V3 plane_n, const Iso& mother, const Support& other, float margin) {
    V3 plane_normal = iso_mul_vec(mplane, plane_n);
    V3 deepest = prox_support_point_toward(other, mother, -plane_normal);
    float distance = dot(plane_normal, mplane.t - deepest);
    if (distance >= -margin) return distance >= 0.f ? NCB_PROXIMITY_INTERSECTING : NCB_PROXIMITY_WITHIN_MARGIN;
    return NCB_PROXIMITY_DISJOINT;
}

// proximity_support_map_support_map_with_params: simplex.reset(CSOPoint::from_shapes(.., dir)) + gjk::closest_points(.., margin,
// exact_dist = false, ..).  axis (in/out): the detector's sep_axis; has_axis false = None.
static __device__ __noinline__ uint8_t proximity_sm_sm(const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, float max_dist, V3& axis,
                                                       bool& has_axis) {
    const float eps_tol = NCB_EPS * 10.0f;
    const float eps_rel = sqrtf(eps_tol);
    V3 dir;
    if (has_axis)
        dir = axis;
    else if (!unit_try_new(m2.t - m1.t, NCB_EPS, dir))
        dir = v3(1.f, 0.f, 0.f);
    Simplex s;
    simplex_init(s, prox_cso(m1, g1, m2, g2, dir));
    V3 proj = simplex_project_origin_and_reduce(s);
    V3 old_dir;
    {
        V3 pd;
        if (!unit_try_new(proj, 0.f, pd)) {
            has_axis = false;
            return NCB_PROXIMITY_INTERSECTING;
        }
        old_dir = -pd;
    }
    float max_bound = NCB_FMAX;
    uint8_t res;
    V3 out_dir;
    int niter = 0;
    for (;;) {
        float old_max_bound = max_bound;
        float dist;
        if (!unit_try_new_and_get(-proj, eps_tol, dir, dist)) {
            has_axis = false;
            return NCB_PROXIMITY_INTERSECTING;  // the origin is on the simplex
        }
        max_bound = dist;
        if (max_bound >= old_max_bound) {
            res = NCB_PROXIMITY_WITHIN_MARGIN, out_dir = old_dir;
            break;
        }
        CSOPoint cso = prox_cso(m1, g1, m2, g2, dir);
        float min_bound = -dot(dir, cso.point);
        if (min_bound > max_dist) {
            res = NCB_PROXIMITY_DISJOINT, out_dir = dir;
            break;
        } else if (min_bound > 0.f && max_bound <= max_dist) {
            res = NCB_PROXIMITY_WITHIN_MARGIN, out_dir = old_dir;
            break;
        } else if (max_bound - min_bound <= eps_rel * max_bound) {
            res = NCB_PROXIMITY_WITHIN_MARGIN, out_dir = dir;
            break;
        }
        if (!simplex_add_point(s, cso)) {
            res = NCB_PROXIMITY_WITHIN_MARGIN, out_dir = dir;
            break;
        }
        old_dir = dir;
        proj = simplex_project_origin_and_reduce(s);
        if (s.dim == 3) {
            if (min_bound >= eps_tol) {
                res = NCB_PROXIMITY_WITHIN_MARGIN, out_dir = old_dir;
                break;
            }
            has_axis = false;
            return NCB_PROXIMITY_INTERSECTING;  // point inside of the CSO
        }
        niter += 1;
        if (niter == 10000) {
            res = NCB_PROXIMITY_DISJOINT, out_dir = v3(1.f, 0.f, 0.f);
            break;
        }
    }
    has_axis = true;
    axis = out_dir;
    return res;
}

// ProximityDetector::update of the detector DefaultProximityDispatcher picks for (object i1, object i2).
NCB_HD uint8_t proximity_pair(const DevObjects& o, const DevHulls& H, uint32_t i1, uint32_t i2, float margin, V3& axis, bool& has_axis) {
    uint32_t t1 = __ldg(&o.type[i1]) & 3u, t2 = __ldg(&o.type[i2]) & 3u;
    if (t1 == NCB_SHAPE_BALL && t2 == NCB_SHAPE_BALL) {
        V3 c1 = v3(__ldg(o.pos + 3 * i1), __ldg(o.pos + 3 * i1 + 1), __ldg(o.pos + 3 * i1 + 2));
        V3 c2 = v3(__ldg(o.pos + 3 * i2), __ldg(o.pos + 3 * i2 + 1), __ldg(o.pos + 3 * i2 + 2));
        return proximity_ball_ball(c1, __ldg(&o.param[i1]).x, c2, __ldg(&o.param[i2]).x, margin);
    }
    if (t1 == NCB_SHAPE_PLANE && t2 == NCB_SHAPE_PLANE) return NCB_PROXIMITY_NONE;
    Shape a = load_shape(o, H, i1, t1), b = load_shape(o, H, i2, t2);
    Iso ma = load_iso(o, i1), mb = load_iso(o, i2);
    if (t1 == NCB_SHAPE_PLANE) return proximity_plane_support_map(ma, a.he, mb, prox_support(b), margin);
    if (t2 == NCB_SHAPE_PLANE) return proximity_plane_support_map(mb, b.he, ma, prox_support(a), margin);
    return proximity_sm_sm(ma, prox_support(a), mb, prox_support(b), margin, axis, has_axis);
}

#ifndef NCB_HOST_SHIM  // tests/host_shim compiles the per-pair functions above for the host; kernels and launchers are CUDA only
// Pairs that involve a sensor leave their contact key for one of the three proximity keys (plane x plane keeps K_NONE: neither
// dispatcher has an algorithm for it, no interaction edge exists).
__global__ void __launch_bounds__(256) k_prox_rekey(const uint2* __restrict__ pairs, uint8_t* __restrict__ keys, uint32_t cap,
                                                    const uint8_t* __restrict__ qkind, const uint32_t* __restrict__ type, const DevCounters* cnt) {
    uint32_t np = min(cnt->n_pairs, cap);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
        uint2 pr = pairs[p];
        if ((__ldg(&qkind[pr.x]) | __ldg(&qkind[pr.y])) == 0) continue;
        if (keys[p] == K_NONE) continue;
        uint32_t t1 = __ldg(&type[pr.x]) & 3u, t2 = __ldg(&type[pr.y]) & 3u;
        uint8_t k = (t1 == NCB_SHAPE_BALL && t2 == NCB_SHAPE_BALL) ? K_PROX_BALL_BALL
                    : (t1 == NCB_SHAPE_PLANE || t2 == NCB_SHAPE_PLANE) ? K_PROX_PLANE
                    : (t1 == NCB_SHAPE_CONVEX_HULL || t2 == NCB_SHAPE_CONVEX_HULL) ? K_PROX_SM_HULL
                                                                                   : K_PROX_SM;
        keys[p] = k;
    }
}

// The four proximity key segments are adjacent: one persistent launch, warps mostly inside a single segment.
__global__ void __launch_bounds__(128) k_proximity(DevObjects o, DevHulls H, const uint2* __restrict__ pairs, const uint32_t* __restrict__ pair_index,
                                                   DevCounters* cnt, uint32_t* __restrict__ manifold_start, uint8_t* __restrict__ manifold_count,
                                                   uint8_t* __restrict__ prox) {
    __shared__ uint32_t hist[4];
    if (threadIdx.x < 4) hist[threadIdx.x] = 0;
    __syncthreads();
    uint32_t seg_begin = cnt->key_start[K_PROX_BALL_BALL];
    uint32_t seg_end = cnt->key_start[K_PROX_SM_HULL] + cnt->key_hist[K_PROX_SM_HULL];
    for (uint32_t p = seg_begin + blockIdx.x * blockDim.x + threadIdx.x; p < seg_end; p += gridDim.x * blockDim.x) {
        uint2 pr = __ldg(&pairs[p]);
        float margin = __ldg(&o.qlimit[pr.x]) + __ldg(&o.qlimit[pr.y]);  // narrow_phase.rs:138
        V3 axis = v3(0.f, 0.f, 0.f);
        bool has_axis = false;  // fresh detector
        uint8_t st = proximity_pair(o, H, pr.x, pr.y, margin, axis, has_axis);
        uint32_t out_index = pair_index ? __ldg(&pair_index[p]) : p;
        prox[out_index] = st;
        manifold_start[out_index] = 0;
        manifold_count[out_index] = 0;
        atomicAdd(&hist[st & 3], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 3 && hist[threadIdx.x]) atomicAdd(&cnt->prox_hist[threadIdx.x], hist[threadIdx.x]);
}

// Stepping world (sim.cu): the detector state of a pair lives in its state slot — the status (Interaction::Proximity's second
// field, Disjoint on a new edge) and the support-map detector's sep_axis (slot_dir: xyz + "Some" flag).  update_proximity
// (narrow_phase.rs:123-143): run the detector, emit ProximityEvent(h1, h2, prev, new) when the status changed, store it.
__global__ void __launch_bounds__(128) k_proximity_persist(DevObjects o, DevHulls H, const uint2* __restrict__ pairs, const uint32_t* __restrict__ slot_of,
                                                           const DevCounters* __restrict__ cnt, float4* __restrict__ slot_dir, uint8_t* __restrict__ slot_prox,
                                                           uint4* __restrict__ events, uint32_t* n_events, uint32_t cap_events) {
    uint32_t seg_begin = cnt->key_start[K_PROX_BALL_BALL];
    uint32_t seg_end = cnt->key_start[K_PROX_SM_HULL] + cnt->key_hist[K_PROX_SM_HULL];
    for (uint32_t p = seg_begin + blockIdx.x * blockDim.x + threadIdx.x; p < seg_end; p += gridDim.x * blockDim.x) {
        uint2 pr = __ldg(&pairs[p]);
        uint32_t slot = __ldg(&slot_of[p]);
        float margin = __ldg(&o.qlimit[pr.x]) + __ldg(&o.qlimit[pr.y]);
        float4 d = slot_dir[slot];
        V3 axis = v3(d.x, d.y, d.z);
        bool has_axis = d.w != 0.f;
        uint8_t st = proximity_pair(o, H, pr.x, pr.y, margin, axis, has_axis);
        uint8_t prev = slot_prox[slot];
        if (st != prev) {
            uint32_t k = atomicAdd(n_events, 1u);
            if (k < cap_events) events[k] = make_uint4(pr.x, pr.y, prev, st);
            slot_prox[slot] = st;
        }
        slot_dir[slot] = make_float4(axis.x, axis.y, axis.z, has_axis ? 1.f : 0.f);
    }
}

// Stage entry: caller-provided pairs, fresh detectors, no sorting.
__global__ void __launch_bounds__(128) k_proximity_batch(DevObjects o, DevHulls H, const uint2* __restrict__ pairs, uint32_t n,
                                                         const float* __restrict__ margins, uint8_t* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint2 pr = __ldg(&pairs[p]);
    float margin = margins ? __ldg(&margins[p]) : __ldg(&o.qlimit[pr.x]) + __ldg(&o.qlimit[pr.y]);
    V3 axis = v3(0.f, 0.f, 0.f);
    bool has_axis = false;
    out[p] = proximity_pair(o, H, pr.x, pr.y, margin, axis, has_axis);
}

#endif  // NCB_HOST_SHIM
}  // namespace

#ifndef NCB_HOST_SHIM
cudaError_t launch_prox_rekey(ncb_ctx* c, uint32_t cap_pairs) {
    k_prox_rekey<<<c->sm_count * 4, 256, 0, c->stream>>>(c->pairs_raw.p, c->keys_raw.p, cap_pairs, c->qkind.p, c->type.p, c->counters.p);
    return cudaGetLastError();
}
cudaError_t launch_proximity_segments(ncb_ctx* c, const DevObjects& o, const uint2* pairs, const uint32_t* pair_index, cudaStream_t s) {
    k_proximity<<<c->sm_count * 8, 128, 0, s>>>(o, c->hulls, pairs, pair_index, c->counters.p, c->manifold_start.p, c->manifold_count.p, c->prox.p);
    return cudaGetLastError();
}
cudaError_t launch_proximity_persistent(ncb_ctx* c, const DevObjects& o, const uint2* pairs, const uint32_t* slot_of, float4* slot_dir,
                                        uint8_t* slot_prox, uint4* events, uint32_t* n_events, uint32_t cap_events) {
    k_proximity_persist<<<c->sm_count * 8, 128, 0, c->stream>>>(o, c->hulls, pairs, slot_of, c->counters.p, slot_dir, slot_prox, events, n_events,
                                                                 cap_events);
    return cudaGetLastError();
}
cudaError_t launch_proximity_batch(ncb_ctx* c, const DevObjects& o, const uint2* pairs, uint32_t n, const float* margins, uint8_t* out) {
    if (n == 0) return cudaSuccess;
    k_proximity_batch<<<(n + 127) / 128, 128, 0, c->stream>>>(o, c->hulls, pairs, n, margins, out);
    return cudaGetLastError();
}

#endif  // NCB_HOST_SHIM

}  // namespace ncb
