// Cooperative EPA: EIGHT lanes per penetrating pair, polytope in shared memory (included by narrow.cu).
//
// Same algorithm, same arithmetic and same decisions as epa_init / epa_step in gjk.cuh (EPA::closest_points, epa3.rs:219-430);
// what changes is who executes what:
//   * the two support maps scan the hull vertices 8 at a time (first maximum = smallest index among the maxima, like the
//     sequential strict `>` scan of point_cloud_support_point.rs);
//   * the faces built around the new vertex (one per silhouette edge, ~5 per step) are built by one lane each;
//   * heap operations and the silhouette flood stay sequential on the group's first lane, but on shared memory;
//   * groups fetch their next pair individually, so nobody waits for the slowest pair of a batch.
// Anything outside the fixed shared-memory capacities, and the rare degenerate start simplices (dimension < 3), is handed
// to the thread-per-pair kernel (k_cc_epa PASS 2 over A.epa_long), which has the large capacities.
#pragma once

#define CE_G 8        // lanes per pair
#define CE_V 24       // vertices
#define CE_F 64       // faces
#define CE_H 64       // heap entries
#define CE_S 32       // silhouette edges
#define CE_STK 64     // flood stack
#define CE_PAIRS 16   // pairs per CTA (128 threads)

struct CoopEpa {
    V3 vpoint[CE_V], vorig1[CE_V], vorig2[CE_V];
    V3 fnormal[CE_F];
    uint8_t fpt[CE_F][3], fadj[CE_F][3], fdel[CE_F];
    float hdist[CE_H];
    uint8_t hid[CE_H];
    uint8_t sil_face[CE_S], sil_opp[CE_S];
    uint8_t stk_face[CE_STK], stk_opp[CE_STK];
    float pend_nd[CE_S];
    uint8_t pend_flag[CE_S];  // 1 inside, 2 closer than the popped face (epa3.rs:393), 4 invalid FaceId (neg_dist > 10 eps)
    int nverts, nfaces, nheap, nsil;
    int status;  // written by lane 0, read by the group
    uint32_t pop_id;
    float pop_nd;
};

enum { CE_CONTINUE = 0, CE_OK = 1, CE_FAIL = 2, CE_DEFER = 3 };

struct SlimSupport {
    int kind;  // 0 cuboid, 1 hull
    V3 he;
    uint32_t nv;
    const float* pts;
};

__device__ __forceinline__ unsigned ce_mask() { return 0xffu << ((threadIdx.x & 31) & ~7); }
__device__ __forceinline__ float ce_bcast(float v, int src) { return __shfl_sync(ce_mask(), v, ((threadIdx.x & 31) & ~7) + src); }
__device__ __forceinline__ uint32_t ce_bcast(uint32_t v, int src) { return __shfl_sync(ce_mask(), v, ((threadIdx.x & 31) & ~7) + src); }
__device__ __forceinline__ int ce_bcast(int v, int src) { return __shfl_sync(ce_mask(), v, ((threadIdx.x & 31) & ~7) + src); }
__device__ __forceinline__ void ce_sync() { __syncwarp(ce_mask()); }

// local_support_point, cooperative for hulls
__device__ __forceinline__ V3 ce_local_support(const SlimSupport& g, V3 dir) {
    if (g.kind == 0) return v3(copysignf(g.he.x, dir.x), copysignf(g.he.y, dir.y), copysignf(g.he.z, dir.z));
    const int gl = threadIdx.x & 7;
    float best_dot = 0.f;
    uint32_t best = 0xffffffffu;
    for (uint32_t i = gl; i < g.nv; i += CE_G) {
        V3 p = v3(__ldg(g.pts + 3 * i), __ldg(g.pts + 3 * i + 1), __ldg(g.pts + 3 * i + 2));
        float d = dot(p, dir);
        if (best == 0xffffffffu || d > best_dot) best_dot = d, best = i;
    }
    const unsigned m = ce_mask();
    for (int off = 4; off; off >>= 1) {
        float od = __shfl_xor_sync(m, best_dot, off);
        uint32_t oi = __shfl_xor_sync(m, best, off);
        if (oi != 0xffffffffu && (best == 0xffffffffu || od > best_dot || (od == best_dot && oi < best))) best_dot = od, best = oi;
    }
    return v3(__ldg(g.pts + 3 * best), __ldg(g.pts + 3 * best + 1), __ldg(g.pts + 3 * best + 2));
}
__device__ __forceinline__ V3 ce_support_point(const SlimSupport& g, const Iso& m, V3 dir) {
    V3 ld = iso_inv_vec(m, dir);
    return iso_mul_point(m, ce_local_support(g, ld));
}
__device__ __forceinline__ CSOPoint ce_cso(const Iso& m1, const SlimSupport& g1, const Iso& m2, const SlimSupport& g2, V3 dir) {
    CSOPoint c;
    c.orig1 = ce_support_point(g1, m1, dir);
    c.orig2 = ce_support_point(g2, m2, -dir);
    c.point = c.orig1 - c.orig2;
    return c;
}

// Face::new (epa3.rs:93-114) into face slot f
__device__ __forceinline__ bool ce_face_new(CoopEpa& e, uint32_t f, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t a0, uint32_t a1, uint32_t a2) {
    V3 A = e.vpoint[p0], B = e.vpoint[p1], C = e.vpoint[p2];
    Loc loc;
    proj_triangle(A, B, C, v3(0.f, 0.f, 0.f), loc);
    e.fpt[f][0] = (uint8_t)p0, e.fpt[f][1] = (uint8_t)p1, e.fpt[f][2] = (uint8_t)p2;
    e.fadj[f][0] = (uint8_t)a0, e.fadj[f][1] = (uint8_t)a1, e.fadj[f][2] = (uint8_t)a2;
    e.fdel[f] = 0;
    V3 n;
    if (!unit_try_new(cross(B - A, C - A), NCB_EPS, n)) n = v3(0.f, 0.f, 0.f);
    e.fnormal[f] = n;
    return loc.kind == LOC_FACE;
}
__device__ __forceinline__ void ce_face_closest_points(const CoopEpa& e, uint32_t f, V3& p1, V3& p2) {
    uint32_t i0 = e.fpt[f][0], i1 = e.fpt[f][1], i2 = e.fpt[f][2];
    Loc loc;
    proj_triangle(e.vpoint[i0], e.vpoint[i1], e.vpoint[i2], v3(0.f, 0.f, 0.f), loc);
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (loc.kind == LOC_FACE) b0 = loc.b0, b1 = loc.b1, b2 = loc.b2;
    p1 = e.vorig1[i0] * b0 + e.vorig1[i1] * b1 + e.vorig1[i2] * b2;
    p2 = e.vorig2[i0] * b0 + e.vorig2[i1] * b1 + e.vorig2[i2] * b2;
}

// ---- sequential parts (first lane of the group), same code as gjk.cuh on the shared-memory arrays ----------------------
__device__ __forceinline__ void ce_heap_sift_up(CoopEpa& e, int start, int pos) {
    float ed = e.hdist[pos];
    uint8_t ei = e.hid[pos];
    while (pos > start) {
        int parent = (pos - 1) / 2;
        if (ed <= e.hdist[parent]) break;
        e.hdist[pos] = e.hdist[parent];
        e.hid[pos] = e.hid[parent];
        pos = parent;
    }
    e.hdist[pos] = ed;
    e.hid[pos] = ei;
}
__device__ __forceinline__ bool ce_heap_push(CoopEpa& e, uint32_t id, float nd) {
    if (e.nheap >= CE_H) return false;
    e.hid[e.nheap] = (uint8_t)id;
    e.hdist[e.nheap] = nd;
    e.nheap++;
    ce_heap_sift_up(e, 0, e.nheap - 1);
    return true;
}
__device__ __forceinline__ bool ce_heap_pop(CoopEpa& e, uint32_t& out_id, float& out_nd) {
    if (e.nheap == 0) return false;
    --e.nheap;
    float item_d = e.hdist[e.nheap];
    uint8_t item_i = e.hid[e.nheap];
    if (e.nheap > 0) {
        float td = e.hdist[0];
        uint8_t ti = e.hid[0];
        int end = e.nheap, pos = 0, child = 1;
        float ed = item_d;
        uint8_t ei = item_i;
        item_d = td;
        item_i = ti;
        while (end >= 2 && child <= end - 2) {
            if (e.hdist[child] <= e.hdist[child + 1]) child += 1;
            e.hdist[pos] = e.hdist[child];
            e.hid[pos] = e.hid[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            e.hdist[pos] = e.hdist[child];
            e.hid[pos] = e.hid[child];
            pos = child;
        }
        e.hdist[pos] = ed;
        e.hid[pos] = ei;
        ce_heap_sift_up(e, 0, pos);
    }
    out_id = item_i;
    out_nd = item_d;
    return true;
}
__device__ __forceinline__ int ce_next_ccw(const CoopEpa& e, uint32_t f, uint32_t id, bool& panicked) {
    if (e.fpt[f][0] == id) return 1;
    if (e.fpt[f][1] == id) return 2;
    if (e.fpt[f][2] != id) panicked = true;
    return 0;
}
__device__ __forceinline__ bool ce_can_be_seen_by(const CoopEpa& e, uint32_t f, uint32_t point, uint32_t opp) {
    V3 p0 = e.vpoint[e.fpt[f][opp]];
    V3 pt = e.vpoint[point];
    if (dot(pt - p0, e.fnormal[f]) >= -(NCB_EPS * 10.0f)) return true;
    V3 p1 = e.vpoint[e.fpt[f][(opp + 1) % 3]], p2 = e.vpoint[e.fpt[f][(opp + 2) % 3]];
    V3 p1p2 = p2 - p1, p1p3 = pt - p1;
    float eps_tol = NCB_EPS * 100.0f;
    return relative_eq(norm_squared(cross(p1p2, p1p3)), 0.f, eps_tol * eps_tol);
}
// compute_silhouette x3 as one LIFO walk (see epa_compute_silhouette3).  Returns CE_CONTINUE / CE_FAIL / CE_DEFER.
__device__ __noinline__ int ce_silhouette(CoopEpa& e, uint32_t point, uint32_t id0, uint32_t opp0, uint32_t id1, uint32_t opp1, uint32_t id2,
                                          uint32_t opp2, bool& panicked) {
    int sp = 0;
    e.stk_face[sp] = (uint8_t)id2, e.stk_opp[sp] = (uint8_t)opp2, sp++;
    e.stk_face[sp] = (uint8_t)id1, e.stk_opp[sp] = (uint8_t)opp1, sp++;
    e.stk_face[sp] = (uint8_t)id0, e.stk_opp[sp] = (uint8_t)opp0, sp++;
    int nsil = 0;
    while (sp > 0) {
        sp--;
        uint32_t id = e.stk_face[sp], opp = e.stk_opp[sp];
        if (e.fdel[id]) continue;
        if (!ce_can_be_seen_by(e, id, point, opp)) {
            if (nsil >= CE_S) return CE_DEFER;
            e.sil_face[nsil] = (uint8_t)id, e.sil_opp[nsil] = (uint8_t)opp, nsil++;
        } else {
            e.fdel[id] = 1;
            uint32_t adj_pt_id1 = (opp + 2) % 3, adj_pt_id2 = opp;
            uint32_t adj1 = e.fadj[id][adj_pt_id1], adj2 = e.fadj[id][adj_pt_id2];
            uint32_t o1 = ce_next_ccw(e, adj1, e.fpt[id][adj_pt_id1], panicked);
            uint32_t o2 = ce_next_ccw(e, adj2, e.fpt[id][adj_pt_id2], panicked);
            if (panicked) return CE_FAIL;
            if (sp + 2 > CE_STK) return CE_DEFER;
            e.stk_face[sp] = (uint8_t)adj2, e.stk_opp[sp] = (uint8_t)o2, sp++;
            e.stk_face[sp] = (uint8_t)adj1, e.stk_opp[sp] = (uint8_t)o1, sp++;
        }
    }
    e.nsil = nsil;
    return CE_CONTINUE;
}

// One pair, start to end.  Returns CE_OK (out1, out2, out_n valid on every lane), CE_FAIL (panicked says whether the
// reference itself would have panicked) or CE_DEFER.
__device__ __noinline__ int ce_run(CoopEpa& e, const Iso& m1, const SlimSupport& g1, const Iso& m2, const SlimSupport& g2, int sdim, CSOPoint* sv,
                                   V3& out1, V3& out2, V3& out_n, bool& panicked) {
    const int gl = threadIdx.x & 7;
    const float eps_tol = NCB_EPS * 100.0f;
    panicked = false;
    if (sdim != 3) return CE_DEFER;  // degenerate start simplices are rare: thread-per-pair kernel
    // ---- initial tetrahedron (epa3.rs:243-283) ----
    {
        V3 dp1 = sv[1].point - sv[0].point, dp2 = sv[2].point - sv[0].point, dp3 = sv[3].point - sv[0].point;
        if (dot(cross(dp1, dp2), dp3) > 0.f) {
            CSOPoint t = sv[1];
            sv[1] = sv[2];
            sv[2] = t;
        }
    }
    if (gl < 4) e.vpoint[gl] = sv[gl].point, e.vorig1[gl] = sv[gl].orig1, e.vorig2[gl] = sv[gl].orig2;
    ce_sync();
    bool in = false;
    if (gl == 0) in = ce_face_new(e, 0, 0, 1, 2, 3, 1, 2);
    if (gl == 1) in = ce_face_new(e, 1, 1, 3, 2, 3, 2, 0);
    if (gl == 2) in = ce_face_new(e, 2, 0, 2, 3, 0, 1, 3);
    if (gl == 3) in = ce_face_new(e, 3, 0, 3, 1, 2, 1, 0);
    if (gl < 4) {
        e.pend_flag[gl] = in ? 1 : 0;
        e.pend_nd[gl] = -dot(e.fnormal[gl], e.vpoint[gl]);
    }
    ce_sync();
    if (gl == 0) {
        e.nverts = 4, e.nfaces = 4, e.nheap = 0, e.nsil = 0;
        int st = CE_CONTINUE;
        for (int j = 0; j < 4 && st == CE_CONTINUE; ++j) {
            if (!e.pend_flag[j]) continue;
            float nd = e.pend_nd[j];
            if (nd > NCB_EPS * 10.0f)
                st = CE_FAIL;  // FaceId::new(..)? fails
            else
                ce_heap_push(e, (uint32_t)j, nd);
        }
        e.pop_id = 0u;
        if (st == CE_CONTINUE && e.nheap == 0) st = CE_FAIL, e.pop_id = 0xffffffffu;  // heap.peek().unwrap() panics in the reference
        e.status = st;
    }
    ce_sync();
    int st = e.status;
    if (st != CE_CONTINUE) {
        panicked = e.pop_id == 0xffffffffu;
        return st;
    }
    float max_dist = NCB_FMAX;
    uint32_t best_id = e.hid[0];
    int niter = 0;
    // ---- expansion loop (epa3.rs:330-425) ----
    for (;;) {
        ce_sync();  // everybody has read the previous turn's status before it is overwritten
        if (gl == 0) {
            uint32_t id = 0;
            float nd = 0.f;
            bool got;
            do {
                got = ce_heap_pop(e, id, nd);
            } while (got && e.fdel[id]);
            e.status = got ? 1 : 0;
            e.pop_id = id, e.pop_nd = nd;
        }
        ce_sync();
        if (!e.status) {  // heap exhausted: the best face so far (epa3.rs:427-429)
            ce_face_closest_points(e, best_id, out1, out2);
            out_n = e.fnormal[best_id];
            return CE_OK;
        }
        const uint32_t fid = e.pop_id;
        const float neg_dist = e.pop_nd;
        const uint32_t fp0 = e.fpt[fid][0], fp1 = e.fpt[fid][1], fp2 = e.fpt[fid][2];
        const uint32_t fa0 = e.fadj[fid][0], fa1 = e.fadj[fid][1], fa2 = e.fadj[fid][2];
        const V3 fnorm = e.fnormal[fid];
        const int nverts = e.nverts;
        if (nverts >= CE_V) return CE_DEFER;
        CSOPoint cso = ce_cso(m1, g1, m2, g2, fnorm);
        const uint32_t support_point_id = (uint32_t)nverts;
        if (gl == 0) {
            e.vpoint[nverts] = cso.point, e.vorig1[nverts] = cso.orig1, e.vorig2[nverts] = cso.orig2;
            e.nverts = nverts + 1;
        }
        float candidate_max_dist = dot(cso.point, fnorm);
        if (candidate_max_dist < max_dist) best_id = fid, max_dist = candidate_max_dist;
        float curr_dist = -neg_dist;
        if (max_dist - curr_dist < eps_tol) {
            ce_sync();
            ce_face_closest_points(e, best_id, out1, out2);
            out_n = e.fnormal[best_id];
            return CE_OK;
        }
        ce_sync();
        if (gl == 0) {
            bool pk = false;
            e.fdel[fid] = 1;
            uint32_t o1 = ce_next_ccw(e, fa0, fp0, pk), o2 = ce_next_ccw(e, fa1, fp1, pk), o3 = ce_next_ccw(e, fa2, fp2, pk);
            int s = pk ? CE_FAIL : ce_silhouette(e, support_point_id, fa0, o1, fa1, o2, fa2, o3, pk);
            e.status = s;
            e.pop_id = pk ? 0xffffffffu : 0u;
        }
        ce_sync();
        st = e.status;
        if (st != CE_CONTINUE) {
            panicked = e.pop_id == 0xffffffffu;
            return st;
        }
        const int nsil = e.nsil;
        if (nsil == 0) return CE_FAIL;
        const uint32_t first_new = (uint32_t)e.nfaces;
        // which silhouette edges still border a live face (epa3.rs:371), as a bit mask over k
        uint32_t vmask = 0;
        for (int r = 0; r * CE_G < nsil; ++r) {
            int k = r * CE_G + gl;
            bool valid = k < nsil && !e.fdel[e.sil_face[k]];
            unsigned b = __ballot_sync(ce_mask(), valid);
            vmask |= ((b >> ((threadIdx.x & 31) & ~7)) & 0xffu) << (r * CE_G);
        }
        const uint32_t count = __popc(vmask);
        if (first_new + count > CE_F) return CE_DEFER;
        for (int r = 0; r * CE_G < nsil; ++r) {
            int k = r * CE_G + gl;
            if (k < nsil && ((vmask >> k) & 1u)) {
                uint32_t efid = e.sil_face[k], eopp = e.sil_opp[k];
                uint32_t new_face_id = first_new + __popc(vmask & ((1u << k) - 1u));
                uint32_t pt_id1 = e.fpt[efid][(eopp + 2) % 3];
                uint32_t pt_id2 = e.fpt[efid][(eopp + 1) % 3];
                bool inside = ce_face_new(e, new_face_id, pt_id1, pt_id2, support_point_id, efid, (new_face_id + 1) & 0xffu, (new_face_id - 1) & 0xffu);
                e.fadj[efid][(eopp + 1) % 3] = (uint8_t)new_face_id;
                uint8_t flag = 0;
                float nd = 0.f;
                if (inside) {
                    float dist = dot(e.fnormal[new_face_id], e.vpoint[pt_id1]);
                    flag = 1;
                    if (dist < curr_dist) flag |= 2;
                    if (-dist > NCB_EPS * 10.0f) flag |= 4;
                    nd = -dist;
                }
                e.pend_flag[k] = flag;
                e.pend_nd[k] = nd;
            }
        }
        ce_sync();
        if (gl == 0) {
            int s = CE_CONTINUE;
            // the reference's loop order: the first edge (ascending k) that ends the search decides how
            for (int k = 0; k < nsil && s == CE_CONTINUE; ++k) {
                if (!((vmask >> k) & 1u)) continue;
                uint8_t fl = e.pend_flag[k];
                if (!(fl & 1)) continue;
                if (fl & 2)
                    s = CE_OK;  // epa3.rs:393-398: the popped face as it was
                else if (fl & 4)
                    s = CE_FAIL;
            }
            if (s == CE_CONTINUE) {
                for (int k = 0; k < nsil; ++k) {
                    if (!((vmask >> k) & 1u) || !(e.pend_flag[k] & 1)) continue;
                    uint32_t nf = first_new + __popc(vmask & ((1u << k) - 1u));
                    if (!ce_heap_push(e, nf, e.pend_nd[k])) {
                        s = CE_DEFER;
                        break;
                    }
                }
            }
            if (s == CE_CONTINUE) {
                if (count == 0)
                    s = CE_FAIL;
                else {
                    e.nfaces = (int)(first_new + count);
                    e.fadj[first_new][2] = (uint8_t)(first_new + count - 1);
                    e.fadj[first_new + count - 1][1] = (uint8_t)first_new;
                    e.nsil = 0;
                }
            }
            e.status = s;
        }
        ce_sync();
        st = e.status;
        if (st == CE_OK) {
            ce_face_closest_points(e, fid, out1, out2);  // topology bytes of the popped face are unchanged (only its deleted flag)
            out_n = fnorm;
            return CE_OK;
        }
        if (st != CE_CONTINUE) return st;
        niter += 1;
        if (niter > 10000) return CE_FAIL;
    }
}
