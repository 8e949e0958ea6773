// Device EPA for one pair per thread: fixed capacity, no recursion, no heap allocation.  Included at the end of gjk.cuh.
//
// Replaces (reference, file:line): query/algorithms/epa3.rs:219-454 (std BinaryHeap + Vec + recursive silhouette flood ->
// fixed arrays + explicit stack); query/contact/contact_support_map_support_map.rs:38-79; utils/ccw_face_normal.rs:26.
//
// The algorithm is written ONCE, over a polytope store `S`; two stores exist:
//   EpaState    48 vertices / 192 faces / 160 heap entries in per-thread arrays (local memory), face normals stored: the rare
//               ball-centre-inside-hull pairs, capsule pairs, and the last resort for a pair beyond the second tier.
//   EpaSmem     lane-strided words of SHARED memory (vertex points, packed topology, the heap, silhouette and flood stack); face
//               normals are RECOMPUTED from the three vertices whenever they are needed (same inputs, same operations, same bits as
//               when Face::new computed them), which is what makes the state small enough for 384 pairs per SM to be resident next
//               to each other.  Two sizes: EpaTier1 (16 / 48 / 24, k_cc_epa_s) for 99 % of the pairs, EpaTier2 (32 / 160 / 96,
//               k_cc_epa_t2) for those that outgrow it.  A pair that exceeds a capacity is flagged (`overflow`) and restarted on the
//               next tier: the restart repeats the same arithmetic, so results do not depend on which store ran.
// Both follow the reference's iteration path exactly (same heap sift order as Rust's BinaryHeap, same flood order, same exits);
// tests/host_shim compiles both for the host and compares them with the oracle bit for bit.
#pragma once

namespace ncb {

struct EpaHeapItem {
    uint32_t id;
    float neg_dist;
};
// per-lane scalars of a run (registers)
struct EpaScalars {
    int nverts, nfaces, nheap, nsil, niter;
    float max_dist;
    EpaHeapItem best_face_id;
    bool overflow, panicked;
#ifdef NCB_EPA_STATS
    int peak_heap, peak_sil, peak_stk;
#endif
};

// utils::ccw_face_normal (the zero vector stands for None, as Face::new stores it)
NCB_HD V3 epa_ccw_normal(V3 A, V3 B, V3 C) {
    V3 n;
    if (!unit_try_new(cross(B - A, C - A), NCB_EPS, n)) n = v3(0.f, 0.f, 0.f);
    return n;
}

#define EPA_MAX_VERTS 48
#define EPA_MAX_FACES 192
#define EPA_MAX_HEAP 160
#define EPA_MAX_STACK 128

// Per-thread polytope in local memory, split hot / cold: vertex CSO points, packed face topology (3 vertex ids + deleted flag |
// 3 neighbour ids), face normals and heap are touched by every expansion step; the original support points (orig1 / orig2) of
// each vertex are read once for the result.  Barycentric coordinates of a face are NOT stored: they are recomputed for the one
// face the result is read from.
struct EpaState : EpaScalars {
    enum { MAXV = EPA_MAX_VERTS, MAXF = EPA_MAX_FACES, MAXH = EPA_MAX_HEAP, MAXSIL = EPA_MAX_STACK, MAXSTK = EPA_MAX_STACK, ADJ_MASK = 0xff };
    static const bool STORED_NORMALS = true;
    V3 vpoint[MAXV];
    uint32_t ftopo[MAXF][2];  // [0] = pts0 | pts1 << 8 | pts2 << 16 | deleted << 24 ; [1] = adj0 | adj1 << 8 | adj2 << 16
    V3 fnormal[MAXF];
    float hdist[MAXH];
    uint8_t hid[MAXH];
    uint8_t sil_face[MAXSIL], sil_opp[MAXSIL];
    uint8_t stk_face[MAXSTK], stk_opp[MAXSTK];
    V3 vorig1[MAXV], vorig2[MAXV];

    NCB_HD V3 vp(uint32_t i) const { return vpoint[i]; }
    NCB_HD void set_vp(uint32_t i, V3 p) { vpoint[i] = p; }
    NCB_HD V3 o1(uint32_t i) const { return vorig1[i]; }
    NCB_HD V3 o2(uint32_t i) const { return vorig2[i]; }
    NCB_HD void set_orig(uint32_t i, V3 a, V3 b) { vorig1[i] = a, vorig2[i] = b; }
    NCB_HD void f_init(uint32_t f, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t a0, uint32_t a1, uint32_t a2) {
        ftopo[f][0] = p0 | (p1 << 8) | (p2 << 16);
        ftopo[f][1] = a0 | (a1 << 8) | (a2 << 16);
    }
    NCB_HD void f_pts(uint32_t f, uint32_t& p0, uint32_t& p1, uint32_t& p2) const {
        uint32_t t = ftopo[f][0];
        p0 = t & 0xffu, p1 = (t >> 8) & 0xffu, p2 = (t >> 16) & 0xffu;
    }
    NCB_HD void f_adjs(uint32_t f, uint32_t& a0, uint32_t& a1, uint32_t& a2) const {
        uint32_t t = ftopo[f][1];
        a0 = t & 0xffu, a1 = (t >> 8) & 0xffu, a2 = (t >> 16) & 0xffu;
    }
    NCB_HD bool f_deleted(uint32_t f) const { return (ftopo[f][0] >> 24) != 0; }
    NCB_HD void f_set_deleted(uint32_t f) { ftopo[f][0] |= 0x01000000u; }
    NCB_HD void f_set_adj(uint32_t f, uint32_t k, uint32_t v) { ftopo[f][1] = (ftopo[f][1] & ~(0xffu << (8 * k))) | (v << (8 * k)); }
    NCB_HD V3 stored_normal(uint32_t f) const { return fnormal[f]; }
    NCB_HD void set_normal(uint32_t f, V3 n) { fnormal[f] = n; }
    NCB_HD float hd(int i) const { return hdist[i]; }
    NCB_HD uint32_t hi(int i) const { return hid[i]; }
    NCB_HD void hset(int i, float d, uint32_t id) { hdist[i] = d, hid[i] = (uint8_t)id; }
    NCB_HD void sil_set(int k, uint32_t f, uint32_t opp) { sil_face[k] = (uint8_t)f, sil_opp[k] = (uint8_t)opp; }
    NCB_HD void sil_get(int k, uint32_t& f, uint32_t& opp) const { f = sil_face[k], opp = sil_opp[k]; }
    NCB_HD void stk_set(int k, uint32_t f, uint32_t opp) { stk_face[k] = (uint8_t)f, stk_opp[k] = (uint8_t)opp; }
    NCB_HD void stk_get(int k, uint32_t& f, uint32_t& opp) const { f = stk_face[k], opp = stk_opp[k]; }
};
static_assert(EPA_MAX_FACES <= 255 && EPA_MAX_VERTS <= 255, "ids are packed in 8 bits");

// The shared-memory store.  Word w of the lane lives at base[w * STRIDE] (STRIDE = threads per CTA, base = shared-memory array +
// thread index): whatever index a lane computes, its bank is its lane id, so the divergent accesses of an expansion step are
// conflict free; all offsets are compile-time constants.
//   WIDE = false (k_cc_epa_s, first tier):  one word per face = pts0 | pts1 << 4 | pts2 << 8 | adj0 << 12 | adj1 << 18 | adj2 << 24 |
//                                           deleted << 30 (<= 16 vertices, <= 64 faces); silhouette / stack entries in one byte.
//   WIDE = true  (k_cc_epa_t2, second tier): two words per face = [pts0 | pts1 << 8 | pts2 << 16 | deleted << 24] [adj0 | adj1 << 8 |
//                                           adj2 << 16]; silhouette / stack entries in 16 bits.
template <int STRIDE, int V, int F, int H, int SIL, int STK, bool WIDE>
struct EpaSmem : EpaScalars {
    enum { MAXV = V, MAXF = F, MAXH = H, MAXSIL = SIL, MAXSTK = STK, ADJ_MASK = WIDE ? 0xff : 63 };
    enum { EB = WIDE ? 2 : 1 };  // bytes per silhouette / stack entry
    enum {
        W_VP = 0,
        W_TP = W_VP + 3 * MAXV,
        W_HD = W_TP + (WIDE ? 2 : 1) * MAXF,
        W_HI = W_HD + MAXH,
        W_SIL = W_HI + (MAXH + 3) / 4,
        W_STK = W_SIL + (MAXSIL * EB + 3) / 4,
        WORDS = W_STK + (MAXSTK * EB + 3) / 4
    };
    static_assert(WIDE || (MAXV <= 16 && MAXF <= 64), "narrow face word: 4-bit vertex ids, 6-bit face ids");
    static_assert(MAXV <= 255 && MAXF <= 255, "8-bit ids");
    static const bool STORED_NORMALS = false;
    uint32_t* base;
    V3 vorig1[MAXV], vorig2[MAXV];  // cold: written once per vertex, three of each read for the result (local memory)

    NCB_HD uint32_t& w(int i) const { return base[i * STRIDE]; }
    NCB_HD uint8_t& b(int w0, int i) const { return reinterpret_cast<uint8_t*>(base + (w0 + (i >> 2)) * STRIDE)[i & 3]; }
    NCB_HD uint16_t& h(int w0, int i) const { return reinterpret_cast<uint16_t*>(base + (w0 + (i >> 1)) * STRIDE)[i & 1]; }
    NCB_HD V3 vp(uint32_t i) const {
        return v3(__uint_as_float(w(W_VP + 3 * i)), __uint_as_float(w(W_VP + 3 * i + 1)), __uint_as_float(w(W_VP + 3 * i + 2)));
    }
    NCB_HD void set_vp(uint32_t i, V3 p) {
        w(W_VP + 3 * i) = __float_as_uint(p.x), w(W_VP + 3 * i + 1) = __float_as_uint(p.y), w(W_VP + 3 * i + 2) = __float_as_uint(p.z);
    }
    NCB_HD V3 o1(uint32_t i) const { return vorig1[i]; }
    NCB_HD V3 o2(uint32_t i) const { return vorig2[i]; }
    NCB_HD void set_orig(uint32_t i, V3 a, V3 c) { vorig1[i] = a, vorig2[i] = c; }
    NCB_HD void f_init(uint32_t f, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t a0, uint32_t a1, uint32_t a2) {
        if (WIDE) {
            w(W_TP + 2 * f) = p0 | (p1 << 8) | (p2 << 16);
            w(W_TP + 2 * f + 1) = a0 | (a1 << 8) | (a2 << 16);
        } else {
            w(W_TP + f) = p0 | (p1 << 4) | (p2 << 8) | (a0 << 12) | (a1 << 18) | (a2 << 24);
        }
    }
    NCB_HD void f_pts(uint32_t f, uint32_t& p0, uint32_t& p1, uint32_t& p2) const {
        if (WIDE) {
            uint32_t t = w(W_TP + 2 * f);
            p0 = t & 0xffu, p1 = (t >> 8) & 0xffu, p2 = (t >> 16) & 0xffu;
        } else {
            uint32_t t = w(W_TP + f);
            p0 = t & 15u, p1 = (t >> 4) & 15u, p2 = (t >> 8) & 15u;
        }
    }
    NCB_HD void f_adjs(uint32_t f, uint32_t& a0, uint32_t& a1, uint32_t& a2) const {
        if (WIDE) {
            uint32_t t = w(W_TP + 2 * f + 1);
            a0 = t & 0xffu, a1 = (t >> 8) & 0xffu, a2 = (t >> 16) & 0xffu;
        } else {
            uint32_t t = w(W_TP + f);
            a0 = (t >> 12) & 63u, a1 = (t >> 18) & 63u, a2 = (t >> 24) & 63u;
        }
    }
    NCB_HD bool f_deleted(uint32_t f) const { return WIDE ? (w(W_TP + 2 * f) >> 24) != 0 : (w(W_TP + f) >> 30) != 0; }
    NCB_HD void f_set_deleted(uint32_t f) {
        if (WIDE)
            w(W_TP + 2 * f) |= 0x01000000u;
        else
            w(W_TP + f) |= 0x40000000u;
    }
    NCB_HD void f_set_adj(uint32_t f, uint32_t k, uint32_t v) {
        if (WIDE) {
            w(W_TP + 2 * f + 1) = (w(W_TP + 2 * f + 1) & ~(0xffu << (8 * k))) | (v << (8 * k));
        } else {
            uint32_t sh = 12 + 6 * k;
            w(W_TP + f) = (w(W_TP + f) & ~(63u << sh)) | (v << sh);
        }
    }
    NCB_HD V3 stored_normal(uint32_t) const { return v3(0.f, 0.f, 0.f); }
    NCB_HD void set_normal(uint32_t, V3) {}
    NCB_HD float hd(int i) const { return __uint_as_float(w(W_HD + i)); }
    NCB_HD uint32_t hi(int i) const { return b(W_HI, i); }
    NCB_HD void hset(int i, float d, uint32_t id) { w(W_HD + i) = __float_as_uint(d), b(W_HI, i) = (uint8_t)id; }
    NCB_HD void ent_set(int w0, int k, uint32_t f, uint32_t opp) {
        if (WIDE)
            h(w0, k) = (uint16_t)(f | (opp << 8));
        else
            b(w0, k) = (uint8_t)(f | (opp << 6));
    }
    NCB_HD void ent_get(int w0, int k, uint32_t& f, uint32_t& opp) const {
        if (WIDE) {
            uint32_t v = h(w0, k);
            f = v & 0xffu, opp = v >> 8;
        } else {
            uint32_t v = b(w0, k);
            f = v & 63u, opp = v >> 6;
        }
    }
    NCB_HD void sil_set(int k, uint32_t f, uint32_t opp) { ent_set(W_SIL, k, f, opp); }
    NCB_HD void sil_get(int k, uint32_t& f, uint32_t& opp) const { ent_get(W_SIL, k, f, opp); }
    NCB_HD void stk_set(int k, uint32_t f, uint32_t opp) { ent_set(W_STK, k, f, opp); }
    NCB_HD void stk_get(int k, uint32_t& f, uint32_t& opp) const { ent_get(W_STK, k, f, opp); }
};
// first tier: 99 % of cfg3's EPA runs (133 words = 532 B per pair; 6 CTAs of 64 threads per SM)
template <int STRIDE>
using EpaTier1 = EpaSmem<STRIDE, 16, 48, 24, 16, 12, false>;
// second tier: the pairs that outgrew the first one, restarted (584 words per pair; 3 CTAs of 32 threads per SM, one wave)
template <int STRIDE>
using EpaTier2 = EpaSmem<STRIDE, 32, 160, 96, 32, 32, true>;

NCB_HD V3 epa_sel3(uint32_t k, V3 a, V3 b, V3 c) { return k == 0 ? a : (k == 1 ? b : c); }
NCB_HD uint32_t epa_sel3(uint32_t k, uint32_t a, uint32_t b, uint32_t c) { return k == 0 ? a : (k == 1 ? b : c); }

// The three vertex points of a face and its normal (stored, or recomputed exactly as Face::new computed it).
template <class S>
NCB_HD V3 epa_face_normal(const S& e, uint32_t f) {
    if (S::STORED_NORMALS) return e.stored_normal(f);
    uint32_t i0, i1, i2;
    e.f_pts(f, i0, i1, i2);
    return epa_ccw_normal(e.vp(i0), e.vp(i1), e.vp(i2));
}

template <class S>
NCB_HD void epa_push_vertex(S& e, const CSOPoint& c) {
    e.set_vp(e.nverts, c.point);
    e.set_orig(e.nverts, c.orig1, c.orig2);
    e.nverts++;
}

// Rust std BinaryHeap<FaceId>: `<=` comes from partial_cmp on neg_dist.
template <class S>
NCB_HD void epa_heap_sift_up(S& e, int pos) {
    float ed = e.hd(pos);
    uint32_t ei = e.hi(pos);
    while (pos > 0) {
        int parent = (pos - 1) / 2;
        float pd = e.hd(parent);
        if (ed <= pd) break;
        e.hset(pos, pd, e.hi(parent));
        pos = parent;
    }
    e.hset(pos, ed, ei);
}
template <class S>
NCB_HD void epa_heap_push(S& e, uint32_t id, float nd) {
    if (e.nheap >= e.MAXH) {
        e.overflow = true;
        return;
    }
    e.hset(e.nheap, nd, id);
    e.nheap++;
#ifdef NCB_EPA_STATS
    if (e.nheap > e.peak_heap) e.peak_heap = e.nheap;
#endif
    epa_heap_sift_up(e, e.nheap - 1);
}
template <class S>
NCB_HD bool epa_heap_pop(S& e, EpaHeapItem& out) {
    if (e.nheap == 0) return false;
    --e.nheap;
    float item_d = e.hd(e.nheap);
    uint32_t item_i = e.hi(e.nheap);
    if (e.nheap > 0) {
        float td = e.hd(0);
        uint32_t ti = e.hi(0);
        // swap(item, data[0]); sift_down_to_bottom(0)
        int end = e.nheap, pos = 0, child = 1;
        float ed = item_d;
        uint32_t ei = item_i;
        item_d = td;
        item_i = ti;
        while (end >= 2 && child <= end - 2) {
            float cl = e.hd(child), cr = e.hd(child + 1);
            if (cl <= cr) child += 1, cl = cr;
            e.hset(pos, cl, e.hi(child));
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            e.hset(pos, e.hd(child), e.hi(child));
            pos = child;
        }
        e.hset(pos, ed, ei);
        epa_heap_sift_up(e, pos);
    }
    out.id = item_i;
    out.neg_dist = item_d;
    return true;
}

// Face::new (epa3.rs:93-114) for the triangle (A, B, C) = the points of (p0, p1, p2): normal + "projection of the origin lies
// inside the face".  false on overflow.
template <class S>
NCB_HD bool epa_face_new(S& e, uint32_t p0, uint32_t p1, uint32_t p2, V3 A, V3 B, V3 C, uint32_t a0, uint32_t a1, uint32_t a2,
                         bool& proj_inside, V3& normal) {
    if (e.nfaces >= e.MAXF) {
        e.overflow = true;
        return false;
    }
    Loc loc;
    proj_triangle_core<false>(A, B, C, v3(0.f, 0.f, 0.f), loc);
    int f = e.nfaces++;
    e.f_init(f, p0, p1, p2, a0, a1, a2);
    normal = epa_ccw_normal(A, B, C);
    e.set_normal(f, normal);
    proj_inside = loc.kind == LOC_FACE;
    return true;
}
// Face::closest_points (epa3.rs:116-126) with the barycentric coordinates recomputed as Face::new computed them.
template <class S>
NCB_HD void epa_face_closest_points(const S& e, uint32_t f, V3& p1, V3& p2) {
    uint32_t i0, i1, i2;
    e.f_pts(f, i0, i1, i2);
    Loc loc;
    proj_triangle_core<false>(e.vp(i0), e.vp(i1), e.vp(i2), v3(0.f, 0.f, 0.f), loc);
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (loc.kind == LOC_FACE) b0 = loc.b0, b1 = loc.b1, b2 = loc.b2;
    p1 = e.o1(i0) * b0 + e.o1(i1) * b1 + e.o1(i2) * b2;
    p2 = e.o2(i0) * b0 + e.o2(i1) * b1 + e.o2(i2) * b2;
}
template <class S>
NCB_HD uint32_t epa_next_ccw(S& e, uint32_t f, uint32_t id) {
    uint32_t i0, i1, i2;
    e.f_pts(f, i0, i1, i2);
    if (i0 == id) return 1;
    if (i1 == id) return 2;
    if (i2 != id) e.panicked = true;  // assert_eq! in the reference
    return 0;
}
// Face::can_be_seen_by (epa3.rs:140-155); `pt` is the point of the vertex `point`
template <class S>
NCB_HD bool epa_can_be_seen_by(const S& e, uint32_t f, V3 pt, uint32_t opp) {
    uint32_t i0, i1, i2;
    e.f_pts(f, i0, i1, i2);
    V3 p0, p1, p2, n;
    if (S::STORED_NORMALS) {
        p0 = e.vp(epa_sel3(opp, i0, i1, i2));
        n = e.stored_normal(f);
        if (dot(pt - p0, n) >= -(NCB_EPS * 10.0f)) return true;
        p1 = e.vp(epa_sel3((opp + 1) % 3, i0, i1, i2)), p2 = e.vp(epa_sel3((opp + 2) % 3, i0, i1, i2));
    } else {
        V3 A = e.vp(i0), B = e.vp(i1), C = e.vp(i2);
        n = epa_ccw_normal(A, B, C);
        p0 = epa_sel3(opp, A, B, C);
        if (dot(pt - p0, n) >= -(NCB_EPS * 10.0f)) return true;
        p1 = epa_sel3((opp + 1) % 3, A, B, C), p2 = epa_sel3((opp + 2) % 3, A, B, C);
    }
    // utils::is_affinely_dependent_triangle(p1, p2, pt)
    V3 p1p2 = p2 - p1, p1p3 = pt - p1;
    float eps_tol = NCB_EPS * 100.0f;
    return relative_eq(norm_squared(cross(p1p2, p1p3)), 0.f, eps_tol * eps_tol);
}
// compute_silhouette (epa3.rs:432-454): the recursion becomes a LIFO of (face, opp) visits in the same order.
template <class S>
NCB_HD void epa_compute_silhouette3(S& e, V3 pt, uint32_t id0, uint32_t opp0, uint32_t id1, uint32_t opp1, uint32_t id2, uint32_t opp2) {
    int sp = 0;
    e.stk_set(sp++, id2, opp2);
    e.stk_set(sp++, id1, opp1);
    e.stk_set(sp++, id0, opp0);
    while (sp > 0) {
        sp--;
        uint32_t id, opp;
        e.stk_get(sp, id, opp);
        if (e.f_deleted(id)) continue;
        if (!epa_can_be_seen_by(e, id, pt, opp)) {
            if (e.nsil >= e.MAXSIL) {
                e.overflow = true;
                return;
            }
            e.sil_set(e.nsil, id, opp);
            e.nsil++;
#ifdef NCB_EPA_STATS
            if (e.nsil > e.peak_sil) e.peak_sil = e.nsil;
#endif
        } else {
            e.f_set_deleted(id);
            uint32_t adj_pt_id1 = (opp + 2) % 3, adj_pt_id2 = opp;
            uint32_t i0, i1, i2, a0, a1, a2;
            e.f_pts(id, i0, i1, i2);
            e.f_adjs(id, a0, a1, a2);
            uint32_t adj1 = epa_sel3(adj_pt_id1, a0, a1, a2), adj2 = epa_sel3(adj_pt_id2, a0, a1, a2);
            uint32_t o1 = epa_next_ccw(e, adj1, epa_sel3(adj_pt_id1, i0, i1, i2));
            uint32_t o2 = epa_next_ccw(e, adj2, epa_sel3(adj_pt_id2, i0, i1, i2));
            if (e.panicked) return;
            if (sp + 2 > e.MAXSTK) {
                e.overflow = true;
                return;
            }
            // visit adj1 first, then adj2
            e.stk_set(sp++, adj2, o2);
            e.stk_set(sp++, adj1, o1);
#ifdef NCB_EPA_STATS
            if (sp > e.peak_stk) e.peak_stk = sp;
#endif
        }
    }
}

// EPA_DONE_OK: the result is read from face `res_face` (epa_result_from_face), or, for EPA_RES_DIRECT, was written directly.
// The three Ok exits of the reference's loop all read one face (the best one, or the popped one as it was cloned), so the
// read-out exists once per kernel instead of once per exit.
enum { EPA_CONTINUE = 0, EPA_DONE_OK = 1, EPA_DONE_FAIL = 2 };
#define EPA_RES_DIRECT 0xffffffffu

#define NCB_EPA_PUSH(ID, ND)                              \
    {                                                     \
        float nd__ = (ND);                                \
        if (nd__ > NCB_EPS * 10.0f) return EPA_DONE_FAIL; \
        epa_heap_push(e, (ID), nd__);                     \
    }

// The result read from a face: closest points + the face's normal.
template <class S>
NCB_HD void epa_result_from_face(const S& e, uint32_t f, V3& out1, V3& out2, V3& out_n) {
    epa_face_closest_points(e, f, out1, out2);
    out_n = epa_face_normal(e, f);
}

// EPA::closest_points, part 1 (epa3.rs:219-328): initial polytope from the GJK simplex.  The reference builds all faces and then
// pushes them on the heap in face order; here face k is pushed right after it is built (building a face does not read the heap
// and a push does not read the faces, and a failing FaceId::new ends the run whichever faces exist), so one face constructor in
// a loop serves both the tetrahedron (4 faces) and the flat (2 faces) start.
// NO_SEGMENT: a segment simplex (its third vertex needs one more support evaluation; never seen on cfg3, where 0.007 % of the EPA
// pairs start from a triangle and the rest from a tetrahedron) is handed to the caller's overflow path instead of being built here.
#define EPA_FACE_SPEC(p0, p1, p2, a0, a1, a2) ((p0) | ((p1) << 4) | ((p2) << 8) | ((a0) << 12) | ((a1) << 16) | ((a2) << 20))
template <bool NO_SEGMENT, class S, class G>
NCB_HD int epa_init_t(S& e, const Iso& m1, const G& g1, const Iso& m2, const G& g2, int sdim, const CSOPoint* sv, V3& out1, V3& out2,
                      V3& out_n, uint32_t& res_face) {
    res_face = EPA_RES_DIRECT;
    e.nverts = e.nfaces = e.nheap = e.nsil = 0;
    e.niter = 0;
    e.overflow = false;
    e.panicked = false;
#ifdef NCB_EPA_STATS
    e.peak_heap = e.peak_sil = e.peak_stk = 0;
#endif
    if (sdim == 0) {
        out1 = v3(0.f, 0.f, 0.f);
        out2 = v3(0.f, 0.f, 0.f);
        out_n = v3(0.f, 1.f, 0.f);
        return EPA_DONE_OK;
    }
    if (sdim == 3) {
        CSOPoint c0 = sv[0], c1 = sv[1], c2 = sv[2], c3 = sv[3];
        V3 dp1 = c1.point - c0.point;
        V3 dp2 = c2.point - c0.point;
        V3 dp3 = c3.point - c0.point;
        bool flip = dot(cross(dp1, dp2), dp3) > 0.f;
        epa_push_vertex(e, c0), epa_push_vertex(e, flip ? c2 : c1), epa_push_vertex(e, flip ? c1 : c2), epa_push_vertex(e, c3);
    } else {
        CSOPoint c0 = sv[0], c1 = sv[1], c2 = sv[2];
        if (sdim == 1) {
            if constexpr (NO_SEGMENT) {
                e.overflow = true;
                return EPA_DONE_FAIL;
            } else {
                V3 dpt = c1.point - c0.point;
                V3 first, second;
                orthonormal_basis(dpt, first, second);
                c2 = cso_from_shapes(m1, g1, m2, g2, first);
            }
        }
        epa_push_vertex(e, c0), epa_push_vertex(e, c1), epa_push_vertex(e, c2);
    }
    const bool tetra = sdim == 3;
    const int nf = tetra ? 4 : 2;
#pragma unroll 1
    for (int k = 0; k < nf; ++k) {
        uint32_t spec;
        if (tetra)
            spec = k == 0 ? EPA_FACE_SPEC(0, 1, 2, 3, 1, 2)
                          : (k == 1 ? EPA_FACE_SPEC(1, 3, 2, 3, 2, 0) : (k == 2 ? EPA_FACE_SPEC(0, 2, 3, 0, 1, 3) : EPA_FACE_SPEC(0, 3, 1, 2, 1, 0)));
        else
            spec = k == 0 ? EPA_FACE_SPEC(0, 1, 2, 1, 1, 1) : EPA_FACE_SPEC(0, 2, 1, 0, 0, 0);
        uint32_t p0 = spec & 15u, p1 = (spec >> 4) & 15u, p2 = (spec >> 8) & 15u;
        bool in;
        V3 n;
        epa_face_new(e, p0, p1, p2, e.vp(p0), e.vp(p1), e.vp(p2), (spec >> 12) & 15u, (spec >> 16) & 15u, (spec >> 20) & 15u, in, n);
        if (tetra) {
            if (in) NCB_EPA_PUSH((uint32_t)k, -dot(n, e.vp((uint32_t)k)));  // dist_k = normal_k . vertices[k]
        } else {
            NCB_EPA_PUSH((uint32_t)k, 0.f);
        }
    }
    e.max_dist = NCB_FMAX;
    if (e.nheap == 0) {  // heap.peek().unwrap() panics in the reference
        e.panicked = true;
        return EPA_DONE_FAIL;
    }
    e.best_face_id.id = e.hi(0);
    e.best_face_id.neg_dist = e.hd(0);
    return EPA_CONTINUE;
}

// EPA::closest_points, part 2: ONE turn of `while let Some(face_id) = self.heap.pop()` (epa3.rs:330-425).
template <class S, class G>
NCB_HD int epa_step_t(S& e, const Iso& m1, const G& g1, const Iso& m2, const G& g2, uint32_t& res_face) {
    const float eps_tol = NCB_EPS * 100.0f;
    EpaHeapItem face_id;
    // `if face.deleted { continue; }` (epa3.rs:334-336): stale heap entries are skipped inside the same turn
    do {
        if (!epa_heap_pop(e, face_id)) {  // heap exhausted: the best face so far (epa3.rs:427-429)
            res_face = e.best_face_id.id;
            return EPA_DONE_OK;
        }
    } while (e.f_deleted(face_id.id));
    uint32_t fid = face_id.id;
    // snapshot of the popped face (the reference clones it before the polytope is edited)
    uint32_t fp0, fp1, fp2, fa0, fa1, fa2;
    e.f_pts(fid, fp0, fp1, fp2);
    e.f_adjs(fid, fa0, fa1, fa2);
    V3 fnorm = epa_face_normal(e, fid);
    if (e.nverts >= e.MAXV) {
        e.overflow = true;
        return EPA_DONE_FAIL;
    }
    CSOPoint cso = cso_from_shapes(m1, g1, m2, g2, fnorm);
    uint32_t support_point_id = (uint32_t)e.nverts;
    epa_push_vertex(e, cso);
    float candidate_max_dist = dot(cso.point, fnorm);
    if (candidate_max_dist < e.max_dist) {
        e.best_face_id = face_id;
        e.max_dist = candidate_max_dist;
    }
    float curr_dist = -face_id.neg_dist;
    if (e.max_dist - curr_dist < eps_tol) {
        res_face = e.best_face_id.id;
        return EPA_DONE_OK;
    }
    e.f_set_deleted(fid);
    uint32_t o1 = epa_next_ccw(e, fa0, fp0);
    uint32_t o2 = epa_next_ccw(e, fa1, fp1);
    uint32_t o3 = epa_next_ccw(e, fa2, fp2);
    if (e.panicked) return EPA_DONE_FAIL;
    // compute_silhouette x3 (epa3.rs:364-366) as ONE LIFO walk: the three roots are stacked in reverse order, so the
    // flood from adj[0] completes before adj[1] is looked at, exactly like the three sequential recursive calls
    epa_compute_silhouette3(e, cso.point, fa0, o1, fa1, o2, fa2, o3);
    if (e.panicked || e.overflow) return EPA_DONE_FAIL;
    uint32_t first_new_face_id = (uint32_t)e.nfaces;
    if (e.nsil == 0) return EPA_DONE_FAIL;
    for (int k = 0; k < e.nsil; ++k) {
        uint32_t efid, eopp;
        e.sil_get(k, efid, eopp);
        if (!e.f_deleted(efid)) {
            uint32_t new_face_id = (uint32_t)e.nfaces;
            uint32_t i0, i1, i2;
            e.f_pts(efid, i0, i1, i2);
            uint32_t pt_id1 = epa_sel3((eopp + 2) % 3, i0, i1, i2);
            uint32_t pt_id2 = epa_sel3((eopp + 1) % 3, i0, i1, i2);
            V3 A = e.vp(pt_id1), B = e.vp(pt_id2);
            bool inside;
            V3 nn;
            // adj = [edge.face_id, new_face_id + 1, new_face_id - 1] (the last two are patched below for the ends)
            if (!epa_face_new(e, pt_id1, pt_id2, support_point_id, A, B, cso.point, efid, (new_face_id + 1) & S::ADJ_MASK,
                              (new_face_id - 1) & S::ADJ_MASK, inside, nn))
                return EPA_DONE_FAIL;
            e.f_set_adj(efid, (eopp + 1) % 3, new_face_id);
            if (inside) {
                float dist = dot(nn, A);
                if (dist < curr_dist) {
                    // the popped face as it was when cloned (epa3.rs:393-398): its vertex ids are unchanged (so are its closest
                    // points and its normal), the deleted flag is not read by the read-out
                    res_face = fid;
                    return EPA_DONE_OK;
                }
                NCB_EPA_PUSH(new_face_id, -dist);  // FaceId::new(new_face_id, -dist)? then heap.push
                if (e.overflow) return EPA_DONE_FAIL;
            }
        }
    }
    if (first_new_face_id == (uint32_t)e.nfaces) return EPA_DONE_FAIL;
    e.f_set_adj(first_new_face_id, 2, (uint32_t)(e.nfaces - 1));
    e.f_set_adj((uint32_t)(e.nfaces - 1), 1, first_new_face_id);
    e.nsil = 0;
    e.niter += 1;
    if (e.niter > 10000) return EPA_DONE_FAIL;
    return EPA_CONTINUE;
}
#undef NCB_EPA_PUSH

// Out-of-line instances for the local-memory store (the kernels that use it call these per lane)
static __device__ __noinline__ int epa_init(EpaState& e, const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, int sdim,
                                            const CSOPoint* sv, V3& out1, V3& out2, V3& out_n) {
    uint32_t res_face;
    return epa_init_t<false>(e, m1, g1, m2, g2, sdim, sv, out1, out2, out_n, res_face);
}
static __device__ __noinline__ int epa_step(EpaState& e, const Iso& m1, const Support& g1, const Iso& m2, const Support& g2, V3& out1, V3& out2,
                                            V3& out_n) {
    uint32_t res_face;
    int st = epa_step_t(e, m1, g1, m2, g2, res_face);
    if (st == EPA_DONE_OK) epa_result_from_face(e, res_face, out1, out2, out_n);
    return st;
}

// EPA::closest_points (epa3.rs:219-430).  false = None (also on capacity overflow, flagged in e.overflow).
static __device__ __noinline__ bool epa_closest_points(EpaState& e, const Iso& m1, const Support& g1, const Iso& m2, const Support& g2,
                                                       int sdim, const CSOPoint* sv, V3& out1, V3& out2, V3& out_n) {
    int st = epa_init(e, m1, g1, m2, g2, sdim, sv, out1, out2, out_n);
    while (st == EPA_CONTINUE) st = epa_step(e, m1, g1, m2, g2, out1, out2, out_n);
    return st == EPA_DONE_OK;
}

// contact_support_map_support_map_with_params (init_dir = None: fresh generator).
// Returns GJK_CLOSEST_POINTS / GJK_NO_INTERSECTION.
static __device__ __noinline__ int contact_sm_sm(EpaState& e, const Iso& m1, const Support& g1, const Iso& m2, const Support& g2,
                                                 float prediction, V3& p1, V3& p2, V3& dir_out, uint32_t* epa_overflow, uint32_t* ref_panics) {
    V3 dir;
    if (!unit_try_new(m2.t - m1.t, NCB_EPS, dir)) dir = v3(1.f, 0.f, 0.f);
    Simplex s;
    int r = gjk_closest_points(m1, g1, m2, g2, prediction, dir, s, p1, p2, dir_out);
    if (r != GJK_INTERSECTION) return r;
    if (epa_closest_points(e, m1, g1, m2, g2, s.dim, s.v, p1, p2, dir_out)) return GJK_CLOSEST_POINTS;
    if (e.overflow) atomicAdd(epa_overflow, 1u);
    if (e.panicked) atomicAdd(ref_panics, 1u);
    dir_out = v3(1.f, 0.f, 0.f);
    return GJK_NO_INTERSECTION;
}

}  // namespace ncb
