// TEST INFRASTRUCTURE ONLY — compiles the per-pair DEVICE functions of ncollide_b200/csrc/dim2.cu (2-D GJK / EPA, the ball queries,
// 2-D features and clipping, the manifold, the AABBs) for the host through the stand-in cuda_runtime.h and runs them in the call
// order of the kernels k_contact2d / k_aabb2d / k_narrow2d, so that the CPU suite can hold the device source bit for bit against
// the oracle (tests/test_dim2.py).  Nothing in the product includes this file.
#include "dim2.cu"

using namespace ncb;
using namespace ncb::d2;

extern "C" {

void shim2_contact(uint64_t n, const uint32_t* type1, const float* param1, const float* pose1, const uint32_t* type2, const float* param2,
                   const float* pose2, const float* poly, const float* poly_nrm, float prediction, uint8_t* found, float* out, uint32_t* flag_counts) {
    const float c1 = cosf((float)(3.14159265358979323846 / 180.0));
    flag_counts[0] = flag_counts[1] = 0;
    for (uint64_t k = 0; k < n; ++k) {
        const float4* p1 = reinterpret_cast<const float4*>(param1) + k;
        const float4* p2 = reinterpret_cast<const float4*>(param2) + k;
        const float4* m1 = reinterpret_cast<const float4*>(pose1) + k;
        const float4* m2 = reinterpret_cast<const float4*>(pose2) + k;
        Operand2 g1 = load_operand(type1[k], *p1, *m1, poly, poly_nrm), g2 = load_operand(type2[k], *p2, *m2, poly, poly_nrm);
        Hit2 h;
        int flags = 0;
        bool ok = contact_of_pair(g1, g2, prediction, c1, h, flags);
        flag_counts[0] += flags & 1, flag_counts[1] += (flags >> 1) & 1;
        found[k] = ok ? 1 : 0;
        float* o = out + 7 * k;
        o[0] = h.w1.x, o[1] = h.w1.y, o[2] = h.w2.x, o[3] = h.w2.y, o[4] = h.n.x, o[5] = h.n.y, o[6] = h.depth;
    }
}

void shim2_proximity(uint64_t n, const uint32_t* type1, const float* param1, const float* pose1, const uint32_t* type2, const float* param2,
                     const float* pose2, const float* poly, const float* margins, uint8_t* out) {
    for (uint64_t k = 0; k < n; ++k) {
        const float4* p1 = reinterpret_cast<const float4*>(param1) + k;
        const float4* p2 = reinterpret_cast<const float4*>(param2) + k;
        const float4* m1 = reinterpret_cast<const float4*>(pose1) + k;
        const float4* m2 = reinterpret_cast<const float4*>(pose2) + k;
        out[k] = proximity_of_pair(load_operand(type1[k], *p1, *m1, poly, nullptr), load_operand(type2[k], *p2, *m2, poly, nullptr), margins[k]);
    }
}

void shim2_ray_cast(uint64_t n, const uint32_t* type, const float* param, const float* pose, const float* poly, const float* rays, uint8_t* found,
                    float* out, uint32_t* feature) {
    for (uint64_t k = 0; k < n; ++k) {
        const float4* p = reinterpret_cast<const float4*>(param) + k;
        const float4* m = reinterpret_cast<const float4*>(pose) + k;
        const float* q = rays + 5 * k;
        RayHit2 h = shape_ray_cast2(load_operand(type[k], *p, *m, poly, nullptr), w2(q[0], q[1]), w2(q[2], q[3]), q[4]);
        found[k] = h.hit ? 1 : 0;
        out[3 * k] = h.toi, out[3 * k + 1] = h.n.x, out[3 * k + 2] = h.n.y;
        feature[k] = h.hit ? h.feature : 0xffffffffu;
    }
}

void shim2_contains_point(uint64_t n, const uint32_t* type, const float* param, const float* pose, const float* poly, const float* pts,
                          uint8_t* out) {
    for (uint64_t k = 0; k < n; ++k) {
        const float4* p = reinterpret_cast<const float4*>(param) + k;
        const float4* m = reinterpret_cast<const float4*>(pose) + k;
        out[k] = shape_contains_point2(load_operand(type[k], *p, *m, poly, nullptr), w2(pts[2 * k], pts[2 * k + 1])) ? 1 : 0;
    }
}

static Operand2 obj(uint32_t i, const float* pos, const float* rot, const uint32_t* type, const float* param, const float* poly, const float* nrm) {
    float4 p = reinterpret_cast<const float4*>(param)[i];
    return load_operand(type[i], p, make_float4(pos[2 * i], pos[2 * i + 1], rot[2 * i], rot[2 * i + 1]), poly, nrm);
}

// k_aabb2d: fat boxes as (lo.x, lo.y, 0, hi.x, hi.y, 0)
void shim2_aabbs(uint32_t n, const float* pos, const float* rot, const uint32_t* type, const float* param, const float* qlimit, const float* poly,
                 const float* nrm, float margin, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        W2 lo, hi;
        aabb_of_shape(obj(i, pos, rot, type, param, poly, nrm), lo, hi);
        float ql = qlimit[i];
        out[6 * i] = (lo.x + (-ql)) + (-margin), out[6 * i + 1] = (lo.y + (-ql)) + (-margin), out[6 * i + 2] = 0.f;
        out[6 * i + 3] = (hi.x + ql) + margin, out[6 * i + 4] = (hi.y + ql) + margin, out[6 * i + 5] = 0.f;
    }
}

// k_narrow2d over given pairs: manifold_off[P + 1], contacts (7 floats), features (2 words); returns the number of contacts
// qkind / prox may be NULL (no sensors): k_narrow2d's body per pair
uint64_t shim2_narrow_sensors(uint32_t n, const float* pos, const float* rot, const uint32_t* type, const float* param, const float* qlimit,
                              const float* ang_pred, const float* poly, const float* nrm, uint64_t n_pairs, const uint32_t* pairs,
                              uint32_t* manifold_off, float* contacts, uint32_t* feats, uint64_t cap, uint32_t* flag_counts, const uint8_t* qkind,
                              uint8_t* prox);
uint64_t shim2_narrow(uint32_t n, const float* pos, const float* rot, const uint32_t* type, const float* param, const float* qlimit,
                      const float* ang_pred, const float* poly, const float* nrm, uint64_t n_pairs, const uint32_t* pairs, uint32_t* manifold_off,
                      float* contacts, uint32_t* feats, uint64_t cap, uint32_t* flag_counts) {
    return shim2_narrow_sensors(n, pos, rot, type, param, qlimit, ang_pred, poly, nrm, n_pairs, pairs, manifold_off, contacts, feats, cap, flag_counts,
                                nullptr, nullptr);
}
uint64_t shim2_narrow_sensors(uint32_t n, const float* pos, const float* rot, const uint32_t* type, const float* param, const float* qlimit,
                              const float* ang_pred, const float* poly, const float* nrm, uint64_t n_pairs, const uint32_t* pairs,
                              uint32_t* manifold_off, float* contacts, uint32_t* feats, uint64_t cap, uint32_t* flag_counts, const uint8_t* qkind,
                              uint8_t* prox) {
    (void)n;
    const float c1 = cosf((float)(3.14159265358979323846 / 180.0));
    uint64_t nc = 0;
    flag_counts[0] = flag_counts[1] = flag_counts[2] = 0;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        Manifold2d mf;
        int flags = 0;
        Operand2 g1 = obj(i1, pos, rot, type, param, poly, nrm), g2 = obj(i2, pos, rot, type, param, poly, nrm);
        if (qkind && (qkind[i1] | qkind[i2])) {
            mf.n = 0, mf.overflow = false;
            prox[p] = (g1.kind == D2_PLANE && g2.kind == D2_PLANE) ? (uint8_t)NCB_PROXIMITY_NONE : proximity_of_pair(g1, g2, qlimit[i1] + qlimit[i2]);
        } else {
            if (prox) prox[p] = (uint8_t)NCB_PROXIMITY_NONE;
            manifold_of_pair(g1, g2, qlimit[i1] + qlimit[i2], cosf(ang_pred[i1]), cosf(ang_pred[i2]), sinf(ang_pred[i1]), sinf(ang_pred[i2]), c1, mf, flags);
        }
        flag_counts[0] += flags & 1, flag_counts[1] += (flags >> 1) & 1, flag_counts[2] += mf.overflow ? 1 : 0;
        manifold_off[p] = (uint32_t)nc;
        for (int k = 0; k < mf.n; ++k, ++nc) {
            if (nc >= cap) continue;
            float* o = contacts + 7 * nc;
            const Hit2& c = mf.c[k];
            o[0] = c.w1.x, o[1] = c.w1.y, o[2] = c.w2.x, o[3] = c.w2.y, o[4] = c.n.x, o[5] = c.n.y, o[6] = c.depth;
            feats[2 * nc] = mf.f1[k], feats[2 * nc + 1] = mf.f2[k];
        }
    }
    manifold_off[n_pairs] = (uint32_t)nc;
    return nc;
}

}  // extern "C"
