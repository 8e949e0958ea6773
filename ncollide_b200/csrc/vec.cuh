// Device-side f32 vector / isometry arithmetic in the evaluation order the reference gets from nalgebra 0.30
// (UnitQuaternion * Vector3, Isometry3 * Point3, Matrix3.abs() * v, 3-vector dot = (a0*b0 + a1*b1) + a2*b2 ...),
// see the call sites cited in DESIGN.md §"Arithmetic contract".  The translation unit is compiled with
// --fmad=false: Rust never contracts a*b+c, and bit-exact AABBs / ray hits depend on it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ncb {

#define NCB_EPS 1.1920929e-7f
#define NCB_FMAX 3.402823466e+38f

struct V3 {
    float x, y, z;
};
struct V2 {
    float x, y;
};
struct Quat {
    float i, j, k, w;
};
struct Iso {
    V3 t;
    Quat q;
};

#define NCB_HD __device__ __forceinline__

NCB_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
NCB_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
NCB_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
NCB_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
NCB_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
NCB_HD V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
NCB_HD float vget(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
NCB_HD void vset(V3& a, int i, float v) {
    if (i == 0)
        a.x = v;
    else if (i == 1)
        a.y = v;
    else
        a.z = v;
}
NCB_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
NCB_HD float norm_squared(V3 a) { return dot(a, a); }
NCB_HD float norm(V3 a) { return sqrtf(norm_squared(a)); }
NCB_HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
NCB_HD V3 normalize(V3 a) { return a / norm(a); }
NCB_HD bool try_normalize(V3 a, float min_norm, V3& out) {
    float n = norm(a);
    if (n <= min_norm) return false;
    out = a / n;
    return true;
}
// Unit::try_new_and_get: Some iff norm_squared > min_norm^2
NCB_HD bool unit_try_new_and_get(V3 a, float min_norm, V3& out, float& n_out) {
    float sq = norm_squared(a);
    if (sq > min_norm * min_norm) {
        float n = sqrtf(sq);
        out = a / n;
        n_out = n;
        return true;
    }
    return false;
}
NCB_HD bool unit_try_new(V3 a, float min_norm, V3& out) {
    float n;
    return unit_try_new_and_get(a, min_norm, out, n);
}
NCB_HD V3 vmin(V3 a, V3 b) { return V3{fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
NCB_HD V3 vmax(V3 a, V3 b) { return V3{fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }

NCB_HD V3 quat_rotate(Quat q, V3 v) {
    V3 qv = V3{q.i, q.j, q.k};
    V3 t = cross(qv, v) * 2.0f;
    V3 c = cross(qv, t);
    return (t * q.w + c) + v;
}
NCB_HD Quat quat_conj(Quat q) { return Quat{-q.i, -q.j, -q.k, q.w}; }
NCB_HD V3 iso_mul_point(const Iso& m, V3 p) { return quat_rotate(m.q, p) + m.t; }
NCB_HD V3 iso_mul_vec(const Iso& m, V3 v) { return quat_rotate(m.q, v); }
NCB_HD V3 iso_inv_point(const Iso& m, V3 p) { return quat_rotate(quat_conj(m.q), p - m.t); }
NCB_HD V3 iso_inv_vec(const Iso& m, V3 v) { return quat_rotate(quat_conj(m.q), v); }

// rotation.to_rotation_matrix().into_inner().abs() * v
NCB_HD V3 absolute_transform_vector(Quat q, V3 v) {
    float i = q.i, j = q.j, k = q.k, w = q.w;
    float ww = w * w, ii = i * i, jj = j * j, kk = k * k;
    float ij = i * j * 2.0f, wk = w * k * 2.0f, wj = w * j * 2.0f;
    float ik = i * k * 2.0f, jk = j * k * 2.0f, wi = w * i * 2.0f;
    float m00 = fabsf(ww + ii - (jj + kk)), m01 = fabsf(ij - wk), m02 = fabsf(wj + ik);
    float m10 = fabsf(wk + ij), m11 = fabsf(ww - ii + jj - kk), m12 = fabsf(jk - wi);
    float m20 = fabsf(ik - wj), m21 = fabsf(wi + jk), m22 = fabsf(ww - (ii + jj) + kk);
    return V3{(m00 * v.x + m01 * v.y) + m02 * v.z, (m10 * v.x + m11 * v.y) + m12 * v.z, (m20 * v.x + m21 * v.y) + m22 * v.z};
}

NCB_HD bool relative_eq(float a, float b, float epsilon = NCB_EPS, float max_relative = NCB_EPS) {
    if (a == b) return true;
    if (isinf(a) || isinf(b)) return false;
    float abs_diff = fabsf(a - b);
    if (abs_diff <= epsilon) return true;
    float aa = fabsf(a), ab = fabsf(b);
    float largest = ab > aa ? ab : aa;
    return abs_diff <= largest * max_relative;
}
NCB_HD bool ulps_eq(float a, float b) {
    if (fabsf(a - b) <= NCB_EPS) return true;
    if (signbit(a) != signbit(b)) return false;
    int ia = __float_as_int(a), ib = __float_as_int(b);
    long long d = (long long)ia - (long long)ib;
    if (d < 0) d = -d;
    return d <= 4;
}
NCB_HD float clampf(float v, float lo, float hi) { return v > lo ? (v < hi ? v : hi) : lo; }
NCB_HD float signumf(float x) {
    if (isnan(x)) return x;
    return signbit(x) ? -1.0f : 1.0f;
}
NCB_HD void orthonormal_basis(V3 v, V3& first, V3& second) {
    V3 a;
    if (fabsf(v.x) > fabsf(v.y))
        a = V3{v.z, 0.0f, -v.x};
    else
        a = V3{0.0f, -v.z, v.y};
    a = normalize(a);
    first = cross(a, v);
    second = a;
}

// Capsule as a SupportMap (capsule.rs:72-85): support_point(m, dir) = m * local_support_point_toward(normalize(m^-1 dir)),
// hh = half height along local y.
NCB_HD V3 capsule_support_point(const Iso& m, float hh, float radius, V3 dir) {
    V3 d = normalize(iso_inv_vec(m, dir));
    return iso_mul_point(m, v3(0.f, copysignf(hh, d.y), 0.f) + d * radius);
}
// AABB of a capsule (aabb_support_map.rs:35-45 -> aabb_utils.rs:9-31): the support points along +-x, +-y, +-z
NCB_HD void capsule_aabb(const Iso& m, float hh, float radius, V3& mins, V3& maxs) {
    maxs = v3(capsule_support_point(m, hh, radius, v3(1.f, 0.f, 0.f)).x, capsule_support_point(m, hh, radius, v3(0.f, 1.f, 0.f)).y,
              capsule_support_point(m, hh, radius, v3(0.f, 0.f, 1.f)).z);
    mins = v3(capsule_support_point(m, hh, radius, v3(-1.f, 0.f, 0.f)).x, capsule_support_point(m, hh, radius, v3(0.f, -1.f, 0.f)).y,
              capsule_support_point(m, hh, radius, v3(0.f, 0.f, -1.f)).z);
}

}  // namespace ncb
