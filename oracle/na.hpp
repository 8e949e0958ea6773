// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or executed by
// the product path (ncollide_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / CPU baseline.
//
// Restatement of the nalgebra 0.30 / approx 0.5 arithmetic that the reference's hot path calls.
// nalgebra / simba / approx are third-party crates that are NOT vendored under /root/reference
// (build/ncollide3d/Cargo.toml:41-43 pins only the caret ranges "0.30" / "0.7" / "0.5"; no Cargo.lock),
// and no Rust toolchain exists in this image, so the published algorithms are restated here and
// parity is anchored on the reference's call sites.  PARITY UNPINNED at the nalgebra boundary for
// everything the reference's own KATs do not cover.
//
// Rust never contracts a*b+c into an FMA: build with -ffp-contract=off, no -ffast-math.
#pragma once
#include <cmath>
#include <cstdint>
#include <cfloat>
#include <cstring>

#include <limits>
#ifndef ORC_REAL
#define ORC_REAL float
#endif
typedef ORC_REAL real;  // f32 on the path; the f64 build exists only for the reference's f64 known-answer tests

namespace orc {

static const real EPS = std::numeric_limits<real>::epsilon();  // N::default_epsilon()
static const real FMAX = std::numeric_limits<real>::max();     // N::max_value()

struct V3 {
    real x, y, z;
    real& operator[](int i) { return (&x)[i]; }
    real operator[](int i) const { return (&x)[i]; }
};
struct V2 {
    real x, y;
};

static inline V3 v3(real x, real y, real z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, real s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator/(V3 a, real s) { return {a.x / s, a.y / s, a.z / s}; }
static inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// nalgebra Matrix::dot, 3x1 special case: (a0*b0 + a1*b1) + a2*b2
static inline real dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline real dot2(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline real perp2(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
static inline V2 sub2(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
static inline real norm_squared(V3 a) { return dot(a, a); }
static inline real norm(V3 a) { return std::sqrt(norm_squared(a)); }
static inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Matrix::normalize(): self.unscale(self.norm())  (component-wise division)
static inline V3 normalize(V3 a) { return a / norm(a); }
// Matrix::try_normalize(min_norm): None iff norm <= min_norm
static inline bool try_normalize(V3 a, real min_norm, V3* out) {
    real n = norm(a);
    if (n <= min_norm) return false;
    *out = a / n;
    return true;
}
// Unit::try_new_and_get(v, min_norm): succeeds iff norm_squared > min_norm^2; returns (v / sqrt(sq), sqrt(sq))
static inline bool unit_try_new_and_get(V3 a, real min_norm, V3* out, real* n_out) {
    real sq = norm_squared(a);
    if (sq > min_norm * min_norm) {
        real n = std::sqrt(sq);
        *out = a / n;
        *n_out = n;
        return true;
    }
    return false;
}
static inline bool unit_try_new(V3 a, real min_norm, V3* out) {
    real n;
    return unit_try_new_and_get(a, min_norm, out, &n);
}
static inline V3 inf(V3 a, V3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
static inline V3 sup(V3 a, V3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }

// UnitQuaternion stored (i, j, k, w) as nalgebra's coords.
struct Quat {
    real i, j, k, w;
};
struct Iso {
    V3 t;
    Quat q;
};
static inline Iso iso_identity() { return Iso{{0, 0, 0}, {0, 0, 0, 1}}; }

// UnitQuaternion * Vector3:  t = (q.ijk x v) * 2;  (t * w + q.ijk x t) + v
static inline V3 quat_rotate(Quat q, V3 v) {
    V3 qv = {q.i, q.j, q.k};
    V3 t = cross(qv, v) * real(2);
    V3 c = cross(qv, t);
    return (t * q.w + c) + v;
}
static inline Quat quat_conj(Quat q) { return {-q.i, -q.j, -q.k, q.w}; }
static inline V3 iso_mul_point(const Iso& m, V3 p) { return quat_rotate(m.q, p) + m.t; }
static inline V3 iso_mul_vec(const Iso& m, V3 v) { return quat_rotate(m.q, v); }
static inline V3 iso_inv_point(const Iso& m, V3 p) { return quat_rotate(quat_conj(m.q), p - m.t); }
static inline V3 iso_inv_vec(const Iso& m, V3 v) { return quat_rotate(quat_conj(m.q), v); }

// UnitQuaternion::to_rotation_matrix (row-major result m[r][c])
static inline void quat_to_matrix(Quat q, real m[3][3]) {
    real i = q.i, j = q.j, k = q.k, w = q.w;
    real ww = w * w, ii = i * i, jj = j * j, kk = k * k;
    real ij = i * j * real(2), wk = w * k * real(2), wj = w * j * real(2);
    real ik = i * k * real(2), jk = j * k * real(2), wi = w * i * real(2);
    m[0][0] = ww + ii - (jj + kk);
    m[0][1] = ij - wk;
    m[0][2] = wj + ik;
    m[1][0] = wk + ij;
    m[1][1] = ww - ii + jj - kk;
    m[1][2] = jk - wi;
    m[2][0] = ik - wj;
    m[2][1] = wi + jk;
    m[2][2] = ww - (ii + jj) + kk;
}
// utils/isometry_ops.rs:43-45: rotation.to_rotation_matrix().into_inner().abs() * v  (gemv, column axpy order)
static inline V3 absolute_transform_vector(const Iso& m, V3 v) {
    real r[3][3];
    quat_to_matrix(m.q, r);
    V3 o;
    for (int a = 0; a < 3; ++a)
        o[a] = (std::fabs(r[a][0]) * v.x + std::fabs(r[a][1]) * v.y) + std::fabs(r[a][2]) * v.z;
    return o;
}

// approx 0.5 relative_eq! with defaults epsilon = max_relative = f32::EPSILON
static inline bool relative_eq(real a, real b, real epsilon = EPS, real max_relative = EPS) {
    if (a == b) return true;
    if (std::isinf(a) || std::isinf(b)) return false;
    real abs_diff = std::fabs(a - b);
    if (abs_diff <= epsilon) return true;
    real aa = std::fabs(a), ab = std::fabs(b);
    real largest = ab > aa ? ab : aa;
    return abs_diff <= largest * max_relative;
}
static inline bool relative_eq_v3(V3 a, V3 b) {
    return relative_eq(a.x, b.x) && relative_eq(a.y, b.y) && relative_eq(a.z, b.z);
}
// approx ulps_eq! defaults: epsilon = f32::EPSILON, max_ulps = 4
static inline bool ulps_eq(real a, real b) {
    if (std::fabs(a - b) <= EPS) return true;
    if (std::signbit(a) != std::signbit(b)) return false;
    int64_t d;
    if (sizeof(real) == 4) {
        int32_t ia, ib;
        std::memcpy(&ia, &a, 4);
        std::memcpy(&ib, &b, 4);
        d = (int64_t)ia - (int64_t)ib;
    } else {
        int64_t ia, ib;
        std::memcpy(&ia, &a, 8);
        std::memcpy(&ib, &b, 8);
        d = ia - ib;
    }
    if (d < 0) d = -d;
    return d <= 4;
}
// na::clamp
static inline real clampf(real v, real lo, real hi) { return v > lo ? (v < hi ? v : hi) : lo; }
// Rust f32::signum
static inline real signum(real x) {
    if (std::isnan(x)) return x;
    return std::signbit(x) ? -real(1) : real(1);
}

// Vector3::orthonormal_subspace_basis(&[v], f):  first callback gets a x v, second gets a.
static inline void orthonormal_basis(V3 v, V3* first, V3* second) {
    V3 a;
    if (std::fabs(v.x) > std::fabs(v.y))
        a = {v.z, real(0), -v.x};
    else
        a = {real(0), -v.z, v.y};
    a = normalize(a);
    *first = cross(a, v);
    *second = a;
}

}  // namespace orc
