// ORACLE (test infrastructure, NOT product code) -- CPU restatement of the subset of
// cgmath 0.17 (third-party, Cargo.toml:20 of the reference; NOT vendored under
// /root/reference) that mgf's hot path calls.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build or load anything under oracle/.
//
// The formulas below restate cgmath 0.17's published algorithms (vector.rs, quaternion.rs,
// matrix.rs, approx 0.3) operation-for-operation in f32, compiled with -ffp-contract=off so
// no FMA is ever formed (rustc never contracts).  They are pinned by the reference's own
// bit-exact unit tests transcribed in oracle/kat.cpp.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace mgfo {

static const float INF = std::numeric_limits<float>::infinity();
static const float F32_EPSILON = 1.1920929e-07f;

// Rust f32::max / f32::min ignore a NaN operand exactly like fmaxf/fminf.
inline float fmax_(float a, float b) { return fmaxf(a, b); }
inline float fmin_(float a, float b) { return fminf(a, b); }
// Rust f32::signum: +1 for +0.0 and positives, -1 for -0.0 and negatives, NaN for NaN.
inline float signum(float x) {
    if (x != x) return x;
    return std::signbit(x) ? -1.0f : 1.0f;
}
inline int32_t f2i(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
inline uint32_t f2u(float f) { uint32_t i; std::memcpy(&i, &f, 4); return i; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// approx 0.3 `ulps_eq!` with default epsilon = f32::EPSILON, max_ulps = 4.
inline bool ulps_eq(float a, float b) {
    if (fabsf(a - b) <= F32_EPSILON) return true;
    if (signum(a) != signum(b)) return false;
    int64_t ia = f2i(a), ib = f2i(b);
    int64_t d = ia - ib;
    if (d < 0) d = -d;
    return d <= 4;
}
// approx 0.3 `relative_eq!(a, b, epsilon = eps)` (max_relative stays f32::EPSILON).
inline bool relative_eq(float a, float b, float eps = F32_EPSILON, float max_rel = F32_EPSILON) {
    if (a == b) return true;
    if (std::isinf(a) || std::isinf(b)) return false;
    float abs_diff = fabsf(a - b);
    if (abs_diff <= eps) return true;
    float aa = fabsf(a), ab = fabsf(b);
    float largest = ab > aa ? ab : aa;
    return abs_diff <= largest * max_rel;
}

struct Vec2 { float x, y; };
inline Vec2 operator+(Vec2 a, Vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline Vec2 operator-(Vec2 a, Vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline Vec2 operator*(float s, Vec2 a) { return {s * a.x, s * a.y}; }

// Vector3<f32> and Point3<f32> share one struct here; the reference's Point/Vector
// distinction carries no arithmetic.
struct Vec3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 v3(float x, float y, float z) { return {x, y, z}; }
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(float s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator/(Vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline Vec3& operator+=(Vec3& a, Vec3 b) { a = a + b; return a; }
inline Vec3& operator-=(Vec3& a, Vec3 b) { a = a - b; return a; }
inline bool operator==(Vec3 a, Vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// cgmath: dot = mul_element_wise().sum() = (x*x' + y*y') + z*z'
inline float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float magnitude2(Vec3 a) { return dot(a, a); }
inline float magnitude(Vec3 a) { return sqrtf(dot(a, a)); }
// cgmath InnerSpace::normalize = normalize_to(1) = self * (1 / magnitude)
inline Vec3 normalize(Vec3 a) { return a * (1.0f / magnitude(a)); }
// num_traits::Zero::is_zero for VectorN: exact comparison with zero.
inline bool is_zero(Vec3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
inline Vec2 truncate(Vec3 a) { return {a.x, a.y}; }

struct Quat { float s; Vec3 v; };
inline Quat quat_one() { return {1.0f, {0.0f, 0.0f, 0.0f}}; }
inline Quat from_sv(float s, Vec3 v) { return {s, v}; }
inline Quat operator+(Quat a, Quat b) { return {a.s + b.s, a.v + b.v}; }
inline Quat operator*(Quat a, float k) { return {a.s * k, a.v * k}; }
inline Quat operator/(Quat a, float k) { return {a.s / k, a.v / k}; }
inline Quat operator*(Quat l, Quat r) {
    return {
        l.s * r.s - l.v.x * r.v.x - l.v.y * r.v.y - l.v.z * r.v.z,
        {l.s * r.v.x + l.v.x * r.s + l.v.y * r.v.z - l.v.z * r.v.y,
         l.s * r.v.y + l.v.y * r.s + l.v.z * r.v.x - l.v.x * r.v.z,
         l.s * r.v.z + l.v.z * r.s + l.v.x * r.v.y - l.v.y * r.v.x}};
}
inline float qdot(Quat a, Quat b) { return a.s * b.s + dot(a.v, b.v); }
inline float qmagnitude2(Quat a) { return qdot(a, a); }
inline Quat qnormalize(Quat a) { return a * (1.0f / sqrtf(qdot(a, a))); }
inline Quat conjugate(Quat a) { return {a.s, -a.v}; }
// Rotation::invert for Quaternion = conjugate / magnitude2
inline Quat qinvert(Quat a) { return conjugate(a) / qmagnitude2(a); }
// Quaternion * Vector3 (Rotation::rotate_vector)
inline Vec3 rotate_vector(Quat q, Vec3 v) {
    Vec3 tmp = cross(q.v, v) + (v * q.s);
    return (cross(q.v, tmp) * 2.0f) + v;
}
inline Vec3 rotate_point(Quat q, Vec3 p) { return rotate_vector(q, p); }
inline Quat from_axis_angle(Vec3 axis, float angle_rad) {
    float h = angle_rad * 0.5f;
    float s = sinf(h), c = cosf(h);
    return from_sv(c, axis * s);
}
// Quaternion::from_arc(src, dst, None)
inline Quat from_arc(Vec3 src, Vec3 dst) {
    float mag_avg = sqrtf(magnitude2(src) * magnitude2(dst));
    float d = dot(src, dst);
    if (ulps_eq(d, mag_avg)) {
        return quat_one();
    } else if (ulps_eq(d, -mag_avg)) {
        Vec3 v = cross(v3(1.0f, 0.0f, 0.0f), src);
        if (ulps_eq(v.x, 0.0f) && ulps_eq(v.y, 0.0f) && ulps_eq(v.z, 0.0f))
            v = cross(v3(0.0f, 1.0f, 0.0f), src);
        v = normalize(v);
        return from_axis_angle(v, 3.14159265358979323846f);
    } else {
        return qnormalize(from_sv(mag_avg + d, cross(src, dst)));
    }
}

// Matrix3<f32>, column-major: c[col] is a column vector.
struct Mat3 { Vec3 c[3]; };
// Matrix3::new(c0r0, c0r1, c0r2, c1r0, ...)
inline Mat3 mat3_new(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
    return {{{a, b, c}, {d, e, f}, {g, h, i}}};
}
inline Mat3 from_cols(Vec3 a, Vec3 b, Vec3 c) { return {{a, b, c}}; }
inline Mat3 mat3_one() { return mat3_new(1, 0, 0, 0, 1, 0, 0, 0, 1); }
inline Mat3 mat3_zero() { return mat3_new(0, 0, 0, 0, 0, 0, 0, 0, 0); }
inline Vec3 row(const Mat3& m, int r) { return {m.c[0][r], m.c[1][r], m.c[2][r]}; }
inline Mat3 operator+(const Mat3& a, const Mat3& b) { return {{a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]}}; }
inline Mat3 operator-(const Mat3& a, const Mat3& b) { return {{a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]}}; }
inline Mat3 operator*(const Mat3& a, float s) { return {{a.c[0] * s, a.c[1] * s, a.c[2] * s}}; }
inline Mat3 operator*(float s, const Mat3& a) { return {{s * a.c[0], s * a.c[1], s * a.c[2]}}; }
inline Vec3 operator*(const Mat3& m, Vec3 v) { return {dot(row(m, 0), v), dot(row(m, 1), v), dot(row(m, 2), v)}; }
inline Mat3 operator*(const Mat3& l, const Mat3& r) {
    return mat3_new(dot(row(l, 0), r.c[0]), dot(row(l, 1), r.c[0]), dot(row(l, 2), r.c[0]),
                    dot(row(l, 0), r.c[1]), dot(row(l, 1), r.c[1]), dot(row(l, 2), r.c[1]),
                    dot(row(l, 0), r.c[2]), dot(row(l, 1), r.c[2]), dot(row(l, 2), r.c[2]));
}
inline Mat3 transpose(const Mat3& m) { return from_cols(row(m, 0), row(m, 1), row(m, 2)); }
inline float determinant(const Mat3& m) {
    return m.c[0].x * (m.c[1].y * m.c[2].z - m.c[2].y * m.c[1].z) -
           m.c[1].x * (m.c[0].y * m.c[2].z - m.c[2].y * m.c[0].z) +
           m.c[2].x * (m.c[0].y * m.c[1].z - m.c[1].y * m.c[0].z);
}
inline bool invert(const Mat3& m, Mat3* out) {
    float det = determinant(m);
    if (det == 0.0f) return false;
    *out = transpose(from_cols(cross(m.c[1], m.c[2]) / det, cross(m.c[2], m.c[0]) / det,
                               cross(m.c[0], m.c[1]) / det));
    return true;
}
// Matrix3::from(Quaternion)
inline Mat3 mat3_from_quat(Quat q) {
    float x2 = q.v.x + q.v.x, y2 = q.v.y + q.v.y, z2 = q.v.z + q.v.z;
    float xx2 = x2 * q.v.x, xy2 = x2 * q.v.y, xz2 = x2 * q.v.z;
    float yy2 = y2 * q.v.y, yz2 = y2 * q.v.z, zz2 = z2 * q.v.z;
    float sy2 = y2 * q.s, sz2 = z2 * q.s, sx2 = x2 * q.s;
    return mat3_new(1.0f - yy2 - zz2, xy2 + sz2, xz2 - sy2,
                    xy2 - sz2, 1.0f - xx2 - zz2, yz2 + sx2,
                    xz2 + sy2, yz2 - sx2, 1.0f - xx2 - yy2);
}

}  // namespace mgfo
