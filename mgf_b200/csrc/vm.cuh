// vm.cuh -- f32 vector / quaternion / matrix arithmetic for the device kernels (and the few
// host-side setup paths of the library).  The operation ORDER of every function is fixed to
// what mgf's math backend (cgmath 0.17) evaluates, because contact/no-contact predicates in
// the narrowphase flip on 1-ulp differences.  The translation units that
// include this header are compiled with --fmad=false and IEEE div/sqrt, so each expression
// rounds exactly like the CPU reference.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define HD __host__ __device__ __forceinline__

namespace mgfb {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct Q4 { float s; V3 v; };
struct M3 { V3 c0, c1, c2; };  // columns

HD V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
HD V3 zero3() { return mk3(0.0f, 0.0f, 0.0f); }
HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
HD V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
HD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
HD float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
HD V3 cross3(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
HD float len2(V3 a) { return dot3(a, a); }
HD float len(V3 a) { return sqrtf(dot3(a, a)); }
HD V3 unit(V3 a) { return a * (1.0f / len(a)); }   // multiply by reciprocal, not divide
HD bool all_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

HD V2 mk2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
HD V2 xy(V3 a) { return mk2(a.x, a.y); }

HD float sgn(float x) {  // +1 for +0, -1 for -0, NaN stays NaN
    if (x != x) return x;
    return signbit(x) ? -1.0f : 1.0f;
}
HD int32_t fbits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    union { float f; int32_t i; } u; u.f = f; return u.i;
#endif
}
// "equal within 4 ulps or within f32::EPSILON" -- the predicate from_arc branches on.
HD bool near_ulps(float a, float b) {
    if (fabsf(a - b) <= 1.1920929e-07f) return true;
    if (sgn(a) != sgn(b)) return false;
    long long d = (long long)fbits(a) - (long long)fbits(b);
    if (d < 0) d = -d;
    return d <= 4;
}
// relative_eq(a, b, epsilon = eps) with max_relative = f32::EPSILON
HD bool near_rel(float a, float b, float eps) {
    if (a == b) return true;
    if (isinf(a) || isinf(b)) return false;
    float ad = fabsf(a - b);
    if (ad <= eps) return true;
    float aa = fabsf(a), ab = fabsf(b);
    float largest = ab > aa ? ab : aa;
    return ad <= largest * 1.1920929e-07f;
}
HD float clampf3(float n, float lo, float hi) { return n < lo ? lo : (n > hi ? hi : n); }

// ---- quaternions (s, v) ----
HD Q4 mkq(float s, V3 v) { Q4 q; q.s = s; q.v = v; return q; }
HD Q4 qident() { return mkq(1.0f, zero3()); }
HD Q4 qadd(Q4 a, Q4 b) { return mkq(a.s + b.s, a.v + b.v); }
HD Q4 qscale(Q4 a, float k) { return mkq(a.s * k, a.v * k); }
HD Q4 qdivs(Q4 a, float k) { return mkq(a.s / k, a.v / k); }
HD Q4 qmul(Q4 l, Q4 r) {
    return mkq(l.s * r.s - l.v.x * r.v.x - l.v.y * r.v.y - l.v.z * r.v.z,
               mk3(l.s * r.v.x + l.v.x * r.s + l.v.y * r.v.z - l.v.z * r.v.y,
                   l.s * r.v.y + l.v.y * r.s + l.v.z * r.v.x - l.v.x * r.v.z,
                   l.s * r.v.z + l.v.z * r.s + l.v.x * r.v.y - l.v.y * r.v.x));
}
HD float qlen2(Q4 a) { return a.s * a.s + dot3(a.v, a.v); }
HD Q4 qunit(Q4 a) { return qscale(a, 1.0f / sqrtf(qlen2(a))); }
HD Q4 qconj(Q4 a) { return mkq(a.s, -a.v); }
HD Q4 qinv(Q4 a) { return qdivs(qconj(a), qlen2(a)); }
HD V3 qrot(Q4 q, V3 v) {
    V3 tmp = cross3(q.v, v) + (v * q.s);
    return (cross3(q.v, tmp) * 2.0f) + v;
}
HD Q4 q_axis_angle(V3 axis, float rad) {
    float h = rad * 0.5f;
    return mkq(cosf(h), axis * sinf(h));
}
// shortest-arc rotation src -> dst
HD Q4 q_from_arc(V3 src, V3 dst) {
    float mag_avg = sqrtf(len2(src) * len2(dst));
    float d = dot3(src, dst);
    if (near_ulps(d, mag_avg)) return qident();
    if (near_ulps(d, -mag_avg)) {
        V3 v = cross3(mk3(1.0f, 0.0f, 0.0f), src);
        if (near_ulps(v.x, 0.0f) && near_ulps(v.y, 0.0f) && near_ulps(v.z, 0.0f)) v = cross3(mk3(0.0f, 1.0f, 0.0f), src);
        return q_axis_angle(unit(v), 3.14159265358979323846f);
    }
    return qunit(mkq(mag_avg + d, cross3(src, dst)));
}

// ---- 3x3 matrices, column-major ----
HD M3 mkm(V3 a, V3 b, V3 c) { M3 m; m.c0 = a; m.c1 = b; m.c2 = c; return m; }
HD M3 m_zero() { return mkm(zero3(), zero3(), zero3()); }
HD M3 m_ident() { return mkm(mk3(1, 0, 0), mk3(0, 1, 0), mk3(0, 0, 1)); }
HD M3 m_diag(float a, float b, float c) { return mkm(mk3(a, 0, 0), mk3(0, b, 0), mk3(0, 0, c)); }
HD V3 mrow0(const M3& m) { return mk3(m.c0.x, m.c1.x, m.c2.x); }
HD V3 mrow1(const M3& m) { return mk3(m.c0.y, m.c1.y, m.c2.y); }
HD V3 mrow2(const M3& m) { return mk3(m.c0.z, m.c1.z, m.c2.z); }
HD V3 mmulv(const M3& m, V3 v) { return mk3(dot3(mrow0(m), v), dot3(mrow1(m), v), dot3(mrow2(m), v)); }
HD M3 mmul(const M3& l, const M3& r) {
    V3 r0 = mrow0(l), r1 = mrow1(l), r2 = mrow2(l);
    return mkm(mk3(dot3(r0, r.c0), dot3(r1, r.c0), dot3(r2, r.c0)),
               mk3(dot3(r0, r.c1), dot3(r1, r.c1), dot3(r2, r.c1)),
               mk3(dot3(r0, r.c2), dot3(r1, r.c2), dot3(r2, r.c2)));
}
HD M3 mtrans(const M3& m) { return mkm(mrow0(m), mrow1(m), mrow2(m)); }
HD M3 madd(const M3& a, const M3& b) { return mkm(a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2); }
HD M3 msub(const M3& a, const M3& b) { return mkm(a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2); }
HD M3 mscale(const M3& a, float s) { return mkm(a.c0 * s, a.c1 * s, a.c2 * s); }
HD M3 smul(float s, const M3& a) { return mkm(s * a.c0, s * a.c1, s * a.c2); }
HD float mdet(const M3& m) {
    return m.c0.x * (m.c1.y * m.c2.z - m.c2.y * m.c1.z) - m.c1.x * (m.c0.y * m.c2.z - m.c2.y * m.c0.z) +
           m.c2.x * (m.c0.y * m.c1.z - m.c1.y * m.c0.z);
}
HD bool minv(const M3& m, M3* out) {
    float det = mdet(m);
    if (det == 0.0f) return false;
    *out = mtrans(mkm(cross3(m.c1, m.c2) / det, cross3(m.c2, m.c0) / det, cross3(m.c0, m.c1) / det));
    return true;
}
HD M3 m_from_q(Q4 q) {
    float x2 = q.v.x + q.v.x, y2 = q.v.y + q.v.y, z2 = q.v.z + q.v.z;
    float xx2 = x2 * q.v.x, xy2 = x2 * q.v.y, xz2 = x2 * q.v.z;
    float yy2 = y2 * q.v.y, yz2 = y2 * q.v.z, zz2 = z2 * q.v.z;
    float sy2 = y2 * q.s, sz2 = z2 * q.s, sx2 = x2 * q.s;
    return mkm(mk3(1.0f - yy2 - zz2, xy2 + sz2, xz2 - sy2), mk3(xy2 - sz2, 1.0f - xx2 - zz2, yz2 + sx2),
               mk3(xz2 + sy2, yz2 - sx2, 1.0f - xx2 - yy2));
}

}  // namespace mgfb
