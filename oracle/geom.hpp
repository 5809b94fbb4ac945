// ORACLE (test infrastructure, NOT product code).  Restates /root/reference/src/geom.rs and
// src/bounds.rs (AABB parts) operation-for-operation.  Citations are file:line of the reference.
#pragma once
#include "cgm.hpp"

namespace mgfo {

static const float COLLISION_EPSILON = 0.000001f;  // geom.rs:27

struct Plane { Vec3 n; float d; };                  // geom.rs:32
struct Ray { Vec3 p, d; };                          // geom.rs:63
struct Segment { Vec3 a, b; };                      // geom.rs:91
struct Triangle { Vec3 a, b, c; };                  // geom.rs:128
struct Rectangle { Vec3 c; Vec3 u[2]; float e[2]; };// geom.rs:216
struct AABB { Vec3 c, r; };                         // geom.rs:257
struct OBB { Vec3 c; Quat q; Vec3 r; };             // geom.rs:272
struct Sphere { Vec3 c; float r; };                 // geom.rs:290
struct Capsule { Vec3 a, d; float r; };             // geom.rs:316
template <class T> struct Moving { T g; Vec3 v; };  // geom.rs:357

// geom.rs:49-58  Plane::from((a, b, c))
inline Plane plane_from_points(Vec3 a, Vec3 b, Vec3 c) {
    Vec3 n = normalize(cross(b - a, c - a));
    return {n, dot(n, a)};
}
inline Plane to_plane(const Triangle& t) { return plane_from_points(t.a, t.b, t.c); }  // geom.rs:182
inline Plane to_plane(const Rectangle& r) {  // geom.rs:240-246
    Vec3 n = cross(r.u[1], r.u[0]);
    return {n, dot(n, r.c)};
}
// geom.rs:227-235 Rectangle::new
inline Rectangle rectangle_new(Vec3 c, Vec3 a0, Vec3 a1) {
    Rectangle r;
    r.c = c;
    r.e[0] = magnitude(a0); r.e[1] = magnitude(a1);
    r.u[0] = a0 / r.e[0];   r.u[1] = a1 / r.e[1];
    return r;
}
inline Vec3 tri_normal(const Triangle& t) { return normalize(cross(t.b - t.a, t.c - t.a)); }  // geom.rs:149
// geom.rs:154-167
inline void tri_barycentric(const Triangle& t, Vec3 p, float* u, float* v_, float* w_) {
    Vec3 v0 = t.b - t.a, v1 = t.c - t.a, v2 = p - t.a;
    float d0 = dot(v0, v0), d1 = dot(v0, v1), d2 = dot(v1, v1), d3 = dot(v2, v0), d4 = dot(v2, v1);
    float denom = d0 * d2 - d1 * d1;
    float v = (d2 * d3 - d1 * d4) / denom;
    float w = (d0 * d4 - d1 * d3) / denom;
    *u = v; *v_ = w; *w_ = 1.0f - v - w;
}

inline float clampf(float n, float mn, float mx) {  // geom.rs:398
    if (n < mn) return mn;
    else if (n > mx) return mx;
    else return n;
}

// geom.rs:408-444
inline bool closest_pts_seg(const Segment& seg1, const Segment& seg2, Vec3* o1, Vec3* o2) {
    Vec3 d1 = seg1.b - seg1.a;
    Vec3 d2 = seg2.b - seg2.a;
    float a = magnitude2(d1);
    float e = magnitude2(d2);
    Vec3 r = seg1.a - seg2.a;
    float f = dot(d2, r);
    float s, t;
    if (a <= COLLISION_EPSILON) {
        if (e <= COLLISION_EPSILON) { s = 0.5f; t = 0.5f; }
        else { s = 0.5f; t = clampf(f / e, 0.0f, 1.0f); }
    } else {
        float c = dot(d1, r);
        if (e <= COLLISION_EPSILON) {
            s = clampf(-c / a, 0.0f, 1.0f); t = 0.0f;
        } else {
            float b = dot(d1, d2);
            float denom = a * e - b * b;
            float s0;
            if (denom != 0.0f) s0 = clampf((b * f - c * e) / denom, 0.0f, 1.0f);
            else return false;
            float t0 = b * s0 + f;
            if (t0 < 0.0f) { s = clampf(-c / a, 0.0f, 1.0f); t = 0.0f; }
            else if (t0 > e) { s = clampf((b - c) / a, 0.0f, 1.0f); t = 1.0f; }
            else { s = s0; t = t0 / e; }
        }
    }
    *o1 = seg1.a + d1 * s;
    *o2 = seg2.a + d2 * t;
    return true;
}

// ---- Shape::center (geom.rs:451-466 and impls) ----
inline Vec3 center(const Sphere& s) { return s.c; }
inline Vec3 center(const Capsule& c) { return c.a + c.d * 0.5f; }   // geom.rs:787
inline Vec3 center(const AABB& a) { return a.c; }
inline Vec3 center(const OBB& o) { return o.c; }
inline Vec3 center(const Rectangle& r) { return r.c; }
inline Vec3 center(const Triangle& t) { return (t.a + t.b + t.c) / 3.0f; }  // geom.rs:640
inline Vec3 center(const Segment& s) { return s.a + (s.b - s.a) * 0.5f; }

// ---- closest_point ----
inline Vec3 closest_point(const Segment& s, Vec3 to) {  // geom.rs:590-603
    Vec3 ab = s.b - s.a;
    float t = dot(ab, to - s.a);
    if (t <= 0.0f) return s.a;
    float denom = dot(ab, ab);
    if (t >= denom) return s.b;
    return s.a + ab * (t / denom);
}
inline Vec3 closest_point(const Triangle& tr, Vec3 to) {  // geom.rs:643-688
    Vec3 ab = tr.b - tr.a, ac = tr.c - tr.a, ap = to - tr.a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) return tr.a;
    Vec3 bp = to - tr.b;
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) return tr.b;
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        float v = d1 / (d1 - d3);
        return tr.a + ab * v;
    }
    Vec3 cp = to - tr.c;
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) return tr.c;
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        float w = d2 / (d2 - d6);
        return tr.a + ac * w;
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        return tr.b + (tr.c - tr.b) * w;
    }
    float denom = 1.0f / (va + vb + vc);
    float v = vb * denom, w = vc * denom;
    return tr.a + ab * v + ac * w;
}
inline Vec3 closest_point(const Sphere& s, Vec3 to) {  // geom.rs:751-755 (sic: |d|^2 / r^2)
    Vec3 d = to - s.c;
    float rat = magnitude2(d) / (s.r * s.r);
    return s.c + d * rat;
}

// ---- Polygon trait (geom.rs:869-923) ----
inline constexpr int num_vertices(const Triangle&) { return 3; }
inline constexpr int num_vertices(const Rectangle&) { return 4; }
inline Vec3 vertex(const Triangle& t, int i) { return i == 0 ? t.a : (i == 1 ? t.b : t.c); }
inline void edge(const Triangle&, int i, int* a, int* b) {
    static const int E[3][2] = {{0, 1}, {1, 2}, {2, 0}};
    *a = E[i][0]; *b = E[i][1];
}
inline Vec3 vertex(const Rectangle& r, int i) {  // geom.rs:906-918
    switch (i) {
        case 0: return r.c + r.u[0] * r.e[0] + r.u[1] * r.e[1];
        case 1: return r.c + r.u[0] * r.e[0] + (-r.u[1]) * r.e[1];
        case 2: return r.c + (-r.u[0]) * r.e[0] + (-r.u[1]) * r.e[1];
        default: return r.c + (-r.u[0]) * r.e[0] + r.u[1] * r.e[1];
    }
}
inline void edge(const Rectangle&, int i, int* a, int* b) {
    static const int E[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
    *a = E[i][0]; *b = E[i][1];
}

// ---- Convex::support (geom.rs:1027-1072) ----
inline Vec3 support(const AABB& a, Vec3 d) {
    return v3(signum(d.x) * a.r.x, signum(d.y) * a.r.y, signum(d.z) * a.r.z) + a.c;
}
inline Vec3 support(const OBB& o, Vec3 d0) {
    Vec3 d = rotate_vector(qinvert(o.q), d0);
    return rotate_point(o.q, v3(signum(d.x) * o.r.x, signum(d.y) * o.r.y, signum(d.z) * o.r.z)) + o.c;
}
// mesh.rs:141-236 ConvexMesh: a point soup.  Only what Convex::support (mesh.rs:223-236) reads is kept: the vertices as stored
// (the reference's support ignores `x`); strict > keeps the FIRST of equally good vertices.
struct ConvexMesh { const float* verts; unsigned n; };
inline Vec3 support(const ConvexMesh& m, Vec3 d) {
    Vec3 best_vert = v3(m.verts[0], m.verts[1], m.verts[2]);
    float best_norm = dot(d, best_vert);
    for (unsigned i = 1; i < m.n; ++i) {
        Vec3 vert = v3(m.verts[3 * i], m.verts[3 * i + 1], m.verts[3 * i + 2]);
        float norm = dot(d, vert);
        if (norm > best_norm) { best_vert = vert; best_norm = norm; }
    }
    return best_vert;
}
inline Vec3 support(const Sphere& s, Vec3 d) { return s.c + d * s.r; }
inline Vec3 support(const Capsule& cp, Vec3 d) {
    Vec3 c = cp.a + cp.d * 0.5f;
    Vec3 u = normalize(cp.d);
    float ud = dot(u, d);
    Vec3 w = d - u * ud;
    if (is_zero(w)) {
        return c + (magnitude(cp.d) * 0.5f + cp.r) * u * signum(ud);
    } else {
        return c + (magnitude(cp.d) * 0.5f + cp.r) * u * signum(ud) + normalize(w) * cp.r;
    }
}

// geom.rs:1138-1145
inline void compute_basis(Vec3 n, Vec3 out[2]) {
    Vec3 b = fabsf(n.x) >= 0.57735f ? v3(n.y, -n.x, 0.0f) : v3(0.0f, n.z, -n.y);
    b = normalize(b);
    out[0] = b;
    out[1] = cross(n, b);
}

// ---- bounds.rs (AABB as Bound) ----
inline AABB aabb_add(const AABB& a, float s) { return {a.c, a.r + v3(s, s, s)}; }   // bounds.rs:91-98
inline AABB aabb_translate(const AABB& a, Vec3 v) { return {a.c + v, a.r}; }          // impl_shape_reqs Add
inline AABB aabb_sub(const AABB& a, Vec3 v) { return {a.c + (-v), a.r}; }             // impl_shape_reqs Sub
struct NanBounds {};
inline AABB aabb_combine(const AABB& a, const AABB& b) {  // bounds.rs:113-130
    Vec3 lower = v3(fmin_(a.c.x - a.r.x, b.c.x - b.r.x), fmin_(a.c.y - a.r.y, b.c.y - b.r.y),
                    fmin_(a.c.z - a.r.z, b.c.z - b.r.z));
    Vec3 upper = v3(fmax_(a.c.x + a.r.x, b.c.x + b.r.x), fmax_(a.c.y + a.r.y, b.c.y + b.r.y),
                    fmax_(a.c.z + a.r.z, b.c.z + b.r.z));
    Vec3 r = (upper - lower) / 2.0f;
    if (!(r.x >= 0.0f) || !(r.y >= 0.0f) || !(r.z >= 0.0f)) throw NanBounds();
    Vec3 c = (upper + lower) / 2.0f;
    return {c, r};
}
inline float surface_area(const AABB& a) { return a.r.x * a.r.y + a.r.y * a.r.z + a.r.z * a.r.x; }  // bounds.rs:132
inline AABB bounds(const AABB& a) { return a; }
inline AABB bounds(const Triangle& t) {  // bounds.rs:137-152
    Vec3 c = (t.a + t.b + t.c) / 3.0f;
    float d0 = fmax_(fabsf(t.a.x - c.x), fmax_(fabsf(t.b.x - c.x), fabsf(t.c.x - c.x)));
    float d1 = fmax_(fabsf(t.a.y - c.y), fmax_(fabsf(t.b.y - c.y), fabsf(t.c.y - c.y)));
    float d2 = fmax_(fabsf(t.a.z - c.z), fmax_(fabsf(t.b.z - c.z), fabsf(t.c.z - c.z)));
    return {c, v3(d0, d1, d2)};
}
inline AABB bounds(const Sphere& s) { return {s.c, v3(s.r, s.r, s.r)}; }  // bounds.rs:170-177
inline AABB bounds(const Capsule& c) {                                       // bounds.rs:179-188
    float r = c.r + magnitude(c.d) * 0.5f;
    return {c.a + c.d * 0.5f, v3(r, r, r)};
}
template <class T> inline AABB bounds(const Moving<T>& m) {  // bounds.rs:60-68
    AABB s = bounds(m.g);
    AABB e = aabb_translate(s, m.v);
    return aabb_combine(s, e);
}

}  // namespace mgfo
