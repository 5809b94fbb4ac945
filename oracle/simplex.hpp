// ORACLE (test infrastructure, NOT product code).  Restates the reference's discrete path:
// src/simplex.rs (GJK closest_point_to_origin :172-200, state objects :224-415, EPA
// compute_contact :456-553), MinkowskiDiff (geom.rs:1099-1133) and the generic
// Penetrates / Contacts impls for Convex x Convex (collision.rs:404-425, 497-519).
//
// One documented deviation: the reference keeps the EPA horizon in a std HashMap whose
// iteration order is randomised per process (simplex.rs:523), so the reference itself is not
// bit-deterministic on ties.  Here the horizon is an insertion-ordered list (the order edges
// were first added), which is one of the orders the reference can take.
#pragma once
#include <vector>
#include "bvh.hpp"

namespace mgfo {

struct SupportPoint { Vec3 p, a, b; };  // geom.rs:1077
// High-water marks of the EPA polytope (Pool slots, horizon edges) since the last reset: test infrastructure for
// sizing the CUDA kernel's fixed-capacity polytope (mgfo_epa_high_water).
inline unsigned& epa_high_water(int which) { static unsigned hw[2] = {0, 0}; return hw[which]; }

template <class S1, class S2>
struct MinkowskiDiff {
    const S1* s1; const S2* s2;
    SupportPoint support_pt(Vec3 axis) const {  // geom.rs:1123-1132
        Vec3 a = support(*s1, axis);
        Vec3 b = support(*s2, -axis);
        return {a - b, a, b};
    }
};

enum SimplexStateId { ST_VERTEX = 1, ST_EDGE = 2, ST_FACE = 3, ST_VOLUME = 4 };

struct Simplex {
    SupportPoint points[4];
    int state;  // doubles as len()

    static Simplex from2(const SupportPoint& a, const SupportPoint& b) {  // simplex.rs:98-114
        Simplex s; SupportPoint z{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        s.points[0] = a; s.points[1] = b; s.points[2] = z; s.points[3] = z; s.state = ST_EDGE;
        return s;
    }

    // FaceSimplex::min_norm (simplex.rs:274-340); returns next state.
    static Vec3 face_min_norm(SupportPoint simp[4], int* next) {
        Vec3 a = simp[0].p, b = simp[1].p, c = simp[2].p;
        Vec3 ab = b - a, ac = c - a, ap = -a;
        float d1 = dot(ab, ap), d2 = dot(ac, ap);
        if (d1 <= 0.0f && d2 <= 0.0f) { *next = ST_EDGE; return simp[0].p; }
        Vec3 bp = -b;
        float d3 = dot(ab, bp), d4 = dot(ac, bp);
        if (d3 >= 0.0f && d4 <= d3) { simp[0] = simp[1]; *next = ST_EDGE; return simp[1].p; }
        float vc = d1 * d4 - d3 * d2;
        if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
            float v = d1 / (d1 - d3);
            *next = ST_FACE; return simp[0].p + ab * v;
        }
        Vec3 cp = -c;
        float d5 = dot(ab, cp), d6 = dot(ac, cp);
        if (d6 >= 0.0f && d5 <= d6) { simp[0] = simp[2]; *next = ST_EDGE; return simp[2].p; }
        float vb = d5 * d2 - d1 * d6;
        if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
            float w = d2 / (d2 - d6);
            simp[1] = simp[2];
            *next = ST_FACE; return simp[0].p + ac * w;
        }
        float va = d3 * d6 - d5 * d4;
        if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
            float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
            simp[0] = simp[2];
            *next = ST_FACE; return simp[1].p + (simp[2].p - simp[1].p) * w;
        }
        float denom = 1.0f / (va + vb + vc);
        float v = vb * denom, w = vc * denom;
        *next = ST_VOLUME; return simp[0].p + ab * v + ac * w;
    }
    static bool origin_outside_plane(Vec3 a, Vec3 b, Vec3 c, Vec3 d) {  // simplex.rs:342-349
        Vec3 ab_x_ac = cross(b - a, c - a);
        float sign_p = dot(-a, ab_x_ac);
        float sign_d = dot(d - a, ab_x_ac);
        return sign_p * sign_d < 0.0f;
    }
    Vec3 min_norm(int* next) {
        SupportPoint* simp = points;
        switch (state) {
            case ST_VERTEX: *next = ST_EDGE; return simp[0].p;  // simplex.rs:224-238
            case ST_EDGE: {                                      // simplex.rs:240-266
                Vec3 ab = simp[1].p - simp[0].p;
                float t = dot(ab, -simp[0].p);
                if (t <= 0.0f) { *next = ST_EDGE; return simp[0].p; }
                float denom = dot(ab, ab);
                if (t >= denom) { simp[0] = simp[1]; *next = ST_EDGE; return simp[1].p; }
                *next = ST_FACE; return simp[0].p + ab * (t / denom);
            }
            case ST_FACE: return face_min_norm(simp, next);
            default: {  // VolumeSimplex (simplex.rs:351-415)
                Vec3 closest_pt = v3(0, 0, 0);
                float best_dist = INF;
                int next_state = ST_VERTEX;
                SupportPoint a = simp[0], b = simp[1], c = simp[2], d = simp[3];
                Vec3 av = a.p, bv = b.p, cv = c.p, dv = d.p;
                auto test = [&](bool outside, SupportPoint n0, SupportPoint n1, SupportPoint n2, SupportPoint n3,
                                bool update_best) {
                    if (!outside) return;
                    SupportPoint ns[4] = {n0, n1, n2, n3};
                    int st;
                    Vec3 p = face_min_norm(ns, &st);
                    float nd = magnitude2(p);
                    if (nd < best_dist) {
                        closest_pt = p;
                        if (update_best) best_dist = nd;
                        next_state = st;
                        for (int i = 0; i < 4; ++i) simp[i] = ns[i];
                    }
                };
                test(origin_outside_plane(av, bv, cv, dv), a, b, c, d, true);
                test(origin_outside_plane(av, cv, dv, bv), a, c, d, b, true);
                test(origin_outside_plane(av, dv, bv, cv), a, d, b, c, true);
                test(origin_outside_plane(bv, dv, cv, av), b, d, c, a, false);  // sic: best_dist not updated
                *next = next_state;
                return closest_pt;
            }
        }
    }
    void add_point(const SupportPoint& p) { points[state - 1] = p; }

    // simplex.rs:172-200
    // The reference loops `loop { .. }` until one of the two exits fires; on some inputs (NaN, or a
    // simplex that cycles between two supports) it never does.  Oracle and device both give up after
    // GJK_MAX_STEPS and say so (`*converged = false`), instead of hanging.
    static constexpr int GJK_MAX_STEPS = 256;
    template <class Shape>
    Vec3 closest_point_to_origin(const Shape& shape, bool* converged = nullptr) {
        Vec3 prev_norm = v3(0, 0, 0);
        if (converged) *converged = true;
        for (int step = 0;; ++step) {
            if (step >= GJK_MAX_STEPS) { if (converged) *converged = false; return v3(0, 0, 0); }
            int next_state;
            Vec3 mn = min_norm(&next_state);
            if (magnitude2(mn) < COLLISION_EPSILON) {
                for (int i = state; i < 4; ++i) {
                    Vec3 m2 = -v3(prev_norm.z, prev_norm.x, prev_norm.y);
                    SupportPoint sp = shape.support_pt(-normalize(m2));
                    prev_norm = -normalize(m2);
                    points[i] = sp;
                }
                state = ST_VOLUME;
                return v3(0, 0, 0);
            }
            SupportPoint sp = shape.support_pt(-normalize(mn));
            Vec3 support_v = sp.p;
            prev_norm = mn;
            if (magnitude2(mn) >= magnitude2(support_v)) return mn;
            state = next_state;
            add_point(sp);
        }
    }

    // simplex.rs:456-553 (EPA)
    template <class S1, class S2>
    Contact compute_contact(const S1& s1, const S2& s2, int* iterations = nullptr) const {
        MinkowskiDiff<S1, S2> diff{&s1, &s2};
        struct Tri { SupportPoint a, b, c; };
        Pool<Tri> tris;
        {
            const SupportPoint &a = points[0], &b = points[1], &c = points[2], &d = points[3];
            tris.push({a, b, c}); tris.push({a, c, d}); tris.push({a, d, b}); tris.push({b, d, c});
        }
        struct Edge { uint32_t ka[3], kb[3]; Vec3 la_a, la_b, lb_a, lb_b; bool live; };
        std::vector<Edge> edges;
        auto key = [](Vec3 p, uint32_t k[3]) { k[0] = f2u(p.x); k[1] = f2u(p.y); k[2] = f2u(p.z); };
        auto add_edge = [&](const SupportPoint& a, const SupportPoint& b) {  // simplex.rs:423-451
            uint32_t ka[3], kb[3]; key(a.p, ka); key(b.p, kb);
            for (Edge& e : edges) {
                if (e.live && e.ka[0] == kb[0] && e.ka[1] == kb[1] && e.ka[2] == kb[2] && e.kb[0] == ka[0] &&
                    e.kb[1] == ka[1] && e.kb[2] == ka[2]) {
                    e.live = false;
                    return;
                }
            }
            // HashMap::insert overwrites an existing [a, b] key in place.
            for (Edge& e : edges) {
                if (e.live && e.ka[0] == ka[0] && e.ka[1] == ka[1] && e.ka[2] == ka[2] && e.kb[0] == kb[0] &&
                    e.kb[1] == kb[1] && e.kb[2] == kb[2]) {
                    e.la_a = a.a; e.la_b = a.b; e.lb_a = b.a; e.lb_b = b.b;
                    return;
                }
            }
            Edge e; for (int i = 0; i < 3; ++i) { e.ka[i] = ka[i]; e.kb[i] = kb[i]; }
            e.la_a = a.a; e.la_b = a.b; e.lb_a = b.a; e.lb_b = b.b; e.live = true;
            edges.push_back(e);
        };
        const int MAX_ITERATIONS = 100;
        for (int iter = 0; iter <= MAX_ITERATIONS; ++iter) {
            float closest_dist = INF; size_t closest_i = 0; Vec3 closest_n = v3(0, 0, 0);
            for (size_t i = 0; i < tris.entries.size(); ++i) {
                if (tris.entries[i].tag != Pool<Tri>::Occupied) continue;
                const Tri& t = tris.entries[i].item;
                Vec3 n = tri_normal(Triangle{t.a.p, t.b.p, t.c.p});
                float dist = fabsf(dot(n, t.a.p));
                if (closest_dist > dist) { closest_dist = dist; closest_i = i; closest_n = n; }
            }
            const Tri ct = tris[closest_i];
            Triangle ct_p{ct.a.p, ct.b.p, ct.c.p}, ct_a{ct.a.a, ct.b.a, ct.c.a};
            SupportPoint sup = diff.support_pt(closest_n);
            float v = dot(closest_n, sup.p) - closest_dist;
            if (v < COLLISION_EPSILON || iter == MAX_ITERATIONS) {
                float u, vv, w;
                tri_barycentric(ct_p, closest_dist * closest_n, &u, &vv, &w);
                Vec3 a = u * ct_a.a + vv * ct_a.b + w * ct_a.c;
                if (iterations) *iterations = iter;
                return Contact{a, a - closest_dist * closest_n, closest_n, 0.0f};
            }
            std::vector<size_t> to_remove;
            for (size_t i = 0; i < tris.entries.size(); ++i) {
                if (tris.entries[i].tag != Pool<Tri>::Occupied) continue;
                const Tri& t = tris.entries[i].item;
                Vec3 n = tri_normal(Triangle{t.a.p, t.b.p, t.c.p});
                if (dot(n, sup.p - t.a.p) > 0.0f) {
                    add_edge(t.a, t.b); add_edge(t.b, t.c); add_edge(t.c, t.a);
                    to_remove.push_back(i);
                }
            }
            if (edges.size() > epa_high_water(1)) epa_high_water(1) = (unsigned)edges.size();
            for (size_t i : to_remove) tris.remove(i);
            for (const Edge& e : edges) {
                if (!e.live) continue;
                SupportPoint a{v3(u2f(e.ka[0]), u2f(e.ka[1]), u2f(e.ka[2])), e.la_a, e.la_b};
                SupportPoint b{v3(u2f(e.kb[0]), u2f(e.kb[1]), u2f(e.kb[2])), e.lb_a, e.lb_b};
                tris.push({sup, a, b});
            }
            edges.clear();
            if (tris.entries.size() > epa_high_water(0)) epa_high_water(0) = (unsigned)tris.entries.size();
        }
        return Contact{};  // unreachable
    }
};

// collision.rs:404-425  Penetrates::separation; returns false for None.
template <class A, class B>
bool separation(const A& a, const B& b, float* out, int* status = nullptr) {
    Vec3 d = v3(1.0f, 0.0f, 0.0f);
    MinkowskiDiff<A, B> diff{&a, &b};
    Simplex simp = Simplex::from2(diff.support_pt(d), diff.support_pt(-d));
    bool ok = true;
    Vec3 min_dist = simp.closest_point_to_origin(diff, &ok);
    if (!ok) { if (status) *status = 3; return false; }
    float mag2 = magnitude2(min_dist);
    if (mag2 < COLLISION_EPSILON) return false;
    *out = sqrtf(mag2);
    return true;
}
// collision.rs:497-519  Contacts for Convex x Convex (discrete, t = 0)
template <class A, class B>
bool gjk_contact(const A& a, const B& b, Contact* out, int* epa_iters = nullptr, int* status = nullptr) {
    Vec3 d = v3(0.0f, 1.0f, 0.0f);
    MinkowskiDiff<A, B> diff{&a, &b};
    Simplex simp = Simplex::from2(diff.support_pt(d), diff.support_pt(-d));
    bool ok = true;
    Vec3 min_dist = simp.closest_point_to_origin(diff, &ok);
    if (!ok) { if (status) *status = 3; return false; }
    float mag2 = magnitude2(min_dist);
    if (mag2 > COLLISION_EPSILON) return false;
    // `tris[closest_i]` panics in the reference when the closest-face search found nothing and slot 0
    // is free (pool.rs:111 "unoccupied"): reported as status 4, never an abort across the C ABI.
    try { *out = simp.compute_contact(a, b, epa_iters); }
    catch (std::out_of_range&) { if (status) *status = 4; return false; }
    return true;
}

}  // namespace mgfo
