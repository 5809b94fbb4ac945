// narrow.cuh -- continuous narrowphase: every `Contacts` implementation on mgf's step path as
// straight-line device functions (no callbacks: a pair test returns its 0..2 hits by value).
// One function per static-shape x moving-shape pair, so each kernel specialisation that calls
// it is divergence-free in the shape dispatch.  Branch structure and floating-point operation
// order follow the reference exactly (citations: file:line of mgf's src/), because the
// hit / no-hit predicates are 1-ulp sensitive.
#pragma once
#include "vm.cuh"

namespace mgfb {

#define MGFB_EPS 0.000001f   /* geom.rs:27 COLLISION_EPSILON */
#define MGFB_INF __int_as_float(0x7f800000)

struct Sph { V3 c; float r; };
struct Cap { V3 a, d; float r; };
struct Pln { V3 n; float d; };
struct Hit { V3 a, b, n; float t; };
struct Hits { int n; Hit h[2]; };

HD Hit mkhit(V3 a, V3 b, V3 n, float t) { Hit h; h.a = a; h.b = b; h.n = n; h.t = t; return h; }
HD Hit flip(Hit h) { return mkhit(h.b, h.a, -h.n, h.t); }   // collision.rs:444-456 Neg for Contact
HD void push(Hits& hs, Hit h) { if (hs.n < 2) hs.h[hs.n] = h; hs.n++; }
#ifdef __CUDA_ARCH__
#define INF_F __int_as_float(0x7f800000)
#else
#define INF_F INFINITY
#endif

// ---- polygons (geom.rs:869-923) ----
struct Tri {
    V3 a, b, c;
    static const int NV = 3;
    HD V3 vertex(int i) const { return i == 0 ? a : (i == 1 ? b : c); }
    HD void edge(int i, int& p, int& q) const { p = i; q = i == 2 ? 0 : i + 1; }
    HD Pln plane() const {  // geom.rs:49-58,182-192
        V3 n = unit(cross3(b - a, c - a));
        Pln pl; pl.n = n; pl.d = dot3(n, a); return pl;
    }
    HD bool contains(V3 p) const {  // collision.rs:85-100
        V3 v = p - a, ac = c - a, ab = b - a;
        float dot1 = dot3(ac, ac), dot2 = dot3(ac, ab), dot3_ = dot3(ac, v), dot4 = dot3(ab, ab), dot5 = dot3(ab, v);
        float invd = 1.0f / (dot1 * dot4 - dot2 * dot2);
        float u = (dot4 * dot3_ - dot2 * dot5) * invd;
        float w = (dot1 * dot5 - dot2 * dot3_) * invd;
        return u >= 0.0f && w >= 0.0f && (u + w) < 1.0f;
    }
};
struct Rct {
    V3 c, u0, u1; float e0, e1;
    static const int NV = 4;
    HD V3 vertex(int i) const {  // geom.rs:906-918
        switch (i) {
            case 0: return c + u0 * e0 + u1 * e1;
            case 1: return c + u0 * e0 + (-u1) * e1;
            case 2: return c + (-u0) * e0 + (-u1) * e1;
            default: return c + (-u0) * e0 + u1 * e1;
        }
    }
    HD void edge(int i, int& p, int& q) const { p = i; q = i == 3 ? 0 : i + 1; }
    HD Pln plane() const { V3 n = cross3(u1, u0); Pln pl; pl.n = n; pl.d = dot3(n, c); return pl; }  // geom.rs:240
    HD bool contains(V3 p) const {  // collision.rs:102-112 (absolute p, n = u0 x u1)
        V3 n = cross3(u0, u1);
        return near_rel(dot3(p, n), dot3(n, c), MGFB_EPS) && fabsf(dot3(p, u0)) <= e0 && fabsf(dot3(p, u1)) <= e1;
    }
};

// geom.rs:590-603 Segment::closest_point
HD V3 seg_closest(V3 sa, V3 sb, V3 to) {
    V3 ab = sb - sa;
    float t = dot3(ab, to - sa);
    if (t <= 0.0f) return sa;
    float denom = dot3(ab, ab);
    if (t >= denom) return sb;
    return sa + ab * (t / denom);
}

// geom.rs:408-444 closest_pts_seg; only the point on segment 1 is ever used by callers.
HD bool seg_seg_closest(V3 a1, V3 b1, V3 a2, V3 b2, V3* p1) {
    V3 d1 = b1 - a1, d2 = b2 - a2;
    float a = len2(d1), e = len2(d2);
    V3 r = a1 - a2;
    float f = dot3(d2, r);
    float s;
    if (a <= MGFB_EPS) {
        s = 0.5f;
    } else {
        float c = dot3(d1, r);
        if (e <= MGFB_EPS) {
            s = clampf3(-c / a, 0.0f, 1.0f);
        } else {
            float b = dot3(d1, d2);
            float denom = a * e - b * b;
            if (denom == 0.0f) return false;
            float s0 = clampf3((b * f - c * e) / denom, 0.0f, 1.0f);
            float t0 = b * s0 + f;
            if (t0 < 0.0f) s = clampf3(-c / a, 0.0f, 1.0f);
            else if (t0 > e) s = clampf3((b - c) / a, 0.0f, 1.0f);
            else s = s0;
        }
    }
    *p1 = a1 + d1 * s;
    return true;
}

// ---- ray casts: Ray has DT = inf (geom.rs:819), callers re-check t <= 1 ----
// collision.rs:249-273
HD bool ray_sphere(V3 p, V3 d, V3 sc, float sr, float* tout, V3* pout) {
    V3 m = p - sc;
    float a = len2(d), b = dot3(m, d), c = len2(m) - sr * sr;
    if (c > 0.0f && b > 0.0f) return false;
    float discr = b * b - a * c;
    if (discr < 0.0f) return false;
    float t = fmaxf((-b - sqrtf(discr)) / a, 0.0f);
    *tout = t; *pout = p + t * d;
    return true;
}
// collision.rs:275-359
HD bool ray_capsule(V3 p, V3 d, V3 ca, V3 cd, float cr, float* tout, V3* pout) {
    V3 m = p - ca;
    float md = dot3(m, cd), nd = dot3(d, cd), dd = dot3(cd, cd), nn = len2(d), mn = dot3(m, d);
    float a = dd * nn - nd * nd;
    float k = len2(m) - cr * cr;
    float t;
    if (fabsf(a) < MGFB_EPS) {
        float b, c;
        if (md < 0.0f) { b = mn; c = k; }
        else if (md > dd) { V3 m2 = p - (ca + cd); b = dot3(m2, d); c = len2(m2) - cr * cr; }
        else return false;
        if (c > 0.0f && b > 0.0f) return false;
        float discr = b * b - nn * c;
        if (discr < 0.0f) return false;
        t = fmaxf((-b - sqrtf(discr)) / nn, 0.0f);
        *tout = t; *pout = p + t * d;
        return true;
    }
    float c = dd * k - md * md;
    float b = dd * mn - nd * md;
    float discr = b * b - a * c;
    if (discr < 0.0f) return false;
    t = (-b - sqrtf(discr)) / a;
    if (t < 0.0f) return false;
    if (md + t * nd < 0.0f) {
        if (mn > 0.0f && k > 0.0f) return false;
        float discr2 = mn * mn - nn * k;
        if (discr2 < 0.0f) return false;
        t = fmaxf((-mn - sqrtf(discr2)) / nn, 0.0f);
    } else if (md + t * nd > dd) {
        V3 m2 = p - (ca + cd);
        float b2 = dot3(m2, d), c2 = len2(m2) - cr * cr;
        if (c2 > 0.0f && b2 > 0.0f) return false;
        float discr2 = b2 * b2 - nn * c2;
        if (discr2 < 0.0f) return false;
        t = fmaxf((-b2 - sqrtf(discr2)) / nn, 0.0f);
    }
    *tout = t; *pout = p + t * d;
    return true;
}

// collision.rs:169-184
HD bool ray_plane(V3 p, V3 d, Pln pl, float DT, float* tout, V3* pout) {
    float denom = dot3(pl.n, d);
    if (denom == 0.0f) return false;
    float t = (pl.d - dot3(pl.n, p)) / denom;
    if (t <= 0.0f || t > DT) return false;
    *tout = t; *pout = p + d * t;
    return true;
}
// collision.rs:202-236
HD bool ray_aabb(V3 p, V3 d, V3 bc, V3 br, float DT, float* tout, V3* pout) {
    float t_min = 0.0f, t_max = __builtin_huge_valf();
    float ps[3] = {p.x, p.y, p.z}, ds[3] = {d.x, d.y, d.z}, cs[3] = {bc.x, bc.y, bc.z}, rs[3] = {br.x, br.y, br.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (fabsf(ds[k]) < MGFB_EPS) {
            if (fabsf(ps[k] - cs[k]) > rs[k]) return false;
        } else {
            float ood = 1.0f / ds[k];
            float t1 = (cs[k] - rs[k] - ps[k]) * ood;
            float t2 = (cs[k] + rs[k] - ps[k]) * ood;
            if (t1 > t2) { t_min = fmaxf(t_min, t2); t_max = fminf(t_max, t1); }
            else { t_min = fmaxf(t_min, t1); t_max = fminf(t_max, t2); }
            if (t_min > t_max) return false;
        }
    }
    if (t_min > DT) return false;
    *tout = t_min; *pout = p + d * t_min;
    return true;
}

// ---- Plane x Moving<Sphere> (collision.rs:521-553) ----
HD bool plane_msphere(Pln pl, Sph s, V3 v, Hit* out) {
    float dist = dot3(pl.n, s.c) - pl.d;
    if (fabsf(dist) <= s.r) {
        *out = mkhit(s.c + (-pl.n) * dist, s.c + (-pl.n) * s.r, pl.n, 0.0f);
        return true;
    }
    float denom = dot3(pl.n, v);
    if (denom * dist >= 0.0f) return false;
    float r = dist > 0.0f ? s.r : -s.r;
    float t = (r - dist) / denom;
    if (t <= 1.0f) {
        V3 q = s.c + t * v - r * pl.n;
        *out = mkhit(q, q, pl.n, t);
        return true;
    }
    return false;
}
// ---- Plane x Moving<Capsule> (collision.rs:555-605) ----
HD bool plane_mcapsule(Pln pl, Cap c, V3 v, Hit* out) {
    float denom = dot3(pl.n, unit(c.d));
    V3 ctr;
    if (fabsf(denom) < MGFB_EPS) {
        ctr = c.a + c.d * 0.5f;
    } else {
        float t = (pl.d - dot3(pl.n, c.a)) / denom;
        if (t > 1.0f) ctr = c.a + c.d;
        else if (t < 0.0f) ctr = c.a;
        else {
            V3 q = c.a + c.d * t;
            float dist = dot3(pl.n, c.a) - pl.d;
            V3 b = (dist < 0.0f ? c.a : c.a + c.d) + (-pl.n) * c.r;
            *out = mkhit(q, b, pl.n, 0.0f);
            return true;
        }
    }
    Sph ms; ms.c = ctr; ms.r = c.r;
    return plane_msphere(pl, ms, v, out);
}

// ---- Polygon x Moving<Sphere> (collision.rs:610-659) ----
template <class Poly>
HD bool poly_msphere(const Poly& poly, Sph s, V3 v, Hit* out) {
    Pln p = poly.plane();
    Hit contact;
    if (!plane_msphere(p, s, v, &contact)) return false;
    if (poly.contains(contact.a)) { *out = contact; return true; }
    float first_t = INF_F;
    V3 tri_p = zero3();
    if (len2(v) == 0.0f) return false;
#pragma unroll
    for (int e = 0; e < Poly::NV; ++e) {
        int ia, ib; poly.edge(e, ia, ib);
        V3 v1 = poly.vertex(ia), v2 = poly.vertex(ib);
        float t; V3 ip;
        if (ray_capsule(s.c, v, v1, v2 - v1, s.r, &t, &ip)) {
            if (t <= 1.0f && t < first_t) {
                first_t = t;
                tri_p = seg_closest(v1, v2, ip);
            }
        }
    }
    if (first_t != INF_F) { *out = mkhit(tri_p, tri_p, p.n, first_t); return true; }
    return false;
}

// collision.rs:667-688
HD float area2d(V2 a, V2 b, V2 c) { return (a.x - c.x) * (b.y - c.y) - (a.y - c.y) * (b.x - c.x); }
HD bool seg2d_hit(V2 a, V2 b, V2 c, V2 d, float* tout) {
    float a1 = area2d(a, b, d), a2 = area2d(a, b, c);
    if (a1 * a2 <= 0.0f) {
        float a3 = area2d(c, d, a);
        float a4 = a3 + a2 - a1;
        if (a3 * a4 <= 0.0f) { *tout = a3 / (a3 - a4); return true; }
    }
    return false;
}

// ---- Polygon x Moving<Capsule> (collision.rs:693-1086): 0, 1 or 2 hits ----
template <class Poly>
HD Hits poly_mcapsule(const Poly& poly, Cap c, V3 v) {
    Hits hs; hs.n = 0;
    Pln p = poly.plane();
    // (1) already piercing the face (:697-719)
    float denom = dot3(p.n, unit(c.d));
    if (fabsf(denom) > MGFB_EPS) {
        float t = (p.d - dot3(p.n, c.a)) / denom;
        if (t <= 1.0f && t >= 0.0f) {
            V3 q = c.a + c.d * t;
            if (poly.contains(q)) {
                V3 b = (dot3(p.n, c.a) - p.d < 0.0f ? c.a : c.a + c.d) + (-p.n) * c.r;
                push(hs, mkhit(q, b, p.n, 0.0f));
                return hs;
            }
        }
    }
    // (2) seed contact from the two end spheres against the plane (:721-764)
    Sph s0; s0.c = c.a; s0.r = c.r;
    Sph s1; s1.c = c.a + c.d; s1.r = c.r;
    bool have = false, checked = false; Hit fc = mkhit(zero3(), zero3(), zero3(), 0.0f); V3 dir = zero3();
    {
        Hit c1, c2;
        if (plane_msphere(p, s0, v, &c1)) {
            if (plane_msphere(p, s1, v, &c2)) {
                if (c2.t < c1.t) { have = true; fc = c2; dir = -c.d; }
                else if (c2.t == 0.0f) {
                    bool in1 = poly.contains(c1.a), in2 = poly.contains(c2.a);
                    if (in1 && in2) { push(hs, c2); push(hs, c1); return hs; }
                    else if (in1) { have = true; fc = c1; dir = c.d; checked = true; }
                    else if (in2) { have = true; fc = c2; dir = -c.d; checked = true; }
                } else { have = true; fc = c1; dir = c.d; }
            } else { have = true; fc = c1; dir = c.d; }
        } else if (plane_msphere(p, s1, v, &c1)) { have = true; fc = c1; dir = -c.d; }
    }
    // (3) silhouette clipping in the plane's 2-D frame (:767-890)
    if (have) {
        V3 sil_v = dir - p.n * dot3(dir, p.n) / len2(p.n);
        Q4 rot = q_from_arc(p.n, mk3(0.0f, 0.0f, 1.0f));
        V2 sil_a = xy(qrot(rot, fc.a + (-p.n) * p.d));
        V2 sil_b = xy(qrot(rot, fc.a + sil_v - p.n * p.d));
        if (checked || poly.contains(fc.a)) {
            push(hs, fc);
            if (fabsf(dot3(dir, p.n)) >= MGFB_EPS) return hs;
            float t_max = 0.0f;
            for (int e = 0; e < Poly::NV; ++e) {
                int ia, ib; poly.edge(e, ia, ib);
                V2 ea = xy(qrot(rot, poly.vertex(ia) - p.n * p.d));
                V2 eb = xy(qrot(rot, poly.vertex(ib) - p.n * p.d));
                float t;
                if (seg2d_hit(sil_a, sil_b, ea, eb, &t)) { if (t_max < t) t_max = t; }
            }
            float tm = t_max == 0.0f ? 1.0f : t_max;
            V3 q = fc.a + sil_v * tm;
            push(hs, mkhit(q, q, p.n, fc.t));
            return hs;
        }
        if (fc.t > 0.0f && fabsf(dot3(dir, p.n)) < MGFB_EPS) {
            float t_min = INF_F, t_max = 0.0f;
            bool found = false;
            for (int e = 0; e < Poly::NV; ++e) {
                int ia, ib; poly.edge(e, ia, ib);
                V2 ea = xy(qrot(rot, poly.vertex(ia) - p.n * p.d));
                V2 eb = xy(qrot(rot, poly.vertex(ib) - p.n * p.d));
                float t;
                if (seg2d_hit(sil_a, sil_b, ea, eb, &t)) {
                    found = true;
                    if (t_min > t) t_min = t;
                    if (t_max < t) t_max = t;
                }
            }
            if (found) {
                float tm = t_max == 0.0f ? 1.0f : t_max;
                V3 q = fc.a + sil_v * t_min;
                push(hs, mkhit(q, q, p.n, fc.t));
                V3 q2 = fc.a + sil_v * tm;
                push(hs, mkhit(q2, q2, p.n, fc.t));
                return hs;
            }
        }
    }
    // (4) ray (capsule origin, v) against the Minkowski sum of polygon and capsule axis (:891-1085)
    unsigned par_mask = 0;
    float best_par_t = INF_F; V3 best_par_1 = zero3(), best_par_2 = zero3();
    for (int e = 0; e < Poly::NV; ++e) {
        int ia, ib; poly.edge(e, ia, ib);
        V3 ea = poly.vertex(ia), eb = poly.vertex(ib);
        V3 ab = eb - ea;
        float ab_cd = dot3(ab, c.d);
        if (fabsf(ab_cd) != len(c.d) * len(ab)) continue;
        par_mask |= (1u << ia) | (1u << ib);
        if (ab_cd < 0.0f) { V3 tmp = ea; ea = eb; eb = tmp; }
        float m_edge = len2(ab);
        float it; V3 ip;
        if (ray_capsule(c.a, v, ea, eb - ea, c.r, &it, &ip)) {
            if (it > fminf(best_par_t, 1.0f)) continue;
            V3 tri_p = seg_closest(ea, eb, ip);
            float m_proj = len2((tri_p + c.d) - ea);
            float c_t = m_proj > m_edge ? (m_proj - m_edge) / (m_proj - len2(tri_p - ea)) : 1.0f;
            best_par_t = it; best_par_1 = tri_p; best_par_2 = tri_p + c.d * c_t;
        } else if (ray_capsule(c.a, v, ea, -c.d, c.r, &it, &ip)) {
            if (it > fminf(best_par_t, 1.0f)) continue;
            V3 d = ip - ea;
            float cap_t = -dot3(d, c.d) / len2(c.d);
            V3 tri_p = seg_closest(ea, ea + (-c.d), ip);
            V3 a2 = tri_p + c.d * cap_t;
            float m_proj = len2((tri_p + c.d) - ea);
            V3 b2 = m_proj > m_edge ? eb : tri_p + c.d;
            best_par_t = it; best_par_1 = a2; best_par_2 = b2;
        }
    }
    float best_sum_t = INF_F; V3 best_sum_p = zero3();
    for (int e = 0; e < Poly::NV; ++e) {
        int ia, ib; poly.edge(e, ia, ib);
        bool a_par = (par_mask >> ia) & 1u, b_par = (par_mask >> ib) & 1u;
        if (a_par && b_par) continue;
        V3 ea = poly.vertex(ia), eb = poly.vertex(ib);
        Tri t0; t0.a = ea + (-c.d); t0.b = ea; t0.c = eb;
        Tri t1; t1.a = ea + (-c.d); t1.b = eb; t1.c = eb + (-c.d);
        Pln pe = t1.plane();
        Sph s; s.c = c.a; s.r = c.r;
        Hit contact;
        if (!plane_msphere(pe, s, v, &contact)) continue;
        if (best_sum_t > contact.t && (t0.contains(contact.a) || t1.contains(contact.b))) {
            V3 d = contact.a - ea;
            float cap_t = -dot3(d, c.d) / len2(c.d);
            best_sum_t = contact.t; best_sum_p = contact.a + c.d * cap_t;
        } else {
            float it; V3 ip;
            if (ray_capsule(c.a, v, ea, eb - ea, c.r, &it, &ip)) {
                if (it <= 1.0f && it <= best_sum_t) { best_sum_t = it; best_sum_p = seg_closest(ea, eb, ip); }
            }
            if (ray_capsule(c.a, v, ea + (-c.d), eb - ea, c.r, &it, &ip)) {
                if (it <= 1.0f && it <= best_sum_t) { best_sum_t = it; best_sum_p = seg_closest(ea, eb, ip + c.d); }
            }
            if (!a_par && ray_capsule(c.a, v, ea, -c.d, c.r, &it, &ip)) {
                if (it <= 1.0f && it <= best_sum_t) { best_sum_t = it; best_sum_p = ea; }
            }
            if (!b_par && ray_capsule(c.a, v, eb, -c.d, c.r, &it, &ip)) {
                if (it <= 1.0f && it <= best_sum_t) { best_sum_t = it; best_sum_p = eb; }
            }
        }
    }
    if (best_sum_t < best_par_t) {
        push(hs, mkhit(best_sum_p, best_sum_p, p.n, best_sum_t));
    } else if (best_par_t != INF_F) {
        push(hs, mkhit(best_par_1, best_par_1, p.n, best_par_t));
        push(hs, mkhit(best_par_2, best_par_2, p.n, best_par_t));
    }
    return hs;
}

// ---- Sphere x Moving<Sphere> (collision.rs:1089-1141) ----
HD bool sphere_msphere(Sph self, Sph s, V3 v, Hit* out) {
    float r = self.r + s.r;
    V3 d = s.c - self.c;
    float l2 = len2(d);
    if (l2 <= r * r) {
        V3 n;
        if (l2 == 0.0f) {
            if (all_zero(v)) return false;
            n = -unit(v);
        } else {
            n = d / sqrtf(l2);
        }
        *out = mkhit(self.c + n * self.r, s.c + (-n) * s.r, n, 0.0f);
        return true;
    }
    if (len2(v) == 0.0f) return false;
    float t; V3 ip;
    if (ray_sphere(self.c, -v, s.c, r, &t, &ip)) {
        if (t <= 1.0f) {
            V3 end_c = s.c + v * t;
            V3 ba = unit(end_c - self.c);
            V3 a = self.c + ba * self.r;
            *out = mkhit(a, a, ba, t);
            return true;
        }
    }
    return false;
}
// ---- Capsule x Moving<Sphere> (collision.rs:1145-1203) ----
HD bool capsule_msphere(Cap self, Sph s, V3 v, Hit* out) {
    float r = self.r + s.r;
    V3 cp = seg_closest(self.a, self.a + self.d, s.c);
    V3 d = s.c - cp;
    float l2 = len2(d);
    if (l2 <= r * r) {
        V3 n;
        if (l2 == 0.0f) {
            if (all_zero(v)) return false;
            n = -unit(v);
        } else {
            n = d / sqrtf(l2);
        }
        *out = mkhit(cp + n * self.r, s.c + (-n) * s.r, n, 0.0f);
        return true;
    }
    if (len2(v) == 0.0f) return false;
    float t; V3 ip;
    if (ray_capsule(s.c, v, self.a, self.d, s.r + self.r, &t, &ip)) {
        if (t <= 1.0f) {
            V3 b = s.c + v * t;
            V3 a = seg_closest(self.a, self.a + self.d, b);
            V3 ba = unit(b - a);
            V3 q = a + ba * self.r;
            *out = mkhit(q, q, ba, t);
            return true;
        }
    }
    return false;
}
// ---- Sphere x Moving<Capsule>: commute (collision.rs:1143) -> Moving<Capsule> x Sphere (:1368) ----
HD bool sphere_mcapsule(Sph self, Cap c, V3 v, Hit* out) {
    Hit h;
    if (!capsule_msphere(c, self, -v, &h)) return false;
    V3 d = v * h.t;
    h.a = h.a + d; h.b = h.b + d;
    *out = flip(h);
    return true;
}
// ---- Capsule x Moving<Capsule> (collision.rs:1205-1356) ----
HD bool capsule_mcapsule(Cap self, Cap c, V3 v, Hit* out) {
    V3 sa = self.a, sb = self.a + self.d;
    V3 p1, p2;
    {
        V3 p, e;
        if (seg_seg_closest(sa, sb, c.a, c.a + v, &p)) {
            if (seg_seg_closest(sa, sb, c.a + c.d, c.a + c.d + v, &e)) { p1 = p; p2 = e; }
            else return false;
        } else { p1 = sa; p2 = sb; }
    }
    {
        V3 q;
        if (seg_seg_closest(p1, p2, c.a, c.a + c.d, &q)) {
            Sph ss; ss.c = q; ss.r = self.r;
            return sphere_mcapsule(ss, c, v, out);
        }
    }
    float d_mag2 = len2(self.d);
    float t1 = dot3(c.a - self.a, self.d) / d_mag2;
    float t2 = dot3(c.a + c.d - self.a, self.d) / d_mag2;
    float t_min, t_max; V3 c_a, c_d;
    if (t1 < t2) { t_min = t1; t_max = t2; c_a = c.a; c_d = c.d; }
    else { t_min = t2; t_max = t1; c_a = c.a + c.d; c_d = -c.d; }
    V3 h = self.a - (c_a + c_d * (-t_min / (t_max - t_min)));
    float h_len = len(h);
    Sph es;
    es.r = c.r;
    if (h_len <= self.r + c.r) {
        if (t_max <= 0.0f) { es.c = c_a + c_d; return capsule_msphere(self, es, v, out); }
        if (t_min >= 1.0f) { es.c = c_a; return capsule_msphere(self, es, v, out); }
        float s_t = (clampf3(t_min, 0.0f, 1.0f) + clampf3(t_max, 0.0f, 1.0f)) * 0.5f;
        float o_t = (s_t - t_min) / (t_max - t_min);
        V3 a_c = self.a + self.d * s_t;
        V3 b_c = c_a + c_d * o_t;
        V3 ab = b_c - a_c;
        V3 n;
        if (all_zero(ab)) {
            if (all_zero(v)) return false;
            n = -unit(v);
        } else n = unit(b_c - a_c);
        *out = mkhit(a_c + n * self.r, b_c + (-n) * c.r, n, 0.0f);
        return true;
    }
    float h_rat = (h_len - self.r - c.r) / h_len;
    float v_comp = dot3(v, h) / (h_len * h_len);
    if (v_comp < h_rat) return false;
    float coll_t = h_rat / v_comp;
    V3 v_travel = v * coll_t;
    float axis_dt = dot3(v_travel, self.d) / d_mag2;
    t_min = t_min + axis_dt;
    t_max = t_max + axis_dt;
    if (t_max <= 0.0f) { es.c = c_a + c_d; return capsule_msphere(self, es, v, out); }
    if (t_min >= 1.0f) { es.c = c_a; return capsule_msphere(self, es, v, out); }
    float s_t = (clampf3(t_min, 0.0f, 1.0f) + clampf3(t_max, 0.0f, 1.0f)) * 0.5f;
    float o_t = (s_t - t_min) / (t_max - t_min);
    V3 a_c = self.a + self.d * s_t;
    V3 b_c = c_a + c_d * o_t + v_travel;
    V3 ab = b_c - a_c;
    V3 n;
    if (all_zero(ab)) {
        if (all_zero(v)) return false;
        n = -unit(v);
    } else n = unit(b_c - a_c);
    *out = mkhit(a_c + n * self.r, b_c + (-n) * c.r, n, coll_t);
    return true;
}

}  // namespace mgfb
