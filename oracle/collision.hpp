// ORACLE (test infrastructure, NOT product code).  Restates /root/reference/src/collision.rs
// (Overlaps / Contains / Intersects / Contacts / LocalContacts, continuous narrowphase)
// operation-for-operation.  Callbacks are C++ callables invoked synchronously, like the
// reference's FnMut closures.  Citations are file:line of the reference.
#pragma once
#include "geom.hpp"

namespace mgfo {

// ---- Overlaps / Contains (collision.rs:17-147) ----
inline bool overlaps(const AABB& a, const AABB& b) {  // collision.rs:22-29
    return fabsf(a.c.x - b.c.x) <= (a.r.x + b.r.x) && fabsf(a.c.y - b.c.y) <= (a.r.y + b.r.y) &&
           fabsf(a.c.z - b.c.z) <= (a.r.z + b.r.z);
}
inline bool contains(const Plane& p, Vec3 pt) {  // collision.rs:79-83
    return relative_eq(dot(p.n, pt), p.d, COLLISION_EPSILON);
}
inline bool contains(const Triangle& t, Vec3 p) {  // collision.rs:85-100
    Vec3 v = p - t.a;
    Vec3 ac = t.c - t.a;
    Vec3 ab = t.b - t.a;
    float dot1 = dot(ac, ac), dot2 = dot(ac, ab), dot3 = dot(ac, v), dot4 = dot(ab, ab), dot5 = dot(ab, v);
    float invd = 1.0f / (dot1 * dot4 - dot2 * dot2);
    float u = (dot4 * dot3 - dot2 * dot5) * invd;
    float vv = (dot1 * dot5 - dot2 * dot3) * invd;
    return u >= 0.0f && vv >= 0.0f && (u + vv) < 1.0f;
}
inline bool contains(const Rectangle& r, Vec3 p) {  // collision.rs:102-112 (sic: absolute p)
    Vec3 n = cross(r.u[0], r.u[1]);
    return relative_eq(dot(p, n), dot(n, r.c), COLLISION_EPSILON) && fabsf(dot(p, r.u[0])) <= r.e[0] &&
           fabsf(dot(p, r.u[1])) <= r.e[1];
}
inline bool contains(const AABB& a, Vec3 p) {  // collision.rs:114-120
    return fabsf(a.c.x - p.x) <= a.r.x && fabsf(a.c.y - p.y) <= a.r.y && fabsf(a.c.z - p.z) <= a.r.z;
}
inline bool contains(const AABB& a, const AABB& rhs) {  // collision.rs:129-135
    Vec3 rhs_max = rhs.c + rhs.r;
    Vec3 rhs_min = rhs.c + (-rhs.r);
    return contains(a, rhs_max) && contains(a, rhs_min);
}

// ---- Intersects (ray casts; Ray has DT = inf, Segment DT = 1: geom.rs:819,843) ----
struct Intersection { Vec3 p; float t; };  // collision.rs:151

// collision.rs:169-184
inline bool intersection(const Ray& r, const Plane& p, Intersection* out, float DT = INF) {
    float denom = dot(p.n, r.d);
    if (denom == 0.0f) return false;
    float t = (p.d - dot(p.n, r.p)) / denom;
    if (t <= 0.0f || t > DT) return false;
    *out = {r.p + r.d * t, t};
    return true;
}
// collision.rs:202-236
inline bool intersection(const Ray& r, const AABB& a, Intersection* out, float DT = INF) {
    float t_min = 0.0f, t_max = INF;
    Vec3 p = r.p, d = r.d;
    for (int dim = 0; dim < 3; ++dim) {
        if (fabsf(d[dim]) < COLLISION_EPSILON) {
            if (fabsf(p[dim] - a.c[dim]) > a.r[dim]) return false;
        } else {
            float ood = 1.0f / d[dim];
            float t1 = (a.c[dim] - a.r[dim] - p[dim]) * ood;
            float t2 = (a.c[dim] + a.r[dim] - p[dim]) * ood;
            if (t1 > t2) { t_min = fmax_(t_min, t2); t_max = fmin_(t_max, t1); }
            else { t_min = fmax_(t_min, t1); t_max = fmin_(t_max, t2); }
            if (t_min > t_max) return false;
        }
    }
    if (t_min > DT) return false;
    *out = {p + d * t_min, t_min};
    return true;
}
// collision.rs:186-200: Particle x Polygon = plane hit that lies inside the face
template <class Poly>
inline bool intersection_poly(const Ray& r, const Poly& poly, Intersection* out, float DT = INF) {
    Intersection i;
    if (intersection(r, to_plane(poly), &i, DT) && contains(poly, i.p)) { *out = i; return true; }
    return false;
}
// collision.rs:238-247 + geom.rs:829-836 (Ray) / :853-861 (Segment): rotate the particle around the box centre,
// then the AABB test.  `segment`: the rotated direction is re-derived as b' - a' (geom.rs:856-858, :848-850).
inline bool intersection(const Ray& r, const OBB& o, Intersection* out, float DT, bool segment) {
    Vec3 p = rotate_vector(o.q, r.p - o.c) + o.c;
    Vec3 d = rotate_vector(o.q, r.d);
    if (segment) { Vec3 b = p + d; d = b - p; }
    return intersection(Ray{p, d}, AABB{o.c, o.r}, out, DT);
}
// collision.rs:249-273
inline bool intersection(const Ray& r, const Sphere& s, Intersection* out, float DT = INF) {
    Vec3 p = r.p, d = r.d;
    Vec3 m = p - s.c;
    float a = magnitude2(d);
    float b = dot(m, d);
    float c = magnitude2(m) - s.r * s.r;
    if (c > 0.0f && b > 0.0f) return false;
    float discr = b * b - a * c;
    if (discr < 0.0f) return false;
    float t = fmax_((-b - sqrtf(discr)) / a, 0.0f);
    if (t > DT) return false;
    *out = {p + t * d, t};
    return true;
}
// collision.rs:275-359
inline bool intersection(const Ray& r, const Capsule& cap, Intersection* out, float DT = INF) {
    Vec3 p = r.p, d = r.d;
    Vec3 m = p - cap.a;
    float md = dot(m, cap.d);
    float nd = dot(d, cap.d);
    float dd = dot(cap.d, cap.d);
    float nn = magnitude2(d);
    float mn = dot(m, d);
    float a = dd * nn - nd * nd;
    float k = magnitude2(m) - cap.r * cap.r;
    if (fabsf(a) < COLLISION_EPSILON) {
        float b, c;
        if (md < 0.0f) { b = mn; c = k; }
        else if (md > dd) {
            Vec3 m2 = p - (cap.a + cap.d);
            b = dot(m2, d); c = magnitude2(m2) - cap.r * cap.r;
        } else {
            return false;  // already colliding
        }
        if (c > 0.0f && b > 0.0f) return false;
        float discr = b * b - nn * c;
        if (discr < 0.0f) return false;
        float t = fmax_((-b - sqrtf(discr)) / nn, 0.0f);
        if (t > DT) return false;
        *out = {p + t * d, t};
        return true;
    }
    float c = dd * k - md * md;
    float b = dd * mn - nd * md;
    float discr = b * b - a * c;
    if (discr < 0.0f) return false;
    float t = (-b - sqrtf(discr)) / a;
    if (t < 0.0f) return false;  // behind ray
    if (md + t * nd < 0.0f) {
        if (mn > 0.0f && k > 0.0f) return false;
        float discr2 = mn * mn - nn * k;
        if (discr2 < 0.0f) return false;
        t = fmax_((-mn - sqrtf(discr2)) / nn, 0.0f);
    } else if (md + t * nd > dd) {
        Vec3 m2 = p - (cap.a + cap.d);
        float b2 = dot(m2, d);
        float c2 = magnitude2(m2) - cap.r * cap.r;
        if (c2 > 0.0f && b2 > 0.0f) return false;
        float discr2 = b2 * b2 - nn * c2;
        if (discr2 < 0.0f) return false;
        t = fmax_((-b2 - sqrtf(discr2)) / nn, 0.0f);
    }
    if (t > DT) return false;
    *out = {p + t * d, t};
    return true;
}

// ---- Contact (collision.rs:431-456) ----
struct Contact { Vec3 a, b, n; float t; };
inline Contact neg(const Contact& c) { return {c.b, c.a, -c.n, c.t}; }

// ---- Plane x Moving<Sphere> (collision.rs:521-553) ----
template <class F>
bool contacts(const Plane& pl, const Moving<Sphere>& sphere, F&& callback) {
    Sphere s = sphere.g; Vec3 v = sphere.v;
    float dist = dot(pl.n, s.c) - pl.d;
    if (fabsf(dist) <= s.r) {
        callback(Contact{s.c + (-pl.n) * dist, s.c + (-pl.n) * s.r, pl.n, 0.0f});
        return true;
    }
    float denom = dot(pl.n, v);
    if (denom * dist >= 0.0f) return false;
    float r = dist > 0.0f ? s.r : -s.r;
    float t = (r - dist) / denom;
    if (t <= 1.0f) {
        Vec3 q = s.c + t * v - r * pl.n;
        callback(Contact{q, q, pl.n, t});
        return true;
    }
    return false;
}
template <class Recv, class Arg>
bool last_contact(const Recv& a, const Arg& b, Contact* out) {  // collision.rs:477-481
    bool any = false;
    contacts(a, b, [&](const Contact& c) { *out = c; any = true; });
    return any;
}

// ---- Plane x Moving<Capsule> (collision.rs:555-605) ----
template <class F>
bool contacts(const Plane& pl, const Moving<Capsule>& capsule, F&& callback) {
    Capsule c = capsule.g; Vec3 v = capsule.v;
    float denom = dot(pl.n, normalize(c.d));
    Vec3 ctr;
    if (fabsf(denom) < COLLISION_EPSILON) {
        ctr = c.a + c.d * 0.5f;
    } else {
        float t = (pl.d - dot(pl.n, c.a)) / denom;
        if (t > 1.0f) ctr = c.a + c.d;
        else if (t < 0.0f) ctr = c.a;
        else {
            Vec3 q = c.a + c.d * t;
            float dist = dot(pl.n, c.a) - pl.d;
            Vec3 b = (dist < 0.0f ? c.a : c.a + c.d) + (-pl.n) * c.r;
            callback(Contact{q, b, pl.n, 0.0f});
            return true;
        }
    }
    Moving<Sphere> ms{Sphere{ctr, c.r}, v};
    return contacts(pl, ms, callback);
}

// ---- Polygon x Moving<Sphere> (collision.rs:610-659) ----
template <class Poly, class F>
bool poly_contacts_sphere(const Poly& self, const Moving<Sphere>& sphere, F&& callback) {
    Sphere s = sphere.g; Vec3 v = sphere.v;
    bool collision = false;
    Plane p = to_plane(self);
    contacts(p, sphere, [&](const Contact& contact) {
        if (contains(self, contact.a)) {
            collision = true;
            callback(contact);
            return;
        }
        float first_t = INF;
        Vec3 tri_p = v3(0.0f, 0.0f, 0.0f);
        if (magnitude2(v) == 0.0f) return;
        Ray ray{s.c, v};
        for (int edge_i = 0; edge_i < num_vertices(self); ++edge_i) {
            int a, b; edge(self, edge_i, &a, &b);
            Vec3 v1 = vertex(self, a), v2 = vertex(self, b);
            Capsule c{v1, v2 - v1, s.r};
            Intersection i;
            if (intersection(ray, c, &i)) {
                if (i.t <= 1.0f && i.t < first_t) {
                    first_t = i.t;
                    tri_p = closest_point(Segment{v1, v2}, i.p);
                }
            }
        }
        if (first_t != INF) {
            collision = true;
            callback(Contact{tri_p, tri_p, p.n, first_t});
        }
    });
    return collision;
}
template <class F> bool contacts(const Triangle& t, const Moving<Sphere>& s, F&& cb) { return poly_contacts_sphere(t, s, cb); }
template <class F> bool contacts(const Rectangle& t, const Moving<Sphere>& s, F&& cb) { return poly_contacts_sphere(t, s, cb); }

// ---- 2D helpers (collision.rs:667-688) ----
inline float signed_2d_tri_area(Vec2 a, Vec2 b, Vec2 c) { return (a.x - c.x) * (b.y - c.y) - (a.y - c.y) * (b.x - c.x); }
inline bool seg_2d_intersect(Vec2 a, Vec2 b, Vec2 c, Vec2 d, Vec2* op, float* ot) {
    float a1 = signed_2d_tri_area(a, b, d);
    float a2 = signed_2d_tri_area(a, b, c);
    if (a1 * a2 <= 0.0f) {
        float a3 = signed_2d_tri_area(c, d, a);
        float a4 = a3 + a2 - a1;
        if (a3 * a4 <= 0.0f) {
            float t = a3 / (a3 - a4);
            *op = a + t * (b - a);
            *ot = t;
            return true;
        }
    }
    return false;
}

// ---- Polygon x Moving<Capsule> (collision.rs:693-1086) ----
template <class Poly, class F>
bool poly_contacts_capsule(const Poly& self, const Moving<Capsule>& capsule, F&& callback) {
    const int NV = num_vertices(self);
    Capsule c = capsule.g; Vec3 v = capsule.v;
    Plane p = to_plane(self);
    // Already colliding?  (collision.rs:697-719)
    float denom = dot(p.n, normalize(c.d));
    if (fabsf(denom) > COLLISION_EPSILON) {
        float t = (p.d - dot(p.n, c.a)) / denom;
        if (t <= 1.0f && t >= 0.0f) {
            Vec3 q = c.a + c.d * t;
            if (contains(self, q)) {
                Vec3 b = (dot(p.n, c.a) - p.d < 0.0f ? c.a : c.a + c.d) + (-p.n) * c.r;
                callback(Contact{q, b, p.n, 0.0f});
                return true;
            }
        }
    }
    // Seed contact from the end spheres (collision.rs:721-764)
    Moving<Sphere> start_sphere{Sphere{c.a, c.r}, v};
    Moving<Sphere> end_sphere{Sphere{c.a + c.d, c.r}, v};
    bool have = false; Contact fc{}; Vec3 dir{}; bool checked_contains = false;
    {
        Contact c1, c2;
        if (last_contact(p, start_sphere, &c1)) {
            if (last_contact(p, end_sphere, &c2)) {
                if (c2.t < c1.t) {
                    have = true; fc = c2; dir = -c.d; checked_contains = false;
                } else {
                    if (c2.t == 0.0f) {
                        bool contains_1 = contains(self, c1.a);
                        bool contains_2 = contains(self, c2.a);
                        if (contains_1 && contains_2) {
                            callback(c2);
                            callback(c1);
                            return true;
                        } else if (contains_1) {
                            have = true; fc = c1; dir = c.d; checked_contains = true;
                        } else if (contains_2) {
                            have = true; fc = c2; dir = -c.d; checked_contains = true;
                        }
                    } else {
                        have = true; fc = c1; dir = c.d; checked_contains = false;
                    }
                }
            } else {
                have = true; fc = c1; dir = c.d; checked_contains = false;
            }
        } else if (last_contact(p, end_sphere, &c1)) {
            have = true; fc = c1; dir = -c.d; checked_contains = false;
        }
    }
    if (have) {  // collision.rs:767-890
        const Contact contact = fc;
        Vec3 silhouette_v = dir - p.n * dot(dir, p.n) / magnitude2(p.n);
        Vec3 n_xy = v3(0.0f, 0.0f, 1.0f);
        Quat plane_rot = from_arc(p.n, n_xy);
        Vec2 silhouette_a = truncate(rotate_vector(plane_rot, contact.a + (-p.n) * p.d));
        Vec2 silhouette_b = truncate(rotate_vector(plane_rot, contact.a + silhouette_v - p.n * p.d));
        if (checked_contains || contains(self, contact.a)) {
            callback(contact);
            if (fabsf(dot(dir, p.n)) >= COLLISION_EPSILON) return true;
            float t_max = 0.0f;
            for (int edge_i = 0; edge_i < NV; ++edge_i) {
                int a, b; edge(self, edge_i, &a, &b);
                Vec2 edge_a = truncate(rotate_vector(plane_rot, vertex(self, a) - p.n * p.d));
                Vec2 edge_b = truncate(rotate_vector(plane_rot, vertex(self, b) - p.n * p.d));
                Vec2 ip; float t;
                if (seg_2d_intersect(silhouette_a, silhouette_b, edge_a, edge_b, &ip, &t)) {
                    if (t_max < t) t_max = t;
                }
            }
            float t_max2 = t_max == 0.0f ? 1.0f : t_max;
            Vec3 q = contact.a + silhouette_v * t_max2;
            callback(Contact{q, q, p.n, contact.t});
            return true;
        }
        if (contact.t > 0.0f && fabsf(dot(dir, p.n)) < COLLISION_EPSILON) {
            float t_min = INF, t_max = 0.0f;
            bool found = false;
            for (int edge_i = 0; edge_i < NV; ++edge_i) {
                int a, b; edge(self, edge_i, &a, &b);
                Vec2 edge_a = truncate(rotate_vector(plane_rot, vertex(self, a) - p.n * p.d));
                Vec2 edge_b = truncate(rotate_vector(plane_rot, vertex(self, b) - p.n * p.d));
                Vec2 ip; float t;
                if (seg_2d_intersect(silhouette_a, silhouette_b, edge_a, edge_b, &ip, &t)) {
                    found = true;
                    if (t_min > t) t_min = t;
                    if (t_max < t) t_max = t;
                }
            }
            if (found) {
                float t_max2 = t_max == 0.0f ? 1.0f : t_max;
                Vec3 q = contact.a + silhouette_v * t_min;
                float t = contact.t;
                callback(Contact{q, q, p.n, t});
                Vec3 q2 = contact.a + silhouette_v * t_max2;
                callback(Contact{q2, q2, p.n, t});
                return true;
            }
        }
    }
    // Minkowski-sum fallback (collision.rs:891-1085)
    if (NV > 64) return false;
    uint64_t parallel_edge_vert = 0;
    float best_par_t = INF; Vec3 best_par_1 = v3(0, 0, 0), best_par_2 = v3(0, 0, 0);
    for (int edge_i = 0; edge_i < NV; ++edge_i) {
        int a, b; edge(self, edge_i, &a, &b);
        Vec3 edge_a = vertex(self, a), edge_b = vertex(self, b);
        Vec3 ab = edge_b - edge_a;
        float ab_cd = dot(ab, c.d);
        if (fabsf(ab_cd) != magnitude(c.d) * magnitude(ab)) continue;  // not parallel
        parallel_edge_vert |= (uint64_t(1) << a);
        parallel_edge_vert |= (uint64_t(1) << b);
        Ray ray{c.a, v};
        if (ab_cd < 0.0f) { Vec3 tmp = edge_a; edge_a = edge_b; edge_b = tmp; }
        Capsule edge_sum{edge_a, edge_b - edge_a, c.r};
        float m_edge = magnitude2(ab);
        Intersection inter;
        if (intersection(ray, edge_sum, &inter)) {
            if (inter.t > fmin_(best_par_t, 1.0f)) continue;
            Vec3 tri_p = closest_point(Segment{edge_a, edge_b}, inter.p);
            float m_proj = magnitude2((tri_p + c.d) - edge_a);
            float c_t = m_proj > m_edge ? (m_proj - m_edge) / (m_proj - magnitude2(tri_p - edge_a)) : 1.0f;
            Vec3 q = tri_p + c.d * c_t;
            best_par_t = inter.t; best_par_1 = tri_p; best_par_2 = q;
        } else if (intersection(ray, Capsule{edge_a, -c.d, c.r}, &inter)) {
            if (inter.t > fmin_(best_par_t, 1.0f)) continue;
            Vec3 d = inter.p - edge_a;
            float capsule_t = -dot(d, c.d) / magnitude2(c.d);
            Vec3 tri_p = closest_point(Segment{edge_a, edge_a + (-c.d)}, inter.p);
            Vec3 a2 = tri_p + c.d * capsule_t;
            float m_proj = magnitude2((tri_p + c.d) - edge_a);
            Vec3 b2 = m_proj > m_edge ? edge_b : tri_p + c.d;
            best_par_t = inter.t; best_par_1 = a2; best_par_2 = b2;
        }
    }
    float best_sum_t = INF; Vec3 best_sum_p = v3(0, 0, 0);
    for (int edge_i = 0; edge_i < NV; ++edge_i) {
        int a, b; edge(self, edge_i, &a, &b);
        bool a_on_parallel_edge = (parallel_edge_vert >> a) & 1;
        bool b_on_parallel_edge = (parallel_edge_vert >> b) & 1;
        if (a_on_parallel_edge && b_on_parallel_edge) continue;
        Vec3 edge_a = vertex(self, a), edge_b = vertex(self, b);
        Triangle tris[2] = {Triangle{edge_a + (-c.d), edge_a, edge_b},
                            Triangle{edge_a + (-c.d), edge_b, edge_b + (-c.d)}};
        Plane pe = to_plane(tris[1]);
        Sphere s{c.a, c.r};
        contacts(pe, Moving<Sphere>{s, v}, [&](const Contact& contact) {
            if (best_sum_t > contact.t && (contains(tris[0], contact.a) || contains(tris[1], contact.b))) {
                Vec3 d = contact.a - edge_a;
                float capsule_t = -dot(d, c.d) / magnitude2(c.d);
                best_sum_t = contact.t; best_sum_p = contact.a + c.d * capsule_t;
            } else {
                Ray ray{c.a, v};
                Capsule bottom_edge{edge_a, edge_b - edge_a, c.r};
                Intersection inter;
                if (intersection(ray, bottom_edge, &inter)) {
                    if (inter.t <= 1.0f && inter.t <= best_sum_t) {
                        Vec3 q = closest_point(Segment{edge_a, edge_b}, inter.p);
                        best_sum_t = inter.t; best_sum_p = q;
                    }
                }
                Capsule top_edge{edge_a + (-c.d), edge_b - edge_a, c.r};
                if (intersection(ray, top_edge, &inter)) {
                    if (inter.t <= 1.0f && inter.t <= best_sum_t) {
                        Vec3 plane_p = inter.p + c.d;
                        Vec3 q = closest_point(Segment{edge_a, edge_b}, plane_p);
                        best_sum_t = inter.t; best_sum_p = q;
                    }
                }
                const Vec3 verts[2] = {edge_a, edge_b};
                const bool is_par[2] = {a_on_parallel_edge, b_on_parallel_edge};
                for (int k = 0; k < 2; ++k) {
                    if (is_par[k]) continue;
                    Capsule cap{verts[k], -c.d, c.r};
                    if (intersection(ray, cap, &inter)) {
                        if (inter.t <= 1.0f && inter.t <= best_sum_t) {
                            best_sum_t = inter.t; best_sum_p = verts[k];
                        }
                    }
                }
            }
        });
    }
    if (best_sum_t < best_par_t) {
        callback(Contact{best_sum_p, best_sum_p, p.n, best_sum_t});
    } else if (best_par_t != INF) {
        callback(Contact{best_par_1, best_par_1, p.n, best_par_t});
        callback(Contact{best_par_2, best_par_2, p.n, best_par_t});
    } else {
        return false;
    }
    return true;
}
template <class F> bool contacts(const Triangle& t, const Moving<Capsule>& s, F&& cb) { return poly_contacts_capsule(t, s, cb); }
template <class F> bool contacts(const Rectangle& t, const Moving<Capsule>& s, F&& cb) { return poly_contacts_capsule(t, s, cb); }

// ---- Sphere x Moving<Sphere> (collision.rs:1089-1141) ----
template <class F>
bool contacts(const Sphere& self, const Moving<Sphere>& sphere, F&& callback) {
    Sphere s = sphere.g; Vec3 v = sphere.v;
    float r = self.r + s.r;
    Vec3 d = s.c - self.c;
    float len = magnitude2(d);
    if (len <= r * r) {
        Vec3 n;
        if (len == 0.0f) {
            if (is_zero(v)) return false;
            n = -normalize(v);
        } else {
            n = d / sqrtf(len);
        }
        callback(Contact{self.c + n * self.r, s.c + (-n) * s.r, n, 0.0f});
        return true;
    }
    float l = magnitude2(v);
    if (l == 0.0f) return false;
    Ray ray{self.c, -v};
    Intersection is;
    if (intersection(ray, Sphere{s.c, r}, &is)) {
        if (is.t <= 1.0f) {
            Vec3 end_c = s.c + v * is.t;
            Vec3 ba = normalize(end_c - self.c);
            Vec3 a = self.c + ba * self.r;
            callback(Contact{a, a, ba, is.t});
            return true;
        }
    }
    return false;
}

// ---- Capsule x Moving<Sphere> (collision.rs:1145-1203) ----
template <class F>
bool contacts(const Capsule& self, const Moving<Sphere>& sphere, F&& callback) {
    Sphere s = sphere.g; Vec3 v = sphere.v;
    float r = self.r + s.r;
    Vec3 closest_pt = closest_point(Segment{self.a, self.a + self.d}, s.c);
    Vec3 d = s.c - closest_pt;
    float len = magnitude2(d);
    if (len <= r * r) {
        Vec3 n;
        if (len == 0.0f) {
            if (is_zero(v)) return false;
            n = -normalize(v);
        } else {
            n = d / sqrtf(len);
        }
        callback(Contact{closest_pt + n * self.r, s.c + (-n) * s.r, n, 0.0f});
        return true;
    }
    float l = magnitude2(v);
    if (l == 0.0f) return false;
    Ray ray{s.c, v};
    Intersection is;
    if (intersection(ray, Capsule{self.a, self.d, s.r + self.r}, &is)) {
        if (is.t <= 1.0f) {
            Vec3 b = s.c + v * is.t;
            Vec3 a = closest_point(Segment{self.a, self.a + self.d}, b);
            Vec3 ba = normalize(b - a);
            Vec3 q = a + ba * self.r;
            callback(Contact{q, q, ba, is.t});
            return true;
        }
    }
    return false;
}

template <class F> bool contacts(const Capsule& self, const Moving<Capsule>& capsule, F&& callback);

// collision.rs:1368-1382  impl Contacts<Arg> for Moving<Recv>
template <class Recv, class Arg, class F>
bool moving_contacts_static(const Moving<Recv>& self, const Arg& rhs, F&& callback) {
    Moving<Arg> rhs_moving{rhs, -self.v};
    return contacts(self.g, rhs_moving, [&](const Contact& c) {
        Vec3 d = self.v * c.t;
        callback(Contact{c.a + d, c.b + d, c.n, c.t});
    });
}
// commute_contacts!{ Sphere, Moving<Capsule> } (collision.rs:1143, 484-494):
// Sphere.contacts(&Moving<Capsule>) = Moving<Capsule>.contacts(&Sphere) negated (-> :1368).
template <class F>
bool contacts(const Sphere& self, const Moving<Capsule>& rhs, F&& callback) {
    return moving_contacts_static(rhs, self, [&](const Contact& c) { callback(neg(c)); });
}

// ---- Capsule x Moving<Capsule> (collision.rs:1205-1356) ----
template <class F>
bool contacts(const Capsule& self, const Moving<Capsule>& capsule, F&& callback) {
    Capsule c = capsule.g; Vec3 v = capsule.v;
    Segment self_seg{self.a, self.a + self.d};
    Vec3 p1, p2;
    {
        Vec3 p, e, dummy;
        if (closest_pts_seg(self_seg, Segment{c.a, c.a + v}, &p, &dummy)) {
            if (closest_pts_seg(self_seg, Segment{c.a + c.d, c.a + c.d + v}, &e, &dummy)) {
                p1 = p; p2 = e;
            } else {
                return false;
            }
        } else {
            p1 = self.a; p2 = self.a + self.d;
        }
    }
    Segment self_seg2{p1, p2};
    {
        Vec3 q, dummy;
        if (closest_pts_seg(self_seg2, Segment{c.a, c.a + c.d}, &q, &dummy)) {
            Sphere ss{q, self.r};
            return contacts(ss, capsule, callback);
        }
    }
    // Parallel capsules (collision.rs:1234-1355)
    float d_mag2 = magnitude2(self.d);
    float t1 = dot(c.a - self.a, self.d) / d_mag2;
    float t2 = dot(c.a + c.d - self.a, self.d) / d_mag2;
    float t_min, t_max; Vec3 c_a, c_d;
    if (t1 < t2) { t_min = t1; t_max = t2; c_a = c.a; c_d = c.d; }
    else { t_min = t2; t_max = t1; c_a = c.a + c.d; c_d = -c.d; }
    Vec3 h = self.a - (c_a + c_d * (-t_min / (t_max - t_min)));
    float h_len = magnitude(h);
    if (h_len <= self.r + c.r) {
        if (t_max <= 0.0f) return contacts(self, Moving<Sphere>{Sphere{c_a + c_d, c.r}, v}, callback);
        if (t_min >= 1.0f) return contacts(self, Moving<Sphere>{Sphere{c_a, c.r}, v}, callback);
        float s_t = (clampf(t_min, 0.0f, 1.0f) + clampf(t_max, 0.0f, 1.0f)) * 0.5f;
        float o_t = (s_t - t_min) / (t_max - t_min);
        Vec3 a_c = self.a + self.d * s_t;
        Vec3 b_c = c_a + c_d * o_t;
        Vec3 ab = b_c - a_c;
        Vec3 n;
        if (is_zero(ab)) {
            if (is_zero(v)) return false;
            n = -normalize(v);
        } else {
            n = normalize(b_c - a_c);
        }
        callback(Contact{a_c + n * self.r, b_c + (-n) * c.r, n, 0.0f});
        return true;
    }
    float h_rat = (h_len - self.r - c.r) / h_len;
    float v_comp = dot(v, h) / (h_len * h_len);
    if (v_comp < h_rat) return false;
    float coll_t = h_rat / v_comp;
    Vec3 v_travel = v * coll_t;
    float axis_t_delta = dot(v_travel, self.d) / d_mag2;
    t_min = t_min + axis_t_delta;
    t_max = t_max + axis_t_delta;
    if (t_max <= 0.0f) return contacts(self, Moving<Sphere>{Sphere{c_a + c_d, c.r}, v}, callback);
    if (t_min >= 1.0f) return contacts(self, Moving<Sphere>{Sphere{c_a, c.r}, v}, callback);
    float s_t = (clampf(t_min, 0.0f, 1.0f) + clampf(t_max, 0.0f, 1.0f)) * 0.5f;
    float o_t = (s_t - t_min) / (t_max - t_min);
    Vec3 a_c = self.a + self.d * s_t;
    Vec3 b_c = c_a + c_d * o_t + v_travel;
    Vec3 ab = b_c - a_c;
    Vec3 n;
    if (is_zero(ab)) {
        if (is_zero(v)) return false;
        n = -normalize(v);
    } else {
        n = normalize(b_c - a_c);
    }
    callback(Contact{a_c + n * self.r, b_c + (-n) * c.r, n, coll_t});
    return true;
}

// commute_contacts!{ Moving<Sphere|Capsule>, Triangle|Rectangle|Plane } (collision.rs:607-608, 661-664)
template <class Shp, class Poly, class F>
bool moving_contacts_poly(const Moving<Shp>& self, const Poly& rhs, F&& callback) {
    return contacts(rhs, self, [&](const Contact& c) { callback(neg(c)); });
}

// collision.rs:1387-1401  impl Contacts<Moving<Arg>> for Moving<Recv>
template <class Recv, class Arg, class F>
bool moving_contacts_moving(const Moving<Recv>& self, const Moving<Arg>& rhs, F&& callback) {
    Vec3 v_a = self.v, v_b = rhs.v;
    return contacts(self.g, Moving<Arg>{rhs.g, v_b - v_a}, [&](const Contact& c) {
        Vec3 a = c.a + v_a * c.t;
        Vec3 b = c.b + v_a * c.t;
        callback(Contact{a, b, c.n, c.t});
    });
}

// ---- LocalContact (collision.rs:1410-1432) ----
struct LocalContact { Vec3 local_a, local_b; Contact global; };

}  // namespace mgfo
