// ORACLE (test infrastructure, NOT product code).  Restates src/compound.rs:232-352 (Compound: an aggregate of Components
// with a displacement and a rotation, kept in a BVH<AABB, Component>) and the pieces of geom.rs / bounds.rs it leans on
// (Volumetric::rotate / rotate_about geom.rs:928-1014, Shape::closest_point geom.rs:751-797, BoundedBy<Sphere> bounds.rs:291-310).
// Citations are file:line of the reference.  Pinned by the reference's own test compound.rs:362-388 (oracle/kat.cpp).
#pragma once
#include <vector>
#include "dynamics.hpp"

namespace mgfo {

// geom.rs:940-986  Volumetric for AABB: the box of the eight rotated corners
inline AABB aabb_rotate(const AABB& s, Quat rot) {
    Vec3 vx = rotate_vector(rot, v3(s.r.x, 0.0f, 0.0f)), vy = rotate_vector(rot, v3(0.0f, s.r.y, 0.0f)), vz = rotate_vector(rot, v3(0.0f, 0.0f, s.r.z));
    Vec3 p[8] = {s.c + (vx + vy + vz), s.c + (vx + vy - vz), s.c + (vx - vy + vz), s.c + (vx - vy - vz),
                 s.c + (-vx + vy + vz), s.c + (-vx + vy - vz), s.c + (-vx - vy + vz), s.c + (-vx - vy - vz)};
    auto fold = [&](int ax, bool mn) {   // p1.min(p2.min(p3.min(...p8)))
        auto get = [&](int i) { return ax == 0 ? p[i].x : (ax == 1 ? p[i].y : p[i].z); };
        float acc = get(7);
        for (int i = 6; i >= 0; --i) acc = mn ? fminf(get(i), acc) : fmaxf(get(i), acc);
        return acc;
    };
    Vec3 lower = v3(fold(0, true), fold(1, true), fold(2, true)), upper = v3(fold(0, false), fold(1, false), fold(2, false));
    return AABB{(upper + lower) / 2.0f, (upper - lower) / 2.0f};
}
// geom.rs:999-1014  Volumetric for Sphere / Capsule, compound.rs:56-63 for Component
inline Component rotate(const Component& k, Quat r) {
    if (k.kind == Component::SPHERE) return k;
    Vec3 c = center(k.c);
    return Component::capsule(Capsule{c + rotate_vector(r, k.c.a - c), rotate_vector(r, k.c.d), k.c.r});
}
// geom.rs:933-937
inline Component rotate_about(Component k, Quat r, Vec3 p) {
    Vec3 c = center(k);
    set_pos(k, p + rotate_vector(r, c - p));
    return rotate(k, r);
}
inline Vec3 closest_point(const Capsule& c, Vec3 to) {   // geom.rs:791-796
    Segment seg{c.a, c.a + c.d};
    return closest_point(Sphere{closest_point(seg, to), c.r}, to);
}
inline Vec3 closest_point(const Component& k, Vec3 to) { return k.kind == Component::SPHERE ? closest_point(k.s, to) : closest_point(k.c, to); }
inline bool intersection(const Ray& r, const Component& k, Intersection* out, float DT) {   // compound.rs:150-157
    return k.kind == Component::SPHERE ? intersection(r, k.s, out, DT) : intersection(r, k.c, out, DT);
}
// compound.rs:163-177 impl_component_collision!{Sphere, Capsule, Triangle, Rectangle}: Recv.contacts(&Moving<Component>)
template <class Recv, class F>
bool contacts_moving_component(const Recv& self, const MovingComponent& rhs, F&& cb) {
    if (rhs.g.kind == Component::SPHERE) return contacts(self, Moving<Sphere>{rhs.g.s, rhs.v}, cb);
    return contacts(self, Moving<Capsule>{rhs.g.c, rhs.v}, cb);
}
// collision.rs:1368-1383 with Arg = Component: Moving<Recv>.contacts(&Component)
template <class Recv, class F>
bool moving_contacts_component(const Moving<Recv>& self, const Component& rhs, F&& cb) {
    MovingComponent rhs_moving{rhs, -self.v};
    return contacts_moving_component(self.g, rhs_moving, [&](const Contact& c) {
        Vec3 d = self.v * c.t;
        cb(Contact{c.a + d, c.b + d, c.n, c.t});
    });
}
inline AABB bounds(const Rectangle& r) {   // bounds.rs:156-168
    Vec3 p1 = r.c + r.u[0] * r.e[0], p2 = r.c + r.u[1] * r.e[1];
    return AABB{r.c, v3(fmaxf(fabsf(p1.x - r.c.x), fabsf(p2.x - r.c.x)), fmaxf(fabsf(p1.y - r.c.y), fabsf(p2.y - r.c.y)),
                        fmaxf(fabsf(p1.z - r.c.z), fabsf(p2.z - r.c.z)))};
}

// compound.rs:232-263
struct Compound {
    Vec3 disp{0, 0, 0};
    Quat rot = quat_one();
    std::vector<size_t> shapes;
    BVH<Component> bvh;
    explicit Compound(const std::vector<Component>& components) {
        for (const Component& c : components) shapes.push_back(bvh.insert(bounds(c), c));
    }
    AABB bounds_aabb() const { return aabb_translate(aabb_rotate(bvh[bvh.root()], rot), disp); }   // :277-281
    Sphere bounds_sphere() const {                                                                  // :283-288, bounds.rs:291-298
        const AABB& b = bvh[bvh.root()];
        return Sphere{b.c + disp, magnitude(b.r)};
    }
    Vec3 closest_point(Vec3 to) const {   // :299-311 (sic: the components are NOT moved by disp / rot)
        Vec3 best_p = v3(0, 0, 0); float best_dist = INF;
        for (size_t s : shapes) {
            Vec3 np = mgfo::closest_point(bvh.pool[s].val, to);
            float nd = magnitude2(to - np);
            if (nd < best_dist) { best_p = np; best_dist = nd; }
        }
        return best_p;
    }
    // :314-337 Intersects<Compound> for P; pos/dir/DT describe the particle (Ray: DT = inf; Segment: DT = 1)
    bool intersection(const Ray& self, float DT, Intersection* out) const {
        Quat conj_rot = conjugate(rot);
        Ray r{rotate_point(conj_rot, self.p + (-disp)) + disp, rotate_vector(conj_rot, self.d)};
        bool have = false; Intersection res{};
        bvh.raytrace(r, [&](const Component& comp, const Intersection& inter) {
            if (inter.t > DT) return;
            Component shape = rotate(comp, rot) + disp;   // (sic: rotate, not rotate_about)
            Intersection hit;
            if (mgfo::intersection(self, shape, &hit, DT)) {
                if (have && hit.t > res.t) return;
                res = hit; have = true;
            }
        });   // the argument is a Ray whatever P is: its own DT is infinite (bvh.rs:345)
        if (have) *out = res;
        return have;
    }
    // :339-357 Contacts<RHS> for Compound, RHS in {Moving<Sphere>, Moving<Capsule>, Moving<Triangle>, Moving<Rectangle>}
    template <class Recv, class F>
    bool contacts(const Moving<Recv>& rhs, F&& callback) const {
        Quat conj_rot = conjugate(rot);
        AABB rhs_bounds = aabb_rotate(bounds(rhs), conj_rot);
        Vec3 rhs_center = rhs_bounds.c;
        Vec3 bounds_disp = rotate_point(conj_rot, rhs_center + (-disp)) + disp;
        rhs_bounds.c = rhs_bounds.c + (bounds_disp - rhs_bounds.c);   // Shape::set_pos, geom.rs:459-462
        bool collided = false;
        bvh.query(rhs_bounds, [&](const Component& comp) {
            Component shape = rotate_about(comp, rot, v3(0, 0, 0)) + disp;
            moving_contacts_component(rhs, shape, [&](const Contact& c) { collided = true; callback(neg(c)); });
        });
        return collided;
    }
};

}  // namespace mgfo
