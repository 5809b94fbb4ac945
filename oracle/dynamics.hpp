// ORACLE (test infrastructure, NOT product code).  Restates the reference's
// src/compound.rs (Component / ComponentConstructor / dispatch), src/mesh.rs (Mesh),
// src/manifold.rs, src/solver.rs and src/physics.rs operation-for-operation.
// Citations are file:line of the reference.
#pragma once
#include <algorithm>
#include <vector>
#include "bvh.hpp"

namespace mgfo {

// ---- compound.rs:33-37 Component ----
struct Component {
    enum Kind { SPHERE = 0, CAPSULE = 1 } kind;
    Sphere s; Capsule c;
    static Component sphere(const Sphere& s) { Component k; k.kind = SPHERE; k.s = s; k.c = Capsule{}; return k; }
    static Component capsule(const Capsule& c) { Component k; k.kind = CAPSULE; k.c = c; k.s = Sphere{}; return k; }
};
inline Vec3 center(const Component& k) { return k.kind == Component::SPHERE ? center(k.s) : center(k.c); }
inline AABB bounds(const Component& k) { return k.kind == Component::SPHERE ? bounds(k.s) : bounds(k.c); }
inline Component operator-(Component k, Vec3 v) {  // compound.rs:88-96 -> impl_shape_reqs Sub: c + -v
    if (k.kind == Component::SPHERE) k.s.c = k.s.c + (-v); else k.c.a = k.c.a + (-v);
    return k;
}
inline Component operator+(Component k, Vec3 v) {
    if (k.kind == Component::SPHERE) k.s.c = k.s.c + v; else k.c.a = k.c.a + v;
    return k;
}
// Shape::set_pos (geom.rs:459-462)
inline void set_pos(Component& k, Vec3 p) { Vec3 disp = p - center(k); k = k + disp; }

// compound.rs:211-228
struct ComponentConstructor { Component::Kind kind; float r, half_h; };
inline Component construct(const ComponentConstructor& cc, Vec3 p, Quat rot) {
    if (cc.kind == Component::SPHERE) return Component::sphere(Sphere{p, cc.r});
    Vec3 d = rotate_vector(rot, v3(0.0f, 1.0f, 0.0f) * cc.half_h);
    return Component::capsule(Capsule{p + (-d), d * 2.0f, cc.r});
}
// compound.rs:42-53
inline void deconstruct(const Component& k, Vec3* x, Quat* q, ComponentConstructor* cc) {
    if (k.kind == Component::SPHERE) {
        *x = k.s.c; *q = quat_one(); *cc = {Component::SPHERE, k.s.r, 0.0f};
    } else {
        float h = magnitude(k.c.d);
        Quat rot = from_arc(v3(0.0f, 1.0f, 0.0f) * h, k.c.d);
        *x = k.c.a + k.c.d * 0.5f; *q = rot; *cc = {Component::CAPSULE, k.c.r, h * 0.5f};
    }
}

typedef Moving<Component> MovingComponent;

// compound.rs:179-190: impl<RHS> Contacts<RHS> for Moving<Component>, RHS = Triangle / Rectangle
template <class Poly, class F>
bool component_contacts_poly(const MovingComponent& self, const Poly& rhs, F&& callback) {
    if (self.g.kind == Component::SPHERE)
        return contacts(rhs, Moving<Sphere>{self.g.s, self.v}, [&](const Contact& c) { callback(neg(c)); });
    return contacts(rhs, Moving<Capsule>{self.g.c, self.v}, [&](const Contact& c) { callback(neg(c)); });
}
// compound.rs:179-190 applied twice with RHS = Moving<Component>, then collision.rs:1387.
// (the two negations cancel.)
template <class F>
bool component_contacts_component(const MovingComponent& self, const MovingComponent& rhs, F&& callback) {
    auto inner = [&](auto const& ga) -> bool {   // ga: Sphere or Capsule of `self`
        Moving<std::decay_t<decltype(ga)>> ma{ga, self.v};
        // rhs.contacts(&ma, |c| callback(-c)) : Moving<Component> as receiver again (compound.rs:184)
        auto cb1 = [&](const Contact& c) { callback(neg(c)); };
        auto cb2 = [&](const Contact& c) { cb1(neg(c)); };
        if (rhs.g.kind == Component::SPHERE)
            return moving_contacts_moving(ma, Moving<Sphere>{rhs.g.s, rhs.v}, cb2);
        return moving_contacts_moving(ma, Moving<Capsule>{rhs.g.c, rhs.v}, cb2);
    };
    if (self.g.kind == Component::SPHERE) return inner(self.g.s);
    return inner(self.g.c);
}
// compound.rs:192-207  LocalContacts<Moving<Component>> for Moving<Component>
template <class F>
bool local_contacts(const MovingComponent& self, const MovingComponent& rhs, F&& callback) {
    return component_contacts_component(self, rhs, [&](const Contact& c) {
        callback(LocalContact{c.a + (-(center(self.g) + self.v * c.t)), c.b + (-(center(rhs.g) + rhs.v * c.t)), c});
    });
}

// ---- mesh.rs:32-73 ----
struct Mesh {
    Vec3 x{0, 0, 0};
    std::vector<Vec3> verts;
    struct Face { size_t a, b, c; };
    std::vector<Face> faces;
    BVH<size_t> bvh;
    size_t push_vert(Vec3 p) { verts.push_back(p); return verts.size() - 1; }
    size_t push_face(size_t a, size_t b, size_t c) {
        Triangle tri{verts[a], verts[b], verts[c]};
        size_t index = faces.size();
        faces.push_back({a, b, c});
        bvh.insert(bounds(tri), index);
        return index;
    }
    void set_pos(Vec3 p) { Vec3 disp = p - x; x += disp; }  // Shape::set_pos via AddAssign (mesh.rs:75-79)
};
inline Vec3 center(const Mesh& m) { return m.x; }
// mesh.rs:115-139  Contacts<RHS> for Mesh, RHS = Moving<Component>
template <class F>
bool contacts(const Mesh& mesh, const MovingComponent& rhs, F&& callback) {
    bool collided = false;
    mesh.bvh.query(aabb_sub(bounds(rhs), mesh.x), [&](const size_t& face_index) {
        Mesh::Face f = mesh.faces[face_index];
        Triangle tri{mesh.verts[f.a] + mesh.x, mesh.verts[f.b] + mesh.x, mesh.verts[f.c] + mesh.x};
        component_contacts_poly(rhs, tri, [&](const Contact& c) {
            collided = true;
            callback(Contact{c.b, c.a, -c.n, c.t});
        });
    });
    return collided;
}
// collision.rs:1490-1506  LocalContacts<Arg> for Moving<Recv>, Arg = Mesh
template <class F>
bool local_contacts(const MovingComponent& self, const Mesh& rhs, F&& callback) {
    return contacts(rhs, self, [&](const Contact& c) {
        Vec3 a_c = center(self.g) + self.v * c.t;
        Vec3 b_c = center(rhs);
        callback(LocalContact{c.b + (-a_c), c.a + (-b_c), neg(c)});
    });
}

// ---- manifold.rs ----
static const float PERSISTENT_THRESHOLD_SQ = 0.5f;  // manifold.rs:38
struct ContactPruner {  // manifold.rs:42-107
    float min_col_time = INF;
    std::vector<LocalContact> contacts;
    void push(const LocalContact& nc) {
        if (nc.global.t < min_col_time - COLLISION_EPSILON) {
            contacts.clear();
            contacts.push_back(nc);
            min_col_time = nc.global.t;
            return;
        } else if (nc.global.t > min_col_time + COLLISION_EPSILON) {
            return;
        }
        for (LocalContact& old : contacts) {
            Vec3 ra = nc.global.a - old.global.a;
            Vec3 rb = nc.global.b - old.global.b;
            if (magnitude2(ra) <= PERSISTENT_THRESHOLD_SQ || magnitude2(rb) <= PERSISTENT_THRESHOLD_SQ) {
                float prev_dist = magnitude2(old.local_a) + magnitude2(old.local_b);
                float new_dist = magnitude2(nc.local_a) + magnitude2(nc.local_b);
                if (prev_dist < new_dist) old = nc;
                return;
            }
        }
        contacts.push_back(nc);
    }
};
struct Manifold {  // manifold.rs:112-118
    float time;
    Vec3 normal;
    Vec3 tangent_vector[2];
    int ncontacts;
    Vec3 local_a[4], local_b[4];
    std::vector<std::pair<Vec3, Vec3>> spill;  // SmallVec spills to the heap beyond 4
    int len() const { return ncontacts; }
    std::pair<Vec3, Vec3> contact(int i) const {
        return i < 4 ? std::make_pair(local_a[i], local_b[i]) : spill[i - 4];
    }
    void push(Vec3 a, Vec3 b) {
        if (ncontacts < 4) { local_a[ncontacts] = a; local_b[ncontacts] = b; } else spill.push_back({a, b});
        ncontacts++;
    }
};
inline Manifold manifold_from(const LocalContact& lc) {  // manifold.rs:120-129
    Manifold m; m.ncontacts = 0;
    m.time = lc.global.t; m.normal = lc.global.n;
    compute_basis(lc.global.n, m.tangent_vector);
    m.push(lc.local_a, lc.local_b);
    return m;
}
inline Manifold manifold_from(const ContactPruner& pr) {  // manifold.rs:131-148
    Manifold m; m.ncontacts = 0;
    Vec3 sum = v3(0.0f, 0.0f, 0.0f);
    for (const LocalContact& lc : pr.contacts) { m.push(lc.local_a, lc.local_b); sum = sum + lc.global.n; }
    Vec3 avg = sum / (float)pr.contacts.size();
    m.time = pr.min_col_time; m.normal = avg;
    compute_basis(avg, m.tangent_vector);
    return m;
}

// ---- physics.rs ----
struct Velocity { Vec3 linear, angular; };
struct RigidBodyInfo { Vec3 x; float restitution, friction, inv_mass; Mat3 inv_moment; };
struct RigidBodyRef {  // physics.rs:159-162
    bool dynamic; size_t i; Vec3 center; float friction;
    static RigidBodyRef Dynamic(size_t i) { return {true, i, {0, 0, 0}, 0.0f}; }
    static RigidBodyRef Static(Vec3 c, float f) { return {false, 0, c, f}; }
};

// physics.rs:30-46
inline Mat3 tensor(const Sphere& s, float m) {
    float i = 0.4f * m * s.r * s.r;
    Mat3 I = mat3_new(i, 0, 0, 0, i, 0, 0, 0, i);
    Vec3 disp = s.c;
    Mat3 outer = from_cols(disp * disp.x, disp * disp.y, disp * disp.z);
    return I + m * (mat3_one() * dot(disp, disp) - outer);
}
// physics.rs:48-84
inline Mat3 tensor(const Capsule& cp, float m) {
    float h = magnitude(cp.d);
    float r = cp.r;
    float mh = m * 2.0f * r / (4.0f * r + 3.0f * h);
    float mc = m * h / (4.0f / 3.0f * r + h);
    float ic_x = 1.0f / 12.0f * mc * (3.0f * r * r + h * h);
    float ic_y = 0.5f * mc * r * r;
    float ic_z = ic_x;
    float is_x = mh * (3.0f * r + 2.0f * h) / 4.0f * h;
    float is_y = 4.0f / 5.0f * mh * r * r;
    float is_z = is_x;
    float i_x = ic_x + is_x, i_y = ic_y + is_y, i_z = ic_z + is_z;
    Vec3 dst = cp.d;
    Vec3 src = v3(0.0f, 1.0f, 0.0f) * h;
    Mat3 rot = mat3_from_quat(from_arc(src, dst));
    Mat3 I = rot * mat3_new(i_x, 0, 0, 0, i_y, 0, 0, 0, i_z) * transpose(rot);
    Vec3 disp = center(cp);
    Mat3 outer = from_cols(disp * disp.x, disp * disp.y, disp * disp.z);
    return I + m * (mat3_one() * dot(disp, disp) - outer);
}
inline Mat3 tensor(const Component& k, float m) { return k.kind == Component::SPHERE ? tensor(k.s, m) : tensor(k.c, m); }
// physics.rs:95-120
inline Mat3 tensor(const OBB& o, float m) {
    float x = o.r.x * 2.0f, y = o.r.y * 2.0f, z = o.r.z * 2.0f;
    float i_x = 1.0f / 12.0f * m * (y * y + z * z);
    float i_y = 1.0f / 12.0f * m * (x * x + z * z);
    float i_z = 1.0f / 12.0f * m * (x * x + y * y);
    Mat3 rot = mat3_from_quat(o.q);
    Mat3 I = rot * mat3_new(i_x, 0, 0, 0, i_y, 0, 0, 0, i_z) * transpose(rot);
    Vec3 disp = o.c;
    Mat3 outer = from_cols(disp * disp.x, disp * disp.y, disp * disp.z);
    return I + m * (mat3_one() * dot(disp, disp) - outer);
}

struct SingularInertia {};

// physics.rs:141-315
struct RigidBodyVec {
    std::vector<Vec3> x; std::vector<Quat> q;
    std::vector<Vec3> v, omega, force, torque;
    std::vector<float> restitution, friction, inv_mass;
    std::vector<Mat3> inv_moment_body, inv_moment;
    std::vector<ComponentConstructor> constructor;
    std::vector<MovingComponent> collider;

    size_t len() const { return x.size(); }
    size_t add_body(const Component& col, float mass, float rest, float fric, Vec3 world_force) {  // :200-218
        size_t id = x.size();
        Vec3 px; Quat pq; ComponentConstructor cc;
        deconstruct(col, &px, &pq, &cc);
        x.push_back(px); q.push_back(pq);
        v.push_back(v3(0, 0, 0)); omega.push_back(v3(0, 0, 0));
        force.push_back(world_force * mass);
        torque.push_back(v3(0, 0, 0));
        restitution.push_back(rest); friction.push_back(fric);
        inv_mass.push_back(1.0f / mass);
        Mat3 im;
        if (!invert(tensor(col - px, mass), &im)) throw SingularInertia();
        inv_moment_body.push_back(im); inv_moment.push_back(im);
        constructor.push_back(cc);
        collider.push_back(MovingComponent{col, v3(0, 0, 0)});
        return id;
    }
    void integrate(float dt) {  // :222-253
        size_t n = x.size();
        for (size_t i = 0; i < n; ++i)
            q[i] = qnormalize(q[i] + from_sv(0.0f, omega[i] * dt) * 0.5f * q[i]);
        for (size_t i = 0; i < n; ++i) {
            Mat3 r = mat3_from_quat(q[i]);
            inv_moment[i] = r * inv_moment_body[i] * transpose(r);
        }
        for (size_t i = 0; i < n; ++i) v[i] += force[i] * inv_mass[i] * dt;
        for (size_t i = 0; i < n; ++i) omega[i] += inv_moment[i] * torque[i] * dt;
        for (size_t i = 0; i < n; ++i) collider[i] = MovingComponent{construct(constructor[i], x[i], q[i]), v[i] * dt};
    }
    void complete_motion() {  // :262-269
        for (size_t i = 0; i < x.size(); ++i) x[i] += collider[i].v;
    }
    void get(const RigidBodyRef& r, Velocity* vel, RigidBodyInfo* info) const {  // :273-304
        if (r.dynamic) {
            size_t i = r.i;
            *vel = {v[i], omega[i]};
            *info = {x[i] + collider[i].v, restitution[i], friction[i], inv_mass[i], inv_moment[i]};
        } else {
            *vel = {v3(0, 0, 0), v3(0, 0, 0)};
            *info = {r.center, 0.0f, r.friction, 0.0f, mat3_zero()};
        }
    }
    void set(const RigidBodyRef& r, const Velocity& vel) {  // :306-314
        if (r.dynamic) { v[r.i] = vel.linear; omega[r.i] = vel.angular; }
    }
};

// ---- solver.rs ----
static const float PENETRATION_SLOP = 0.05f;  // solver.rs:277
static const float BAUMGARTE = 0.2f;          // solver.rs:278
inline float solver_clamp(float n, float mn, float mx) {  // solver.rs:281
    if (n < mn) return mn; else if (n > mx) return mx; else return n;
}
struct ContactState { float bias, normal_mass, normal_impulse, tangent_mass[2], tangent_impulse[2]; };  // :256
struct ContactConstraint {  // solver.rs:82-93
    RigidBodyRef obj_a, obj_b;
    Manifold manifold;
    float friction;
    std::vector<ContactState> states;

    // solver.rs:101-191
    static ContactConstraint make(const RigidBodyVec& pool, RigidBodyRef obj_a, RigidBodyRef obj_b,
                                  const Manifold& manifold, float dt) {
        Velocity va_, vb_; RigidBodyInfo ia, ib;
        pool.get(obj_a, &va_, &ia);
        pool.get(obj_b, &vb_, &ib);
        Vec3 va = va_.linear, oa = va_.angular, vb = vb_.linear, ob = vb_.angular;
        Vec3 xa = ia.x, xb = ib.x;
        float restitution = fmax_(ia.restitution, ib.restitution);
        float friction = sqrtf(ia.friction * ib.friction);
        ContactConstraint cc{obj_a, obj_b, manifold, friction, {}};
        for (int k = 0; k < manifold.len(); ++k) {
            Vec3 ra = manifold.contact(k).first, rb = manifold.contact(k).second;
            Vec3 ca = ra + xa, cb = rb + xb;
            Vec3 ra_cn = cross(ra, manifold.normal), rb_cn = cross(rb, manifold.normal);
            float pen = dot(cb - ca, manifold.normal);
            Vec3 dv = vb + cross(ob, rb) - va - cross(oa, ra);
            float rel_v = dot(dv, manifold.normal);
            float bias = -BAUMGARTE / dt * (pen > 0.0f ? 0.0f : pen + PENETRATION_SLOP) +
                         (rel_v < -1.0f ? -restitution * rel_v : 0.0f);
            float normal_mass = 1.0f / (ia.inv_mass + dot(ra_cn, ia.inv_moment * ra_cn) + ib.inv_mass +
                                        dot(rb_cn, ib.inv_moment * rb_cn));
            float tangent_mass[2];
            for (int i = 0; i < 2; ++i) {
                Vec3 ra_ct = cross(ra, manifold.tangent_vector[i]);
                Vec3 rb_ct = cross(rb, manifold.tangent_vector[i]);
                tangent_mass[i] = 1.0f / (ia.inv_mass + dot(ra_ct, ia.inv_moment * ra_ct) + ib.inv_mass +
                                          dot(rb_ct, ib.inv_moment * rb_ct));
            }
            cc.states.push_back(ContactState{bias, normal_mass, 0.0f, {tangent_mass[0], tangent_mass[1]}, {0.0f, 0.0f}});
        }
        return cc;
    }
    // solver.rs:203-252
    void solve(RigidBodyVec& pool) {
        Velocity va_, vb_; RigidBodyInfo ia, ib;
        pool.get(obj_a, &va_, &ia);
        pool.get(obj_b, &vb_, &ib);
        Vec3 va = va_.linear, oa = va_.angular, vb = vb_.linear, ob = vb_.angular;
        for (size_t k = 0; k < states.size(); ++k) {
            ContactState& st = states[k];
            Vec3 ra = manifold.contact((int)k).first, rb = manifold.contact((int)k).second;
            Vec3 dv = vb + cross(ob, rb) - va - cross(oa, ra);
            for (int i = 0; i < 2; ++i) {
                float lambda = -dot(dv, manifold.tangent_vector[i]) * st.tangent_mass[i];
                float max_lambda = friction * st.normal_impulse;
                float prev_impulse = st.tangent_impulse[i];
                st.tangent_impulse[i] = solver_clamp(-max_lambda, max_lambda, prev_impulse + lambda);  // sic
                Vec3 impulse = manifold.tangent_vector[i] * lambda;
                va -= impulse * ia.inv_mass;
                oa -= ia.inv_moment * cross(ra, impulse);
                vb += impulse * ib.inv_mass;
                ob += ib.inv_moment * cross(rb, impulse);
            }
            Vec3 dv2 = vb + cross(ob, rb) - va - cross(oa, ra);
            float vn = dot(dv2, manifold.normal);
            float lambda = st.normal_mass * (-vn + st.bias);
            float prev_impulse = st.normal_impulse;
            st.normal_impulse = fmax_(prev_impulse + lambda, 0.0f);
            lambda = st.normal_impulse - prev_impulse;
            Vec3 impulse = manifold.normal * lambda;
            va -= impulse * ia.inv_mass;
            oa -= ia.inv_moment * cross(ra, impulse);
            vb += impulse * ib.inv_mass;
            ob += ib.inv_moment * cross(rb, impulse);
        }
        pool.set(obj_a, Velocity{va, oa});
        pool.set(obj_b, Velocity{vb, ob});
    }
};
struct Solver {  // solver.rs:53-79
    std::vector<ContactConstraint> constraints;
    void add_constraint(const ContactConstraint& c) { constraints.push_back(c); }
    void solve(RigidBodyVec& cs, size_t iters) {
        for (size_t it = 0; it < iters; ++it)
            for (ContactConstraint& c : constraints) c.solve(cs);
    }
};

}  // namespace mgfo
