// ORACLE (test infrastructure, NOT product code) -- C entry points so tests/ and bench.py's
// cpu_baseline / --impl reference legs can drive the CPU restatement through ctypes with the
// same POD structs as include/mgfb.h.  Nothing in mgf_b200/ may link or load this library.
#include <chrono>
#include <cstring>
#include <new>
#include "../include/mgfb.h"
#include "simplex.hpp"
#include "world.hpp"
#include "compound.hpp"

using namespace mgfo;

namespace {
Vec3 p3(const float* p) { return v3(p[0], p[1], p[2]); }
Sphere to_sphere(const mgfb_shape& s) { return Sphere{p3(s.p), s.p[3]}; }
Capsule to_capsule(const mgfb_shape& s) { return Capsule{p3(s.p), p3(s.p + 3), s.p[6]}; }
Triangle to_tri(const mgfb_shape& s) { return Triangle{p3(s.p), p3(s.p + 3), p3(s.p + 6)}; }
Rectangle to_rect(const mgfb_shape& s) { return Rectangle{p3(s.p), {p3(s.p + 3), p3(s.p + 6)}, {s.p[9], s.p[10]}}; }
Plane to_plane_s(const mgfb_shape& s) { return Plane{p3(s.p), s.p[3]}; }
Component to_component(const mgfb_shape& s) {
    return s.kind == MGFB_SPHERE ? Component::sphere(to_sphere(s)) : Component::capsule(to_capsule(s));
}
void put3(float* o, Vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
void put_contact(mgfb_contact* o, const Contact& c) { put3(o->a, c.a); put3(o->b, c.b); put3(o->n, c.n); o->t = c.t; }
}  // namespace

extern "C" {

// ---- BVH<AABB, u32> exactly as src/bvh.rs builds it (incremental insert / remove / balance): the checker of mgfb_bvh_*
void* mgfo_bvh_create() { return new BVH<uint32_t>(); }
void mgfo_bvh_destroy(void* h) { delete static_cast<BVH<uint32_t>*>(h); }
uint32_t mgfo_bvh_insert(void* h, const float* box, uint32_t value) {
    return (uint32_t)static_cast<BVH<uint32_t>*>(h)->insert(AABB{p3(box), p3(box + 3)}, value);
}
void mgfo_bvh_remove(void* h, uint32_t index) { static_cast<BVH<uint32_t>*>(h)->remove(index); }
// values in the reference's DFS callback order; returns the number of callbacks (may exceed cap)
uint32_t mgfo_bvh_query(void* h, const float* box, uint32_t* out, uint32_t cap) {
    uint32_t n = 0;
    static_cast<BVH<uint32_t>*>(h)->query(AABB{p3(box), p3(box + 3)}, [&](const uint32_t& v) { if (n < cap) out[n] = v; ++n; });
    return n;
}
uint32_t mgfo_bvh_raytrace(void* h, uint32_t particle_kind, const float* q, uint32_t* out, mgfb_intersection* hits, uint32_t cap) {
    const bool seg = particle_kind == MGFB_SEGMENT;
    Ray r{p3(q), seg ? p3(q + 3) - p3(q) : p3(q + 3)};
    uint32_t n = 0;
    static_cast<BVH<uint32_t>*>(h)->raytrace(r, [&](const uint32_t& v, const Intersection& it) {
        if (n < cap) { out[n] = v; put3(hits[n].p, it.p); hits[n].t = it.t; }
        ++n;
    }, seg ? 1.0f : INF);
    return n;
}

// Same contract as mgfb_intersections_batch (include/mgfb.h): Intersects<RHS> for Ray / Segment (collision.rs:163-373).
int32_t mgfo_intersections_batch(uint32_t particle_kind, const float* particles, const mgfb_shape* shapes, uint32_t n,
                                 mgfb_intersection* out, uint32_t* hit) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* q = particles + 6 * i;
        const bool seg = particle_kind == MGFB_SEGMENT;
        // geom.rs:818-851: Ray{p, d}: pos = p, dir = d, DT = inf; Segment{a, b}: pos = a, dir = b - a, DT = 1
        Ray r{p3(q), seg ? p3(q + 3) - p3(q) : p3(q + 3)};
        const float DT = seg ? 1.0f : INF;
        const mgfb_shape& S = shapes[i];
        Intersection it{v3(0, 0, 0), 0.0f};
        bool ok = false;
        switch (S.kind) {
            case MGFB_PLANE: ok = intersection(r, to_plane_s(S), &it, DT); break;
            case MGFB_TRIANGLE: ok = intersection_poly(r, to_tri(S), &it, DT); break;
            case MGFB_RECTANGLE: ok = intersection_poly(r, to_rect(S), &it, DT); break;
            case MGFB_AABB: ok = intersection(r, AABB{p3(S.p), p3(S.p + 3)}, &it, DT); break;
            case MGFB_OBB: ok = intersection(r, OBB{p3(S.p), Quat{S.p[6], p3(S.p + 7)}, p3(S.p + 3)}, &it, DT, seg); break;
            case MGFB_SPHERE:   // Moving<Sphere> (v != 0) is the capsule swept by the sphere (collision.rs:361-373)
                if (S.v[0] != 0.0f || S.v[1] != 0.0f || S.v[2] != 0.0f) ok = intersection(r, Capsule{p3(S.p), p3(S.v), S.p[3]}, &it, DT);
                else ok = intersection(r, to_sphere(S), &it, DT);
                break;
            case MGFB_CAPSULE: ok = intersection(r, to_capsule(S), &it, DT); break;
            default: return 1;
        }
        hit[i] = ok ? 1u : 0u;
        if (ok) { put3(out[i].p, it.p); out[i].t = it.t; } else { out[i].p[0] = out[i].p[1] = out[i].p[2] = 0.0f; out[i].t = 0.0f; }
    }
    return 0;
}

// Same contract as mgfb_contacts_batch (include/mgfb.h).
int32_t mgfo_contacts_batch(uint32_t pair_kind, const mgfb_shape* recv, const mgfb_shape* arg, uint32_t n, mgfb_contact* out,
                            mgfb_local_contact* out_local, uint32_t* counts) {
    for (uint32_t i = 0; i < n; ++i) {
        const mgfb_shape &R = recv[i], &A = arg[i];
        Vec3 v = p3(A.v);
        uint32_t cnt = 0;
        auto cb = [&](const Contact& c) { if (cnt < 2) put_contact(&out[2 * i + cnt], c); cnt++; };
        auto lcb = [&](const LocalContact& lc) {
            if (cnt < 2) {
                put_contact(&out[2 * i + cnt], lc.global);
                if (out_local) { put3(out_local[2 * i + cnt].local_a, lc.local_a); put3(out_local[2 * i + cnt].local_b, lc.local_b); put_contact(&out_local[2 * i + cnt].global, lc.global); }
            }
            cnt++;
        };
        switch (pair_kind) {
            case MGFB_SPHERE_X_MSPHERE: contacts(to_sphere(R), Moving<Sphere>{to_sphere(A), v}, cb); break;
            case MGFB_CAPSULE_X_MSPHERE: contacts(to_capsule(R), Moving<Sphere>{to_sphere(A), v}, cb); break;
            case MGFB_SPHERE_X_MCAPSULE: contacts(to_sphere(R), Moving<Capsule>{to_capsule(A), v}, cb); break;
            case MGFB_CAPSULE_X_MCAPSULE: contacts(to_capsule(R), Moving<Capsule>{to_capsule(A), v}, cb); break;
            case MGFB_PLANE_X_MSPHERE: contacts(to_plane_s(R), Moving<Sphere>{to_sphere(A), v}, cb); break;
            case MGFB_PLANE_X_MCAPSULE: contacts(to_plane_s(R), Moving<Capsule>{to_capsule(A), v}, cb); break;
            case MGFB_TRI_X_MSPHERE: contacts(to_tri(R), Moving<Sphere>{to_sphere(A), v}, cb); break;
            case MGFB_TRI_X_MCAPSULE: contacts(to_tri(R), Moving<Capsule>{to_capsule(A), v}, cb); break;
            case MGFB_RECT_X_MSPHERE: contacts(to_rect(R), Moving<Sphere>{to_sphere(A), v}, cb); break;
            case MGFB_RECT_X_MCAPSULE: contacts(to_rect(R), Moving<Capsule>{to_capsule(A), v}, cb); break;
            case MGFB_MCOMP_X_MCOMP:
                local_contacts(MovingComponent{to_component(R), p3(R.v)}, MovingComponent{to_component(A), v}, lcb);
                break;
            case MGFB_MCOMP_X_TRI: {
                // collision.rs:1490 + mesh.rs:119-137 for one face; arg.v carries mesh.x
                MovingComponent self{to_component(R), p3(R.v)};
                Triangle tri = to_tri(A);
                Vec3 mesh_x = v;
                component_contacts_poly(self, tri, [&](const Contact& c0) {
                    Contact c{c0.b, c0.a, -c0.n, c0.t};
                    Vec3 a_c = center(self.g) + self.v * c.t;
                    lcb(LocalContact{c.b + (-a_c), c.a + (-mesh_x), neg(c)});
                });
                break;
            }
            default: return MGFB_ERR_INVALID_ARG;
        }
        counts[i] = cnt;
    }
    return MGFB_OK;
}

// ---- Compound (compound.rs:232-352): the checker of mgfb_compound_* ----
void* mgfo_compound_create(const mgfb_shape* comps, uint32_t n) {
    std::vector<Component> v;
    for (uint32_t i = 0; i < n; ++i) v.push_back(to_component(comps[i]));
    try { return new Compound(v); } catch (NanBounds&) { return nullptr; }
}
void mgfo_compound_destroy(void* h) { delete static_cast<Compound*>(h); }
void mgfo_compound_set_transform(void* h, const float* disp, const float* rot) {
    Compound* c = static_cast<Compound*>(h);
    c->disp = p3(disp); c->rot = Quat{rot[0], p3(rot + 1)};
}
void mgfo_compound_bounds(const void* h, float* aabb, float* sphere) {
    const Compound* c = static_cast<const Compound*>(h);
    AABB b = c->bounds_aabb(); Sphere s = c->bounds_sphere();
    put3(aabb, b.c); put3(aabb + 3, b.r); put3(sphere, s.c); sphere[3] = s.r;
}
void mgfo_compound_closest_points(const void* h, const float* to, uint32_t n, float* out) {
    const Compound* c = static_cast<const Compound*>(h);
    for (uint32_t i = 0; i < n; ++i) put3(out + 3 * i, c->closest_point(p3(to + 3 * i)));
}
void mgfo_compound_intersections_batch(const void* h, uint32_t particle_kind, const float* particles, uint32_t n, mgfb_intersection* out, uint32_t* hit) {
    const Compound* c = static_cast<const Compound*>(h);
    for (uint32_t i = 0; i < n; ++i) {
        const float* q = particles + 6 * i;
        const bool seg = particle_kind == MGFB_SEGMENT;
        Ray r{p3(q), seg ? p3(q + 3) - p3(q) : p3(q + 3)};
        Intersection it{v3(0, 0, 0), 0.0f};
        bool ok = c->intersection(r, seg ? 1.0f : INF, &it);
        hit[i] = ok ? 1u : 0u;
        put3(out[i].p, ok ? it.p : v3(0, 0, 0)); out[i].t = ok ? it.t : 0.0f;
    }
}
int32_t mgfo_compound_contacts_batch(const void* h, const mgfb_shape* rhs, uint32_t n, uint32_t slots, mgfb_contact* out, uint32_t* counts) {
    const Compound* c = static_cast<const Compound*>(h);
    std::memset(out, 0, (size_t)n * slots * sizeof(mgfb_contact));
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t cnt = 0;
        auto cb = [&](const Contact& k) { if (cnt < slots) put_contact(&out[(size_t)i * slots + cnt], k); ++cnt; };
        const mgfb_shape& R = rhs[i];
        Vec3 v = p3(R.v);
        switch (R.kind) {
            case MGFB_SPHERE: c->contacts(Moving<Sphere>{to_sphere(R), v}, cb); break;
            case MGFB_CAPSULE: c->contacts(Moving<Capsule>{to_capsule(R), v}, cb); break;
            case MGFB_TRIANGLE: c->contacts(Moving<Triangle>{to_tri(R), v}, cb); break;
            case MGFB_RECTANGLE: c->contacts(Moving<Rectangle>{to_rect(R), v}, cb); break;
            default: return MGFB_ERR_INVALID_ARG;
        }
        counts[i] = cnt;
    }
    return MGFB_OK;
}

// Same contract as mgfb_manifolds_prune: ContactPruner::push per group in order, then Manifold::from(pruner) (manifold.rs:42-148).
int32_t mgfo_manifolds_prune(const mgfb_local_contact* contacts, const uint32_t* offsets, uint32_t ngroups, float* time, float* normal, float* tangent,
                             uint32_t* ncontacts, float* local_a, float* local_b) {
    for (uint32_t g = 0; g < ngroups; ++g) {
        ContactPruner pruner;
        for (uint32_t k = offsets[g]; k < offsets[g + 1]; ++k) {
            const mgfb_local_contact& c = contacts[k];
            pruner.push(LocalContact{p3(c.local_a), p3(c.local_b), Contact{p3(c.global.a), p3(c.global.b), p3(c.global.n), c.global.t}});
        }
        Manifold m = manifold_from(pruner);
        if (time) time[g] = m.time;
        put3(normal + 3 * g, m.normal); put3(tangent + 6 * g, m.tangent_vector[0]); put3(tangent + 6 * g + 3, m.tangent_vector[1]);
        ncontacts[g] = (uint32_t)m.len();
        for (int i = 0; i < 4; ++i) {
            put3(local_a + 12 * g + 3 * i, i < m.len() ? m.contact(i).first : v3(0, 0, 0));
            put3(local_b + 12 * g + 3 * i, i < m.len() ? m.contact(i).second : v3(0, 0, 0));
        }
    }
    return MGFB_OK;
}

// ---- world handle ----
struct mgfo_world { World w; };

mgfo_world* mgfo_world_create(float fat_margin) {
    mgfo_world* h = new (std::nothrow) mgfo_world();
    if (h) h->w.fat_margin = fat_margin;
    return h;
}
void mgfo_world_destroy(mgfo_world* h) { delete h; }

int32_t mgfo_world_add_bodies(mgfo_world* h, uint32_t n, const mgfb_shape* shapes, const float* mass, const float* rest,
                              const float* fric, const float* world_force) {
    try {
        for (uint32_t i = 0; i < n; ++i)
            h->w.add_body(to_component(shapes[i]), mass[i], rest[i], fric[i], p3(world_force + 3 * i));
    } catch (SingularInertia&) { return MGFB_ERR_SINGULAR_INERTIA; } catch (NanBounds&) { return MGFB_ERR_NAN_BOUNDS; }
    return MGFB_OK;
}
int32_t mgfo_world_set_terrain(mgfo_world* h, const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces, const float x[3]) {
    Mesh m;
    for (uint32_t i = 0; i < nverts; ++i) m.push_vert(p3(verts + 3 * i));
    for (uint32_t i = 0; i < nfaces; ++i) m.push_face(faces[3 * i], faces[3 * i + 1], faces[3 * i + 2]);
    m.set_pos(p3(x));
    h->w.terrain = m;
    return MGFB_OK;
}
uint32_t mgfo_world_count(const mgfo_world* h) { return (uint32_t)h->w.bodies.len(); }
void mgfo_world_get_state(const mgfo_world* h, float* x, float* q, float* v, float* omega) {
    const RigidBodyVec& b = h->w.bodies;
    for (size_t i = 0; i < b.len(); ++i) {
        if (x) put3(x + 3 * i, b.x[i]);
        if (q) { q[4 * i] = b.q[i].s; put3(q + 4 * i + 1, b.q[i].v); }
        if (v) put3(v + 3 * i, b.v[i]);
        if (omega) put3(omega + 3 * i, b.omega[i]);
    }
}
void mgfo_world_set_velocity(mgfo_world* h, uint32_t first, uint32_t n, const float* v, const float* omega) {
    for (uint32_t i = 0; i < n; ++i) h->w.bodies.set(RigidBodyRef::Dynamic(first + i), Velocity{p3(v + 3 * i), p3(omega + 3 * i)});
}
void mgfo_world_get_colliders(const mgfo_world* h, mgfb_shape* out) {
    const RigidBodyVec& b = h->w.bodies;
    for (size_t i = 0; i < b.len(); ++i) {
        mgfb_shape s; std::memset(&s, 0, sizeof(s));
        const MovingComponent& m = b.collider[i];
        if (m.g.kind == Component::SPHERE) { s.kind = MGFB_SPHERE; put3(s.p, m.g.s.c); s.p[3] = m.g.s.r; }
        else { s.kind = MGFB_CAPSULE; put3(s.p, m.g.c.a); put3(s.p + 3, m.g.c.d); s.p[6] = m.g.c.r; }
        put3(s.v, m.v);
        out[i] = s;
    }
}
void mgfo_world_get_inv_moment(const mgfo_world* h, float* out) {
    const RigidBodyVec& b = h->w.bodies;
    for (size_t i = 0; i < b.len(); ++i) for (int c = 0; c < 3; ++c) put3(out + 9 * i + 3 * c, b.inv_moment[i].c[c]);
}
// Checkpoint support (the checker of mgfb_bodies_get_fat_bounds / mgfb_bodies_set_state): the stored fat boxes are the
// leaves bvh[bvh_ids[i]] (world.rs:180, 235-238).
void mgfo_world_get_fat_bounds(const mgfo_world* h, float* out) {
    const World& w = h->w;
    for (size_t i = 0; i < w.bodies.len(); ++i) { const AABB& b = w.bvh[w.bvh_ids[i]]; put3(out + 6 * i, b.c); put3(out + 6 * i + 3, b.r); }
}
// Overwrites the pub fields x, q, collider (physics.rs:142-154), v, omega, and rebuilds the body BVH from the given fat
// boxes, inserted in body order like add_body does (world.rs:180-182).  inv_moment follows q as in integrate
// (physics.rs:231-232).  Any pointer may be NULL.
int32_t mgfo_world_set_state(mgfo_world* h, const float* x, const float* q, const float* v, const float* omega, const mgfb_shape* colliders,
                             const float* fat) {
    World& w = h->w;
    RigidBodyVec& b = w.bodies;
    for (size_t i = 0; i < b.len(); ++i) {
        if (x) b.x[i] = p3(x + 3 * i);
        if (q) {
            b.q[i] = Quat{q[4 * i], p3(q + 4 * i + 1)};
            Mat3 r = mat3_from_quat(b.q[i]);
            b.inv_moment[i] = r * b.inv_moment_body[i] * transpose(r);
        }
        if (v) b.v[i] = p3(v + 3 * i);
        if (omega) b.omega[i] = p3(omega + 3 * i);
        if (colliders) {
            const mgfb_shape& s = colliders[i];
            MovingComponent& m = b.collider[i];
            if ((s.kind == MGFB_SPHERE) != (m.g.kind == Component::SPHERE)) return MGFB_ERR_INVALID_ARG;
            if (s.kind == MGFB_SPHERE) m.g.s.c = p3(s.p); else { m.g.c.a = p3(s.p); m.g.c.d = p3(s.p + 3); }
            m.v = p3(s.v);
        }
    }
    if (fat) {
        try {
            w.bvh.clear();   // a fresh BVH::new, then the inserts add_body makes (world.rs:180-182)
            for (size_t i = 0; i < b.len(); ++i) w.bvh_ids[i] = w.bvh.insert(AABB{p3(fat + 6 * i), p3(fat + 6 * i + 3)}, i);
        } catch (NanBounds&) { return MGFB_ERR_NAN_BOUNDS; }
    }
    return MGFB_OK;
}
void mgfo_world_integrate(mgfo_world* h, float dt) { h->w.bodies.integrate(dt); }
void mgfo_world_complete_motion(mgfo_world* h) { h->w.bodies.complete_motion(); }

// World::step in the reference's own order (world.rs:227-294).
int32_t mgfo_world_step(mgfo_world* h, float dt, uint32_t iters, uint32_t nsteps) {
    try { for (uint32_t s = 0; s < nsteps; ++s) h->w.step(dt, iters); } catch (NanBounds&) { return MGFB_ERR_NAN_BOUNDS; }
    return MGFB_OK;
}
// World::step up to (not including) solver.solve; returns the number of constraints.
int32_t mgfo_world_build(mgfo_world* h, float dt, uint32_t* count) {
    try { h->w.build_constraints(dt); } catch (NanBounds&) { return MGFB_ERR_NAN_BOUNDS; }
    *count = (uint32_t)h->w.last.size();
    return MGFB_OK;
}
void mgfo_world_constraints(const mgfo_world* h, uint32_t* a, int32_t* b, uint32_t* face, uint32_t* sub) {
    const auto& L = h->w.last;
    for (size_t k = 0; k < L.size(); ++k) { a[k] = L[k].a; b[k] = L[k].b; face[k] = L[k].face; sub[k] = L[k].sub; }
}
// Manifolds of the last build, in insertion order, in mgfb_manifolds layout.
void mgfo_world_manifolds(const mgfo_world* h, int32_t* obj_a, int32_t* obj_b, float* static_center, float* static_friction,
                          float* normal, float* tangent, uint32_t* ncontacts, float* local_a, float* local_b) {
    const auto& L = h->w.last;
    for (size_t k = 0; k < L.size(); ++k) {
        const ContactConstraint& c = L[k].c;
        obj_a[k] = c.obj_a.dynamic ? (int32_t)c.obj_a.i : -1;
        obj_b[k] = c.obj_b.dynamic ? (int32_t)c.obj_b.i : -1;
        const RigidBodyRef& st = c.obj_b.dynamic ? c.obj_a : c.obj_b;
        put3(static_center + 3 * k, st.center); static_friction[k] = st.friction;
        put3(normal + 3 * k, c.manifold.normal);
        put3(tangent + 6 * k, c.manifold.tangent_vector[0]); put3(tangent + 6 * k + 3, c.manifold.tangent_vector[1]);
        ncontacts[k] = (uint32_t)c.manifold.len();
        for (int i = 0; i < 4; ++i) {
            Vec3 a = i < c.manifold.len() ? c.manifold.contact(i).first : v3(0, 0, 0);
            Vec3 b = i < c.manifold.len() ? c.manifold.contact(i).second : v3(0, 0, 0);
            put3(local_a + 12 * k + 3 * i, a); put3(local_b + 12 * k + 3 * i, b);
        }
    }
}
// O(n^2) count of {(i, j): j < i, tight_i overlaps stored fat_j} over the LEAF boxes of the last
// build.  The reference's BVH query can return fewer: its internal nodes are rounded unions
// ((hi+lo)/2, (hi-lo)/2 in bounds.rs:113-130), so a leaf that only TOUCHES the query box can be
// pruned at a parent.  Such pairs are >= fat_margin apart and never produce a contact.
uint64_t mgfo_world_brute_pairs(const mgfo_world* h) {
    const World& w = h->w;
    uint64_t cnt = 0;
    for (size_t i = 1; i < w.bodies.len(); ++i) {
        AABB t = bounds(w.bodies.collider[i]);
        for (size_t j = 0; j < i; ++j) cnt += overlaps(t, w.bvh[w.bvh_ids[j]]) ? 1 : 0;
    }
    return cnt;
}
void mgfo_world_stats(const mgfo_world* h, uint64_t* candidate_pairs, uint64_t* terrain_candidates) {
    *candidate_pairs = h->w.candidate_pairs; *terrain_candidates = h->w.terrain_candidates;
}
// solver.solve with a caller-chosen order over the constraints of the last build.
void mgfo_world_solve_order(mgfo_world* h, const uint32_t* perm, uint32_t n, uint32_t iters) {
    std::vector<uint32_t> p(perm, perm + n);
    h->w.solve_in_order(p, iters);
}
// Solver::new + add_constraint(ContactConstraint::new(..))* in the order perm (NULL = as given) + solve.
int32_t mgfo_world_solve_manifolds(mgfo_world* h, const mgfb_manifolds* m, float dt, uint32_t iters, const uint32_t* perm,
                                   float* normal_impulse_out) {
    Solver solver;
    std::vector<uint32_t> order(m->n);
    for (uint32_t k = 0; k < m->n; ++k) order[k] = perm ? perm[k] : k;
    for (uint32_t r = 0; r < m->n; ++r) {
        uint32_t k = order[r];
        Manifold mf; mf.ncontacts = 0; mf.time = 0.0f;
        mf.normal = p3(m->normal + 3 * k);
        mf.tangent_vector[0] = p3(m->tangent + 6 * k); mf.tangent_vector[1] = p3(m->tangent + 6 * k + 3);
        for (uint32_t c = 0; c < m->ncontacts[k]; ++c) mf.push(p3(m->local_a + 12 * k + 3 * c), p3(m->local_b + 12 * k + 3 * c));
        auto ref = [&](int32_t o) {
            return o >= 0 ? RigidBodyRef::Dynamic((size_t)o) : RigidBodyRef::Static(p3(m->static_center + 3 * k), m->static_friction[k]);
        };
        solver.add_constraint(ContactConstraint::make(h->w.bodies, ref(m->obj_a[k]), ref(m->obj_b[k]), mf, dt));
    }
    solver.solve(h->w.bodies, iters);
    if (normal_impulse_out) {
        std::memset(normal_impulse_out, 0, (size_t)m->n * 16);
        for (uint32_t r = 0; r < m->n; ++r) {
            uint32_t k = order[r];
            for (size_t c = 0; c < solver.constraints[r].states.size(); ++c) normal_impulse_out[4 * k + c] = solver.constraints[r].states[c].normal_impulse;
        }
    }
    return MGFB_OK;
}

// Timed run for the CPU baseline: `nsteps` reference-order steps, returns seconds and counts.
double mgfo_world_time_steps(mgfo_world* h, float dt, uint32_t iters, uint32_t nsteps, uint64_t* constraint_iters, uint64_t* pairs) {
    uint64_t ci = 0, pr = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t s = 0; s < nsteps; ++s) {
        h->w.step(dt, iters);
        ci += (uint64_t)h->w.last.size() * iters;
        pr += h->w.candidate_pairs + h->w.terrain_candidates;
    }
    auto t1 = std::chrono::steady_clock::now();
    *constraint_iters = ci; *pairs = pr;
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- discrete path (collision.rs:404-425, 497-519) ----
static const float* g_convex_pool = nullptr; static uint32_t g_convex_n = 0;
// the vertex pool MGFB_CONVEX_MESH shapes index (same contract as mgfb_convex_vertices_set); the caller keeps it alive
void mgfo_convex_vertices_set(const float* verts, uint32_t n) { g_convex_pool = verts; g_convex_n = n; }
static bool gjk_dispatch(const mgfb_shape& A, const mgfb_shape& B, Contact* c, int* it, int* st);
// largest Pool slot count / horizon size any EPA run reached since the last call with reset != 0
int32_t mgfo_epa_high_water(uint32_t* slots, uint32_t* edges, int32_t reset) {
    if (slots) *slots = epa_high_water(0);
    if (edges) *edges = epa_high_water(1);
    if (reset) { epa_high_water(0) = 0; epa_high_water(1) = 0; }
    return 0;
}
int32_t mgfo_gjk_batch(const mgfb_shape* a, const mgfb_shape* b, uint32_t n, mgfb_contact* out, uint32_t* hit, uint32_t* epa_iters) {
    for (uint32_t i = 0; i < n; ++i) {
        Contact c{}; int it = 0;
        int st = 0;
        bool ok = gjk_dispatch(a[i], b[i], &c, &it, &st);
        hit[i] = ok ? 1u : (uint32_t)st;   // 0 no contact, 1 contact, 3 GJK step cap (same codes as mgfb_gjk_batch)
        if (epa_iters) epa_iters[i] = (uint32_t)it;
        if (ok) put_contact(&out[i], c); else std::memset(&out[i], 0, sizeof(mgfb_contact));
    }
    return MGFB_OK;
}
static bool sep_dispatch(const mgfb_shape& A, const mgfb_shape& B, float* d, int* st);
// Penetrates::separation (collision.rs:404-425) for a batch: some[i] = 1 and sep[i] = distance for Some(d).
int32_t mgfo_separation_batch(const mgfb_shape* a, const mgfb_shape* b, uint32_t n, float* sep, uint32_t* some) {
    for (uint32_t i = 0; i < n; ++i) {
        float d = 0.0f; int st = 0;
        bool ok = sep_dispatch(a[i], b[i], &d, &st);
        some[i] = ok ? 1u : (uint32_t)st; sep[i] = ok ? d : 0.0f;
    }
    return MGFB_OK;
}
}  // extern "C"

namespace {
AABB to_aabb(const mgfb_shape& s) { return AABB{p3(s.p), p3(s.p + 3)}; }
OBB to_obb(const mgfb_shape& s) { return OBB{p3(s.p), Quat{s.p[6], p3(s.p + 7)}, p3(s.p + 3)}; }
ConvexMesh to_mesh(const mgfb_shape& s) { return ConvexMesh{g_convex_pool + 3 * (size_t)s.p[0], (unsigned)s.p[1]}; }
bool mesh_ok(const mgfb_shape& s) { return s.kind != MGFB_CONVEX_MESH || (g_convex_pool && s.p[1] >= 1.0f && (uint64_t)s.p[0] + (uint64_t)s.p[1] <= g_convex_n); }
template <class SA>
bool gjk_with(const SA& a, const mgfb_shape& B, Contact* c, int* it, int* st) {
    switch (B.kind) {
        case MGFB_SPHERE: return gjk_contact(a, to_sphere(B), c, it, st);
        case MGFB_CAPSULE: return gjk_contact(a, to_capsule(B), c, it, st);
        case MGFB_AABB: return gjk_contact(a, to_aabb(B), c, it, st);
        case MGFB_OBB: return gjk_contact(a, to_obb(B), c, it, st);
        case MGFB_CONVEX_MESH: return gjk_contact(a, to_mesh(B), c, it, st);
        default: return false;
    }
}
}  // namespace
static bool gjk_dispatch(const mgfb_shape& A, const mgfb_shape& B, Contact* c, int* it, int* st) {
    if (!mesh_ok(A) || !mesh_ok(B)) return false;
    switch (A.kind) {
        case MGFB_SPHERE: return gjk_with(to_sphere(A), B, c, it, st);
        case MGFB_CAPSULE: return gjk_with(to_capsule(A), B, c, it, st);
        case MGFB_AABB: return gjk_with(to_aabb(A), B, c, it, st);
        case MGFB_OBB: return gjk_with(to_obb(A), B, c, it, st);
        case MGFB_CONVEX_MESH: return gjk_with(to_mesh(A), B, c, it, st);
        default: return false;
    }
}
namespace {
template <class SA>
bool sep_with(const SA& a, const mgfb_shape& B, float* d, int* st) {
    switch (B.kind) {
        case MGFB_SPHERE: return separation(a, to_sphere(B), d, st);
        case MGFB_CAPSULE: return separation(a, to_capsule(B), d, st);
        case MGFB_AABB: return separation(a, to_aabb(B), d, st);
        case MGFB_OBB: return separation(a, to_obb(B), d, st);
        case MGFB_CONVEX_MESH: return separation(a, to_mesh(B), d, st);
        default: return false;
    }
}
}  // namespace
static bool sep_dispatch(const mgfb_shape& A, const mgfb_shape& B, float* d, int* st) {
    if (!mesh_ok(A) || !mesh_ok(B)) return false;
    switch (A.kind) {
        case MGFB_SPHERE: return sep_with(to_sphere(A), B, d, st);
        case MGFB_CAPSULE: return sep_with(to_capsule(A), B, d, st);
        case MGFB_AABB: return sep_with(to_aabb(A), B, d, st);
        case MGFB_OBB: return sep_with(to_obb(A), B, d, st);
        case MGFB_CONVEX_MESH: return sep_with(to_mesh(A), B, d, st);
        default: return false;
    }
}
