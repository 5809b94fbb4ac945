// ORACLE (test infrastructure, NOT product code).  Restates the reference's per-step driver
// mgf_demo/world.rs:118-150 (box terrain), :178-184 (add_body), :227-294 (World::step) and the
// scene generators mgf_demo/balls.rs:67-96 / capsules.rs:67-95.
//
// Besides the reference's own insertion order (ORDER_REFERENCE: bodies ascending, terrain
// contacts in mesh-BVH DFS order, then body pairs in body-BVH DFS order), the world can replay
// a caller-supplied constraint order (Solver::add_constraint order is caller-chosen in mgf's
// API, solver.rs:66).  This is how the CUDA path's colour-major order is checked bit-for-bit.
#pragma once
#include <cstdint>
#include "dynamics.hpp"

namespace mgfo {

// One generated constraint with its identity, so an external order can be replayed.
struct ConstraintRecord {
    uint32_t a;        // dynamic body index i
    int32_t b;         // dynamic body index j (< i), or -1 for the static terrain
    uint32_t face;     // terrain face index (b == -1), else 0
    uint32_t sub;      // k-th contact emitted for this (body, face) pair (0 or 1), else 0
    ContactConstraint c;
};

struct World {
    RigidBodyVec bodies;
    std::vector<size_t> bvh_ids;
    BVH<size_t> bvh;
    Mesh terrain;
    float fat_margin = 0.25f;   // world.rs:181,237
    std::vector<ConstraintRecord> last;   // constraints of the most recent step, insertion order
    uint64_t candidate_pairs = 0;         // body-body candidates (after j<i filter) of the last step
    uint64_t terrain_candidates = 0;      // (body, face) candidates of the last step

    // world.rs:178-184
    size_t add_body(const Component& col, float mass, float rest, float fric, Vec3 world_force) {
        size_t id = bodies.add_body(col, mass, rest, fric, world_force);
        AABB b = bounds(bodies.collider[id]);
        size_t bvh_id = bvh.insert(aabb_add(b, fat_margin), id);
        bvh_ids.push_back(bvh_id);
        return id;
    }

    // world.rs:227-291 up to (not including) solver.solve: fills `last`.
    void build_constraints(float dt) {
        last.clear();
        candidate_pairs = 0; terrain_candidates = 0;
        bodies.complete_motion();
        bodies.integrate(dt);
        for (size_t i = 0; i < bodies.len(); ++i) {
            const MovingComponent& collider = bodies.collider[i];
            AABB b = bounds(collider);
            if (!contains(bvh[bvh_ids[i]], b)) {
                bvh.remove(bvh_ids[i]);
                bvh_ids[i] = bvh.insert(aabb_add(b, fat_margin), i);
            }
            // terrain (world.rs:240-253): one constraint per LocalContact, no pruning.
            {
                bool collided = false; (void)collided;
                terrain.bvh.query(aabb_sub(b, terrain.x), [&](const size_t& face_index) {
                    terrain_candidates++;
                    Mesh::Face f = terrain.faces[face_index];
                    Triangle tri{terrain.verts[f.a] + terrain.x, terrain.verts[f.b] + terrain.x,
                                 terrain.verts[f.c] + terrain.x};
                    uint32_t sub = 0;
                    component_contacts_poly(collider, tri, [&](const Contact& c0) {
                        Contact c{c0.b, c0.a, -c0.n, c0.t};                 // mesh.rs:129-134
                        Vec3 a_c = center(collider.g) + collider.v * c.t;  // collision.rs:1497
                        Vec3 b_c = center(terrain);
                        LocalContact lc{c.b + (-a_c), c.a + (-b_c), neg(c)};
                        last.push_back(ConstraintRecord{
                            (uint32_t)i, -1, (uint32_t)face_index, sub++,
                            ContactConstraint::make(bodies, RigidBodyRef::Dynamic(i),
                                                    RigidBodyRef::Static(center(terrain), 0.0f),
                                                    manifold_from(lc), dt)});
                    });
                });
            }
            if (i == 0) continue;  // world.rs:256
            bvh.query(b, [&](const size_t& j) {
                if (j >= i) return;  // world.rs:266
                candidate_pairs++;
                ContactPruner pruner;
                local_contacts(collider, bodies.collider[j], [&](const LocalContact& lc) { pruner.push(lc); });
                Manifold m = manifold_from(pruner);
                if (m.len() == 0) return;
                last.push_back(ConstraintRecord{(uint32_t)i, (int32_t)j, 0, 0,
                                                ContactConstraint::make(bodies, RigidBodyRef::Dynamic(i),
                                                                        RigidBodyRef::Dynamic(j), m, dt)});
            });
        }
    }

    // world.rs:293 with the reference's own order.
    void step(float dt, size_t iters) {
        build_constraints(dt);
        for (size_t it = 0; it < iters; ++it)
            for (ConstraintRecord& r : last) r.c.solve(bodies);
    }
    // Same step, but the constraints are solved in the order given by `perm`
    // (perm[k] = index into `last` of the k-th constraint to solve).
    void solve_in_order(const std::vector<uint32_t>& perm, size_t iters) {
        for (size_t it = 0; it < iters; ++it)
            for (uint32_t k : perm) last[k].c.solve(bodies);
    }
};

// mgf_demo/world.rs:118-150: open box, 8 verts, 10 faces, floor at y = pos.y, scaled by
// (hx, wall_h, hz).  The reference is hx = hz = 10, wall_h = 10, pos = (0,-10,0).
inline void make_box_terrain(Mesh& m, float hx, float wall_h, float hz, Vec3 pos) {
    const float V[8][3] = {{-hx, 0, -hz}, {-hx, 0, hz}, {hx, 0, hz}, {hx, 0, -hz},
                           {-hx, wall_h, -hz}, {-hx, wall_h, hz}, {hx, wall_h, hz}, {hx, wall_h, -hz}};
    for (auto& p : V) m.push_vert(v3(p[0], p[1], p[2]));
    const int F[10][3] = {{0, 1, 3}, {1, 2, 3}, {0, 5, 1}, {0, 4, 5}, {0, 3, 7},
                          {0, 7, 4}, {2, 6, 3}, {3, 6, 7}, {1, 5, 2}, {2, 5, 6}};
    for (auto& f : F) m.push_face(f[0], f[1], f[2]);
    m.set_pos(pos);
}

}  // namespace mgfo
