// layout.cuh -- HBM data layout of bodies, work lists and constraints.
// Everything is struct-of-arrays of 16-byte vectors so that a warp touching 32 consecutive
// items issues full 512-byte coalesced requests; the one gathered record (BodyVel) is a
// 64-byte, 64-byte-aligned record = exactly two 32-byte DRAM sectors.
#pragma once
#include "narrow.cuh"

namespace mgfb {

// v, omega, inv_mass and the WORLD inverse inertia of one body: the only body data the solver
// gathers (physics.rs:273-288 `get`).  The solver rewrites the first 32 bytes (v, omega).
struct __align__(64) BodyVel {
    float4 a;  // v.x v.y v.z   w.x
    float4 b;  // w.y w.z inv_mass I.c0.x
    float4 c;  // I.c0.y I.c0.z I.c1.x I.c1.y
    float4 d;  // I.c1.z I.c2.x I.c2.y I.c2.z
};

// Collider = Moving<Component> (physics.rs:154) + its ComponentConstructor (physics.rs:153).
struct __align__(16) Collider {
    float4 p0;  // sphere: c.xyz, r      capsule: a.xyz, r
    float4 p1;  // sphere: unused        capsule: d.xyz ; w = kind as int bits (0 sphere, 1 capsule)
    float4 v;   // Moving.1 = v*dt (delta) ; w = ctor half_h (capsule)
};

struct __align__(16) Box { float4 c, r; };  // AABB centre / half extents (w unused)

HD int col_kind(const Collider& k) { return (int)fbits(k.p1.w); }
HD V3 f4v(float4 f) { return mk3(f.x, f.y, f.z); }
HD float4 v4(V3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
HD float ibits(int i) {
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    union { float f; int32_t i; } u; u.i = i; return u.f;
#endif
}
HD Sph col_sphere(const Collider& k) { Sph s; s.c = f4v(k.p0); s.r = k.p0.w; return s; }
HD Cap col_capsule(const Collider& k) { Cap c; c.a = f4v(k.p0); c.d = f4v(k.p1); c.r = k.p0.w; return c; }
HD V3 col_center(const Collider& k) {  // Shape::center: geom.rs:747,787
    return col_kind(k) == 0 ? f4v(k.p0) : f4v(k.p0) + f4v(k.p1) * 0.5f;
}

HD M3 vel_inertia(const BodyVel& r) {
    return mkm(mk3(r.b.w, r.c.x, r.c.y), mk3(r.c.z, r.c.w, r.d.x), mk3(r.d.y, r.d.z, r.d.w));
}
HD void vel_set_inertia(BodyVel& r, const M3& m) {
    r.b.w = m.c0.x; r.c.x = m.c0.y; r.c.y = m.c0.z; r.c.z = m.c1.x; r.c.w = m.c1.y;
    r.d.x = m.c1.z; r.d.y = m.c2.x; r.d.z = m.c2.y; r.d.w = m.c2.z;
}

// One contact produced by the narrowphase, before it becomes a constraint (64 B + identity).
struct ContactList {
    int* a;          // body i
    int* b;          // body j (< i) or -1 = terrain
    uint32_t* face;  // terrain face (b == -1)
    uint32_t* sub;   // k-th contact of the (body, face) pair
    float4* la;      // local_a.xyz ; w = normal.x
    float4* lb;      // local_b.xyz ; w = normal.y
    float4* nt;      // normal.z, t, unused, unused
};

// Constraint rows in SOLVE ORDER (group-major).  Single-contact rows (everything World::step
// produces) live entirely here; rows with 2..4 contacts keep contact 0 here and the rest in
// the `extra` arrays at index row*3 + (k-1).
struct ConstraintRows {
    int2* ab;        // body a, body b (-1 = static)
    float4* n;       // normal.xyz, bias
    float4* t0;      // tangent0.xyz, tangent_mass0
    float4* t1;      // tangent1.xyz, tangent_mass1
    float4* ra;      // ra.xyz, normal_mass
    float4* rb;      // rb.xyz, ncontacts as int bits
    float* impulse;  // normal_impulse accumulator (read+written every iteration)
    // contacts 1..3 of multi-contact manifolds
    float4* xra;     // ra.xyz, normal_mass
    float4* xrb;     // rb.xyz, bias
    float4* xtm;     // tangent_mass0, tangent_mass1, normal_impulse, unused
};


// Device-resident counters of one step.  Zeroed (except the sticky ones) at step start.
struct Counters {
    unsigned pairs[4];       // body-pair candidates per kind: [ki*2+kj], k = 0 sphere / 1 capsule
    unsigned tpairs[2];      // (body, face) terrain candidates per body kind
    unsigned contacts;       // contacts emitted = constraints
    unsigned tcontacts;      // of which terrain
    unsigned grid_entries;
    unsigned fat_refreshes;
    unsigned max_fat_bits;   // float bits of the largest stored fat-box half extent
    unsigned ngroups;
    unsigned remaining;      // uncoloured constraints
    unsigned bar;            // grid barrier arrival counter
    unsigned rounds;         // colouring rounds taken
    unsigned max_tight_bits; // float bits of the largest tight (swept) box half extent
    unsigned n_total;        // owned bodies + ghost bodies received this step (= n when not tiled)
    unsigned n_edge;         // owned bodies sent to the left neighbour as ghosts this step
    unsigned blocks_done;    // last-block-done counter of k_ghost_send
    unsigned xr_bits;        // ordered-float bits of max(fat.c.x + fat.r.x) over owned bodies
    unsigned n_phases;       // non-empty groups (solver phases per iteration)
    unsigned n_int_phases;   // of which interior (before the boundary exchange); = n_phases when not tiled
    unsigned n_int_rows;     // constraints in interior phases
    unsigned df_links;       // dataflow solver: total (body, row) incidences = sum of rows per body
    unsigned colour_fallback; // k_colour_df ran out of colours: k_order (which stacks colours beyond 64) redoes the step's colouring
    unsigned bp_path;        // which broadphase path ran: 0 plain sweep (cache off), 1 coherent, 2 rebuild, 3 sweep chosen by the cache (bpcache.cuh)
    unsigned n_ref;          // bodies whose stored fat box was replaced this step and that are listed for the coherent broadphase (bpcache.cuh)
    // sticky until the host clears them
    unsigned overflow;       // bit0 pairs, bit1 tpairs, bit2 contacts, bit3 grid entries, bit4 groups, bit5 ghosts
    unsigned nan_bounds;     // AABB::combine assert (bounds.rs:125-127)
    unsigned steps_done;
    unsigned comm_error;     // tiled mode: bit0 neighbour timed out, bit1 tile thinner than its two ghost layers
    // running totals since the last mgfb_step_totals(reset)
    unsigned long long acc_constraints;
    unsigned long long acc_pairs;        // body-body + terrain candidates
    unsigned long long acc_groups;
    unsigned long long acc_steps;
};
// One side of a row's inbox (dataflow solver, kernels.cuh): written by the row's predecessor with ONE 32-byte store.
struct __align__(32) Inbox { float4 lo, hi; };   // v.xyz, tag | omega.xyz, tag

enum { OVF_PAIRS = 1, OVF_TPAIRS = 2, OVF_CONTACTS = 4, OVF_GRID = 8, OVF_GROUPS = 16, OVF_GHOSTS = 32 };
enum { COMM_TIMEOUT = 1, COMM_TILE_TOO_THIN = 2 };
struct PairLists { int2* p[4]; };

struct BodyArrays {
    float4* x;        // position
    float4* q;        // s, x, y, z
    BodyVel* vel;
    float4* force;    // force.xyz, restitution
    float4* torque;   // torque.xyz, friction
    float4* imb;      // inv_moment_body: 3 float4 per body (columns)
    Collider* col;
    Box* tight;
    Box* fat;
    unsigned* gid;    // global body id (= index when the world is not tiled): orders pairs (j < i, world.rs:266)
    // coherent broadphase (bpcache.cuh), or NULL: which bodies replaced their stored fat box this step
    unsigned char* ref_flag; unsigned* ref_list; unsigned ref_cap;
};

}  // namespace mgfb
