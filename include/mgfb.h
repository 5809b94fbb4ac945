/* mgfb.h -- C ABI of the B200-native implementation of mgf's per-step physics hot path.
 *
 * This is the drop-in boundary: every entry point replaces one piece of
 * the reference's Rust API on that path; the reference item (file:line under the mgf source
 * tree) is cited on each declaration.  A Rust `mgf-sys` crate binds these 1:1 (see
 * INTEGRATION.md).  POD only, no callbacks across the boundary, no unwinding: every function
 * returns an int32 status and mgfb_last_error() gives the message.
 *
 * Conventions
 *   - all reals are IEEE f32; vectors are 3 packed floats (x,y,z); quaternions are 4 floats
 *     (s, x, y, z) like cgmath's Quaternion{s, v}; matrices are 9 floats column-major.
 *   - a ctx owns one CUDA device, one stream and all device memory; it is NOT thread-safe.
 *   - host pointers passed in are only read/written during the call.
 *   - the reference's closures invoked "0..n times in a defined order" become output arrays
 *     plus per-item counts, in the same per-item order.
 *   - there is no CPU fallback: without a CUDA device mgfb_ctx_create fails with MGFB_ERR_CUDA.
 */
#ifndef MGFB_H
#define MGFB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MGFB_ABI_VERSION 1

enum mgfb_status {
    MGFB_OK = 0,
    MGFB_ERR_INVALID_ARG = 1,      /* assert!(radius > 0.0) geom.rs:300,328; bad index (Rust bounds panic) */
    MGFB_ERR_SINGULAR_INERTIA = 2, /* .invert().unwrap() physics.rs:212 */
    MGFB_ERR_CAPACITY = 3,         /* a device work list overflowed even after growing */
    MGFB_ERR_CUDA = 4,
    MGFB_ERR_NAN_BOUNDS = 5,       /* AABB::combine asserts r >= 0, bounds.rs:125-127 */
    MGFB_ERR_STATE = 6,            /* call order violated (e.g. step without terrain is fine; replay without a step is not) */
    MGFB_ERR_TILE = 7              /* tiled world: a neighbour tile did not answer in time, or the tile is thinner than its ghost layers */
};

enum mgfb_shape_kind {
    MGFB_SPHERE = 0,    /* geom.rs:290   p = c[3], r                     */
    MGFB_CAPSULE = 1,   /* geom.rs:316   p = a[3], d[3], r               */
    MGFB_TRIANGLE = 2,  /* geom.rs:128   p = a[3], b[3], c[3]            */
    MGFB_RECTANGLE = 3, /* geom.rs:216   p = c[3], u0[3], u1[3], e0, e1  */
    MGFB_PLANE = 4,     /* geom.rs:32    p = n[3], d                     */
    MGFB_AABB = 5,      /* geom.rs:257   p = c[3], r[3]                  */
    MGFB_OBB = 6,       /* geom.rs:272   p = c[3], r[3], q[4] (s,x,y,z)  */
    MGFB_CONVEX_MESH = 7 /* mesh.rs:141  p = first vertex, vertex count (whole numbers < 2^24) in the pool of mgfb_convex_vertices_set;
                            GJK / EPA only */
};

/* A shape (optionally swept: geom.rs:357 Moving<T>(T, Vector3)). 64 bytes. */
typedef struct mgfb_shape {
    uint32_t kind;      /* mgfb_shape_kind */
    float p[12];
    float v[3];         /* Moving<T>.1 ; zero for a static shape */
} mgfb_shape;

/* collision.rs:431 Contact */
typedef struct mgfb_contact { float a[3], b[3], n[3], t; } mgfb_contact;
/* collision.rs:1410 LocalContact */
typedef struct mgfb_local_contact { float local_a[3], local_b[3]; mgfb_contact global; } mgfb_local_contact;

/* Which `Contacts` impl a homogeneous batch runs: one divergence-free kernel specialisation each. */
enum mgfb_pair_kind {
    MGFB_SPHERE_X_MSPHERE = 0,   /* collision.rs:1089 Sphere.contacts(&Moving<Sphere>)   */
    MGFB_CAPSULE_X_MSPHERE = 1,  /* collision.rs:1145 Capsule.contacts(&Moving<Sphere>)  */
    MGFB_SPHERE_X_MCAPSULE = 2,  /* collision.rs:1143 commute -> :1368 -> :1145          */
    MGFB_CAPSULE_X_MCAPSULE = 3, /* collision.rs:1205                                    */
    MGFB_PLANE_X_MSPHERE = 4,    /* collision.rs:521                                     */
    MGFB_PLANE_X_MCAPSULE = 5,   /* collision.rs:555                                     */
    MGFB_TRI_X_MSPHERE = 6,      /* collision.rs:610 (Poly = Triangle)                   */
    MGFB_TRI_X_MCAPSULE = 7,     /* collision.rs:693 (Poly = Triangle)                   */
    MGFB_RECT_X_MSPHERE = 8,     /* collision.rs:610 (Poly = Rectangle)                  */
    MGFB_RECT_X_MCAPSULE = 9,    /* collision.rs:693 (Poly = Rectangle)                  */
    MGFB_MCOMP_X_MCOMP = 10,     /* compound.rs:192 Moving<Component>.local_contacts(&Moving<Component>) */
    MGFB_MCOMP_X_TRI = 11,       /* collision.rs:1490 + mesh.rs:119-137: body vs one terrain triangle,
                                    recv = Moving component, arg = triangle already offset by mesh.x;
                                    local_b is relative to arg.v (pass mesh.x there)     */
    MGFB_PAIR_KIND_COUNT = 12
};

/* How the device walks the constraint rows of Solver::solve (solver.rs:72-78).  Both produce the
 * reference's sequential sweep over the same row order, bit for bit. */
enum mgfb_solver_schedule {
    MGFB_SCHEDULE_DATAFLOW = 0,  /* default: per-body version counters, a row runs as soon as both its bodies are ready */
    MGFB_SCHEDULE_PHASES = 1,    /* one grid-wide barrier per colour per iteration */
    MGFB_SCHEDULE_PHASES_JP = 2  /* as PHASES, and the colouring itself by barrier-synchronised Jones-Plassmann rounds
                                    (priorities = identity hash only) instead of the chain colouring (priorities =
                                    geometric class, then hash): a different, equally valid proper colouring */
};

/* The order in which World::step hands its constraints to the Solver (Gauss-Seidel results depend on it).
 *   COLOURED   (default) the device's colour-major order, exported by mgfb_step_constraints so that it can be replayed.
 *   REFERENCE  the demo world's own order (world.rs:233-291): body by body, terrain contacts in the mesh BVH's callback order,
 *              then body pairs in the body BVH's callback order.  Both trees are replayed on the host exactly as bvh.rs grows
 *              and re-balances them, step after step, so the state after every step is BIT-IDENTICAL to the reference's own
 *              World::step.  The tree walk is sequential like the reference's: for parity runs and worlds up to ~10^5 bodies;
 *              one GPU, synchronous steps only. */
enum mgfb_step_order { MGFB_STEP_ORDER_COLOURED = 0, MGFB_STEP_ORDER_REFERENCE = 1 };

/* solver.rs:265-279 ContactConstraintParams, manifold.rs:27-39 PruningParams, world.rs:181 */
typedef struct mgfb_config {
    int32_t device;                 /* CUDA device ordinal */
    float penetration_slop;         /* 0.05  solver.rs:277 */
    float baumgarte;                /* 0.2   solver.rs:278 */
    float persistent_threshold_sq;  /* 0.5   manifold.rs:38 */
    float fat_margin;               /* 0.25  world.rs:181,237 */
    uint32_t initial_body_capacity; /* 0 = default */
    uint32_t max_cooperative_ctas;  /* 0 = one CTA per SM; lower it when several contexts must share one GPU */
    uint32_t tile_timeout_ms;       /* 0 = 20000: how long a tile waits for a neighbour before MGFB_ERR_TILE */
    uint32_t solver_schedule;       /* mgfb_solver_schedule; coloured order only (as-given order always uses phases); every tile
                                       of a tiled world must use the same value */
    uint32_t step_order;            /* mgfb_step_order: in which order mgfb_step adds the constraints to the solver */
} mgfb_config;
void mgfb_config_default(mgfb_config* cfg);

typedef struct mgfb_ctx mgfb_ctx;
int32_t mgfb_abi_version(void);
int32_t mgfb_ctx_create(const mgfb_config* cfg, mgfb_ctx** out);
void mgfb_ctx_destroy(mgfb_ctx* ctx);
const char* mgfb_last_error(const mgfb_ctx* ctx);   /* valid until the next call on ctx; ctx may be NULL */
int32_t mgfb_synchronize(mgfb_ctx* ctx);

/* ---------------- RigidBodyVec (physics.rs:141-315) ---------------- */
/* RigidBodyVec::add_body (physics.rs:200-218) for n bodies.  shapes[i].kind in {SPHERE, CAPSULE};
 * shapes[i].v is ignored (a new collider is Moving::sweep(collider, 0)).  world_force is the
 * per-unit-mass force (force = world_force * mass).  *first_id receives the index of the first
 * new body (RigidBodyRef::Dynamic(id)).  Also inserts the fat AABB (bounds + fat_margin) that
 * the demo world keeps in its body BVH (world.rs:178-184).  A context holds fewer than 2^29 bodies (ghosts included) and
 * 2^27 constraints per step (bits of the pair and chain-link words); more is MGFB_ERR_CAPACITY. */
int32_t mgfb_bodies_add(mgfb_ctx* ctx, uint32_t n, const mgfb_shape* shapes, const float* mass,
                        const float* restitution, const float* friction, const float* world_force /* n*3 */,
                        uint32_t* first_id);
int32_t mgfb_bodies_count(const mgfb_ctx* ctx, uint32_t* n);
/* pub fields x, q (physics.rs:142-143) and ConstrainedSet::get's Velocity (physics.rs:273-281).
 * Any output pointer may be NULL. */
int32_t mgfb_bodies_get_state(mgfb_ctx* ctx, uint32_t first, uint32_t n, float* x /* n*3 */, float* q /* n*4 */,
                              float* v /* n*3 */, float* omega /* n*3 */);
/* ConstrainedSet::set (physics.rs:306-314) for Dynamic(first..first+n). */
int32_t mgfb_bodies_set_velocity(mgfb_ctx* ctx, uint32_t first, uint32_t n, const float* v, const float* omega);
/* pub field collider: Vec<Moving<Component>> (physics.rs:154) / colliders() (:256). */
int32_t mgfb_bodies_get_colliders(mgfb_ctx* ctx, uint32_t first, uint32_t n, mgfb_shape* out);
/* world inverse inertia (physics.rs:152,232), 9 floats column-major per body. */
int32_t mgfb_bodies_get_inv_moment(mgfb_ctx* ctx, uint32_t first, uint32_t n, float* out /* n*9 */);
/* The stored fat AABB of each body in the demo world's body BVH (world.rs:180-181, refreshed lazily at :235-238;
 * `bvh[bvh_ids[i]]`, bvh.rs:483): centre[3], half extents[3] per body. */
int32_t mgfb_bodies_get_fat_bounds(mgfb_ctx* ctx, uint32_t first, uint32_t n, float* boxes /* n*6 */);
/* Writes the pub fields x, q, collider (physics.rs:142-154), the velocities (ConstrainedSet::set, physics.rs:306) and the
 * stored fat AABBs of bodies first..first+n: together with get_state / get_colliders / get_fat_bounds this saves and
 * restores a world bit for bit (checkpoint, or loading one state into two implementations).  Any pointer may be NULL
 * (that field is kept).  A collider keeps its Component kind and radius (MGFB_ERR_INVALID_ARG otherwise); colliders[i].v is
 * Moving.1, the displacement complete_motion adds next (physics.rs:262-269).  The world inverse inertia follows q
 * (physics.rs:231-232). */
int32_t mgfb_bodies_set_state(mgfb_ctx* ctx, uint32_t first, uint32_t n, const float* x /* n*3 */, const float* q /* n*4 */,
                              const float* v /* n*3 */, const float* omega /* n*3 */, const mgfb_shape* colliders /* n */,
                              const float* fat_boxes /* n*6 */);
/* RigidBodyVec::integrate (physics.rs:222-253) */
int32_t mgfb_integrate(mgfb_ctx* ctx, float dt);
/* RigidBodyVec::complete_motion (physics.rs:262-269) */
int32_t mgfb_complete_motion(mgfb_ctx* ctx);

/* ---------------- Mesh (mesh.rs:32-139) ---------------- */
/* Mesh::new + push_vert* + push_face* + set_pos(x) (mesh.rs:40-73, geom.rs:459).  Builds the
 * device triangle index that replaces the mesh's BVH<AABB,usize>.  The mesh becomes the
 * world's terrain (world.rs:72,150).  Face winding decides the normal (world.rs:137-141). */
int32_t mgfb_terrain_set(mgfb_ctx* ctx, const float* verts /* nverts*3 */, uint32_t nverts,
                         const uint32_t* faces /* nfaces*3 */, uint32_t nfaces, const float x[3]);

/* ---------------- Contacts / LocalContacts (collision.rs:471,1441) ---------------- */
/* Runs `recv[i].contacts(&arg[i], cb)` for i in 0..n on the device, one specialised kernel per
 * pair_kind.  out holds 2 slots per pair (the most any continuous impl emits); counts[i] says
 * how many are valid, in callback order.  For MGFB_MCOMP_* kinds `out_local` (n*2) is filled as
 * well (may be NULL otherwise). */
int32_t mgfb_contacts_batch(mgfb_ctx* ctx, uint32_t pair_kind, const mgfb_shape* recv, const mgfb_shape* arg,
                            uint32_t n, mgfb_contact* out /* n*2 */, mgfb_local_contact* out_local /* n*2 or NULL */,
                            uint32_t* counts /* n */);

/* ---------------- ray casts: Intersects<RHS> for Particle (collision.rs:163-373) ---------------- */
enum mgfb_particle_kind {
    MGFB_RAY = 0,      /* geom.rs:341,818  Ray{p, d}:     6 floats p[3], d[3];  pos = p, dir = d,     DT = inf */
    MGFB_SEGMENT = 1   /* geom.rs:349,842  Segment{a, b}: 6 floats a[3], b[3];  pos = a, dir = b - a, DT = 1   */
};
/* collision.rs:151 Intersection */
typedef struct mgfb_intersection { float p[3]; float t; } mgfb_intersection;
/* `particles[i].intersection(&shapes[i])` (Option<Intersection>): hit[i] = 1 and out[i] when Some.  Shapes: PLANE
 * (collision.rs:169), TRIANGLE / RECTANGLE (:186, the plane hit must lie inside the face), AABB (:202), OBB (:238),
 * SPHERE (:249), CAPSULE (:275), and Moving<Sphere> = a SPHERE with v != 0 (:361, the capsule its sweep covers).
 * One thread per query, shapes of mixed kinds allowed (a warp diverges on the kind). */
int32_t mgfb_intersections_batch(mgfb_ctx* ctx, uint32_t particle_kind, const float* particles /* n*6 */, const mgfb_shape* shapes /* n */,
                                 uint32_t n, mgfb_intersection* out /* n */, uint32_t* hit /* n */);

/* ---------------- BVH<AABB, V> (src/bvh.rs:86-369), V = u32 ----------------
 * The tree callers outside World::step use (ray picking, Compound, user code).  The reference grows it one insert
 * at a time; here the leaf set lives on the device and the tree over it is rebuilt there (Morton order, implicit
 * binary tree) before the first query after a change.  Preserved: WHICH leaves a query / raytrace reports (exactly
 * the leaves whose own box passes the reference's test; the reference may additionally prune a leaf that merely
 * touches, because its parents are rounded unions, bounds.rs:113-130).  Not preserved: the callback order (here
 * ascending Morton order of the leaf centres, deterministic) and the numeric value of the indices insert returns
 * (stable leaf handles, reused last-freed-first like pool.rs; the reference numbers its internal nodes too).
 * Boxes are 6 floats: centre[3], half extents[3] (geom.rs:257 AABB{c, r}). */
typedef struct mgfb_bvh mgfb_bvh;
int32_t mgfb_bvh_create(mgfb_ctx* ctx, mgfb_bvh** out);                                   /* BVH::new        bvh.rs:88  */
void mgfb_bvh_destroy(mgfb_bvh* bvh);
/* BVH::insert(&key, value) -> index, n at a time (bvh.rs:125).  Half extents must be >= 0 (bounds.rs:125-127). */
int32_t mgfb_bvh_insert(mgfb_bvh* bvh, const float* boxes /* n*6 */, const uint32_t* values /* n */, uint32_t n, uint32_t* indices /* n */);
/* BVH::remove(index) (bvh.rs:225); an index that holds no leaf is MGFB_ERR_INVALID_ARG (the reference panics, pool.rs:100). */
int32_t mgfb_bvh_remove(mgfb_bvh* bvh, const uint32_t* indices, uint32_t n);
/* Index<usize> / get_leaf (bvh.rs:270, :483): the leaf's box and value; either output may be NULL. */
int32_t mgfb_bvh_get(const mgfb_bvh* bvh, uint32_t index, float* box /* 6 */, uint32_t* value);
int32_t mgfb_bvh_len(const mgfb_bvh* bvh, uint32_t* n);
/* BVH::query(&arg, cb) for nq argument boxes (bvh.rs:283-310): values[offsets[q] .. offsets[q+1]) are the values of the
 * leaves overlapping boxes[q] (AABB::overlaps, collision.rs:22-29, closed).  *total = all results; when it exceeds
 * `capacity` nothing is written to `values` and the call returns MGFB_ERR_CAPACITY (offsets and *total are valid). */
int32_t mgfb_bvh_query_batch(mgfb_bvh* bvh, const float* boxes /* nq*6 */, uint32_t nq, uint32_t* offsets /* nq+1 */, uint32_t* values,
                             uint32_t capacity, uint32_t* total);
/* BVH::raytrace(&arg, cb) for nq Rays / Segments (bvh.rs:340-369): every leaf whose box the particle intersects
 * (collision.rs:202-236), with the Intersection the callback would receive. */
int32_t mgfb_bvh_raytrace_batch(mgfb_bvh* bvh, uint32_t particle_kind, const float* particles /* nq*6 */, uint32_t nq, uint32_t* offsets /* nq+1 */,
                                uint32_t* values, mgfb_intersection* hits, uint32_t capacity, uint32_t* total);

/* ---------------- Compound (src/compound.rs:232-352) ----------------
 * An aggregate of Components (compound.rs:33-37: a Sphere or a Capsule, given in the compound's own frame) with a
 * displacement `disp` and a rotation `rot` (assumed normalised, like the reference assumes).  Compound::new inserts the
 * components one by one into a BVH<AABB, Component>; that tree is grown here exactly as bvh.rs grows it, so the queries
 * below visit the same leaves in the same ORDER as the reference's callbacks (the last contact delivered is what
 * `last_contact`, collision.rs:477, returns).  Intersects<Component> itself (compound.rs:150-157) needs no entry point of its
 * own: a Component is a SPHERE or a CAPSULE shape of mgfb_intersections_batch. */
typedef struct mgfb_compound mgfb_compound;
/* Compound::new(components) (compound.rs:247-262); disp = 0, rot = identity. */
int32_t mgfb_compound_create(mgfb_ctx* ctx, const mgfb_shape* components /* n, SPHERE | CAPSULE */, uint32_t n, mgfb_compound** out);
void mgfb_compound_destroy(mgfb_compound* c);
/* the pub fields disp and rot (compound.rs:236-238, AddAssign / SubAssign :265-275); rot = (s, x, y, z) */
int32_t mgfb_compound_set_transform(mgfb_compound* c, const float disp[3], const float rot[4]);
/* BoundedBy<AABB> (compound.rs:277-281: the tree's root box rotated, + disp) and BoundedBy<Sphere> (:283-288); either may be NULL.
 * An empty compound is MGFB_ERR_STATE (the reference panics: "BVH is empty", bvh.rs:263). */
int32_t mgfb_compound_bounds(const mgfb_compound* c, float aabb[6] /* centre, half extents */, float sphere[4] /* centre, radius */);
/* Shape::closest_point (compound.rs:299-311) for n points.  Like the reference, the components are taken as stored (disp and
 * rot are not applied) and ties keep the earlier component. */
int32_t mgfb_compound_closest_points(mgfb_compound* c, const float* to /* n*3 */, uint32_t n, float* out /* n*3 */);
/* Intersects<Compound> for Ray / Segment (compound.rs:314-337): the earliest hit over the components the traced tree reaches. */
int32_t mgfb_compound_intersections_batch(mgfb_compound* c, uint32_t particle_kind, const float* particles /* n*6 */, uint32_t n,
                                          mgfb_intersection* out /* n */, uint32_t* hit /* n */);
/* `compound.contacts(&rhs[i], cb)` (compound.rs:339-357) for RHS = Moving<Sphere | Capsule | Triangle | Rectangle> (shape + v):
 * out[i*slots ..] receives the contacts in callback order (a on the compound's component, b on the rhs), counts[i] how many
 * callbacks fired (only the first `slots` are stored; 2 * n_components always suffices). */
int32_t mgfb_compound_contacts_batch(mgfb_compound* c, const mgfb_shape* rhs /* n */, uint32_t n, uint32_t slots, mgfb_contact* out /* n*slots */,
                                     uint32_t* counts /* n */);

/* ---------------- discrete path: GJK + EPA (simplex.rs:172-553) ---------------- */
/* The vertices of every ConvexMesh (mesh.rs:141-189: a closed convex point soup, `ConvexMesh::from(verts)` / `push`) the
 * GJK calls below may name: one pool per context, replaced by each call; a MGFB_CONVEX_MESH shape is a slice of it.  Only
 * Convex::support (mesh.rs:223-236) reads a ConvexMesh on this path: the vertices as stored, the first of equally good ones. */
int32_t mgfb_convex_vertices_set(mgfb_ctx* ctx, const float* verts /* n*3 */, uint32_t n);
/* `a[i].contacts(&b[i], cb)` for static convex pairs through the generic impl for Convex + Volumetric shapes
 * (collision.rs:497-519): GJK seeded along +-y (simplex.rs:172-200), then EPA (simplex.rs:456-553, at most
 * 101 iterations); the contact has t = 0, `a` on shape a, `b = a - depth * n`.  Shapes: SPHERE, CAPSULE, AABB,
 * OBB (the Convex implementors of geom.rs:1027-1072) and CONVEX_MESH (mesh.rs:223).  status[i]: 0 = no contact (the callback is not invoked),
 * 1 = contact in out[i], 2 = EPA polytope outgrew the device's fixed capacity (254 faces), 3 = GJK did not
 * converge in 256 steps (the reference's loop has no other exit there: NaN input, or separated polytopes whose
 * support point never satisfies |min|^2 >= |support|^2, simplex.rs:195), 4 = the reference panics (EPA indexes a
 * free Pool slot, pool.rs:111).  epa_iters may be NULL. */
int32_t mgfb_gjk_batch(mgfb_ctx* ctx, const mgfb_shape* a, const mgfb_shape* b, uint32_t n, mgfb_contact* out /* n */,
                       uint32_t* status /* n */, uint32_t* epa_iters /* n or NULL */);
/* Penetrates::separation (collision.rs:404-425): GJK seeded along +-x; status[i] = 1 and separation[i] =
 * distance for Some(d), status[i] = 0 for None (the shapes overlap). */
int32_t mgfb_separation_batch(mgfb_ctx* ctx, const mgfb_shape* a, const mgfb_shape* b, uint32_t n, float* separation /* n */,
                              uint32_t* status /* n */);

/* ---------------- Solver / ContactConstraint (solver.rs:53-262) ---------------- */
/* A batch of ContactConstraint::new inputs (solver.rs:101): obj_a/obj_b, Manifold.
 * obj < 0 means RigidBodyRef::Static{center, friction} taken from static_center/static_friction.
 * Manifolds hold up to 4 contacts (manifold.rs:117 SmallVec<[_;4]>). */
typedef struct mgfb_manifolds {
    uint32_t n;
    const int32_t* obj_a;          /* n */
    const int32_t* obj_b;          /* n */
    const float* static_center;    /* n*3, used when obj_b < 0 (or obj_a < 0) */
    const float* static_friction;  /* n */
    const float* normal;           /* n*3   Manifold.normal */
    const float* tangent;          /* n*6   Manifold.tangent_vector[2] */
    const uint32_t* ncontacts;     /* n     1..4 */
    const float* local_a;          /* n*4*3 Manifold.contacts[k].0 */
    const float* local_b;          /* n*4*3 Manifold.contacts[k].1 */
} mgfb_manifolds;

/* ContactPruner::new + push(contact)* + Manifold::from(pruner) (manifold.rs:42-148) for `ngroups` pairs of objects: group g
 * owns contacts[offsets[g] .. offsets[g+1]) in push order (what a caller collects from the LocalContacts callbacks of every
 * collider pair of two objects).  push keeps the earliest collision time (+-1e-6), merges a contact with a kept one when
 * either pair of global points is within persistent_threshold_sq (mgfb_config; the one further from the centres survives) and
 * appends it otherwise; the manifold's normal is the mean of the kept normals (not renormalised), its tangents
 * compute_basis(normal) (geom.rs:1138-1145).  Outputs are laid out like mgfb_manifolds, so they feed mgfb_solver_solve as they
 * are; ncontacts[g] is the pruner's length (0 for an empty group: the normal is then 0/0 = NaN like the reference's; above 4 only
 * the first 4 points are stored -- the reference's SmallVec spills to the heap there, a solver row here holds 4).  `time` may be NULL. */
int32_t mgfb_manifolds_prune(mgfb_ctx* ctx, const mgfb_local_contact* contacts, const uint32_t* offsets /* ngroups+1 */, uint32_t ngroups,
                             float* time /* ngroups */, float* normal /* ngroups*3 */, float* tangent /* ngroups*6 */, uint32_t* ncontacts /* ngroups */,
                             float* local_a /* ngroups*12 */, float* local_b /* ngroups*12 */);

enum mgfb_solve_order {
    /* Bit-identical to the reference's sequential Gauss-Seidel over the list as given
     * (solver.rs:72-78): constraints are level-scheduled, level(c) = 1 + max level of any earlier
     * constraint sharing a dynamic body; levels run in order, a level runs in parallel. */
    MGFB_ORDER_AS_GIVEN = 0,
    /* Throughput order: the device greedy-colours the constraint graph and solves colour by
     * colour.  This equals the reference's sequential sweep over the list permuted colour-major;
     * the permutation is returned so a caller (or the oracle) can reproduce it exactly. */
    MGFB_ORDER_COLOURED = 1
};

typedef struct mgfb_solve_stats {
    uint32_t constraints;      /* constraints solved per iteration */
    uint32_t contacts;         /* contact points (sum of ncontacts) */
    uint32_t groups;           /* levels or colours per iteration */
    uint32_t iterations;
    float solve_ms;            /* device time of the solve kernel (CUDA events) */
    uint32_t reserved[3];
} mgfb_solve_stats;

/* Solver::new + add_constraint(ContactConstraint::new(&bodies, a, b, manifold, dt))* +
 * solve(&mut bodies, iters) (solver.rs:59-78,101).  perm_out (n, may be NULL) receives the
 * solve order: perm_out[k] = index in `m` of the k-th constraint solved in each iteration.
 * normal_impulse_out (n*4, may be NULL) receives ContactState.normal_impulse after the solve. */
int32_t mgfb_solver_solve(mgfb_ctx* ctx, const mgfb_manifolds* m, float dt, uint32_t iters, uint32_t order,
                          uint32_t* perm_out, float* normal_impulse_out, mgfb_solve_stats* stats);

/* ---------------- World::step (mgf_demo/world.rs:227-294) ---------------- */
typedef struct mgfb_step_stats {
    uint32_t bodies;
    uint32_t candidate_pairs;     /* body-body AABB candidates with j < i (world.rs:261-268) */
    uint32_t terrain_candidates;  /* (body, face) candidates from the mesh query (mesh.rs:121) */
    uint32_t constraints;         /* ContactConstraints added to the solver */
    uint32_t terrain_constraints;
    uint32_t groups;              /* colours (or levels) per solver iteration */
    uint32_t iterations;
    uint32_t fat_refreshes;       /* bodies whose stored fat AABB was replaced (world.rs:235-238) */
    float step_ms;                /* device time of the whole step */
    float solve_ms;               /* device time of the solve kernel alone */
    uint32_t overflow;            /* nonzero if a work list had to be regrown and the step rerun */
    uint32_t colouring_rounds;    /* sweeps of the chain colouring, or Jones-Plassmann rounds (diagnostic) */
    uint32_t ghosts;              /* tiled world: bodies received from the right neighbour this step */
    uint32_t boundary_constraints;/* tiled world: constraints between an owned body and a ghost */
    uint32_t phases;              /* non-empty groups = grid-wide phases per solver iteration */
    uint32_t broadphase_path;     /* how the pair set was found: 0 grid + sweep, 1 coherent (cached superset filtered, replaced bodies queried),
                                   * 2 cache rebuilt, 3 grid + sweep chosen by the cache (too much moved); the set is the same */
} mgfb_step_stats;

/* One World::step(dt) with `iters` solver iterations (the demo hard-codes 20, world.rs:293).
 * The body-pair set is the reference's (fat-AABB query, j < i); constraints are solved in the
 * device's colour-major order (MGFB_ORDER_COLOURED), exported by mgfb_step_constraints. */
int32_t mgfb_step(mgfb_ctx* ctx, float dt, uint32_t iters, mgfb_step_stats* stats /* may be NULL */);
/* Enqueue `nsteps` steps with no host synchronisation in between (stats of the last one). */
int32_t mgfb_step_n(mgfb_ctx* ctx, float dt, uint32_t iters, uint32_t nsteps, mgfb_step_stats* stats);

/* One step with a CUDA event between its phases: where the step's device time goes (diagnostic; the events cost a few
 * microseconds, so mgfb_step's own step_ms is the number to quote).  MGFB_PHASE_TERRAIN runs on a
 * side stream beside BODY_GRID..NARROW_BODIES and is timed there. */
enum mgfb_phase {
    MGFB_PHASE_INTEGRATE = 0,      /* complete_motion + integrate + swept / fat AABBs (physics.rs:222-269, world.rs:235-238)  */
    MGFB_PHASE_BODY_GRID = 1,      /* broadphase index over the stored fat boxes (replaces the incremental tree, bvh.rs)       */
    MGFB_PHASE_PAIR_SWEEP = 2,     /* BVH::query + j < i filter (world.rs:261-268)                                            */
    MGFB_PHASE_NARROW_BODIES = 3,  /* Moving<Component> x Moving<Component> LocalContacts (compound.rs:192-207)               */
    MGFB_PHASE_TERRAIN = 4,        /* Mesh::contacts: mesh query + Polygon x Moving<Sphere|Capsule> (mesh.rs:115-139)         */
    MGFB_PHASE_COLOURING = 5,      /* solve order: chain colouring, group scan, row scatter                                   */
    MGFB_PHASE_BUILD_ROWS = 6,     /* ContactConstraint::new (solver.rs:101-191)                                              */
    MGFB_PHASE_SOLVE = 7,          /* Solver::solve (solver.rs:72-78)                                                         */
    MGFB_PHASE_COUNT = 8
};
typedef struct mgfb_phase_profile {
    float phase_ms[8];          /* indexed by mgfb_phase */
    uint32_t pairs[4];          /* body-pair candidates per shape pair: [receiver kind * 2 + argument kind], 0 = sphere, 1 = capsule */
    uint32_t terrain_pairs[2];  /* (body, face) candidates per body kind */
    uint32_t body_contacts;     /* LocalContacts emitted by the body-body narrowphase */
    uint32_t terrain_contacts;  /* LocalContacts emitted by the terrain narrowphase */
} mgfb_phase_profile;
int32_t mgfb_step_profile(mgfb_ctx* ctx, float dt, uint32_t iters, mgfb_step_stats* stats /* may be NULL */, mgfb_phase_profile* out);

/* Pipelined World::step for callers that move state across PCIe every step (rendering, logging, a host-side
 * controller).  mgfb_step_enqueue queues  [v_in, omega_in -> device]  ->  the step  ->  [x, q, v, omega -> host]
 * and returns at once; mgfb_step_wait blocks until the OLDEST queued step's outputs are in the caller's buffers
 * and returns its stats.  Up to four steps may be in flight: the transfers of step k (separate copy streams, both
 * directions) overlap the kernels of the steps behind it.  Same kernels and bit-identical results as mgfb_step.
 * v_in/omega_in (n*3 each, both or neither) are applied before the step: MGFB_INPUT_SET overwrites the velocities
 * (ConstrainedSet::set, physics.rs:304), MGFB_INPUT_ADD adds to them (what `bodies.v[i] += dv` on the pub fields does
 * between two steps of the reference: external impulses).  Any output pointer may be NULL.  Host buffers must be page-locked and must not be touched until the
 * matching wait returns.  Work lists are sized once (16 pairs / 16 contacts per body); when a step overflows one all the
 * same, its mgfb_step_wait drains the pipeline, regrows the lists, re-runs that step from after its integration and
 * re-queues the younger steps whole (stats.overflow != 0 tells; results are unchanged).  Only a TILED world cannot regrow
 * (its neighbours hold pointers into the lists): there the wait returns MGFB_ERR_CAPACITY.  On a tiled world every rank
 * must enqueue and wait the same sequence of steps. */
enum mgfb_input_mode { MGFB_INPUT_SET = 0, MGFB_INPUT_ADD = 1 };
int32_t mgfb_step_enqueue(mgfb_ctx* ctx, float dt, uint32_t iters, uint32_t input_mode, const float* v_in, const float* omega_in,
                          float* x_out, float* q_out, float* v_out, float* omega_out);
int32_t mgfb_step_wait(mgfb_ctx* ctx, mgfb_step_stats* stats /* may be NULL */);

/* Identity of the constraints of the most recent step, in the order they were solved:
 * body_a = i, body_b = j (< i) or -1 for terrain, face = terrain face index, sub = k-th contact
 * of that (body, face) pair.  Arrays hold stats.constraints entries; any may be NULL. */
int32_t mgfb_step_constraints(mgfb_ctx* ctx, uint32_t capacity, uint32_t* body_a, int32_t* body_b,
                              uint32_t* face, uint32_t* sub, uint32_t* colour, uint32_t* count);

/* Running totals over all steps since the last reset (device-side accumulators, one readback):
 * steps completed, ContactConstraints solved per iteration summed over steps, candidate pairs
 * (body-body + terrain) summed over steps, colour groups summed over steps, and the number of
 * CUDA kernels this ctx launched.  Any pointer may be NULL. */
int32_t mgfb_step_totals(mgfb_ctx* ctx, uint64_t* steps, uint64_t* constraints, uint64_t* candidate_pairs, uint64_t* groups,
                         uint64_t* kernel_launches, int32_t reset);

/* ---------------- one world tiled across GPUs (no counterpart in mgf) ----------------
 * Each GPU (one process per GPU, or one ctx per GPU) owns a slab of the bodies.  Once per step the
 * right neighbour's bodies that can touch this tile arrive as "ghosts" (written by the neighbour's
 * kernel straight into this ctx's body arrays over NVLink peer memory); constraints between an owned
 * body and a ghost are solved here, inside the same persistent solver kernel: with the default
 * dataflow schedule the constraint chain of a boundary body simply continues on the neighbour GPU
 * (each hand-over is one 32-byte peer store into the next constraint's inbox); with
 * MGFB_SCHEDULE_PHASES the ghost velocities are exchanged once per iteration between barrier phases.
 * The executed constraint order is a valid sequential order
 * (per iteration: every tile's interior constraints, then every tile's boundary constraints), exported
 * by mgfb_step_constraints with GLOBAL body ids so a caller (or the oracle) can replay it.
 *
 *   setup:  add this tile's bodies; mgfb_bodies_set_gid; mgfb_terrain_set (same mesh on every tile);
 *           mgfb_tile_export; exchange the descriptors between ranks (host plumbing, e.g.
 *           torch.distributed all_gather); mgfb_tile_connect; then every rank calls mgfb_step /
 *           mgfb_step_n the same number of times.
 *   rules:  tiles are ordered along x; rank r's bodies must only ever touch bodies of ranks r-1, r, r+1
 *           and no body may touch both neighbours (MGFB_ERR_TILE otherwise): tiles must stay wider than
 *           two ghost layers.  Ownership is fixed until the caller re-tiles. */
/* The global id of bodies first..first+n (default: the local index).  Pairs are generated as
 * (i, j) with gid_j < gid_i (world.rs:266), so the global ids define which body is the receiver. */
int32_t mgfb_bodies_set_gid(mgfb_ctx* ctx, uint32_t first, uint32_t n, const uint32_t* gids);
typedef struct mgfb_tile_desc { uint64_t opaque[112]; } mgfb_tile_desc;   /* 896 bytes, plain data: send it to the other ranks */
/* Freezes the body set, reserves room for `ghost_capacity` ghosts and describes this tile's device
 * memory (pointers for contexts of the same process, cudaIpc handles for other processes). */
int32_t mgfb_tile_export(mgfb_ctx* ctx, uint32_t ghost_capacity, mgfb_tile_desc* out);
/* descs[0..nranks) in tile order along x; maps the neighbours' memory (rank-1, rank+1). */
int32_t mgfb_tile_connect(mgfb_ctx* ctx, uint32_t rank, uint32_t nranks, const mgfb_tile_desc* descs);

/* The dataflow solver hands velocities from one constraint row to the next with single 32-byte stores (payload + tag in one
 * sector) instead of flag + fence; that such a store is never observed half-written is checked, not assumed: at
 * mgfb_ctx_create on local memory and at mgfb_tile_connect across each NVLink peer mapping (writers overwrite records while
 * readers on other SMs poll them).  A torn read makes the context fall back to MGFB_SCHEDULE_PHASES (tiled: MGFB_ERR_TILE
 * asking for that schedule).  This call runs `rounds` more rounds of the local test (0 = none) and returns the totals so
 * far: *torn must be 0; *observed counts distinct whole records the readers saw. */
int32_t mgfb_selftest_handover(mgfb_ctx* ctx, uint32_t rounds, uint32_t* torn, uint32_t* observed);

/* Device-resident staging used by multi-GPU drivers and benchmarks: raw device pointers to the
 * SoA body arrays (for NCCL send/recv issued by the host plumbing).  See DESIGN.md. */
typedef struct mgfb_device_view {
    void* x;          /* float4[n]: x.xyz, w unused */
    void* q;          /* float4[n]: s, x, y, z */
    void* vel;        /* 64-byte records: v[3], omega[3], inv_mass, inv_moment[9] */
    void* collider;   /* mgfb_shape-like device records, 48 bytes: see DESIGN.md */
    uint32_t n;
    void* stream;     /* cudaStream_t */
} mgfb_device_view;
int32_t mgfb_device_view_get(mgfb_ctx* ctx, mgfb_device_view* out);

#ifdef __cplusplus
}
#endif
#endif /* MGFB_H */
