//! Raw binding of include/mgfb.h (libmgfb.so).  Kept identical to the block in INTEGRATION.md section 1 by
//! tests/test_abi.py::test_rust_binding_matches_integration_md; not compiled in this repository's CI (no Rust toolchain in the image).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)] #[derive(Copy, Clone)]
pub struct mgfb_shape { pub kind: u32, pub p: [f32; 12], pub v: [f32; 3] }          // 64 bytes
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct mgfb_contact { pub a: [f32; 3], pub b: [f32; 3], pub n: [f32; 3], pub t: f32 }
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct mgfb_local_contact { pub local_a: [f32; 3], pub local_b: [f32; 3], pub global: mgfb_contact }
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct mgfb_intersection { pub p: [f32; 3], pub t: f32 }                          // collision.rs:151
#[repr(C)] #[derive(Copy, Clone)]
pub struct mgfb_config {
    pub device: i32, pub penetration_slop: f32, pub baumgarte: f32,
    pub persistent_threshold_sq: f32, pub fat_margin: f32,
    pub initial_body_capacity: u32, pub max_cooperative_ctas: u32, pub tile_timeout_ms: u32,
    pub solver_schedule: u32,   // 0 = dataflow (default), 1 = barrier phases, 2 = phases + Jones-Plassmann colouring
    pub step_order: u32,        // 0 = coloured (default, exported for replay), 1 = the reference's own order (bit-identical World::step)
}
#[repr(C)]
pub struct mgfb_manifolds {
    pub n: u32,
    pub obj_a: *const i32, pub obj_b: *const i32,
    pub static_center: *const f32, pub static_friction: *const f32,
    pub normal: *const f32, pub tangent: *const f32, pub ncontacts: *const u32,
    pub local_a: *const f32, pub local_b: *const f32,
}
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct mgfb_step_stats {
    pub bodies: u32, pub candidate_pairs: u32, pub terrain_candidates: u32, pub constraints: u32,
    pub terrain_constraints: u32, pub groups: u32, pub iterations: u32, pub fat_refreshes: u32,
    pub step_ms: f32, pub solve_ms: f32, pub overflow: u32, pub colouring_rounds: u32,
    pub ghosts: u32, pub boundary_constraints: u32, pub phases: u32, pub broadphase_path: u32,
}
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct mgfb_phase_profile { pub phase_ms: [f32; 8], pub pairs: [u32; 4], pub terrain_pairs: [u32; 2], pub body_contacts: u32, pub terrain_contacts: u32 }
#[repr(C)] #[derive(Copy, Clone)]
pub struct mgfb_tile_desc { pub opaque: [u64; 112] }
#[repr(C)]
pub struct mgfb_device_view { pub x: *mut c_void, pub q: *mut c_void, pub vel: *mut c_void, pub collider: *mut c_void, pub n: u32, pub stream: *mut c_void }
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct mgfb_solve_stats { pub constraints: u32, pub contacts: u32, pub groups: u32, pub iterations: u32, pub solve_ms: f32, pub reserved: [u32; 3] }
pub enum mgfb_ctx {}
pub enum mgfb_bvh {}
pub enum mgfb_compound {}

#[link(name = "mgfb")]
extern "C" {
    pub fn mgfb_config_default(cfg: *mut mgfb_config);
    pub fn mgfb_ctx_create(cfg: *const mgfb_config, out: *mut *mut mgfb_ctx) -> i32;
    pub fn mgfb_ctx_destroy(ctx: *mut mgfb_ctx);
    pub fn mgfb_last_error(ctx: *const mgfb_ctx) -> *const c_char;
    pub fn mgfb_bodies_add(ctx: *mut mgfb_ctx, n: u32, shapes: *const mgfb_shape, mass: *const f32,
                           restitution: *const f32, friction: *const f32, world_force: *const f32, first_id: *mut u32) -> i32;
    pub fn mgfb_bodies_count(ctx: *const mgfb_ctx, n: *mut u32) -> i32;
    pub fn mgfb_bodies_get_state(ctx: *mut mgfb_ctx, first: u32, n: u32, x: *mut f32, q: *mut f32, v: *mut f32, omega: *mut f32) -> i32;
    pub fn mgfb_bodies_set_velocity(ctx: *mut mgfb_ctx, first: u32, n: u32, v: *const f32, omega: *const f32) -> i32;
    pub fn mgfb_bodies_get_colliders(ctx: *mut mgfb_ctx, first: u32, n: u32, out: *mut mgfb_shape) -> i32;
    pub fn mgfb_integrate(ctx: *mut mgfb_ctx, dt: f32) -> i32;
    pub fn mgfb_complete_motion(ctx: *mut mgfb_ctx) -> i32;
    pub fn mgfb_terrain_set(ctx: *mut mgfb_ctx, verts: *const f32, nverts: u32, faces: *const u32, nfaces: u32, x: *const f32) -> i32;
    pub fn mgfb_contacts_batch(ctx: *mut mgfb_ctx, pair_kind: u32, recv: *const mgfb_shape, arg: *const mgfb_shape, n: u32,
                               out: *mut mgfb_contact, out_local: *mut mgfb_local_contact, counts: *mut u32) -> i32;
    pub fn mgfb_solver_solve(ctx: *mut mgfb_ctx, m: *const mgfb_manifolds, dt: f32, iters: u32, order: u32,
                             perm_out: *mut u32, normal_impulse_out: *mut f32, stats: *mut mgfb_solve_stats) -> i32;
    pub fn mgfb_intersections_batch(ctx: *mut mgfb_ctx, particle_kind: u32, particles: *const f32, shapes: *const mgfb_shape, n: u32,
                                    out: *mut mgfb_intersection, hit: *mut u32) -> i32;
    pub fn mgfb_bvh_create(ctx: *mut mgfb_ctx, out: *mut *mut mgfb_bvh) -> i32;
    pub fn mgfb_bvh_destroy(bvh: *mut mgfb_bvh);
    pub fn mgfb_bvh_insert(bvh: *mut mgfb_bvh, boxes: *const f32, values: *const u32, n: u32, indices: *mut u32) -> i32;
    pub fn mgfb_bvh_remove(bvh: *mut mgfb_bvh, indices: *const u32, n: u32) -> i32;
    pub fn mgfb_bvh_get(bvh: *const mgfb_bvh, index: u32, box6: *mut f32, value: *mut u32) -> i32;
    pub fn mgfb_bvh_len(bvh: *const mgfb_bvh, n: *mut u32) -> i32;
    pub fn mgfb_bvh_query_batch(bvh: *mut mgfb_bvh, boxes: *const f32, nq: u32, offsets: *mut u32, values: *mut u32, capacity: u32, total: *mut u32) -> i32;
    pub fn mgfb_bvh_raytrace_batch(bvh: *mut mgfb_bvh, particle_kind: u32, particles: *const f32, nq: u32, offsets: *mut u32, values: *mut u32,
                                   hits: *mut mgfb_intersection, capacity: u32, total: *mut u32) -> i32;
    pub fn mgfb_step(ctx: *mut mgfb_ctx, dt: f32, iters: u32, stats: *mut mgfb_step_stats) -> i32;
    pub fn mgfb_step_n(ctx: *mut mgfb_ctx, dt: f32, iters: u32, nsteps: u32, stats: *mut mgfb_step_stats) -> i32;
    pub fn mgfb_step_enqueue(ctx: *mut mgfb_ctx, dt: f32, iters: u32, input_mode: u32, v_in: *const f32, omega_in: *const f32,
                             x_out: *mut f32, q_out: *mut f32, v_out: *mut f32, omega_out: *mut f32) -> i32;   // page-locked buffers
    pub fn mgfb_step_wait(ctx: *mut mgfb_ctx, stats: *mut mgfb_step_stats) -> i32;
    pub fn mgfb_step_constraints(ctx: *mut mgfb_ctx, capacity: u32, body_a: *mut u32, body_b: *mut i32,
                                 face: *mut u32, sub: *mut u32, colour: *mut u32, count: *mut u32) -> i32;
    pub fn mgfb_step_profile(ctx: *mut mgfb_ctx, dt: f32, iters: u32, stats: *mut mgfb_step_stats, out: *mut mgfb_phase_profile) -> i32;
    pub fn mgfb_step_totals(ctx: *mut mgfb_ctx, steps: *mut u64, constraints: *mut u64, candidate_pairs: *mut u64, groups: *mut u64,
                            kernel_launches: *mut u64, reset: i32) -> i32;
    pub fn mgfb_abi_version() -> i32;
    pub fn mgfb_synchronize(ctx: *mut mgfb_ctx) -> i32;
    // checkpoint: the pub fields + velocities + the body BVH's stored fat boxes, out and back in, bit for bit
    pub fn mgfb_bodies_get_inv_moment(ctx: *mut mgfb_ctx, first: u32, n: u32, out: *mut f32) -> i32;
    pub fn mgfb_bodies_get_fat_bounds(ctx: *mut mgfb_ctx, first: u32, n: u32, boxes: *mut f32) -> i32;
    pub fn mgfb_bodies_set_state(ctx: *mut mgfb_ctx, first: u32, n: u32, x: *const f32, q: *const f32, v: *const f32, omega: *const f32,
                                 colliders: *const mgfb_shape, fat_boxes: *const f32) -> i32;
    // ContactPruner + Manifold::from (manifold.rs:42-148), batched
    pub fn mgfb_manifolds_prune(ctx: *mut mgfb_ctx, contacts: *const mgfb_local_contact, offsets: *const u32, ngroups: u32, time: *mut f32,
                                normal: *mut f32, tangent: *mut f32, ncontacts: *mut u32, local_a: *mut f32, local_b: *mut f32) -> i32;
    // discrete path: GJK + EPA (collision.rs:404-425, 497-519), ConvexMesh vertices (mesh.rs:141)
    pub fn mgfb_convex_vertices_set(ctx: *mut mgfb_ctx, verts: *const f32, n: u32) -> i32;
    pub fn mgfb_gjk_batch(ctx: *mut mgfb_ctx, a: *const mgfb_shape, b: *const mgfb_shape, n: u32, out: *mut mgfb_contact,
                          status: *mut u32, epa_iters: *mut u32) -> i32;
    pub fn mgfb_separation_batch(ctx: *mut mgfb_ctx, a: *const mgfb_shape, b: *const mgfb_shape, n: u32, separation: *mut f32, status: *mut u32) -> i32;
    // Compound (compound.rs:232-352)
    pub fn mgfb_compound_create(ctx: *mut mgfb_ctx, components: *const mgfb_shape, n: u32, out: *mut *mut mgfb_compound) -> i32;
    pub fn mgfb_compound_destroy(c: *mut mgfb_compound);
    pub fn mgfb_compound_set_transform(c: *mut mgfb_compound, disp: *const f32, rot: *const f32) -> i32;
    pub fn mgfb_compound_bounds(c: *const mgfb_compound, aabb: *mut f32, sphere: *mut f32) -> i32;
    pub fn mgfb_compound_closest_points(c: *mut mgfb_compound, to: *const f32, n: u32, out: *mut f32) -> i32;
    pub fn mgfb_compound_intersections_batch(c: *mut mgfb_compound, particle_kind: u32, particles: *const f32, n: u32,
                                             out: *mut mgfb_intersection, hit: *mut u32) -> i32;
    pub fn mgfb_compound_contacts_batch(c: *mut mgfb_compound, rhs: *const mgfb_shape, n: u32, slots: u32, out: *mut mgfb_contact, counts: *mut u32) -> i32;
    // one world tiled across GPUs (no counterpart in mgf)
    pub fn mgfb_bodies_set_gid(ctx: *mut mgfb_ctx, first: u32, n: u32, gids: *const u32) -> i32;
    pub fn mgfb_tile_export(ctx: *mut mgfb_ctx, ghost_capacity: u32, out: *mut mgfb_tile_desc) -> i32;
    pub fn mgfb_tile_connect(ctx: *mut mgfb_ctx, rank: u32, nranks: u32, descs: *const mgfb_tile_desc) -> i32;
    pub fn mgfb_selftest_handover(ctx: *mut mgfb_ctx, rounds: u32, torn: *mut u32, observed: *mut u32) -> i32;
    pub fn mgfb_device_view_get(ctx: *mut mgfb_ctx, out: *mut mgfb_device_view) -> i32;
}
