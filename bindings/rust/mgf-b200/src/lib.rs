//! Safe wrappers that keep mgf's surface for the hot path:
//! `GpuWorld` = `RigidBodyVec` + the demo `World::step` (physics.rs:141-269, mgf_demo/world.rs:227-294),
//! `GpuSolver` = `Solver` (solver.rs:53-78), `contacts_batch` = `Contacts::contacts` over many pairs (collision.rs:471).
//! Written against include/mgfb.h 1:1; this image has no Rust toolchain, so the crate is NOT compiled or tested here --
//! the Python mirror (mgf_b200/api.py) is what the test-suite drives through the same C ABI.
pub mod sys;

use std::ffi::CStr;
use std::ptr;

/// The reference's panics on this path, as values ("Error behaviour").
#[derive(Debug)]
pub enum Error {
    InvalidArg(String),
    SingularInertia(String), // physics.rs:212 `.invert().unwrap()`
    Capacity(String),
    Cuda(String),            // includes "no CUDA device": there is no CPU fallback
    NanBounds(String),       // bounds.rs:125-127 `assert!(r >= 0)`
    State(String),
}

fn check(ctx: *const sys::mgfb_ctx, code: i32) -> Result<(), Error> {
    if code == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(sys::mgfb_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(match code {
        1 => Error::InvalidArg(msg),
        2 => Error::SingularInertia(msg),
        3 => Error::Capacity(msg),
        4 => Error::Cuda(msg),
        5 => Error::NanBounds(msg),
        _ => Error::State(msg),
    })
}

/// `Component` (compound.rs:33) flattened the way the library takes it.
pub fn sphere(c: [f32; 3], r: f32) -> sys::mgfb_shape {
    let mut s = sys::mgfb_shape { kind: 0, p: [0.0; 12], v: [0.0; 3] };
    s.p[0] = c[0]; s.p[1] = c[1]; s.p[2] = c[2]; s.p[3] = r;
    s
}
pub fn capsule(a: [f32; 3], d: [f32; 3], r: f32) -> sys::mgfb_shape {
    let mut s = sys::mgfb_shape { kind: 1, p: [0.0; 12], v: [0.0; 3] };
    s.p[..3].copy_from_slice(&a); s.p[3..6].copy_from_slice(&d); s.p[6] = r;
    s
}

/// State of all bodies, struct of arrays like `RigidBodyVec`'s pub fields.
#[derive(Default, Clone)]
pub struct BodyState { pub x: Vec<[f32; 3]>, pub q: Vec<[f32; 4]>, pub v: Vec<[f32; 3]>, pub omega: Vec<[f32; 3]> }

pub struct GpuWorld { ctx: *mut sys::mgfb_ctx, n: u32 }
// mgfb_ctx is not thread-safe: GpuWorld may move between threads but is not Sync.
unsafe impl Send for GpuWorld {}

impl GpuWorld {
    /// `RigidBodyVec::new()` on CUDA device `device`.
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut cfg = unsafe { std::mem::zeroed::<sys::mgfb_config>() };
        unsafe { sys::mgfb_config_default(&mut cfg) };
        cfg.device = device;
        let mut ctx = ptr::null_mut();
        check(ptr::null(), unsafe { sys::mgfb_ctx_create(&cfg, &mut ctx) })?;
        Ok(GpuWorld { ctx, n: 0 })
    }
    /// `RigidBodyVec::add_body(collider, mass, restitution, friction, world_force) -> RigidBodyRef::Dynamic(i)` (physics.rs:200).
    pub fn add_body(&mut self, collider: sys::mgfb_shape, mass: f32, restitution: f32, friction: f32, world_force: [f32; 3]) -> Result<usize, Error> {
        let mut first = 0u32;
        check(self.ctx, unsafe { sys::mgfb_bodies_add(self.ctx, 1, &collider, &mass, &restitution, &friction, world_force.as_ptr(), &mut first) })?;
        self.n += 1;
        Ok(first as usize)
    }
    /// Many bodies in one call (the per-call cost is a host-to-device copy).
    pub fn add_bodies(&mut self, colliders: &[sys::mgfb_shape], mass: &[f32], restitution: &[f32], friction: &[f32], world_force: &[[f32; 3]]) -> Result<usize, Error> {
        let n = colliders.len();
        assert!(mass.len() == n && restitution.len() == n && friction.len() == n && world_force.len() == n);
        let mut first = 0u32;
        check(self.ctx, unsafe {
            sys::mgfb_bodies_add(self.ctx, n as u32, colliders.as_ptr(), mass.as_ptr(), restitution.as_ptr(), friction.as_ptr(), world_force.as_ptr() as *const f32, &mut first)
        })?;
        self.n += n as u32;
        Ok(first as usize)
    }
    /// `Mesh::new` + `push_vert` / `push_face` + `set_pos` (mesh.rs:40-73): the static terrain of the demo world (world.rs:118-150).
    pub fn set_terrain(&mut self, verts: &[[f32; 3]], faces: &[[u32; 3]], pos: [f32; 3]) -> Result<(), Error> {
        check(self.ctx, unsafe { sys::mgfb_terrain_set(self.ctx, verts.as_ptr() as *const f32, verts.len() as u32, faces.as_ptr() as *const u32, faces.len() as u32, pos.as_ptr()) })
    }
    /// `RigidBodyVec::integrate(dt)` (physics.rs:222) and `complete_motion()` (:262) on their own.
    pub fn integrate(&mut self, dt: f32) -> Result<(), Error> { check(self.ctx, unsafe { sys::mgfb_integrate(self.ctx, dt) }) }
    pub fn complete_motion(&mut self) -> Result<(), Error> { check(self.ctx, unsafe { sys::mgfb_complete_motion(self.ctx) }) }
    /// `World::step(dt)` (world.rs:227-294) with the demo's 20 solver iterations (world.rs:293).
    pub fn step(&mut self, dt: f32) -> Result<sys::mgfb_step_stats, Error> { self.step_iters(dt, 20) }
    pub fn step_iters(&mut self, dt: f32, iters: u32) -> Result<sys::mgfb_step_stats, Error> {
        let mut st = sys::mgfb_step_stats::default();
        check(self.ctx, unsafe { sys::mgfb_step(self.ctx, dt, iters, &mut st) })?;
        Ok(st)
    }
    /// The pub fields `x`, `q` (+ `v`, `omega`) of `RigidBodyVec`, read back.
    pub fn state(&mut self) -> Result<BodyState, Error> {
        let n = self.n as usize;
        let mut s = BodyState { x: vec![[0.0; 3]; n], q: vec![[0.0; 4]; n], v: vec![[0.0; 3]; n], omega: vec![[0.0; 3]; n] };
        check(self.ctx, unsafe {
            sys::mgfb_bodies_get_state(self.ctx, 0, self.n, s.x.as_mut_ptr() as *mut f32, s.q.as_mut_ptr() as *mut f32, s.v.as_mut_ptr() as *mut f32, s.omega.as_mut_ptr() as *mut f32)
        })?;
        Ok(s)
    }
    /// `ConstrainedSet::set` for a range of bodies (physics.rs:304-314).
    pub fn set_velocity(&mut self, first: usize, v: &[[f32; 3]], omega: &[[f32; 3]]) -> Result<(), Error> {
        assert_eq!(v.len(), omega.len());
        check(self.ctx, unsafe { sys::mgfb_bodies_set_velocity(self.ctx, first as u32, v.len() as u32, v.as_ptr() as *const f32, omega.as_ptr() as *const f32) })
    }
    /// `ShapeA.contacts(&Moving(ShapeB, v), |c| ..)` for many pairs of one kind (collision.rs:521-1532); `counts[i]` contacts were
    /// delivered for pair i (0, 1 or 2), in callback order, at `out[2 * i ..]`.
    pub fn contacts_batch(&mut self, pair_kind: u32, recv: &[sys::mgfb_shape], arg: &[sys::mgfb_shape]) -> Result<(Vec<sys::mgfb_contact>, Vec<u32>), Error> {
        assert_eq!(recv.len(), arg.len());
        let n = recv.len();
        let mut out = vec![sys::mgfb_contact::default(); 2 * n];
        let mut counts = vec![0u32; n];
        check(self.ctx, unsafe { sys::mgfb_contacts_batch(self.ctx, pair_kind, recv.as_ptr(), arg.as_ptr(), n as u32, out.as_mut_ptr(), ptr::null_mut(), counts.as_mut_ptr()) })?;
        Ok((out, counts))
    }
    pub fn raw(&mut self) -> *mut sys::mgfb_ctx { self.ctx }
}
impl Drop for GpuWorld { fn drop(&mut self) { unsafe { sys::mgfb_ctx_destroy(self.ctx) } } }

/// `Solver` (solver.rs:53-78): `add_constraint` collects `(obj_a, obj_b, Manifold)`, `solve` runs
/// `ContactConstraint::new` + the sequential-impulse sweep on the device, in the order the constraints were added
/// (`MGFB_ORDER_AS_GIVEN`: bit-identical to the reference's loop) or colour-major (`MGFB_ORDER_COLOURED`).
#[derive(Default)]
pub struct GpuSolver {
    obj_a: Vec<i32>, obj_b: Vec<i32>, static_center: Vec<f32>, static_friction: Vec<f32>,
    normal: Vec<f32>, tangent: Vec<f32>, ncontacts: Vec<u32>, local_a: Vec<f32>, local_b: Vec<f32>,
}
/// `Manifold` (manifold.rs:104-118): up to four contacts sharing one normal and tangent basis.
pub struct Manifold { pub normal: [f32; 3], pub tangent: [[f32; 3]; 2], pub local_a: Vec<[f32; 3]>, pub local_b: Vec<[f32; 3]> }
/// `RigidBodyRef` (physics.rs:157-166).
pub enum BodyRef { Dynamic(usize), Static { center: [f32; 3], friction: f32 } }

impl GpuSolver {
    pub fn new() -> Self { Self::default() }
    pub fn add_constraint(&mut self, a: BodyRef, b: BodyRef, m: &Manifold) {
        assert!(!m.local_a.is_empty() && m.local_a.len() <= 4 && m.local_a.len() == m.local_b.len());
        let (mut center, mut fric) = ([0.0f32; 3], 0.0f32);
        let mut idx = |r: BodyRef| match r { BodyRef::Dynamic(i) => i as i32, BodyRef::Static { center: c, friction: f } => { center = c; fric = f; -1 } };
        let (ia, ib) = (idx(a), idx(b));
        self.obj_a.push(ia); self.obj_b.push(ib);
        self.static_center.extend_from_slice(&center); self.static_friction.push(fric);
        self.normal.extend_from_slice(&m.normal);
        self.tangent.extend_from_slice(&m.tangent[0]); self.tangent.extend_from_slice(&m.tangent[1]);
        self.ncontacts.push(m.local_a.len() as u32);
        for k in 0..4 {
            let (la, lb) = if k < m.local_a.len() { (m.local_a[k], m.local_b[k]) } else { ([0.0; 3], [0.0; 3]) };
            self.local_a.extend_from_slice(&la); self.local_b.extend_from_slice(&lb);
        }
    }
    /// `Solver::solve(&mut bodies, iters)` (solver.rs:72): velocities of `world`'s bodies are updated in place on the device.
    pub fn solve(&mut self, world: &mut GpuWorld, dt: f32, iters: u32, as_given: bool) -> Result<sys::mgfb_solve_stats, Error> {
        let m = sys::mgfb_manifolds {
            n: self.obj_a.len() as u32, obj_a: self.obj_a.as_ptr(), obj_b: self.obj_b.as_ptr(),
            static_center: self.static_center.as_ptr(), static_friction: self.static_friction.as_ptr(),
            normal: self.normal.as_ptr(), tangent: self.tangent.as_ptr(), ncontacts: self.ncontacts.as_ptr(),
            local_a: self.local_a.as_ptr(), local_b: self.local_b.as_ptr(),
        };
        let mut st = sys::mgfb_solve_stats::default();
        check(world.ctx, unsafe { sys::mgfb_solver_solve(world.ctx, &m, dt, iters, if as_given { 0 } else { 1 }, ptr::null_mut(), ptr::null_mut(), &mut st) })?;
        Ok(st)
    }
}
