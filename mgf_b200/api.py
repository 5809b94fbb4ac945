"""Host-side mirror of mgf's API for the step path, over the C ABI (include/mgfb.h).

Names follow the reference: ``World`` is mgf_demo/world.rs ``World`` (bodies: RigidBodyVec,
terrain: Mesh, ``step``); shape constructors mirror src/geom.rs.  All arithmetic of the path
runs in the CUDA library; this module only marshals numpy arrays.
"""
import ctypes as C

import numpy as np

from . import _lib as L


class MgfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mgfb error {code}: {msg}")
        self.code = code


# ---------------------------------------------------------------- shapes (geom.rs)
def make_shapes(n):
    return np.zeros(n, dtype=L.SHAPE_DTYPE)


def _one(kind, p, v=(0.0, 0.0, 0.0)):
    s = make_shapes(1)
    s["kind"] = kind
    s["p"][0, :len(p)] = np.asarray(p, dtype=np.float32)
    s["v"][0] = np.asarray(v, dtype=np.float32)
    return s


def sphere(c, r, v=(0, 0, 0)):
    """geom.rs:290 Sphere{c, r} (optionally Moving, geom.rs:357)."""
    return _one(L.SPHERE, [*c, r], v)


def capsule(a, d, r, v=(0, 0, 0)):
    """geom.rs:316 Capsule{a, d, r}."""
    return _one(L.CAPSULE, [*a, *d, r], v)


def triangle(a, b, c, v=(0, 0, 0)):
    """geom.rs:128 Triangle{a, b, c}."""
    return _one(L.TRIANGLE, [*a, *b, *c], v)


def rectangle(c, u0, u1, e0, e1):
    """geom.rs:216 Rectangle{c, u, e} (u must be normalised, as in the reference)."""
    return _one(L.RECTANGLE, [*c, *u0, *u1, e0, e1])


def plane(n, d):
    """geom.rs:32 Plane{n, d}."""
    return _one(L.PLANE, [*n, d])


class Context:
    """One mgfb_ctx: one CUDA device, one stream, all device memory."""

    def __init__(self, device=0, **cfg):
        self.lib = L.load()
        c = L.Config()
        self.lib.mgfb_config_default(C.byref(c))
        c.device = device
        for k, v in cfg.items():
            setattr(c, k, v)
        h = C.c_void_p()
        st = self.lib.mgfb_ctx_create(C.byref(c), C.byref(h))
        if st != L.OK:
            raise MgfbError(st, (self.lib.mgfb_last_error(None) or b"").decode())
        self.h = h
        self.cfg = c

    def selftest_handover(self, rounds=0):
        """(torn, observed) 32-byte hand-overs seen by the self-tests so far (+ `rounds` more local rounds); torn must be 0."""
        t = C.c_uint32(); o = C.c_uint32()
        self.check(self.lib.mgfb_selftest_handover(self.h, rounds, C.byref(t), C.byref(o)))
        return t.value, o.value

    def check(self, st):
        if st != L.OK:
            raise MgfbError(st, (self.lib.mgfb_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.mgfb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def aabb(c, r):
    """geom.rs:257 AABB{c, r}."""
    return _one(L.AABB, [*c, *r])


def obb(c, r, q=(1.0, 0.0, 0.0, 0.0)):
    """geom.rs:272 OBB{c, q, r}; q = (s, x, y, z)."""
    return _one(L.OBB, [*c, *r, *q])


def convex_mesh(first, count):
    """mesh.rs:141 ConvexMesh as a slice [first, first + count) of the context's vertex pool (convex_vertices_set)."""
    return _one(L.CONVEX_MESH, [float(first), float(count)])


def convex_vertices_set(ctx, verts):
    """The vertex pool CONVEX_MESH shapes index (mgfb_convex_vertices_set)."""
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
    ctx.check(ctx.lib.mgfb_convex_vertices_set(ctx.h, L.ptr(verts), len(verts)))


def gjk_batch(ctx, a, b):
    """`a[i].contacts(&b[i], cb)` through the discrete GJK + EPA path (collision.rs:497-519).
    Returns (contacts[n], status[n], epa_iterations[n]); status 1 = a contact was delivered."""
    a = np.ascontiguousarray(a, dtype=L.SHAPE_DTYPE); b = np.ascontiguousarray(b, dtype=L.SHAPE_DTYPE)
    n = len(a)
    assert len(b) == n
    out = np.zeros(n, dtype=L.CONTACT_DTYPE); status = np.zeros(n, np.uint32); iters = np.zeros(n, np.uint32)
    ctx.check(ctx.lib.mgfb_gjk_batch(ctx.h, L.ptr(a), L.ptr(b), n, L.ptr(out), L.ptr(status), L.ptr(iters)))
    return out, status, iters


def separation_batch(ctx, a, b):
    """Penetrates::separation (collision.rs:404-425): (distance[n], is_some[n])."""
    a = np.ascontiguousarray(a, dtype=L.SHAPE_DTYPE); b = np.ascontiguousarray(b, dtype=L.SHAPE_DTYPE)
    n = len(a)
    assert len(b) == n
    sep = np.zeros(n, np.float32); status = np.zeros(n, np.uint32)
    ctx.check(ctx.lib.mgfb_separation_batch(ctx.h, L.ptr(a), L.ptr(b), n, L.ptr(sep), L.ptr(status)))
    return sep, status


def intersections_batch(ctx, particle_kind, particles, shapes):
    """`particles[i].intersection(&shapes[i])` (collision.rs:163-373).  particles: (n, 6) f32, Ray = (p, d),
    Segment = (a, b).  Returns (intersections[n] {p, t}, hit[n])."""
    particles = np.ascontiguousarray(particles, dtype=np.float32).reshape(-1, 6)
    shapes = np.ascontiguousarray(shapes, dtype=L.SHAPE_DTYPE)
    n = len(shapes)
    assert len(particles) == n
    out = np.zeros(n, dtype=L.INTERSECTION_DTYPE); hit = np.zeros(n, np.uint32)
    ctx.check(ctx.lib.mgfb_intersections_batch(ctx.h, particle_kind, L.ptr(particles), L.ptr(shapes), n, L.ptr(out), L.ptr(hit)))
    return out, hit


def manifolds_prune(ctx, contacts, offsets):
    """ContactPruner::push over every group's LocalContacts in order + Manifold::from (manifold.rs:42-148).
    Returns a dict laid out like mgfb_manifolds: time, normal, tangent, ncontacts, local_a, local_b."""
    contacts = np.ascontiguousarray(contacts, dtype=L.LOCAL_CONTACT_DTYPE)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    g = len(offsets) - 1
    d = dict(time=np.zeros(g, np.float32), normal=np.zeros((g, 3), np.float32), tangent=np.zeros((g, 6), np.float32), ncontacts=np.zeros(g, np.uint32),
             local_a=np.zeros((g, 12), np.float32), local_b=np.zeros((g, 12), np.float32))
    ctx.check(ctx.lib.mgfb_manifolds_prune(ctx.h, L.ptr(contacts), L.ptr(offsets), g, *[L.ptr(d[k]) for k in ("time", "normal", "tangent", "ncontacts", "local_a", "local_b")]))
    return d


def contacts_batch(ctx, pair_kind, recv, arg, want_local=False):
    """`recv[i].contacts(&arg[i], cb)` for a homogeneous batch (collision.rs:471).

    Returns (contacts[n,2], counts[n]) and additionally local contacts[n,2] when want_local."""
    recv = np.ascontiguousarray(recv, dtype=L.SHAPE_DTYPE)
    arg = np.ascontiguousarray(arg, dtype=L.SHAPE_DTYPE)
    n = len(recv)
    assert len(arg) == n
    out = np.zeros((n, 2), dtype=L.CONTACT_DTYPE)
    loc = np.zeros((n, 2), dtype=L.LOCAL_CONTACT_DTYPE) if want_local else None
    counts = np.zeros(n, dtype=np.uint32)
    ctx.check(ctx.lib.mgfb_contacts_batch(ctx.h, pair_kind, L.ptr(recv), L.ptr(arg), n, L.ptr(out), L.ptr(loc), L.ptr(counts)))
    return (out, counts, loc) if want_local else (out, counts)


class BVH:
    """src/bvh.rs BVH<AABB, u32>: insert / remove / get / len / query / raytrace, batched (include/mgfb.h).
    Boxes are (n, 6) f32 rows (centre, half extents)."""

    def __init__(self, ctx):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(ctx.lib.mgfb_bvh_create(ctx.h, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.mgfb_bvh_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001
            pass

    def __len__(self):
        n = C.c_uint32()
        self.ctx.check(self.ctx.lib.mgfb_bvh_len(self.h, C.byref(n)))
        return n.value

    def insert(self, boxes, values):
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 6)
        values = np.ascontiguousarray(values, dtype=np.uint32).reshape(-1)
        assert len(values) == len(boxes)
        idx = np.zeros(len(boxes), np.uint32)
        self.ctx.check(self.ctx.lib.mgfb_bvh_insert(self.h, L.ptr(boxes), L.ptr(values), len(boxes), L.ptr(idx)))
        return idx

    def remove(self, indices):
        indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        self.ctx.check(self.ctx.lib.mgfb_bvh_remove(self.h, L.ptr(indices), len(indices)))

    def get(self, index):
        box = np.zeros(6, np.float32); val = C.c_uint32()
        self.ctx.check(self.ctx.lib.mgfb_bvh_get(self.h, int(index), L.ptr(box), C.byref(val)))
        return box, val.value

    def _batch(self, call, nq, want_hits):
        offsets = np.zeros(nq + 1, np.uint32); total = C.c_uint32()
        cap = max(1024, 16 * nq)
        for _ in range(2):
            values = np.zeros(cap, np.uint32)
            hits = np.zeros(cap, dtype=L.INTERSECTION_DTYPE) if want_hits else None
            st = call(offsets, values, hits, cap, total)
            if st == L.ERR_CAPACITY and total.value > cap:
                cap = total.value
                continue
            self.ctx.check(st)
            break
        n = total.value
        return (offsets, values[:n], hits[:n]) if want_hits else (offsets, values[:n])

    def query(self, boxes):
        """BVH::query for every row of `boxes`: (offsets[nq+1], values) -- CSR of the overlapping leaves' values."""
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 6)
        lib = self.ctx.lib
        return self._batch(lambda off, val, hits, cap, tot: lib.mgfb_bvh_query_batch(self.h, L.ptr(boxes), len(boxes), L.ptr(off), L.ptr(val), cap,
                                                                                       C.byref(tot)), len(boxes), False)

    def raytrace(self, particle_kind, particles):
        """BVH::raytrace for every Ray / Segment: (offsets[nq+1], values, intersections)."""
        particles = np.ascontiguousarray(particles, dtype=np.float32).reshape(-1, 6)
        lib = self.ctx.lib
        return self._batch(lambda off, val, hits, cap, tot: lib.mgfb_bvh_raytrace_batch(self.h, particle_kind, L.ptr(particles), len(particles), L.ptr(off),
                                                                                          L.ptr(val), L.ptr(hits), cap, C.byref(tot)), len(particles), True)


class Compound:
    """src/compound.rs Compound: components (spheres / capsules in the compound's frame) + disp + rot, queried in batches."""

    def __init__(self, ctx, components):
        self.ctx = ctx
        comps = np.ascontiguousarray(components, dtype=L.SHAPE_DTYPE)
        self.n = len(comps)
        h = C.c_void_p()
        ctx.check(ctx.lib.mgfb_compound_create(ctx.h, L.ptr(comps), len(comps), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.mgfb_compound_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_transform(self, disp, rot):
        d = np.ascontiguousarray(disp, dtype=np.float32); r = np.ascontiguousarray(rot, dtype=np.float32)
        assert d.shape == (3,) and r.shape == (4,)
        self.ctx.check(self.ctx.lib.mgfb_compound_set_transform(self.h, L.ptr(d), L.ptr(r)))

    def bounds(self):
        """(AABB centre | half extents, bounding sphere centre | radius)."""
        a = np.zeros(6, np.float32); s = np.zeros(4, np.float32)
        self.ctx.check(self.ctx.lib.mgfb_compound_bounds(self.h, L.ptr(a), L.ptr(s)))
        return a, s

    def closest_points(self, to):
        to = np.ascontiguousarray(to, dtype=np.float32).reshape(-1, 3)
        out = np.zeros_like(to)
        self.ctx.check(self.ctx.lib.mgfb_compound_closest_points(self.h, L.ptr(to), len(to), L.ptr(out)))
        return out

    def intersections(self, particle_kind, particles):
        particles = np.ascontiguousarray(particles, dtype=np.float32).reshape(-1, 6)
        out = np.zeros(len(particles), dtype=L.INTERSECTION_DTYPE); hit = np.zeros(len(particles), np.uint32)
        self.ctx.check(self.ctx.lib.mgfb_compound_intersections_batch(self.h, particle_kind, L.ptr(particles), len(particles), L.ptr(out), L.ptr(hit)))
        return out, hit

    def contacts(self, rhs, slots=None):
        """compound.contacts(&rhs[i]): (contacts[n, slots] in callback order, counts[n])."""
        rhs = np.ascontiguousarray(rhs, dtype=L.SHAPE_DTYPE)
        slots = slots or max(2 * self.n, 1)
        out = np.zeros((len(rhs), slots), dtype=L.CONTACT_DTYPE); counts = np.zeros(len(rhs), np.uint32)
        self.ctx.check(self.ctx.lib.mgfb_compound_contacts_batch(self.h, L.ptr(rhs), len(rhs), slots, L.ptr(out), L.ptr(counts)))
        return out, counts


class World:
    """mgf_demo/world.rs World restricted to the physics: bodies + terrain + step."""

    def __init__(self, ctx=None, device=0, **cfg):
        self.ctx = ctx or Context(device=device, **cfg)
        self.lib = self.ctx.lib

    # -- RigidBodyVec::add_body (physics.rs:200) + World::add_body (world.rs:178)
    def add_bodies(self, shapes, mass, restitution, friction, world_force):
        shapes = np.ascontiguousarray(shapes, dtype=L.SHAPE_DTYPE)
        n = len(shapes)
        f32 = lambda a, shape: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float32), shape))
        mass, restitution, friction = f32(mass, (n,)), f32(restitution, (n,)), f32(friction, (n,))
        world_force = f32(world_force, (n, 3))
        first = C.c_uint32()
        self.ctx.check(self.lib.mgfb_bodies_add(self.ctx.h, n, L.ptr(shapes), L.ptr(mass), L.ptr(restitution), L.ptr(friction),
                                                L.ptr(world_force), C.byref(first)))
        return first.value

    def __len__(self):
        n = C.c_uint32()
        self.ctx.check(self.lib.mgfb_bodies_count(self.ctx.h, C.byref(n)))
        return n.value

    # -- Mesh (mesh.rs:40-73) as the world's terrain
    def set_terrain(self, verts, faces, x=(0.0, 0.0, 0.0)):
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        faces = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        x = np.asarray(x, dtype=np.float32)
        self.ctx.check(self.lib.mgfb_terrain_set(self.ctx.h, L.ptr(verts), len(verts), L.ptr(faces), len(faces), L.ptr(x)))

    def state(self, first=0, n=None):
        n = len(self) - first if n is None else n
        x = np.zeros((n, 3), np.float32); q = np.zeros((n, 4), np.float32)
        v = np.zeros((n, 3), np.float32); w = np.zeros((n, 3), np.float32)
        self.ctx.check(self.lib.mgfb_bodies_get_state(self.ctx.h, first, n, L.ptr(x), L.ptr(q), L.ptr(v), L.ptr(w)))
        return x, q, v, w

    def set_velocity(self, first, v, omega):
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 3)
        omega = np.ascontiguousarray(omega, dtype=np.float32).reshape(-1, 3)
        self.ctx.check(self.lib.mgfb_bodies_set_velocity(self.ctx.h, first, len(v), L.ptr(v), L.ptr(omega)))

    def colliders(self, first=0, n=None):
        n = len(self) - first if n is None else n
        out = np.zeros(n, dtype=L.SHAPE_DTYPE)
        self.ctx.check(self.lib.mgfb_bodies_get_colliders(self.ctx.h, first, n, L.ptr(out)))
        return out

    def inv_moment(self, first=0, n=None):
        n = len(self) - first if n is None else n
        out = np.zeros((n, 9), np.float32)
        self.ctx.check(self.lib.mgfb_bodies_get_inv_moment(self.ctx.h, first, n, L.ptr(out)))
        return out

    def fat_bounds(self, first=0, n=None):
        """Stored fat AABBs of the body BVH (world.rs:180, 235-238): (n, 6) = centre, half extents."""
        n = len(self) - first if n is None else n
        out = np.zeros((n, 6), np.float32)
        self.ctx.check(self.lib.mgfb_bodies_get_fat_bounds(self.ctx.h, first, n, L.ptr(out)))
        return out

    def set_state(self, first=0, x=None, q=None, v=None, omega=None, colliders=None, fat=None):
        """Write the pub fields x, q, collider, the velocities and the stored fat boxes (mgfb_bodies_set_state)."""
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (x, q, v, omega)]
        col = None if colliders is None else np.ascontiguousarray(colliders, dtype=L.SHAPE_DTYPE)
        fb = None if fat is None else np.ascontiguousarray(fat, dtype=np.float32)
        ns = {len(a) for a in (*arrs, col, fb) if a is not None}
        assert len(ns) == 1, "all given arrays must describe the same bodies"
        self.ctx.check(self.lib.mgfb_bodies_set_state(self.ctx.h, first, ns.pop(), *[L.ptr(a) for a in arrs], L.ptr(col), L.ptr(fb)))

    def snapshot(self):
        """Everything World::step carries from one step to the next: x, q, v, omega, collider (incl. the pending
        displacement complete_motion adds), stored fat boxes.  restore() of it continues bit-identically."""
        x, q, v, w = self.state()
        return dict(x=x, q=q, v=v, omega=w, colliders=self.colliders(), fat=self.fat_bounds())

    def restore(self, snap):
        self.set_state(0, snap["x"], snap["q"], snap["v"], snap["omega"], snap["colliders"], snap["fat"])

    def integrate(self, dt):
        self.ctx.check(self.lib.mgfb_integrate(self.ctx.h, dt))

    def complete_motion(self):
        self.ctx.check(self.lib.mgfb_complete_motion(self.ctx.h))

    # -- World::step (world.rs:227-294)
    def step(self, dt, iters=20, nsteps=1):
        st = L.StepStats()
        self.ctx.check(self.lib.mgfb_step_n(self.ctx.h, dt, iters, nsteps, C.byref(st)))
        return st.as_dict()

    def step_profile(self, dt, iters=20):
        """One step with events between its phases: (stats, {phase name: device ms, + per-kind pair / contact counts})."""
        st = L.StepStats(); pr = L.PhaseProfile()
        self.ctx.check(self.lib.mgfb_step_profile(self.ctx.h, dt, iters, C.byref(st), C.byref(pr)))
        d = dict(zip(L.PHASES, list(pr.phase_ms)))
        d.update(pairs=list(pr.pairs), terrain_pairs=list(pr.terrain_pairs), body_contacts=pr.body_contacts, terrain_contacts=pr.terrain_contacts)
        return st.as_dict(), d

    # -- pipelined step: transfers of step k overlap the kernels of step k+1 (mgfb_step_enqueue / mgfb_step_wait)
    def step_enqueue(self, dt, iters=20, v_in=None, omega_in=None, x_out=None, q_out=None, v_out=None, omega_out=None, add=False):
        """Queue [v_in, omega_in -> device] -> step -> [state -> host] and return at once.  All arrays must be
        C-contiguous float32 in PAGE-LOCKED memory and stay untouched until the matching step_wait().  add=True: the
        inputs are ADDED to the velocities (external impulses) instead of replacing them."""
        self.ctx.check(self.lib.mgfb_step_enqueue(self.ctx.h, dt, iters, L.INPUT_ADD if add else L.INPUT_SET, L.ptr(v_in), L.ptr(omega_in), L.ptr(x_out), L.ptr(q_out),
                                                  L.ptr(v_out), L.ptr(omega_out)))

    def step_wait(self):
        """Block until the oldest queued step's outputs are in the host buffers; returns its stats."""
        st = L.StepStats()
        self.ctx.check(self.lib.mgfb_step_wait(self.ctx.h, C.byref(st)))
        return st.as_dict()

    # -- one world tiled across GPUs (mgfb.h "one world tiled across GPUs"; driver in tiling.py)
    def set_gid(self, gids, first=0):
        gids = np.ascontiguousarray(gids, dtype=np.uint32)
        self.ctx.check(self.lib.mgfb_bodies_set_gid(self.ctx.h, first, len(gids), L.ptr(gids)))

    def tile_export(self, ghost_capacity):
        """Freeze the body set and describe this tile's device memory; returns the descriptor bytes."""
        d = L.TileDesc()
        self.ctx.check(self.lib.mgfb_tile_export(self.ctx.h, ghost_capacity, C.byref(d)))
        return bytes(d)

    def tile_connect(self, rank, descs):
        """descs: the descriptor bytes of every tile, in tile order along x."""
        arr = (L.TileDesc * len(descs))()
        for k, b in enumerate(descs):
            C.memmove(C.byref(arr[k]), b, C.sizeof(L.TileDesc))
        self.ctx.check(self.lib.mgfb_tile_connect(self.ctx.h, rank, len(descs), arr))

    def totals(self, reset=False):
        """Running totals since the last reset (see mgfb_step_totals)."""
        v = [C.c_uint64() for _ in range(5)]
        self.ctx.check(self.lib.mgfb_step_totals(self.ctx.h, *[C.byref(x) for x in v], 1 if reset else 0))
        return dict(zip(("steps", "constraints", "candidate_pairs", "groups", "kernel_launches"), [x.value for x in v]))

    def constraints(self):
        """Identity of the last step's constraints in solve order."""
        cnt = C.c_uint32()
        st = self.lib.mgfb_step_constraints(self.ctx.h, 0, None, None, None, None, None, C.byref(cnt))
        m = cnt.value
        a = np.zeros(m, np.uint32); b = np.zeros(m, np.int32); face = np.zeros(m, np.uint32)
        sub = np.zeros(m, np.uint32); colour = np.zeros(m, np.uint32)
        if m:
            self.ctx.check(self.lib.mgfb_step_constraints(self.ctx.h, m, L.ptr(a), L.ptr(b), L.ptr(face), L.ptr(sub), L.ptr(colour),
                                                          C.byref(cnt)))
        elif st not in (L.OK, L.ERR_CAPACITY):
            self.ctx.check(st)
        return a, b, face, sub, colour

    # -- Solver::solve over caller-built manifolds (solver.rs:53-79, 101)
    def solve_manifolds(self, obj_a, obj_b, normal, tangent, ncontacts, local_a, local_b, dt, iters, order=L.ORDER_AS_GIVEN,
                        static_center=None, static_friction=None):
        n = len(obj_a)
        arrs = dict(
            obj_a=np.ascontiguousarray(obj_a, np.int32), obj_b=np.ascontiguousarray(obj_b, np.int32),
            normal=np.ascontiguousarray(normal, np.float32).reshape(n, 3), tangent=np.ascontiguousarray(tangent, np.float32).reshape(n, 6),
            ncontacts=np.ascontiguousarray(ncontacts, np.uint32), local_a=np.ascontiguousarray(local_a, np.float32).reshape(n, 12),
            local_b=np.ascontiguousarray(local_b, np.float32).reshape(n, 12),
            static_center=None if static_center is None else np.ascontiguousarray(static_center, np.float32).reshape(n, 3),
            static_friction=None if static_friction is None else np.ascontiguousarray(static_friction, np.float32))
        m = L.Manifolds(n=n, **{k: L.ptr(v) for k, v in arrs.items()})
        perm = np.zeros(n, np.uint32); imp = np.zeros((n, 4), np.float32)
        st = L.SolveStats()
        self.ctx.check(self.lib.mgfb_solver_solve(self.ctx.h, C.byref(m), dt, iters, order, L.ptr(perm), L.ptr(imp), C.byref(st)))
        stats = {k: getattr(st, k) for k, _ in st._fields_ if k != "reserved"}
        return perm, imp, stats
