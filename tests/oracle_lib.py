"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: nothing
under mgf_b200/ imports this; it is the checker, never the thing measured or shipped."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ODIR, "liboracle.so")

import sys
sys.path.insert(0, ROOT)
from mgf_b200 import _lib as L  # POD struct layouts only (numpy dtypes); does not load libmgfb

_P = C.c_void_p
_lib = None


def build():
    subprocess.run(["make", "-C", ODIR, "-s", "all"], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    srcs = [os.path.join(ODIR, f) for f in os.listdir(ODIR) if f.endswith((".hpp", ".cpp"))]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        build()
    lib = C.CDLL(SO)
    lib.mgfo_bvh_create.restype = _P; lib.mgfo_bvh_create.argtypes = []
    lib.mgfo_bvh_destroy.restype = None; lib.mgfo_bvh_destroy.argtypes = [_P]
    lib.mgfo_bvh_insert.restype = C.c_uint32; lib.mgfo_bvh_insert.argtypes = [_P, _P, C.c_uint32]
    lib.mgfo_bvh_remove.restype = None; lib.mgfo_bvh_remove.argtypes = [_P, C.c_uint32]
    lib.mgfo_bvh_query.restype = C.c_uint32; lib.mgfo_bvh_query.argtypes = [_P, _P, _P, C.c_uint32]
    lib.mgfo_bvh_raytrace.restype = C.c_uint32; lib.mgfo_bvh_raytrace.argtypes = [_P, C.c_uint32, _P, _P, _P, C.c_uint32]
    lib.mgfo_intersections_batch.restype = C.c_int32
    lib.mgfo_intersections_batch.argtypes = [C.c_uint32, _P, _P, C.c_uint32, _P, _P]
    lib.mgfo_contacts_batch.restype = C.c_int32
    lib.mgfo_contacts_batch.argtypes = [C.c_uint32, _P, _P, C.c_uint32, _P, _P, _P]
    lib.mgfo_compound_create.restype = _P; lib.mgfo_compound_create.argtypes = [_P, C.c_uint32]
    lib.mgfo_compound_destroy.restype = None; lib.mgfo_compound_destroy.argtypes = [_P]
    lib.mgfo_compound_set_transform.restype = None; lib.mgfo_compound_set_transform.argtypes = [_P, _P, _P]
    lib.mgfo_compound_bounds.restype = None; lib.mgfo_compound_bounds.argtypes = [_P, _P, _P]
    lib.mgfo_compound_closest_points.restype = None; lib.mgfo_compound_closest_points.argtypes = [_P, _P, C.c_uint32, _P]
    lib.mgfo_compound_intersections_batch.restype = None; lib.mgfo_compound_intersections_batch.argtypes = [_P, C.c_uint32, _P, C.c_uint32, _P, _P]
    lib.mgfo_compound_contacts_batch.restype = C.c_int32; lib.mgfo_compound_contacts_batch.argtypes = [_P, _P, C.c_uint32, C.c_uint32, _P, _P]
    lib.mgfo_manifolds_prune.restype = C.c_int32; lib.mgfo_manifolds_prune.argtypes = [_P, _P, C.c_uint32, _P, _P, _P, _P, _P, _P]
    lib.mgfo_world_create.restype = _P
    lib.mgfo_world_create.argtypes = [C.c_float]
    lib.mgfo_world_destroy.argtypes = [_P]
    lib.mgfo_world_add_bodies.restype = C.c_int32
    lib.mgfo_world_add_bodies.argtypes = [_P, C.c_uint32, _P, _P, _P, _P, _P]
    lib.mgfo_world_set_terrain.restype = C.c_int32
    lib.mgfo_world_set_terrain.argtypes = [_P, _P, C.c_uint32, _P, C.c_uint32, _P]
    lib.mgfo_world_count.restype = C.c_uint32
    lib.mgfo_world_count.argtypes = [_P]
    lib.mgfo_world_get_state.argtypes = [_P, _P, _P, _P, _P]
    lib.mgfo_world_set_velocity.argtypes = [_P, C.c_uint32, C.c_uint32, _P, _P]
    lib.mgfo_world_get_colliders.argtypes = [_P, _P]
    lib.mgfo_world_get_inv_moment.argtypes = [_P, _P]
    lib.mgfo_world_get_fat_bounds.argtypes = [_P, _P]
    lib.mgfo_world_set_state.restype = C.c_int32
    lib.mgfo_world_set_state.argtypes = [_P, _P, _P, _P, _P, _P, _P]
    lib.mgfo_world_integrate.argtypes = [_P, C.c_float]
    lib.mgfo_world_complete_motion.argtypes = [_P]
    lib.mgfo_world_step.restype = C.c_int32
    lib.mgfo_world_step.argtypes = [_P, C.c_float, C.c_uint32, C.c_uint32]
    lib.mgfo_world_build.restype = C.c_int32
    lib.mgfo_world_build.argtypes = [_P, C.c_float, C.POINTER(C.c_uint32)]
    lib.mgfo_world_constraints.argtypes = [_P, _P, _P, _P, _P]
    lib.mgfo_world_manifolds.argtypes = [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]
    lib.mgfo_world_brute_pairs.restype = C.c_uint64
    lib.mgfo_world_brute_pairs.argtypes = [_P]
    lib.mgfo_world_stats.argtypes = [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.mgfo_world_solve_order.argtypes = [_P, _P, C.c_uint32, C.c_uint32]
    lib.mgfo_world_solve_manifolds.restype = C.c_int32
    lib.mgfo_world_solve_manifolds.argtypes = [_P, C.POINTER(L.Manifolds), C.c_float, C.c_uint32, _P, _P]
    lib.mgfo_world_time_steps.restype = C.c_double
    lib.mgfo_world_time_steps.argtypes = [_P, C.c_float, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.mgfo_convex_vertices_set.restype = None; lib.mgfo_convex_vertices_set.argtypes = [_P, C.c_uint32]
    lib.mgfo_gjk_batch.restype = C.c_int32
    lib.mgfo_gjk_batch.argtypes = [_P, _P, C.c_uint32, _P, _P, _P]
    lib.mgfo_separation_batch.restype = C.c_int32
    lib.mgfo_separation_batch.argtypes = [_P, _P, C.c_uint32, _P, _P]
    _lib = lib
    return lib


_convex_pool = None


def convex_vertices_set(verts):
    """Oracle of mgfb_convex_vertices_set: the vertex pool MGFB_CONVEX_MESH shapes index."""
    global _convex_pool
    _convex_pool = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)   # kept alive here
    load().mgfo_convex_vertices_set(L.ptr(_convex_pool), len(_convex_pool))


def gjk_batch(a, b):
    """Oracle of mgfb_gjk_batch: (contacts[n], hit[n], epa_iterations[n])."""
    lib = load()
    a = np.ascontiguousarray(a, dtype=L.SHAPE_DTYPE); b = np.ascontiguousarray(b, dtype=L.SHAPE_DTYPE)
    n = len(a)
    out = np.zeros(n, dtype=L.CONTACT_DTYPE); hit = np.zeros(n, np.uint32); iters = np.zeros(n, np.uint32)
    assert lib.mgfo_gjk_batch(L.ptr(a), L.ptr(b), n, L.ptr(out), L.ptr(hit), L.ptr(iters)) == 0
    return out, hit, iters


def separation_batch(a, b):
    """Oracle of mgfb_separation_batch: (distance[n], is_some[n])."""
    lib = load()
    a = np.ascontiguousarray(a, dtype=L.SHAPE_DTYPE); b = np.ascontiguousarray(b, dtype=L.SHAPE_DTYPE)
    n = len(a)
    sep = np.zeros(n, np.float32); some = np.zeros(n, np.uint32)
    assert lib.mgfo_separation_batch(L.ptr(a), L.ptr(b), n, L.ptr(sep), L.ptr(some)) == 0
    return sep, some


def manifolds_prune(contacts, offsets):
    """Oracle of mgfb_manifolds_prune."""
    lib = load()
    contacts = np.ascontiguousarray(contacts, dtype=L.LOCAL_CONTACT_DTYPE)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    g = len(offsets) - 1
    d = dict(time=np.zeros(g, np.float32), normal=np.zeros((g, 3), np.float32), tangent=np.zeros((g, 6), np.float32), ncontacts=np.zeros(g, np.uint32),
             local_a=np.zeros((g, 12), np.float32), local_b=np.zeros((g, 12), np.float32))
    assert lib.mgfo_manifolds_prune(L.ptr(contacts), L.ptr(offsets), g, *[L.ptr(d[k]) for k in ("time", "normal", "tangent", "ncontacts", "local_a", "local_b")]) == 0
    return d


def contacts_batch(pair_kind, recv, arg, want_local=False):
    lib = load()
    recv = np.ascontiguousarray(recv, dtype=L.SHAPE_DTYPE)
    arg = np.ascontiguousarray(arg, dtype=L.SHAPE_DTYPE)
    n = len(recv)
    out = np.zeros((n, 2), dtype=L.CONTACT_DTYPE)
    loc = np.zeros((n, 2), dtype=L.LOCAL_CONTACT_DTYPE) if want_local else None
    counts = np.zeros(n, dtype=np.uint32)
    st = lib.mgfo_contacts_batch(pair_kind, L.ptr(recv), L.ptr(arg), n, L.ptr(out), L.ptr(loc), L.ptr(counts))
    assert st == 0
    return (out, counts, loc) if want_local else (out, counts)


def intersections_batch(particle_kind, particles, shapes):
    lib = load()
    particles = np.ascontiguousarray(particles, dtype=np.float32).reshape(-1, 6)
    shapes = np.ascontiguousarray(shapes, dtype=L.SHAPE_DTYPE)
    n = len(shapes)
    out = np.zeros(n, dtype=L.INTERSECTION_DTYPE); hit = np.zeros(n, np.uint32)
    st = lib.mgfo_intersections_batch(particle_kind, L.ptr(particles), L.ptr(shapes), n, L.ptr(out), L.ptr(hit))
    assert st == 0
    return out, hit


class OracleBVH:
    """src/bvh.rs BVH<AABB, u32> on the CPU restatement (incremental insert / remove / balance)."""

    def __init__(self):
        self.lib = load()
        self.h = self.lib.mgfo_bvh_create()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.mgfo_bvh_destroy(self.h)
            self.h = None

    def insert(self, box, value):
        box = np.ascontiguousarray(box, dtype=np.float32)
        return self.lib.mgfo_bvh_insert(self.h, L.ptr(box), int(value))

    def remove(self, index):
        self.lib.mgfo_bvh_remove(self.h, int(index))

    def query(self, box, cap=65536):
        box = np.ascontiguousarray(box, dtype=np.float32)
        out = np.zeros(cap, np.uint32)
        n = self.lib.mgfo_bvh_query(self.h, L.ptr(box), L.ptr(out), cap)
        assert n <= cap
        return out[:n]

    def raytrace(self, kind, particle, cap=65536):
        particle = np.ascontiguousarray(particle, dtype=np.float32)
        out = np.zeros(cap, np.uint32); hits = np.zeros(cap, dtype=L.INTERSECTION_DTYPE)
        n = self.lib.mgfo_bvh_raytrace(self.h, kind, L.ptr(particle), L.ptr(out), L.ptr(hits), cap)
        assert n <= cap
        return out[:n], hits[:n]


class OracleCompound:
    """src/compound.rs Compound on the CPU restatement (same methods as mgf_b200.Compound)."""

    def __init__(self, components):
        self.lib = load()
        comps = np.ascontiguousarray(components, dtype=L.SHAPE_DTYPE)
        self.n = len(comps)
        self.h = self.lib.mgfo_compound_create(L.ptr(comps), len(comps))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.mgfo_compound_destroy(self.h)
            self.h = None

    def set_transform(self, disp, rot):
        d = np.ascontiguousarray(disp, dtype=np.float32); r = np.ascontiguousarray(rot, dtype=np.float32)
        self.lib.mgfo_compound_set_transform(self.h, L.ptr(d), L.ptr(r))

    def bounds(self):
        a = np.zeros(6, np.float32); s = np.zeros(4, np.float32)
        self.lib.mgfo_compound_bounds(self.h, L.ptr(a), L.ptr(s))
        return a, s

    def closest_points(self, to):
        to = np.ascontiguousarray(to, dtype=np.float32).reshape(-1, 3)
        out = np.zeros_like(to)
        self.lib.mgfo_compound_closest_points(self.h, L.ptr(to), len(to), L.ptr(out))
        return out

    def intersections(self, particle_kind, particles):
        particles = np.ascontiguousarray(particles, dtype=np.float32).reshape(-1, 6)
        out = np.zeros(len(particles), dtype=L.INTERSECTION_DTYPE); hit = np.zeros(len(particles), np.uint32)
        self.lib.mgfo_compound_intersections_batch(self.h, particle_kind, L.ptr(particles), len(particles), L.ptr(out), L.ptr(hit))
        return out, hit

    def contacts(self, rhs, slots=None):
        rhs = np.ascontiguousarray(rhs, dtype=L.SHAPE_DTYPE)
        slots = slots or max(2 * self.n, 1)
        out = np.zeros((len(rhs), slots), dtype=L.CONTACT_DTYPE); counts = np.zeros(len(rhs), np.uint32)
        assert self.lib.mgfo_compound_contacts_batch(self.h, L.ptr(rhs), len(rhs), slots, L.ptr(out), L.ptr(counts)) == 0
        return out, counts


class OracleWorld:
    """mgf_demo World::step on the CPU restatement."""

    def __init__(self, fat_margin=0.25):
        self.lib = load()
        self.h = self.lib.mgfo_world_create(fat_margin)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.mgfo_world_destroy(self.h)
            self.h = None

    def add_bodies(self, shapes, mass, restitution, friction, world_force):
        shapes = np.ascontiguousarray(shapes, dtype=L.SHAPE_DTYPE)
        n = len(shapes)
        f32 = lambda a, shape: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float32), shape))
        st = self.lib.mgfo_world_add_bodies(self.h, n, L.ptr(shapes), L.ptr(f32(mass, (n,))), L.ptr(f32(restitution, (n,))),
                                            L.ptr(f32(friction, (n,))), L.ptr(f32(world_force, (n, 3))))
        assert st == 0, st

    def set_terrain(self, verts, faces, x=(0, 0, 0)):
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        faces = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        x = np.asarray(x, dtype=np.float32)
        self.lib.mgfo_world_set_terrain(self.h, L.ptr(verts), len(verts), L.ptr(faces), len(faces), L.ptr(x))

    def __len__(self):
        return self.lib.mgfo_world_count(self.h)

    def state(self):
        n = len(self)
        x = np.zeros((n, 3), np.float32); q = np.zeros((n, 4), np.float32)
        v = np.zeros((n, 3), np.float32); w = np.zeros((n, 3), np.float32)
        self.lib.mgfo_world_get_state(self.h, L.ptr(x), L.ptr(q), L.ptr(v), L.ptr(w))
        return x, q, v, w

    def set_velocity(self, first, v, omega):
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 3)
        omega = np.ascontiguousarray(omega, dtype=np.float32).reshape(-1, 3)
        self.lib.mgfo_world_set_velocity(self.h, first, len(v), L.ptr(v), L.ptr(omega))

    def colliders(self):
        out = np.zeros(len(self), dtype=L.SHAPE_DTYPE)
        self.lib.mgfo_world_get_colliders(self.h, L.ptr(out))
        return out

    def inv_moment(self):
        out = np.zeros((len(self), 9), np.float32)
        self.lib.mgfo_world_get_inv_moment(self.h, L.ptr(out))
        return out

    def fat_bounds(self):
        out = np.zeros((len(self), 6), np.float32)
        self.lib.mgfo_world_get_fat_bounds(self.h, L.ptr(out))
        return out

    def snapshot(self):
        x, q, v, w = self.state()
        return dict(x=x, q=q, v=v, omega=w, colliders=self.colliders(), fat=self.fat_bounds())

    def restore(self, snap):
        """Load a state saved by snapshot() -- of this oracle or of mgf_b200.World (same layout)."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        col = np.ascontiguousarray(snap["colliders"], dtype=L.SHAPE_DTYPE)
        assert len(col) == len(self)
        st = self.lib.mgfo_world_set_state(self.h, L.ptr(f(snap["x"])), L.ptr(f(snap["q"])), L.ptr(f(snap["v"])), L.ptr(f(snap["omega"])),
                                           L.ptr(col), L.ptr(f(snap["fat"])))
        assert st == 0, st

    def integrate(self, dt):
        self.lib.mgfo_world_integrate(self.h, dt)

    def complete_motion(self):
        self.lib.mgfo_world_complete_motion(self.h)

    def step(self, dt, iters=20, nsteps=1):
        st = self.lib.mgfo_world_step(self.h, dt, iters, nsteps)
        assert st == 0, st

    def build(self, dt):
        cnt = C.c_uint32()
        st = self.lib.mgfo_world_build(self.h, dt, C.byref(cnt))
        assert st == 0, st
        return cnt.value

    def constraints(self, m):
        a = np.zeros(m, np.uint32); b = np.zeros(m, np.int32); face = np.zeros(m, np.uint32); sub = np.zeros(m, np.uint32)
        self.lib.mgfo_world_constraints(self.h, L.ptr(a), L.ptr(b), L.ptr(face), L.ptr(sub))
        return a, b, face, sub

    def manifolds(self, m):
        d = dict(obj_a=np.zeros(m, np.int32), obj_b=np.zeros(m, np.int32), static_center=np.zeros((m, 3), np.float32),
                 static_friction=np.zeros(m, np.float32), normal=np.zeros((m, 3), np.float32), tangent=np.zeros((m, 6), np.float32),
                 ncontacts=np.zeros(m, np.uint32), local_a=np.zeros((m, 12), np.float32), local_b=np.zeros((m, 12), np.float32))
        self.lib.mgfo_world_manifolds(self.h, *[L.ptr(d[k]) for k in ("obj_a", "obj_b", "static_center", "static_friction", "normal",
                                                                       "tangent", "ncontacts", "local_a", "local_b")])
        return d

    def stats(self):
        a = C.c_uint64(); b = C.c_uint64()
        self.lib.mgfo_world_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def brute_pairs(self):
        return self.lib.mgfo_world_brute_pairs(self.h)

    def solve_order(self, perm, iters):
        perm = np.ascontiguousarray(perm, dtype=np.uint32)
        self.lib.mgfo_world_solve_order(self.h, L.ptr(perm), len(perm), iters)

    def solve_manifolds(self, d, dt, iters, perm=None):
        n = len(d["obj_a"])
        m = L.Manifolds(n=n, **{k: L.ptr(np.ascontiguousarray(v)) for k, v in d.items()})
        keep = [np.ascontiguousarray(v) for v in d.values()]  # noqa: F841 keep alive
        m = L.Manifolds(n=n, **{k: L.ptr(v) for k, v in zip(d.keys(), keep)})
        imp = np.zeros((n, 4), np.float32)
        p = None if perm is None else np.ascontiguousarray(perm, dtype=np.uint32)
        st = self.lib.mgfo_world_solve_manifolds(self.h, C.byref(m), dt, iters, L.ptr(p), L.ptr(imp))
        assert st == 0
        return imp

    def time_steps(self, dt, iters, nsteps):
        ci = C.c_uint64(); pr = C.c_uint64()
        sec = self.lib.mgfo_world_time_steps(self.h, dt, iters, nsteps, C.byref(ci), C.byref(pr))
        return sec, ci.value, pr.value
