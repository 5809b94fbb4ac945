"""ctypes binding of include/mgfb.h (the C ABI of the CUDA library).

The library is built in-tree by ``__graft_entry__.build()`` into ``mgf_b200/lib/libmgfb.so``.
There is no fallback: if the library is missing this module raises, and if no CUDA device is
present ``mgfb_ctx_create`` fails with MGFB_ERR_CUDA.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmgfb.so")

# enum mgfb_status
OK, ERR_INVALID_ARG, ERR_SINGULAR_INERTIA, ERR_CAPACITY, ERR_CUDA, ERR_NAN_BOUNDS, ERR_STATE, ERR_TILE = range(8)
# enum mgfb_shape_kind
SPHERE, CAPSULE, TRIANGLE, RECTANGLE, PLANE, AABB, OBB, CONVEX_MESH = range(8)
# enum mgfb_pair_kind
(SPHERE_X_MSPHERE, CAPSULE_X_MSPHERE, SPHERE_X_MCAPSULE, CAPSULE_X_MCAPSULE, PLANE_X_MSPHERE, PLANE_X_MCAPSULE,
 TRI_X_MSPHERE, TRI_X_MCAPSULE, RECT_X_MSPHERE, RECT_X_MCAPSULE, MCOMP_X_MCOMP, MCOMP_X_TRI) = range(12)
ORDER_AS_GIVEN, ORDER_COLOURED = 0, 1

SHAPE_DTYPE = np.dtype([("kind", "<u4"), ("p", "<f4", (12,)), ("v", "<f4", (3,))])
CONTACT_DTYPE = np.dtype([("a", "<f4", (3,)), ("b", "<f4", (3,)), ("n", "<f4", (3,)), ("t", "<f4")])
INTERSECTION_DTYPE = np.dtype([("p", "<f4", (3,)), ("t", "<f4")])   # collision.rs:151 Intersection
RAY, SEGMENT = 0, 1
INPUT_SET, INPUT_ADD = 0, 1
PHASES = ("integrate", "body_grid", "pair_sweep", "narrow_bodies", "terrain", "colouring", "build_rows", "solve")   # enum mgfb_phase
LOCAL_CONTACT_DTYPE = np.dtype([("local_a", "<f4", (3,)), ("local_b", "<f4", (3,)), ("global", CONTACT_DTYPE)])
assert SHAPE_DTYPE.itemsize == 64 and CONTACT_DTYPE.itemsize == 40 and LOCAL_CONTACT_DTYPE.itemsize == 64


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("penetration_slop", C.c_float), ("baumgarte", C.c_float),
                ("persistent_threshold_sq", C.c_float), ("fat_margin", C.c_float),
                ("initial_body_capacity", C.c_uint32), ("max_cooperative_ctas", C.c_uint32), ("tile_timeout_ms", C.c_uint32),
                ("solver_schedule", C.c_uint32), ("step_order", C.c_uint32)]


SCHEDULE_DATAFLOW, SCHEDULE_PHASES, SCHEDULE_PHASES_JP = 0, 1, 2
STEP_ORDER_COLOURED, STEP_ORDER_REFERENCE = 0, 1


class Manifolds(C.Structure):
    _fields_ = [("n", C.c_uint32), ("obj_a", C.c_void_p), ("obj_b", C.c_void_p), ("static_center", C.c_void_p),
                ("static_friction", C.c_void_p), ("normal", C.c_void_p), ("tangent", C.c_void_p),
                ("ncontacts", C.c_void_p), ("local_a", C.c_void_p), ("local_b", C.c_void_p)]


class SolveStats(C.Structure):
    _fields_ = [("constraints", C.c_uint32), ("contacts", C.c_uint32), ("groups", C.c_uint32),
                ("iterations", C.c_uint32), ("solve_ms", C.c_float), ("reserved", C.c_uint32 * 3)]


class StepStats(C.Structure):
    _fields_ = [("bodies", C.c_uint32), ("candidate_pairs", C.c_uint32), ("terrain_candidates", C.c_uint32),
                ("constraints", C.c_uint32), ("terrain_constraints", C.c_uint32), ("groups", C.c_uint32),
                ("iterations", C.c_uint32), ("fat_refreshes", C.c_uint32), ("step_ms", C.c_float),
                ("solve_ms", C.c_float), ("overflow", C.c_uint32), ("colouring_rounds", C.c_uint32), ("ghosts", C.c_uint32),
                ("boundary_constraints", C.c_uint32), ("phases", C.c_uint32), ("broadphase_path", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PhaseProfile(C.Structure):
    _fields_ = [("phase_ms", C.c_float * 8), ("pairs", C.c_uint32 * 4), ("terrain_pairs", C.c_uint32 * 2),
                ("body_contacts", C.c_uint32), ("terrain_contacts", C.c_uint32)]


class TileDesc(C.Structure):
    _fields_ = [("opaque", C.c_uint64 * 112)]


class DeviceView(C.Structure):
    _fields_ = [("x", C.c_void_p), ("q", C.c_void_p), ("vel", C.c_void_p), ("collider", C.c_void_p),
                ("n", C.c_uint32), ("stream", C.c_void_p)]


# every symbol include/mgfb.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("mgfb_abi_version", C.c_int32, []),
    ("mgfb_config_default", None, [C.POINTER(Config)]),
    ("mgfb_ctx_create", C.c_int32, [C.POINTER(Config), C.POINTER(_P)]),
    ("mgfb_ctx_destroy", None, [_P]),
    ("mgfb_last_error", C.c_char_p, [_P]),
    ("mgfb_synchronize", C.c_int32, [_P]),
    ("mgfb_bodies_add", C.c_int32, [_P, C.c_uint32, _P, _P, _P, _P, _P, C.POINTER(C.c_uint32)]),
    ("mgfb_bodies_count", C.c_int32, [_P, C.POINTER(C.c_uint32)]),
    ("mgfb_bodies_get_state", C.c_int32, [_P, C.c_uint32, C.c_uint32, _P, _P, _P, _P]),
    ("mgfb_bodies_set_velocity", C.c_int32, [_P, C.c_uint32, C.c_uint32, _P, _P]),
    ("mgfb_bodies_get_colliders", C.c_int32, [_P, C.c_uint32, C.c_uint32, _P]),
    ("mgfb_bodies_get_inv_moment", C.c_int32, [_P, C.c_uint32, C.c_uint32, _P]),
    ("mgfb_bodies_get_fat_bounds", C.c_int32, [_P, C.c_uint32, C.c_uint32, _P]),
    ("mgfb_bodies_set_state", C.c_int32, [_P, C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, _P]),
    ("mgfb_integrate", C.c_int32, [_P, C.c_float]),
    ("mgfb_complete_motion", C.c_int32, [_P]),
    ("mgfb_terrain_set", C.c_int32, [_P, _P, C.c_uint32, _P, C.c_uint32, _P]),
    ("mgfb_contacts_batch", C.c_int32, [_P, C.c_uint32, _P, _P, C.c_uint32, _P, _P, _P]),
    ("mgfb_manifolds_prune", C.c_int32, [_P, _P, _P, C.c_uint32, _P, _P, _P, _P, _P, _P]),
    ("mgfb_solver_solve", C.c_int32, [_P, C.POINTER(Manifolds), C.c_float, C.c_uint32, C.c_uint32, _P, _P, C.POINTER(SolveStats)]),
    ("mgfb_step", C.c_int32, [_P, C.c_float, C.c_uint32, C.POINTER(StepStats)]),
    ("mgfb_step_n", C.c_int32, [_P, C.c_float, C.c_uint32, C.c_uint32, C.POINTER(StepStats)]),
    ("mgfb_step_profile", C.c_int32, [_P, C.c_float, C.c_uint32, C.POINTER(StepStats), C.POINTER(PhaseProfile)]),
    ("mgfb_step_enqueue", C.c_int32, [_P, C.c_float, C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, _P]),
    ("mgfb_step_wait", C.c_int32, [_P, C.POINTER(StepStats)]),
    ("mgfb_step_constraints", C.c_int32, [_P, C.c_uint32, _P, _P, _P, _P, _P, C.POINTER(C.c_uint32)]),
    ("mgfb_step_totals", C.c_int32, [_P] + [C.POINTER(C.c_uint64)] * 5 + [C.c_int32]),
    ("mgfb_device_view_get", C.c_int32, [_P, C.POINTER(DeviceView)]),
    ("mgfb_intersections_batch", C.c_int32, [_P, C.c_uint32, _P, _P, C.c_uint32, _P, _P]),
    ("mgfb_bvh_create", C.c_int32, [_P, C.POINTER(_P)]),
    ("mgfb_bvh_destroy", None, [_P]),
    ("mgfb_bvh_insert", C.c_int32, [_P, _P, _P, C.c_uint32, _P]),
    ("mgfb_bvh_remove", C.c_int32, [_P, _P, C.c_uint32]),
    ("mgfb_bvh_get", C.c_int32, [_P, C.c_uint32, _P, C.POINTER(C.c_uint32)]),
    ("mgfb_bvh_len", C.c_int32, [_P, C.POINTER(C.c_uint32)]),
    ("mgfb_bvh_query_batch", C.c_int32, [_P, _P, C.c_uint32, _P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("mgfb_bvh_raytrace_batch", C.c_int32, [_P, C.c_uint32, _P, C.c_uint32, _P, _P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("mgfb_compound_create", C.c_int32, [_P, _P, C.c_uint32, C.POINTER(_P)]),
    ("mgfb_compound_destroy", None, [_P]),
    ("mgfb_compound_set_transform", C.c_int32, [_P, _P, _P]),
    ("mgfb_compound_bounds", C.c_int32, [_P, _P, _P]),
    ("mgfb_compound_closest_points", C.c_int32, [_P, _P, C.c_uint32, _P]),
    ("mgfb_compound_intersections_batch", C.c_int32, [_P, C.c_uint32, _P, C.c_uint32, _P, _P]),
    ("mgfb_compound_contacts_batch", C.c_int32, [_P, _P, C.c_uint32, C.c_uint32, _P, _P]),
    ("mgfb_convex_vertices_set", C.c_int32, [_P, _P, C.c_uint32]),
    ("mgfb_gjk_batch", C.c_int32, [_P, _P, _P, C.c_uint32, _P, _P, _P]),
    ("mgfb_separation_batch", C.c_int32, [_P, _P, _P, C.c_uint32, _P, _P]),
    ("mgfb_bodies_set_gid", C.c_int32, [_P, C.c_uint32, C.c_uint32, _P]),
    ("mgfb_tile_export", C.c_int32, [_P, C.c_uint32, C.POINTER(TileDesc)]),
    ("mgfb_tile_connect", C.c_int32, [_P, C.c_uint32, C.c_uint32, _P]),
    ("mgfb_selftest_handover", C.c_int32, [_P, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
]

_lib = None


def load():
    """Load libmgfb.so (once) and set the prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MGFB_LIB", LIB_PATH)   # development builds (e.g. -DMGFB_DF_PROFILE); still a CUDA library, never a fallback
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  mgf_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the ABI and the header disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
