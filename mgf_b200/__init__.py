"""mgf_b200 -- B200-native implementation of mgf's per-step physics hot path.

Host-side mirror of the reference's API for that path (RigidBodyVec / Contacts / Solver /
World::step), over the C ABI in include/mgfb.h.  See DESIGN.md.
"""
from . import _lib
from .api import (BVH, Compound, Context, MgfbError, World, aabb, capsule, contacts_batch, convex_mesh, convex_vertices_set, gjk_batch, intersections_batch, make_shapes, manifolds_prune, obb, plane, rectangle,
                  separation_batch, sphere, triangle)

__all__ = ["BVH", "Compound", "Context", "MgfbError", "World", "sphere", "capsule", "triangle", "rectangle", "plane", "make_shapes",
           "contacts_batch", "convex_mesh", "convex_vertices_set", "gjk_batch", "intersections_batch", "manifolds_prune", "separation_batch", "aabb", "obb", "_lib"]
