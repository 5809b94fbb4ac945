"""Writes tests/golden/reference_kat.json: the known-answer vectors of the reference's OWN unit
tests for the continuous narrowphase (src/collision.rs `mod tests`), as literals.

Each case is one `recv.contacts(&arg, cb)` call: inputs are the exact literals of the Rust
test, `expect` lists what the test asserts.  mode "eq" = assert_eq! on f32 (bit-exact),
"rel" = assert_relative_eq!(.., epsilon = eps) with approx-0.3 semantics (max_relative =
f32::EPSILON).  The same file checks the CPU oracle (-m "not gpu") and the CUDA kernels
(-m gpu) through mgfb_contacts_batch.  Nothing is computed here: run it only to regenerate
the JSON after editing the literals.
"""
import json
import os

SPH, CAP, TRI, RECT, PLANE = 0, 1, 2, 3, 4
K = dict(SPHERE_X_MSPHERE=0, CAPSULE_X_MSPHERE=1, SPHERE_X_MCAPSULE=2, CAPSULE_X_MCAPSULE=3, PLANE_X_MSPHERE=4,
         PLANE_X_MCAPSULE=5, TRI_X_MSPHERE=6, TRI_X_MCAPSULE=7, RECT_X_MSPHERE=8, RECT_X_MCAPSULE=9, MCOMP_X_MCOMP=10)
E = 0.000001        # geom::COLLISION_EPSILON
FE = 1.1920929e-07  # f32::EPSILON (assert_relative_eq! default)


def sphere(c, r, v=(0, 0, 0)): return dict(kind=SPH, p=[*c, r], v=list(v))
def capsule(a, d, r, v=(0, 0, 0)): return dict(kind=CAP, p=[*a, *d, r], v=list(v))
def tri(a, b, c): return dict(kind=TRI, p=[*a, *b, *c], v=[0, 0, 0])
def rect(c, u0, u1, e0, e1): return dict(kind=RECT, p=[*c, *u0, *u1, e0, e1], v=[0, 0, 0])


def eq(i, field, val): return dict(i=i, field=field, mode="eq", val=val)
def rel(i, field, val, eps=E): return dict(i=i, field=field, mode="rel", val=val, eps=eps)


cases = []


def case(name, src, kind, recv, arg, count=None, min_count=None, expect=()):
    cases.append(dict(name=name, src=src, pair_kind=K[kind], recv=recv, arg=arg, count=count, min_count=min_count, expect=list(expect)))


# ---- spheres::test_rect_collision  collision.rs:1699-1758
floor_r = rect((0, 1, 0), (1, 0, 0), (0, 0, 1), 3.0, 3.0)
case("rect_sphere_center", "collision.rs:1705-1717", "RECT_X_MSPHERE", floor_r, sphere((0, 13, 0), 2.0, (0, -10, 0)), count=1,
     expect=[eq(0, "a", [0, 1, 0]), eq(0, "b", [0, 1, 0]), eq(0, "t", 1.0), eq(0, "n", [0, 1, 0])])
case("rect_sphere_center_2s", "collision.rs:1724-1736", "RECT_X_MSPHERE", floor_r, sphere((0, 13, 0), 2.0, (0, -20, 0)), count=1,
     expect=[eq(0, "a", [0, 1, 0]), eq(0, "b", [0, 1, 0]), eq(0, "t", 0.5), eq(0, "n", [0, 1, 0])])
case("rect_sphere_corner", "collision.rs:1737-1749", "RECT_X_MSPHERE", floor_r, sphere((0, 13, 0), 2.0, (0, -10, 3)), count=1,
     expect=[eq(0, "a", [0, 1, 3]), eq(0, "b", [0, 1, 3]), eq(0, "t", 1.0), eq(0, "n", [0, 1, 0])])
case("rect_sphere_miss_corner", "collision.rs:1750-1757", "RECT_X_MSPHERE", floor_r, sphere((0, 13, 0), 2.0, (0, -10, 3.00001)), count=0)

# ---- spheres::test_tri_collision  collision.rs:1761-1814   (a, c, b literal order in the test)
floor_t = tri((1, 1, 0), (0, 1, -1), (0, 1, 1))
case("tri_sphere_center", "collision.rs:1767-1779", "TRI_X_MSPHERE", floor_t, sphere((0, 13, 0), 2.0, (0, -10, 0)), count=1,
     expect=[eq(0, "a", [0, 1, 0]), eq(0, "b", [0, 1, 0]), eq(0, "t", 1.0), eq(0, "n", [0, 1, 0])])
case("tri_sphere_corner", "collision.rs:1780-1792", "TRI_X_MSPHERE", floor_t, sphere((0, 13, 0), 2.0, (0, -10, 1)), count=1,
     expect=[rel(0, "a", [0, 1, 1]), rel(0, "b", [0, 1, 1]), dict(i=0, field="t", mode="one_minus_lt", val=E), eq(0, "n", [0, 1, 0])])
case("tri_sphere_miss_corner", "collision.rs:1793-1800", "TRI_X_MSPHERE", floor_t, sphere((0, 13, 0), 2.0, (0, -10, 1.00001)), count=0)
case("tri_sphere_edge", "collision.rs:1801-1813", "TRI_X_MSPHERE", floor_t, sphere((0, 13, 0), 2.0, (0.5, -10, 0.5)), count=1,
     expect=[eq(0, "a", [0.5, 1, 0.5]), eq(0, "b", [0.5, 1, 0.5]), eq(0, "t", 1.0), eq(0, "n", [0, 1, 0])])

# ---- spheres::test_moving_spheres_collision collision.rs:1675-1696 (Moving x Moving, :1387)
case("msphere_msphere", "collision.rs:1675-1696", "MCOMP_X_MCOMP", sphere((-3, 0, 0), 1.0, (1, 0, 0)),
     sphere((3, 0, 0), 2.0, (-2, 0, 0)), count=1, expect=[eq(0, "t", 1.0), eq(0, "a", [-1, 0, 0]), eq(0, "b", [-1, 0, 0]), eq(0, "n", [1, 0, 0])])

# ---- capsules::test_moving_sphere_collision collision.rs:1853-1874
case("capsule_msphere", "collision.rs:1854-1869", "CAPSULE_X_MSPHERE", capsule((4, 3, 5.5), (0, 1, 0), 2.0),
     sphere((0, 3, 5.5), 1.0, (1, 0, 0)), count=1, expect=[eq(0, "t", 1.0), eq(0, "a", [2, 3, 5.5]), eq(0, "b", [2, 3, 5.5])])

# ---- capsules::test_moving_capsule_collision collision.rs:1877-1980
case("capcap_1", "collision.rs:1878-1894", "CAPSULE_X_MCAPSULE", capsule((4, 3, 5.5), (0, 1, 0), 2.0),
     capsule((0, 3, 5.5), (0, 1, 0), 1.0, (1, 0, 0)), count=1, expect=[eq(0, "t", 1.0), eq(0, "a", [2, 3.5, 5.5]), eq(0, "b", [2, 3.5, 5.5])])
case("capcap_2", "collision.rs:1895-1911", "CAPSULE_X_MCAPSULE", capsule((4, 3, 5.5), (0, 1, 0), 1.0),
     capsule((0, 3, 5.5), (0, 1, 0), 2.0, (1, 0, 0)), count=1, expect=[eq(0, "a", [3, 3.5, 5.5]), eq(0, "b", [3, 3.5, 5.5]), eq(0, "t", 1.0)])
case("capcap_end_to_end", "collision.rs:1912-1928", "CAPSULE_X_MCAPSULE", capsule((1, 0, 0), (1, 0, 0), 1.0),
     capsule((-2, 0, 0), (-1, 0, 0), 1.0, (2, 0, 0)), count=1, expect=[eq(0, "a", [0, 0, 0]), eq(0, "b", [0, 0, 0]), eq(0, "t", 0.5)])
case("capcap_overlapping_parallel", "collision.rs:1929-1945", "CAPSULE_X_MCAPSULE", capsule((0, 0, 0), (1, 0, 0), 1.0),
     capsule((0, 0, 0), (-1, 0, 0), 1.0, (2, 0, 0)), count=1, expect=[eq(0, "a", [-1, 0, 0]), eq(0, "b", [1, 0, 0]), eq(0, "t", 0.0)])
case("capcap_offset_parallel_1", "collision.rs:1946-1962", "CAPSULE_X_MCAPSULE", capsule((4, 3, 5.5), (0, 1, 0), 2.0),
     capsule((0, 2, 5.5), (0, 1, 0), 1.0, (1, 0, 0)), count=1, expect=[eq(0, "t", 1.0), eq(0, "a", [2, 3, 5.5]), eq(0, "b", [2, 3, 5.5])])
case("capcap_offset_parallel_2", "collision.rs:1963-1979", "CAPSULE_X_MCAPSULE", capsule((4, 3, 5.5), (0, 1, 0), 2.0),
     capsule((0, 2.5, 5.5), (0, 1, 0), 1.0, (1, 0, 0)), count=1, expect=[eq(0, "t", 1.0), eq(0, "a", [2, 3.25, 5.5]), eq(0, "b", [2, 3.25, 5.5])])

# ---- capsules::test_rect_collision collision.rs:1983-2003
case("rect_capsule_level", "collision.rs:1984-2001", "RECT_X_MCAPSULE", floor_r, capsule((1, 13, 0), (3, 0, 0), 2.0, (0, -10, 0)), min_count=2,
     expect=[eq(0, "t", 1.0), rel(0, "a", [1, 1, 0]), rel(1, "a", [3, 1, 0])])

# ---- capsules::test_tri_collision collision.rs:2006-2268
def tc(name, src, a, d, r, v, floor=floor_t, **kw):
    case(name, src, "TRI_X_MCAPSULE", floor, capsule(a, d, r, v), **kw)


tc("tri_cap_clip_edge", "collision.rs:2012-2024", (0.9, 3, 1), (0, 0, -2), 1.0, (0, -1, 0), min_count=2,
   expect=[eq(0, "t", 1.0), rel(0, "a", [0.9, 1, 0.1]), rel(1, "a", [0.9, 1, -0.1])])
tc("tri_cap_clip_off_center_1", "collision.rs:2026-2037", (0.9, 3, 0), (0, 0, 2), 1.0, (0, -1, 0), min_count=2,
   expect=[eq(0, "t", 1.0), rel(0, "a", [0.9, 1, 0.0]), rel(1, "a", [0.9, 1, 0.1])])
tc("tri_cap_clip_off_center_2", "collision.rs:2039-2050", (0.9, 3, 0), (0, 0, -2), 1.0, (0, -1, 0), min_count=2,
   expect=[eq(0, "t", 1.0), rel(0, "a", [0.9, 1, 0.0]), rel(1, "a", [0.9, 1, -0.1])])
tc("tri_cap_through_center", "collision.rs:2052-2063", (0.9, 2, 0), (1, 0, 0), 1.0, (0, -1, 0), min_count=2,
   expect=[eq(0, "t", 0.0), rel(0, "a", [0.9, 1, 0.0]), rel(1, "a", [1.0, 1, 0.0])])
tc("tri_cap_tilted_center", "collision.rs:2065-2078 (last_contact)", (0.5, 4, 0), (-1, -0.5, 0), 1.0, (0, -2, 0), min_count=1,
   expect=[eq(-1, "t", 0.81598306), rel(-1, "a", [0, 1, 0])])
tc("tri_cap_tilted_skew", "collision.rs:2093-2103 (last_contact)", (0.5, 4, 0), (-1, -1, 2), 1.0, (0, -2, 0), min_count=1,
   expect=[rel(-1, "a", [0, 1, 1]), eq(-1, "t", 0.7022774)])
tc("tri_cap_parallel_edge_1", "collision.rs:2104-2116", (-1, 2, 2), (0, 0, -2), 1.0, (0, -1, 0), count=2,
   expect=[eq(0, "t", 1.0), rel(0, "a", [0, 1, 1]), rel(1, "a", [0, 1, 0])])
tc("tri_cap_parallel_edge_2", "collision.rs:2118-2129", (-1, 4, 2), (0, -2, -2), 1.0, (0, -1, 0), count=1,
   expect=[eq(0, "t", 1.0), rel(0, "a", [0, 1, 0])])
# :2130-2141 -- the reference does not clear `contacts` before this call and still asserts
# len() == 1, i.e. this call must emit nothing.
tc("tri_cap_parallel_edge_3_emits_nothing", "collision.rs:2130-2141", (-1, 4, 0), (0, 2, -2), 1.0, (0, -1, 0), count=0)
tc("tri_cap_parallel_edge_4", "collision.rs:2143-2155", (-1, 2, 2), (0, 0, -4), 1.0, (0, -1, 0), count=2,
   expect=[eq(0, "t", 1.0), rel(0, "a", [0, 1, 1]), rel(1, "a", [0, 1, -1])])
tc("tri_cap_parallel_edge_5", "collision.rs:2157-2169", (-1, 2, -2), (0, 0, 4), 1.0, (0, -1, 0), count=2,
   expect=[eq(0, "t", 1.0), rel(0, "a", [0, 1, -1]), rel(1, "a", [0, 1, 1])])
floor2 = tri((1, 1, 0), (0, 1, 2), (0, 1, -2))
tc("tri2_cap_parallel_edge", "collision.rs:2171-2188", (-0.5, 2, 0.5), (0, 0, -1), 0.5, (0, -1, 0), floor=floor2, count=2,
   expect=[eq(0, "t", 1.0), rel(0, "a", [0, 1, 0.5]), rel(1, "a", [0, 1, -0.5])])
tc("tri2_cap_perp_edge_1", "collision.rs:2190-2201", (-1, 2, 0), (-3, 0, 0), 1.0, (0, -1, 0), floor=floor2, count=1,
   expect=[eq(0, "t", 1.0), rel(0, "a", [0, 1, 0])])
tc("tri2_cap_perp_edge_2", "collision.rs:2203-2214", (-4, 2, 0), (3, 0, 0), 1.0, (0, -1, 0), floor=floor2, count=1,
   expect=[eq(0, "t", 1.0), rel(0, "a", [0, 1, 0])])
tc("tri2_cap_next_to_vert", "collision.rs:2216-2227", (2, 2, 1), (0, 0, -2), 1.0, (0, -1, 0), floor=floor2, count=1,
   expect=[eq(0, "t", 1.0), rel(0, "a", [1, 1, 0])])
tc("tri2_cap_next_to_vert_skewed", "collision.rs:2229-2240", (2, 2, 1), (0, -1, -2), 1.0, (0, -1, 0), floor=floor2, count=1,
   expect=[eq(0, "t", 0.5), rel(0, "a", [1, 1, 0])])
tc("tri2_cap_intersects_plane_1", "collision.rs:2242-2253", (0, 4, 0), (-2, -4, 0), 1.0, (0, -1, 0), floor=floor2, count=1,
   expect=[rel(0, "t", 0.7639319, FE), rel(0, "a", [0, 1, 0])])
tc("tri2_cap_intersects_plane_2", "collision.rs:2255-2266", (-1, 2, 0), (-1, -2, 0), 1.0, (0, -1, 0), floor=floor2, count=1,
   expect=[rel(0, "t", 1.0, FE), rel(0, "a", [0, 1, 0])])

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kat.json")
    with open(out, "w") as f:
        json.dump(dict(source="maplant/mgf v1.4.0 src/collision.rs mod tests", cases=cases), f, indent=1)
    print(f"wrote {len(cases)} cases to {out}")
