"""Writes tests/golden/reference_ray_kat.json: the known-answer vectors of the reference's own
`test_ray_intersections` (src/collision.rs:1543-1637), as literals.  Each case is one
`ray.intersection(&capsule).unwrap()`; `expect` lists what the Rust test asserts ("eq" =
assert_eq! on f32, "rel" = assert_relative_eq! with the given epsilon, "along" = the test's
`r.p + r.d*t` compared with relative_eq).  Ray directions given as `.normalize()` in the test are
stored un-normalised with normalize=true: the checker applies cgmath's normalize in f32
(v * (1/|v|)).  Nothing is computed here."""
import json
import os

E = 0.000001        # geom::COLLISION_EPSILON
FE = 1.1920929e-07  # f32::EPSILON (assert_relative_eq! default)
cases = []


def case(name, src, cap, p, d, normalize=False, expect=()):
    cases.append(dict(name=name, src=src, capsule=dict(a=cap[0], d=cap[1], r=cap[2]), p=list(p), d=list(d), normalize=normalize, expect=list(expect)))


c1 = ((0.0, 0.0, 0.0), (1.0, 0.0, 0.0), 1.0)
c2 = ((0.0, 0.0, 0.0), (0.0, 2.0, 0.0), 2.0)
case("side_from_right", "collision.rs:1544-1557", c1, (1.0, -3.0, 0.0), (-0.25, 1.0, 0.0), True,
     [dict(field="p", mode="rel", val=[0.5, -1.0, 0.0], eps=E), dict(field="along", mode="rel", val=[0.5, -1.0, 0.0], eps=E)])
case("side_from_left", "collision.rs:1558-1571", c1, (0.0, -3.0, 0.0), (0.25, 1.0, 0.0), True,
     [dict(field="p", mode="rel", val=[0.5, -1.0, 0.0], eps=E), dict(field="along", mode="rel", val=[0.5, -1.0, 0.0], eps=E)])
case("vertical_capsule_side", "collision.rs:1572-1583", c2, (4.0, 1.0, 0.0), (-1.0, 0.0, 0.0), False,
     [dict(field="p", mode="eq", val=[2.0, 1.0, 0.0]), dict(field="t", mode="eq", val=2.0)])
case("axis_far_cap", "collision.rs:1584-1595", c1, (3.0, 0.0, 0.0), (-1.0, 0.0, 0.0), False,
     [dict(field="p", mode="eq", val=[2.0, 0.0, 0.0]), dict(field="t", mode="eq", val=1.0)])
case("axis_near_cap", "collision.rs:1596-1607", c1, (-2.0, 0.0, 0.0), (1.0, 0.0, 0.0), False,
     [dict(field="p", mode="eq", val=[-1.0, 0.0, 0.0]), dict(field="t", mode="eq", val=1.0)])
case("offset_near_cap", "collision.rs:1608-1624", c1, (-2.0, 0.5, 0.0), (1.0, 0.0, 0.0), False,
     [dict(field="p", mode="rel", val=[-0.8660254037844386, 0.5, 0.0], eps=FE), dict(field="t", mode="rel", val=1.13397459621556196, eps=E)])
case("offset_far_cap", "collision.rs:1625-1636", c1, (3.0, 0.5, 0.0), (-1.0, 0.0, 0.0), False,
     [dict(field="p", mode="rel", val=[1.8660254037844386, 0.5, 0.0], eps=FE), dict(field="t", mode="rel", val=1.13397459621556196, eps=E)])

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_ray_kat.json")
with open(out, "w") as f:
    json.dump(dict(source="maplant/mgf src/collision.rs test_ray_intersections (1543-1637)", cases=cases), f, indent=1)
print(f"wrote {len(cases)} cases to {out}")
