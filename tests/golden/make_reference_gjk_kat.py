"""Writes tests/golden/reference_gjk_kat.json: the reference's OWN known answers for the discrete
GJK/EPA path, as literals -- src/collision.rs `mod tests`: spheres::test_sphere_penetration
(:1646-1672, Penetrates::separation) and obbs::test_obb_collision (:1822-1843, Contacts for
Convex x Convex -> GJK + EPA).  Nothing is computed here except box4's quaternion, which the Rust
test builds with Quaternion::from_arc(x, y): cgmath 0.17 gives from_sv(1 + 0, x cross y).normalize()
= (1, 0, 0, 1) * (1 / sqrt(2)) in f32."""
import json
import os

import numpy as np

SPH, AABB, OBB = 0, 5, 6
inv_sqrt2 = float(np.float32(1.0) * (np.float32(1.0) / np.sqrt(np.float32(2.0))))


def sphere(c, r): return dict(kind=SPH, p=[*c, r])
def obb(c, r, q=(1.0, 0.0, 0.0, 0.0)): return dict(kind=OBB, p=[*c, *r, *q])


box1 = obb((0, 0, 0), (1, 1, 1))
box2 = obb((0, 1, 0), (1, 1.5, 1))
box3 = obb((0, 4.1, 0), (1, 1.5, 1))
box4 = obb((0, 2, 0), (1.7, 1.5, 1), (inv_sqrt2, 0.0, 0.0, inv_sqrt2))
s1 = sphere((0, 0, 0), 1.0)
cases = [
    # op "contacts": a.last_contact(&b); expect: hit, and assert_eq! on single components
    dict(name="obb_obb_stacked", src="collision.rs:1824-1829", op="contacts", a=box1, b=box2, hit=1, expect=[["a", 1, 1.0], ["b", 1, -0.5]]),
    dict(name="obb_obb_stacked_commuted", src="collision.rs:1830-1832", op="contacts", a=box2, b=box1, hit=1, expect=[["b", 1, 1.0], ["a", 1, -0.5]]),
    dict(name="obb_obb_apart", src="collision.rs:1833-1834", op="contacts", a=box1, b=box3, hit=0, expect=[]),
    dict(name="obb_obb_rotated", src="collision.rs:1835-1842", op="contacts", a=box1, b=box4, hit=1, expect=[["a", 1, 1.0], ["b", 1, 0.30000007]]),
    # op "separation": a.separation(&b) -> None / Some(d)
    dict(name="sphere_sphere_overlap", src="collision.rs:1656-1659", op="separation", a=s1, b=sphere((2, 0, 0), 1.5), some=0, val=None),
    dict(name="sphere_sphere_overlap_commuted", src="collision.rs:1660-1663", op="separation", a=sphere((2, 0, 0), 1.5), b=s1, some=0, val=None),
    dict(name="sphere_sphere_apart", src="collision.rs:1668-1671", op="separation", a=s1, b=sphere((2, 0, 0), 0.75), some=1, val=0.25),
]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_gjk_kat.json")
json.dump(dict(source="maplant/mgf v1.4.0 src/collision.rs mod tests (spheres::test_sphere_penetration, obbs::test_obb_collision)", cases=cases),
          open(out, "w"), indent=1)
print(out, len(cases))
