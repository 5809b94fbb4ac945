"""CPU-only: the oracle's Compound (oracle/compound.hpp, src/compound.rs:232-352) against the reference's own test
(compound.rs:362-388) and hand-derived answers.  The same literals pin the CUDA path in test_gpu_compound.py."""
import numpy as np

import oracle_lib
from compound_cases import from_arc_x_to_y, random_compound, random_rhs, reference_test_compound
from mgf_b200 import _lib as L
from mgf_b200 import api


def test_reference_test_compound():
    c = oracle_lib.OracleCompound(reference_test_compound())
    test_sphere = api.sphere((0.0, 8.0, 0.0), 1.0, v=(0.0, -1.5, 0.0))
    out, counts = c.contacts(test_sphere)
    assert counts[0] == 0                                   # compound.rs:371: no contact while the spheres sit at x = -5, +5
    c.set_transform((0, 0, 0), from_arc_x_to_y())           # rotated: one sphere is now at (0, 5, 0)
    out, counts = c.contacts(test_sphere)
    assert counts[0] >= 1
    last = out[0, counts[0] - 1]                            # last_contact (collision.rs:477-481)
    assert abs(last["t"] - 0.6666663) <= 1e-6 * max(1.0, abs(last["t"]))          # assert_relative_eq!, COLLISION_EPSILON
    assert np.allclose(last["a"], [0.0, 6.0, 0.0], atol=1e-6, rtol=1e-6)
    c.set_transform((0, 0, 0), (1, 0, 0, 0))
    rect = api.rectangle((0.0, -2.0, 0.0), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), 6.0, 6.0)
    rect["v"][0] = (0.0, 3.0, 0.0)
    out, counts = c.contacts(rect)
    assert counts[0] >= 1                                    # compound.rs:386: .unwrap() must find a contact


def test_bounds_closest_point_and_rays_hand_derived():
    c = oracle_lib.OracleCompound(reference_test_compound())
    aabb, sph = c.bounds()
    assert np.array_equal(aabb, np.array([0, 0, 0, 6, 1, 1], np.float32))         # union of the two leaf boxes
    assert np.allclose(sph, [0, 0, 0, np.sqrt(38.0)])                             # bounds.rs:291-298: radius = |r|
    c.set_transform((1, 2, 3), (1, 0, 0, 0))
    aabb, sph = c.bounds()
    assert np.array_equal(aabb, np.array([1, 2, 3, 6, 1, 1], np.float32))
    # closest_point ignores disp / rot (compound.rs:299-311) and uses Sphere::closest_point's |d|^2 / r^2 ratio (geom.rs:751-755)
    p = c.closest_points([[7.0, 0.0, 0.0]])[0]
    assert np.allclose(p, [5 + 2 * 4, 0, 0])
    # a ray along +x from the far left hits the left sphere first (disp = (1, 2, 3))
    c.set_transform((0, 0, 0), (1, 0, 0, 0))
    out, hit = c.intersections(L.RAY, [[-10.0, 0.0, 0.0, 1.0, 0.0, 0.0]])
    assert hit[0] == 1 and np.allclose(out[0]["p"], [-6, 0, 0]) and np.isclose(out[0]["t"], 4.0)
    out, hit = c.intersections(L.SEGMENT, [[-10.0, 0.0, 0.0, -7.0, 0.0, 0.0]])    # too short
    assert hit[0] == 0


def test_random_compounds_are_self_consistent():
    """Every contact the compound reports is also reported by the component it came from, queried alone (the tree only decides
    WHICH components are tried and in what order)."""
    rng = np.random.default_rng(5)
    comps = random_compound(rng, 7)
    whole = oracle_lib.OracleCompound(comps)
    singles = [oracle_lib.OracleCompound(comps[i:i + 1]) for i in range(len(comps))]
    rhs = random_rhs(rng, 400)
    out, counts = whole.contacts(rhs)
    assert counts.sum() > 50
    per = [s.contacts(rhs, slots=out.shape[1]) for s in singles]
    for i in range(len(rhs)):
        got = [out[i, k].tobytes() for k in range(counts[i])]
        pool = [o[i, k].tobytes() for o, cnt in per for k in range(cnt[i])]
        assert sorted(got) == sorted(pool)
