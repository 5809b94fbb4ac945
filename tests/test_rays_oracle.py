"""CPU: the oracle's ray casts (Intersects<RHS> for Ray / Segment, collision.rs:163-373) against
the reference's own test vectors (tests/golden/reference_ray_kat.json = collision.rs:1543-1637),
plus hand-derived checks for the shapes no reference test covers (labelled OURS: for those the
oracle's authority is line-by-line fidelity to the cited source)."""
import numpy as np

import mgf_b200
import oracle_lib
import ray_cases
from mgf_b200 import _lib as L


def test_oracle_matches_reference_ray_vectors():
    fails = ray_cases.check_kat(oracle_lib.intersections_batch)
    assert not fails, "\n".join(fails)


def test_oracle_ray_hand_derived_ours():
    f = np.float32
    rays = np.array([[0, 5, 0, 0, -1, 0],      # down onto the plane y = 1
                     [0, 5, 0, 0, 1, 0],       # away from it: t <= 0 -> None
                     [0, 5, 0, 0, -2, 0],      # AABB top face at y = 1, |d| = 2 -> t = 2
                     [5, 0, 0, -1, 0, 0],      # sphere r = 1 at origin -> t = 4
                     [0.25, 5, 0.25, 0, -1, 0],  # triangle in the plane y = 0 (inside)
                     [3, 5, 3, 0, -1, 0]], f)    # same plane, outside the triangle
    shapes = np.concatenate([mgf_b200.plane((0, 1, 0), 1.0), mgf_b200.plane((0, 1, 0), 1.0), mgf_b200.aabb((0, 0, 0), (1, 1, 1)),
                             mgf_b200.sphere((0, 0, 0), 1.0), mgf_b200.triangle((0, 0, 0), (0, 0, 1), (1, 0, 0)),
                             mgf_b200.triangle((0, 0, 0), (0, 0, 1), (1, 0, 0))])
    out, hit = oracle_lib.intersections_batch(L.RAY, rays, shapes)
    assert hit.tolist() == [1, 0, 1, 1, 1, 0]
    assert out["t"][[0, 2, 3, 4]].tolist() == [4.0, 2.0, 4.0, 5.0]
    assert out["p"][0].tolist() == [0.0, 1.0, 0.0] and out["p"][3].tolist() == [1.0, 0.0, 0.0]
    # Segment: DT = 1 (geom.rs:843): the same geometry given as a -> b hits only when b reaches the shape
    segs = np.array([[0, 5, 0, 0, 2, 0], [0, 5, 0, 0, 0, 0]], f)    # stops at y = 2 (short of the plane) / reaches y = 0
    out, hit = oracle_lib.intersections_batch(L.SEGMENT, segs, np.concatenate([mgf_b200.plane((0, 1, 0), 1.0)] * 2))
    assert hit.tolist() == [0, 1] and out["t"][1] == f(0.8)
    # Moving<Sphere> = the capsule its sweep covers (collision.rs:361-373)
    ms = mgf_b200.sphere((0, 0, 0), 1.0); ms["v"][0] = (1, 0, 0)
    a, ha = oracle_lib.intersections_batch(L.RAY, np.array([[3, 0, 0, -1, 0, 0]], f), ms)
    b, hb = oracle_lib.intersections_batch(L.RAY, np.array([[3, 0, 0, -1, 0, 0]], f), mgf_b200.capsule((0, 0, 0), (1, 0, 0), 1.0))
    assert ha[0] == hb[0] == 1 and a["t"][0] == b["t"][0] == f(1.0)
