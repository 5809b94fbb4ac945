"""Shared inputs of the Compound tests (src/compound.rs:232-352): the reference's own test (compound.rs:362-388) as literals,
plus seeded random compounds, transforms and queries."""
import numpy as np

from mgf_b200 import _lib as L
from mgf_b200 import api


def reference_test_compound():
    """compound.rs:364-368: two unit spheres at (-5, 0, 0) and (5, 0, 0)."""
    return np.concatenate([api.sphere((-5.0, 0.0, 0.0), 1.0), api.sphere((5.0, 0.0, 0.0), 1.0)])


def from_arc_x_to_y():
    """Quaternion::from_arc(x, y, None).normalize() (compound.rs:373-375), evaluated in f32 like cgmath: (mag_avg + dot, src x dst) normalised."""
    q = np.array([1.0, 0.0, 0.0, 1.0], np.float32)            # s = 1 + 0, v = x cross y = z
    inv = np.float32(1.0) / np.sqrt(np.float32(q[0] * q[0] + q[3] * q[3]))
    q = (q * inv).astype(np.float32)
    inv2 = np.float32(1.0) / np.sqrt(np.float32(q[0] * q[0]) + np.float32(np.float32(q[1] * q[1] + q[2] * q[2]) + q[3] * q[3]))   # .normalize() again: s*s + v.v
    return (q * inv2).astype(np.float32)


def random_quats(rng, n):
    q = rng.normal(size=(n, 4)).astype(np.float32)
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


def random_compound(rng, ncomp, spread=3.0):
    comps = np.zeros(ncomp, dtype=L.SHAPE_DTYPE)
    for i in range(ncomp):
        if rng.random() < 0.5:
            comps[i] = api.sphere(rng.uniform(-spread, spread, 3), rng.uniform(0.3, 1.0))[0]
        else:
            comps[i] = api.capsule(rng.uniform(-spread, spread, 3), rng.uniform(-1.5, 1.5, 3), rng.uniform(0.2, 0.8))[0]
    return comps


def random_rhs(rng, n, reach=6.0):
    """Moving<Sphere | Capsule | Triangle | Rectangle> aimed roughly at the origin."""
    out = np.zeros(n, dtype=L.SHAPE_DTYPE)
    for i in range(n):
        k = i % 4
        p = rng.uniform(-reach, reach, 3)
        v = (-p * rng.uniform(0.2, 1.2) + rng.normal(size=3) * 0.5)
        if k == 0:
            out[i] = api.sphere(p, rng.uniform(0.3, 1.0), v)[0]
        elif k == 1:
            out[i] = api.capsule(p, rng.uniform(-1.5, 1.5, 3), rng.uniform(0.2, 0.8), v)[0]
        elif k == 2:
            out[i] = api.triangle(p + rng.uniform(-2, 2, 3), p + rng.uniform(-2, 2, 3), p + rng.uniform(-2, 2, 3), v)[0]
        else:
            u0 = rng.normal(size=3); u0 /= np.linalg.norm(u0)
            u1 = np.cross(u0, rng.normal(size=3)); u1 /= np.linalg.norm(u1)
            s = api.rectangle(p, u0.astype(np.float32), u1.astype(np.float32), rng.uniform(0.5, 3.0), rng.uniform(0.5, 3.0))
            s["v"][0] = v.astype(np.float32)
            out[i] = s[0]
    return out


def random_particles(rng, n, reach=8.0):
    p = rng.uniform(-reach, reach, (n, 3)).astype(np.float32)
    target = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    rays = np.concatenate([p, (target - p) * rng.uniform(0.3, 2.0, (n, 1)).astype(np.float32)], axis=1).astype(np.float32)   # Ray{p, d}
    segs = np.concatenate([p, target + (target - p) * rng.uniform(-0.5, 0.5, (n, 1)).astype(np.float32)], axis=1).astype(np.float32)   # Segment{a, b}
    return rays, segs
