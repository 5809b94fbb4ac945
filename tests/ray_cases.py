"""Ray-cast parity cases shared by the CPU (oracle vs the reference's golden vectors) and GPU
(CUDA vs oracle, bit for bit) tests of mgfb_intersections_batch."""
import json
import os

import numpy as np

import kat_check
import mgf_b200
from mgf_b200 import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))


def load_ray_kat():
    with open(os.path.join(HERE, "golden", "reference_ray_kat.json")) as f:
        return json.load(f)["cases"]


def kat_inputs():
    """(particles[n,6], shapes[n], cases): the reference test's rays (normalised like cgmath: v * (1/|v|), in f32)."""
    cases = load_ray_kat()
    parts = np.zeros((len(cases), 6), np.float32)
    shapes = np.zeros(len(cases), dtype=L.SHAPE_DTYPE)
    for i, c in enumerate(cases):
        d = np.asarray(c["d"], np.float32)
        if c["normalize"]:
            mag = np.sqrt(np.float32(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2])))
            d = d * np.float32(np.float32(1.0) / mag)
        parts[i, :3] = c["p"]; parts[i, 3:] = d
        shapes[i] = mgf_b200.capsule(c["capsule"]["a"], c["capsule"]["d"], c["capsule"]["r"])[0]
    return parts, shapes, cases


def check_kat(batch_fn):
    """batch_fn(kind, particles, shapes) -> (out, hit); returns failure strings."""
    parts, shapes, cases = kat_inputs()
    out, hit = batch_fn(L.RAY, parts, shapes)
    fails = []
    for i, c in enumerate(cases):
        if not hit[i]:
            fails.append(f'{c["name"]} ({c["src"]}): no intersection'); continue
        for e in c["expect"]:
            if e["field"] == "along":   # r.p + r.d * t, evaluated in f32 like the test
                got = parts[i, :3] + parts[i, 3:] * np.float32(out[i]["t"])
            else:
                got = np.atleast_1d(out[i][e["field"]])
            want = np.atleast_1d(np.asarray(e["val"], np.float32))
            ok = bool(np.all(got == want)) if e["mode"] == "eq" else all(kat_check.relative_eq(g, w, e["eps"]) for g, w in zip(got, want))
            if not ok:
                fails.append(f'{c["name"]} ({c["src"]}): {e["field"]} = {np.asarray(got).tolist()} want {want.tolist()} ({e["mode"]})')
    return fails


def random_queries(n, seed):
    """Seeded mixed batch: every shape kind the API takes, rays/segments aimed roughly at the shapes
    (about half hit), plus degenerate directions (axis-parallel, zero components) for the AABB slab branches."""
    rng = np.random.default_rng(seed)
    shapes = np.zeros(n, dtype=L.SHAPE_DTYPE)
    kinds = rng.integers(0, 8, n)   # 0 plane 1 tri 2 rect 3 aabb 4 obb 5 sphere 6 capsule 7 moving sphere
    centre = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    for i in range(n):
        k = kinds[i]; c = centre[i]
        if k == 0:
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            shapes[i] = mgf_b200.plane(nrm.astype(np.float32), np.float32(np.dot(nrm, c)))[0]
        elif k == 1:
            shapes[i] = mgf_b200.triangle(c + rng.uniform(-1, 1, 3), c + rng.uniform(-1, 1, 3), c + rng.uniform(-1, 1, 3))[0]
        elif k == 2:
            u0 = rng.normal(size=3); u0 /= np.linalg.norm(u0)
            u1 = np.cross(u0, rng.normal(size=3)); u1 /= np.linalg.norm(u1)
            shapes[i] = mgf_b200.rectangle(c, u0, u1, rng.uniform(0.3, 1.5), rng.uniform(0.3, 1.5))[0]
        elif k == 3:
            shapes[i] = mgf_b200.aabb(c, rng.uniform(0.2, 1.2, 3))[0]
        elif k == 4:
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            shapes[i] = mgf_b200.obb(c, rng.uniform(0.2, 1.2, 3), q)[0]
        elif k == 5:
            shapes[i] = mgf_b200.sphere(c, rng.uniform(0.2, 1.2))[0]
        elif k == 6:
            shapes[i] = mgf_b200.capsule(c, rng.uniform(-1, 1, 3), rng.uniform(0.2, 0.8))[0]
        else:
            s = mgf_b200.sphere(c, rng.uniform(0.2, 0.8)); s["v"][0] = rng.uniform(-1, 1, 3); shapes[i] = s[0]
    origin = (centre + rng.normal(size=(n, 3)) * 3.0).astype(np.float32)
    target = (centre + rng.normal(size=(n, 3)) * 0.8).astype(np.float32)
    direction = (target - origin).astype(np.float32)
    deg = rng.random(n) < 0.15                 # axis-parallel / sparse directions
    direction[deg] *= (rng.random((int(deg.sum()), 3)) < 0.5)
    inside = rng.random(n) < 0.05              # origins inside the shape
    origin[inside] = centre[inside]
    rays = np.concatenate([origin, direction], axis=1).astype(np.float32)
    segs = np.concatenate([origin, origin + direction * rng.uniform(0.3, 1.6, (n, 1)).astype(np.float32)], axis=1).astype(np.float32)
    return rays, segs, shapes
