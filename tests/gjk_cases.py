"""Shared by the CPU and GPU GJK tests: the reference's golden vectors and the seeded random pairs
of SURVEY.md 8(d) C5 (centres U(-1,1)^3 * 2, sizes U(0.3,1), unit quaternions, LCG seed 7)."""
import json
import os

import numpy as np

from mgf_b200 import _lib as L
from mgf_b200 import scenes

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    return json.load(open(os.path.join(HERE, "golden", "reference_gjk_kat.json")))["cases"]


def to_shape(d):
    s = np.zeros(1, dtype=L.SHAPE_DTYPE)
    s["kind"] = d["kind"]
    s["p"][0, :len(d["p"])] = np.asarray(d["p"], dtype=np.float32)
    return s


def check_golden(run_contacts, run_separation):
    """run_contacts(a, b) -> (contacts, status, iters); run_separation(a, b) -> (sep, some)."""
    n = 0
    for c in golden():
        a, b = to_shape(c["a"]), to_shape(c["b"])
        if c["op"] == "contacts":
            out, status, _ = run_contacts(a, b)
            assert int(status[0]) == c["hit"], (c["name"], status)
            for field, comp, val in c["expect"]:
                got = out[0][field][comp]
                assert np.float32(got) == np.float32(val), f"{c['name']} ({c['src']}): {field}[{comp}] = {got!r}, reference asserts {val!r}"
                n += 1
        else:
            sep, some = run_separation(a, b)
            assert int(some[0]) == c["some"], (c["name"], some)
            if c["some"]:
                assert np.float32(sep[0]) == np.float32(c["val"]), (c["name"], sep)
            n += 1
    return n


def random_shapes(n, kind, seed):
    u = scenes.lcg_uniform_fast(16 * n, seed).reshape(n, 16)
    s = np.zeros(n, dtype=L.SHAPE_DTYPE)
    s["kind"] = kind
    c = (u[:, 0:3] * np.float32(2) - np.float32(1)) * np.float32(2)
    size = np.float32(0.3) + u[:, 3:6] * np.float32(0.7)
    q = u[:, 6:10] * np.float32(2) - np.float32(1)
    q = (q / np.sqrt((q.astype(np.float64) ** 2).sum(axis=1, keepdims=True))).astype(np.float32)
    s["p"][:, 0:3] = c
    if kind == L.SPHERE:
        s["p"][:, 3] = size[:, 0]
    elif kind == L.CAPSULE:
        d = (u[:, 10:13] * np.float32(2) - np.float32(1)) * size[:, 1:2]
        s["p"][:, 0:3] = c - d * np.float32(0.5)
        s["p"][:, 3:6] = d
        s["p"][:, 6] = size[:, 0] * np.float32(0.6)
    elif kind == L.AABB:
        s["p"][:, 3:6] = size
    else:
        s["p"][:, 3:6] = size
        s["p"][:, 6:10] = q
    return s


KINDS = [L.SPHERE, L.CAPSULE, L.AABB, L.OBB]


def mixed_pairs(n, seed=7):
    """n pairs cycling through all 16 (kind_a, kind_b) combinations."""
    a = np.zeros(n, dtype=L.SHAPE_DTYPE); b = np.zeros(n, dtype=L.SHAPE_DTYPE)
    for k in range(16):
        idx = np.arange(k, n, 16)
        if len(idx) == 0:
            continue
        a[idx] = random_shapes(len(idx), KINDS[k // 4], seed + 2 * k)
        b[idx] = random_shapes(len(idx), KINDS[k % 4], seed + 2 * k + 1)
    return a, b


def convex_mesh_pool(n_meshes, seed=31):
    """n_meshes convex point soups (mesh.rs:141 ConvexMesh): box corners (8 vertices), tetrahedra and 12..40 random points on an
    ellipsoid, each rotated and placed like the random shapes above.  Returns (vertex pool (V, 3) f32, shapes[n_meshes])."""
    rng = np.random.default_rng(seed)
    verts = []; shapes = np.zeros(n_meshes, dtype=L.SHAPE_DTYPE)
    shapes["kind"] = L.CONVEX_MESH
    first = 0
    for i in range(n_meshes):
        c = rng.uniform(-2, 2, 3); size = rng.uniform(0.3, 1.0, 3)
        kind = i % 3
        if kind == 0:
            pts = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float64) * size
        elif kind == 1:
            pts = rng.normal(size=(4, 3)) * size
        else:
            p = rng.normal(size=(int(rng.integers(12, 41)), 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
            pts = p * size
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        pts = (pts @ R.T + c).astype(np.float32)
        verts.append(pts)
        shapes["p"][i, 0] = first; shapes["p"][i, 1] = len(pts)
        first += len(pts)
    return np.concatenate(verts).astype(np.float32), shapes


def mesh_pairs(n, seed=7):
    """n pairs: ConvexMesh against each of the five kinds (both orders), cycling.  Returns (pool, a, b)."""
    pool, meshes = convex_mesh_pool(2 * n, seed + 100)
    a = np.zeros(n, dtype=L.SHAPE_DTYPE); b = np.zeros(n, dtype=L.SHAPE_DTYPE)
    for k in range(9):
        idx = np.arange(k, n, 9)
        if len(idx) == 0:
            continue
        if k < 4:        # mesh x plain
            a[idx] = meshes[idx]; b[idx] = random_shapes(len(idx), KINDS[k], seed + 40 + k)
        elif k < 8:      # plain x mesh
            a[idx] = random_shapes(len(idx), KINDS[k - 4], seed + 50 + k); b[idx] = meshes[n + idx]
        else:            # mesh x mesh
            a[idx] = meshes[idx]; b[idx] = meshes[n + idx]
    return pool, a, b
