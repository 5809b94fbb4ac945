"""CPU: the oracle's GJK / EPA against the reference's own golden vectors (collision.rs tests)."""
import numpy as np

import gjk_cases
import oracle_lib


def test_oracle_gjk_matches_reference_vectors(oracle):
    assert gjk_cases.check_golden(oracle_lib.gjk_batch, oracle_lib.separation_batch) == 9


def test_oracle_gjk_random_pairs_sane(oracle):
    a, b = gjk_cases.mixed_pairs(640)
    out, hit, iters = oracle_lib.gjk_batch(a, b)
    sep, some = oracle_lib.separation_batch(a, b)
    assert set(np.unique(hit).tolist()) <= {0, 1, 3, 4} and set(np.unique(some).tolist()) <= {0, 1, 3}
    assert (hit == 1).sum() > 50 and (hit == 0).sum() > 50 and (some == 1).sum() > 50
    # NOTE: no cross-check between the two entry points is asserted.  The reference's GJK is seeded
    # differently for contacts (+-y) and separation (+-x), treats a flat tetrahedron as "origin inside"
    # (sign_p * sign_d < 0 is false when sign_d == 0, simplex.rs:342-349) and can cycle forever on
    # separated polytopes (exit test |min|^2 >= |support|^2, simplex.rs:195); the port reproduces all of
    # that -- status 3 marks the pairs where the reference would never return.
    n = out["n"][hit == 1]
    assert np.isfinite(n).all()
    assert iters.max() <= 100


def test_oracle_convex_mesh_support_and_gjk(oracle):
    """ConvexMesh (mesh.rs:141-236) has no test in the reference: OUR sanity checks of the restatement.  (The reference's GJK
    is itself inexact -- on random sphere / box pairs it agrees with the analytic answer ~85 % of the time, for the AABB
    support as much as for the mesh support -- so only unambiguous configurations are asserted.)"""
    from mgf_b200 import api
    corners = np.array([[sx * 1.0, sy * 0.5, sz * 0.25] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) + np.float32([0.5, 0, 0])
    corners += np.random.default_rng(1).uniform(-0.03, 0.03, corners.shape).astype(np.float32)   # no two vertices tie along the +-y seed
    tet = (np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], np.float64) @ np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]])).astype(np.float32)   # regular, rotated off the axes
    pool = np.concatenate([corners, tet])
    oracle_lib.convex_vertices_set(pool)
    box_mesh, tet_mesh = api.convex_mesh(0, 8), api.convex_mesh(8, 4)
    rng = np.random.default_rng(3)
    n = 200
    rep = lambda s: np.repeat(s, n)
    inside = np.concatenate([api.sphere(np.float32([0.5, 0, 0]) + rng.uniform(-0.5, 0.5, 3) * np.float32([1, 0.5, 0.25]), rng.uniform(0.3, 0.6)) for _ in range(n)])
    far = np.concatenate([api.sphere(np.float32([0.5, 0, 0]) + d / np.linalg.norm(d) * rng.uniform(4, 6), rng.uniform(0.2, 1.0))
                          for d in rng.normal(size=(n, 3))])
    hit_in = oracle_lib.gjk_batch(rep(box_mesh), inside)[1]; hit_far = oracle_lib.gjk_batch(rep(box_mesh), far)[1]
    assert (hit_in == 1).mean() > 0.9 and (hit_far == 1).sum() == 0, ((hit_in == 1).mean(), (hit_far == 1).sum())
    sep, some = oracle_lib.separation_batch(rep(box_mesh), far)
    assert (some[some <= 1] == 1).all() and (sep[some == 1] > 1.5).all()
    # a regular tetrahedron and a sphere out along (vertex 0 - centroid): the closest point is that vertex
    v0 = tet[0].astype(np.float64); cen = tet.astype(np.float64).mean(axis=0)
    c = v0 + (v0 - cen) / np.linalg.norm(v0 - cen) * 3.0
    sep, some = oracle_lib.separation_batch(tet_mesh, api.sphere(c, 0.5))
    assert some[0] == 1 and sep[0] > 1.0, (some, sep)   # (the reference's exit test |min|^2 >= |support|^2, simplex.rs:195, stops early on polytopes: the distance itself is not asserted)
    out, hit, _ = oracle_lib.gjk_batch(tet_mesh, api.sphere(cen, 0.3))
    assert hit[0] == 1 and np.isfinite(out[0]["n"]).all()
    # a mesh slice outside the pool is refused (no contact), like `verts[0]` on an empty mesh would panic in the reference
    assert oracle_lib.gjk_batch(api.convex_mesh(10, 40), api.sphere((0, 0, 0), 1.0))[1][0] == 0
