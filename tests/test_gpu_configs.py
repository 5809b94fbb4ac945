"""BASELINE.json's configs at FULL size on the GPU, bit for bit against the oracle replaying the
device's constraint order (same method as test_gpu_step.py, minus the O(n^2) brute-force pair
count): C2 = 100 000 spheres in a box, C3 = 50 000 capsules over a 20 000-triangle mesh floor,
C5 = 200 000 mixed sphere/capsule bodies over a 200 000-triangle mesh (its GJK half is
test_gpu_gjk.py).  C4 (2 M spheres over 8 GPUs) is the tiled path: test_gpu_tiled*.py + bench.py."""
import numpy as np
import pytest

import mgf_b200
import oracle_lib
from mgf_b200 import scenes
from test_gpu_step import DT, _assert_state_equal, _pair

pytestmark = pytest.mark.gpu


def _lockstep_full(g, o, iters, nsteps, what):
    total = 0
    for s in range(nsteps):
        st = g.step(DT, iters)
        m = o.build(DT)
        assert st["constraints"] == m, f"{what} step {s}: {st['constraints']} constraints vs oracle {m}"
        cand, tcand = o.stats()
        assert cand <= st["candidate_pairs"], (cand, st["candidate_pairs"])   # the tree may prune touching leaves, never a contact
        assert tcand == st["terrain_candidates"], (tcand, st["terrain_candidates"])
        ga, gb, gf, gs, gc = g.constraints()
        oa, ob, of, osub = o.constraints(m)
        key = lambda a, b, f, s_: (a.astype(np.uint64) << np.uint64(42)) | ((b.astype(np.int64) + 1).astype(np.uint64) << np.uint64(21)) | \
            (f.astype(np.uint64) << np.uint64(1)) | s_.astype(np.uint64)
        # (a < 2^21, b + 1 < 2^21, face < 2^20 at these sizes)
        ok_, gk = key(oa, ob, of, osub), key(ga, gb, gf, gs)
        order = np.argsort(ok_, kind="stable")
        pos = np.searchsorted(ok_[order], gk)
        assert np.all(pos < m) and np.array_equal(ok_[order][pos], gk), f"{what} step {s}: constraint sets differ"
        perm = order[pos].astype(np.uint32)
        assert len(np.unique(perm)) == m
        assert np.all(np.diff(gc.astype(np.int64)) >= 0), "rows are not colour-major"
        o.solve_order(perm, iters)
        total += m
        _assert_state_equal(g, o, f"{what} step {s}")
    return total


def test_c2_100k_spheres_full_size():
    bodies, terrain, iters = scenes.build_config("C2pile")
    g, o = _pair(bodies, terrain)
    total = _lockstep_full(g, o, iters, 3, "C2 100k spheres")
    assert total > 800000


def test_c2_settled_window_full_size_from_snapshot():
    """BASELINE.md's C2 as written: 100 000 spheres dropped from the balls.rs lattice into the 160 x 160 box; the window
    the benchmark times starts at step 600 (the disordered pile).  The GPU runs the 600 steps, its state is loaded into the
    oracle (snapshot = x, q, v, omega, colliders, stored fat boxes) and the two are lock-stepped from there."""
    bodies, terrain, iters = scenes.build_config("C2")
    g, o = _pair(bodies, terrain)
    g.step(DT, iters, nsteps=600)
    snap = g.snapshot()
    assert np.isfinite(snap["x"]).all() and snap["x"][:, 1].min() > -10.6
    o.restore(snap)
    total = _lockstep_full(g, o, iters, 3, "C2 settled window (steps 600..602)")
    assert total > 300000, total


def test_c3_50k_capsules_on_20k_triangle_mesh_full_size():
    bodies, terrain, iters = scenes.config_c3()
    assert len(bodies[0]) == 50000 and len(terrain[1]) == 20000
    g, o = _pair(bodies, terrain)
    total = _lockstep_full(g, o, iters, 3, "C3 50k capsules / 20k triangles")
    assert total > 300000


def test_c5_200k_mixed_bodies_on_200k_triangle_mesh_full_size():
    bodies, terrain, iters = scenes.config_c5()
    assert len(bodies[0]) == 200000 and len(terrain[1]) == 199712
    g, o = _pair(bodies, terrain)
    total = _lockstep_full(g, o, iters, 2, "C5 200k mixed / 200k triangles")
    assert total > 1000000


def test_c2_long_run_stays_finite_and_deterministic():
    """100 steps of C2 (the north star's horizon) twice: identical bits run to run (the colouring's
    priorities depend only on constraint identity, never on list order or atomics), all finite, and
    nothing has fallen through the floor."""
    bodies, terrain, iters = scenes.build_config("C2pile")
    runs = []
    for _ in range(2):
        g = mgf_b200.World(device=0)
        g.add_bodies(*bodies); g.set_terrain(*terrain)
        g.step(DT, iters, nsteps=100)
        runs.append(g.state())
        g.ctx.close()
    for a, b in zip(*runs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    x = runs[0][0]
    assert np.isfinite(x).all() and x[:, 1].min() > -10.6
