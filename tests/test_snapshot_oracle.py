"""CPU-only: the oracle's checkpoint (mgfo_world_get_fat_bounds / mgfo_world_set_state) restores World::step state exactly.
A restored world's body BVH is a fresh tree, so the reference's DFS callback ORDER differs from the original world's; the
constraint SET must not, and replaying one common order must give bit-identical state (same argument as DESIGN.md section 2)."""
import numpy as np

import oracle_lib
from mgf_b200 import scenes


def _perm_to(order_of, a, b, face, sub):
    index = {k: i for i, k in enumerate(zip(a.tolist(), b.tolist(), face.tolist(), sub.tolist()))}
    return np.array([index[k] for k in order_of], dtype=np.uint32)


def test_oracle_snapshot_restore_continues_bit_identically():
    bodies, terrain, iters = scenes.balls_scene(6, 0, 0.05, 3), scenes.box_terrain(), 10   # jittered: columns topple, bodies spin
    dt = np.float32(1 / 60)
    a = oracle_lib.OracleWorld(); a.add_bodies(*bodies); a.set_terrain(*terrain)
    a.step(dt, iters, 200)                      # through the first impacts: fat boxes have been refreshed, bodies rotate
    snap = a.snapshot()
    assert np.abs(snap["omega"]).max() > 0 and not np.array_equal(snap["fat"][:, :3], snap["x"])
    b = oracle_lib.OracleWorld(); b.add_bodies(*bodies); b.set_terrain(*terrain); b.restore(snap)
    for x, y in zip(a.snapshot().values(), b.snapshot().values()):
        assert np.array_equal(np.ascontiguousarray(x).view(np.uint8), np.ascontiguousarray(y).view(np.uint8))
    assert np.array_equal(a.inv_moment().view(np.uint32), b.inv_moment().view(np.uint32))
    total = 0
    for _ in range(6):
        ma, mb = a.build(dt), b.build(dt)
        assert ma == mb
        ka = list(zip(*[v.tolist() for v in a.constraints(ma)]))
        assert sorted(ka) == sorted(zip(*[v.tolist() for v in b.constraints(mb)]))
        a.solve_order(np.arange(ma, dtype=np.uint32), iters)
        b.solve_order(_perm_to(ka, *b.constraints(mb)), iters)
        for x, y in zip(a.state(), b.state()):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
        assert np.array_equal(a.fat_bounds().view(np.uint32), b.fat_bounds().view(np.uint32))
        total += ma
    assert total > 0
