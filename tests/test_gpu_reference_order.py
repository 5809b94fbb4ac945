"""GPU: World::step in the REFERENCE's own constraint order (mgfb_config.step_order = MGFB_STEP_ORDER_REFERENCE, csrc/reforder.cuh)
against the oracle's native World::step (world.rs:227-294: bodies ascending, terrain contacts in mesh-BVH callback order, body
pairs in body-BVH callback order) -- nothing is exported or replayed: both worlds simply step, and the full state must agree BIT
FOR BIT after every step.  This is the north star's "matches the reference CPU path after N steps" with tolerance 0, through
impacts and settling, where the coloured order (a different but equally valid Gauss-Seidel order) drifts apart chaotically."""
import numpy as np
import pytest

import mgf_b200
import oracle_lib
from mgf_b200 import _lib as L
from mgf_b200 import scenes

pytestmark = pytest.mark.gpu
DT = np.float32(1.0 / 60.0)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _pair(bodies, terrain):
    g = mgf_b200.World(device=0, step_order=L.STEP_ORDER_REFERENCE)
    o = oracle_lib.OracleWorld()
    for w in (g, o):
        w.add_bodies(*bodies)
        if terrain is not None:
            w.set_terrain(*terrain)
    return g, o


def _run(g, o, iters, nsteps, what):
    total = 0
    for s in range(nsteps):
        st = g.step(DT, iters)
        o.step(DT, iters)
        cand, tcand = o.stats()
        assert (st["candidate_pairs"], st["terrain_candidates"]) == (cand, tcand), f"{what} step {s}: candidates {st['candidate_pairs']}, {st['terrain_candidates']} vs the reference's trees {cand}, {tcand}"
        for name, sg, so in zip("x q v omega".split(), g.state(), o.state()):
            bad = np.nonzero((_bits(sg) != _bits(so)).any(axis=1))[0]
            assert len(bad) == 0, f"{what} step {s}: {name} differs for {len(bad)} bodies, first {bad[:5].tolist()}"
        total += st["constraints"]
    return total


def test_c1_300_steps_reference_order_bit_identical_to_native_world_step():
    """BASELINE configs[0] (512 spheres in the demo box, 10 iterations) from rest through free fall, the impact at step ~145 and
    the settling: 300 steps, max error of x, q, v, omega = 0."""
    bodies, terrain, iters = scenes.build_config("C1")
    g, o = _pair(bodies, terrain)
    total = _run(g, o, iters, 300, "C1")
    assert total > 40000


def test_jittered_pile_and_capsules_on_a_mesh_reference_order():
    shapes, mass, rest, fric, force = scenes.balls_scene(num=6, jitter=0.2, seed=3)
    shapes["p"][:, 1] -= 24.0
    g, o = _pair((shapes, mass, rest, fric, force), scenes.box_terrain(4.0, 10.0, 4.0))
    assert _run(g, o, 20, 120, "jittered pile") > 20000
    bodies, terrain, iters = scenes.config_c3(scale=0.12)          # capsules on a height field: 2-contact faces, rotating bodies
    g, o = _pair(bodies, terrain)
    assert _run(g, o, iters, 40, "capsules on a mesh") > 1500


def test_reference_order_keeps_the_trees_history_bodies_added_mid_run_and_restore():
    """The order is a function of the trees' HISTORY: bodies added between steps are inserted into the tree as it is then
    (world.rs:178-184), and a restored snapshot rebuilds it in body order -- in both implementations."""
    shapes, mass, rest, fric, force = scenes.balls_scene(num=5, jitter=0.2, seed=7)
    shapes["p"][:, 1] -= 26.0
    g, o = _pair((shapes, mass, rest, fric, force), scenes.box_terrain(4.0, 10.0, 4.0))
    _run(g, o, 10, 60, "before")
    more = scenes.balls_scene(num=3, jitter=0.1, seed=9)
    more[0]["p"][:, 1] -= 20.0
    for w in (g, o):
        w.add_bodies(*more)
    assert _run(g, o, 10, 60, "after adding bodies") > 5000
    snap = g.snapshot()
    for w in (g, o):
        w.restore(snap)
    assert _run(g, o, 10, 30, "after restore") > 2000
