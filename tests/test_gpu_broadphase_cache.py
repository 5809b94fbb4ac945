"""GPU: the temporally coherent broadphase (csrc/bpcache.cuh) against the plain grid + sweep it replaces.  Two worlds hold the same
scene -- one with the cache (default), one created under MGFB_BROADPHASE=sweep -- and step side by side: the candidate-pair count,
the constraint count and every bit of x, q, v, omega must agree after EVERY step (the pair SET is what world.rs:261-268 defines;
the solve order depends only on the constraints' identities, not on list order), through all three paths the device chooses
between: rebuild, coherent steps with bodies replaced / drifting / running away, and the plain sweep when too much moves."""
import os

import numpy as np
import pytest

import mgf_b200
from mgf_b200 import scenes

pytestmark = pytest.mark.gpu
DT = np.float32(1.0 / 60.0)


def _world(bodies, terrain, sweep):
    old = os.environ.get("MGFB_BROADPHASE")
    if sweep:
        os.environ["MGFB_BROADPHASE"] = "sweep"
    else:
        os.environ.pop("MGFB_BROADPHASE", None)
    try:
        w = mgf_b200.World(device=0)     # the switch is read when the context is created
    finally:
        if old is None:
            os.environ.pop("MGFB_BROADPHASE", None)
        else:
            os.environ["MGFB_BROADPHASE"] = old
    w.add_bodies(*bodies)
    w.set_terrain(*terrain)
    return w


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _side_by_side(a, b, iters, nsteps, what, paths):
    for s in range(nsteps):
        sa, sb = a.step(DT, iters), b.step(DT, iters)
        assert sb["broadphase_path"] == 0
        paths[sa["broadphase_path"]] = paths.get(sa["broadphase_path"], 0) + 1
        for key in ("candidate_pairs", "terrain_candidates", "constraints", "terrain_constraints", "fat_refreshes"):
            assert sa[key] == sb[key], f"{what} step {s}: {key} {sa[key]} (cache, path {sa['broadphase_path']}) vs {sb[key]} (sweep)"
        for name, xa, xb in zip("x q v omega".split(), a.state(), b.state()):
            bad = np.nonzero((_bits(xa) != _bits(xb)).any(axis=1))[0]
            assert len(bad) == 0, f"{what} step {s}: {name} differs for {len(bad)} bodies, first {bad[:5].tolist()}"


def test_settling_pile_cache_equals_sweep_every_step():
    """20 k spheres: a squeezed pile relaxes (most bodies move at first: plain sweep by decision), settles (coherent steps, a rebuild
    whenever runaways accumulate), is kicked (a block of bodies thrown upwards: replaced every step, drifting past the cells'
    allowance into the overflow list) and settles again."""
    bodies = scenes.pile_xyz(40, 12, 42, jitter=0.03, seed=5)
    terrain = scenes.box_terrain(30.0, 30.0, 30.0)
    a, b = _world(bodies, terrain, False), _world(bodies, terrain, True)
    paths = {}
    _side_by_side(a, b, 10, 150, "relaxing pile", paths)
    n = len(bodies[0])
    v = np.zeros((n, 3), np.float32); w = np.zeros((n, 3), np.float32)
    for wld in (a, b):
        x, q, v0, w0 = wld.state()
        v[:] = v0; w[:] = w0
        v[: n // 40, 1] += np.float32(6.0)       # 2.5 % of the bodies: few enough for the cache to stay on, fast enough to run away
        wld.set_velocity(0, v, w)
    _side_by_side(a, b, 10, 120, "kicked pile", paths)
    assert paths.get(1, 0) > 100, paths            # coherent steps dominate once the pile is calm
    assert paths.get(2, 0) >= 2, paths             # rebuilt more than once
    assert paths.get(3, 0) >= 1, paths             # and the plain sweep was chosen while everything moved
    a.ctx.close(); b.ctx.close()


def test_cache_survives_restore_and_added_bodies():
    """Host-side invalidation: a restored snapshot and bodies added mid-run make the next step rebuild; results stay identical."""
    bodies = scenes.pile_xyz(16, 8, 16, jitter=0.03, seed=9)
    terrain = scenes.box_terrain(12.0, 20.0, 12.0)
    a, b = _world(bodies, terrain, False), _world(bodies, terrain, True)
    paths = {}
    _side_by_side(a, b, 10, 60, "before the snapshot", paths)
    snap = a.snapshot()
    _side_by_side(a, b, 10, 20, "after the snapshot", paths)
    a.restore(snap); b.restore(snap)
    st = a.step(DT, 10); b.step(DT, 10)
    assert st["broadphase_path"] in (2, 3), st
    _side_by_side(a, b, 10, 20, "after the restore", paths)
    extra = scenes.pile_xyz(4, 2, 4, jitter=0.02, seed=11)
    extra[0]["p"][:, 1] += np.float32(12.0)
    for wld in (a, b):
        wld.add_bodies(*extra)
    st = a.step(DT, 10); b.step(DT, 10)
    assert st["broadphase_path"] in (2, 3), st
    _side_by_side(a, b, 10, 60, "after adding bodies", paths)
    assert paths.get(1, 0) > 50, paths
    a.ctx.close(); b.ctx.close()
