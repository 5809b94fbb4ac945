"""CPU-only tests: the oracle against the reference's golden vectors, plus hand-derived
known answers for the parts of the path the reference never tests (solver, manifold,
integration) -- those are OURS, not the reference's, and say so."""
import os
import subprocess

import numpy as np

import kat_check
import oracle_lib
import mgf_b200.api as api
from mgf_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_unit_tests_transcribed_in_cpp():
    """oracle/kat.cpp = the reference's #[test] functions that touch the path, bit-exact where
    the Rust test uses assert_eq!.  This is what pins the cgmath restatement."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "kat"], check=True)
    r = subprocess.run([os.path.join(ROOT, "oracle", "kat")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


def test_reference_vectors_through_batch_api(oracle):
    failures = kat_check.check_cases(oracle_lib.contacts_batch)
    assert not failures, "\n".join(failures)


def test_hand_derived_head_on_impulse(oracle):
    """OURS (the reference has no solver test): two unit spheres, unit mass, touching, closing
    head-on along x at 2 each, restitution 0.3, no gravity, dt = 1/60, one iteration.
    pen = 0 -> Baumgarte term = -(0.2*60)*(0 + 0.05) = -0.6; rel_v = -4 < -1 -> bias = -0.6 + 0.3*4 = 0.6;
    normal_mass = 1/(1+1); lambda = 0.5*(4 + 0.6) = 2.3; v0 = 2 - 2.3, v1 = -2 + 2.3."""
    w = oracle_lib.OracleWorld()
    shapes = np.concatenate([api.sphere((-1, 0, 0), 1.0), api.sphere((1, 0, 0), 1.0)])
    w.add_bodies(shapes, 1.0, 0.3, 0.5, (0.0, 0.0, 0.0))
    w.set_velocity(0, [[2.0, 0, 0], [-2.0, 0, 0]], np.zeros((2, 3)))
    d = dict(obj_a=np.array([0], np.int32), obj_b=np.array([1], np.int32), static_center=np.zeros((1, 3), np.float32),
             static_friction=np.zeros(1, np.float32), normal=np.array([[1, 0, 0]], np.float32),
             tangent=np.array([[0, 0, -1, 0, 1, 0]], np.float32), ncontacts=np.array([1], np.uint32),
             local_a=np.zeros((1, 12), np.float32), local_b=np.zeros((1, 12), np.float32))
    d["local_a"][0, :3] = [1, 0, 0]
    d["local_b"][0, :3] = [-1, 0, 0]
    imp = w.solve_manifolds(d, 1.0 / 60.0, 1)
    _, _, v, om = w.state()
    assert np.allclose(v[:, 0], [-0.3, 0.3], atol=1e-5)
    assert np.allclose(imp[0, 0], 2.3, atol=1e-5)
    assert np.all(om == 0)


def test_hand_derived_sphere_resting_on_floor(oracle):
    """OURS: a sphere on the demo floor yields ONE terrain constraint against
    Static{center: mesh.x, friction: 0} with normal = -n_tri.  With
    penetration below PENETRATION_SLOP the Baumgarte bias -(0.2/dt)*(pen + 0.05) is negative
    and larger than the approach speed, so the accumulated impulse clamps to 0 and the body
    keeps sinking (reference behaviour: a resting sphere settles ~0.07 deep); beyond the slop
    the contact pushes back."""
    dt = np.float32(1.0 / 60.0)
    for y0, pushes in ((-9.5, False), (-9.62, True)):
        w = oracle_lib.OracleWorld()
        w.add_bodies(api.sphere((1.0, y0, 2.0), 0.5), 2.0, 0.3, 0.6, (0.0, -9.8, 0.0))
        w.set_terrain(*scenes.box_terrain())
        m = w.build(dt)
        assert m == 1
        d = w.manifolds(m)
        assert d["obj_a"][0] == 0 and d["obj_b"][0] == -1
        assert np.array_equal(d["normal"][0], np.array([0, -1, 0], np.float32))
        assert np.allclose(d["static_center"][0], [0, -10, 0])
        w.solve_order(np.array([0], np.uint32), 20)
        _, _, v, _ = w.state()
        free = np.float32(-9.8) * np.float32(2.0) * np.float32(0.5) * dt   # force * inv_mass * dt
        if pushes:
            assert v[0, 1] > 0.0
        else:
            assert v[0, 1] == free


def test_solve_order_identity_equals_step(oracle):
    """Replaying the constraints in their own insertion order is the reference step."""
    bodies, terrain, iters = scenes.build_config("C1")
    a = oracle_lib.OracleWorld(); b = oracle_lib.OracleWorld()
    for w in (a, b):
        w.add_bodies(*bodies); w.set_terrain(*terrain)
    a.step(1 / 60, iters, 160); b.step(1 / 60, iters, 160)
    for _ in range(5):
        a.step(1 / 60, iters)
        m = b.build(1 / 60)
        b.solve_order(np.arange(m, dtype=np.uint32), iters)
    for sa, sb in zip(a.state(), b.state()):
        assert np.array_equal(sa.view(np.uint32), sb.view(np.uint32))


def test_constraint_identities_unique(oracle):
    bodies, terrain, iters = scenes.build_config("C1")
    w = oracle_lib.OracleWorld()
    w.add_bodies(*bodies); w.set_terrain(*terrain)
    w.step(1 / 60, iters, 200)
    m = w.build(1 / 60)
    cand, tcand = w.stats()
    a, b, face, sub = w.constraints(m)
    assert m > 0 and cand >= np.count_nonzero(b >= 0)
    assert len(set(zip(a.tolist(), b.tolist(), face.tolist(), sub.tolist()))) == m


def test_lcg_fast_matches_slow():
    assert np.array_equal(scenes.lcg_uniform(1000, 7), scenes.lcg_uniform_fast(1000, 7))
    u = scenes.lcg_uniform_fast(200000, 1)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01


def test_oracle_contact_pruner_hand_derived():
    """ContactPruner::push / Manifold::from (manifold.rs:72-148) have no test in the reference: hand-derived answers (OURS)."""
    from pruner_cases import lc
    x, y = (1.0, 0.0, 0.0), (0.0, 1.0, 0.0)
    far = np.concatenate([lc((0, 0, 0), (0, 0, 0), y, 0.0, (1, 0, 0), (0, 0, 0)), lc((3, 0, 0), (3, 0, 0), x, 0.0, (0, 2, 0), (0, 0, 0))])
    m = oracle_lib.manifolds_prune(far, [0, 2])
    assert m["ncontacts"][0] == 2 and np.array_equal(m["normal"][0], np.float32([0.5, 0.5, 0.0]))          # mean, not renormalised
    assert np.array_equal(m["local_a"][0][:6], np.float32([1, 0, 0, 0, 2, 0]))
    near = np.concatenate([lc((0, 0, 0), (0, 0, 0), y, 0.0, (1, 0, 0), (0, 0, 0)), lc((0.5, 0, 0), (9, 9, 9), x, 0.0, (0, 2, 0), (0, 0, 0))])
    m = oracle_lib.manifolds_prune(near, [0, 2])       # |a - a'|^2 = 0.25 <= 0.5: merged; the newcomer is further from the centres and wins
    assert m["ncontacts"][0] == 1 and np.array_equal(m["normal"][0], np.float32(x)) and np.array_equal(m["local_a"][0][:3], np.float32([0, 2, 0]))
    m = oracle_lib.manifolds_prune(near[::-1].copy(), [0, 2])   # other order: the incumbent is further out and stays
    assert m["ncontacts"][0] == 1 and np.array_equal(m["normal"][0], np.float32(x))
    times = np.concatenate([lc((0, 0, 0), (0, 0, 0), y, 0.5, (1, 0, 0), (0, 0, 0)), lc((5, 0, 0), (5, 0, 0), x, 0.2, (1, 0, 0), (0, 0, 0)),
                            lc((9, 0, 0), (9, 0, 0), y, 0.9, (1, 0, 0), (0, 0, 0))])
    m = oracle_lib.manifolds_prune(times, [0, 3])      # the earlier hit replaces, the later one is dropped
    assert m["ncontacts"][0] == 1 and m["time"][0] == np.float32(0.2) and np.array_equal(m["normal"][0], np.float32(x))
    m = oracle_lib.manifolds_prune(times[:0], [0, 0])
    assert m["ncontacts"][0] == 0 and np.isnan(m["normal"][0]).all() and np.isinf(m["time"][0])
