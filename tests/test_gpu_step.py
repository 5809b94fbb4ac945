"""GPU parity of World::step, through the C ABI.  The CUDA path solves constraints in its own
colour-major order and exports that order; the oracle replays the step with exactly that
order (Solver::add_constraint order is caller-chosen, solver.rs:66) and the full body state
must then agree BIT FOR BIT, step after step.  That checks, in one go: the pair set of the
broadphase, every narrowphase contact, ContactConstraint::new, the solver and integration."""
import numpy as np
import pytest

import oracle_lib
import mgf_b200
from mgf_b200 import _lib as L
from mgf_b200 import scenes

pytestmark = pytest.mark.gpu
DT = np.float32(1.0 / 60.0)


SCHEDULE = [L.SCHEDULE_DATAFLOW]


@pytest.fixture(autouse=True, params=[L.SCHEDULE_DATAFLOW, L.SCHEDULE_PHASES, L.SCHEDULE_PHASES_JP], ids=["dataflow", "phases", "phases_jp"])
def _schedule(request):
    """Every test runs under both solver schedules (include/mgfb.h mgfb_solver_schedule): body version
    counters without grid barriers, and one grid barrier per colour."""
    SCHEDULE[0] = request.param
    yield


def _pair(bodies, terrain):
    g = mgf_b200.World(device=0, solver_schedule=SCHEDULE[0])
    o = oracle_lib.OracleWorld()
    for w in (g, o):
        w.add_bodies(*bodies)
        if terrain is not None:
            w.set_terrain(*terrain)
    return g, o


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _assert_state_equal(g, o, what):
    for name, sg, so in zip("x q v omega".split(), g.state(), o.state()):
        bad = np.nonzero((_bits(sg) != _bits(so)).any(axis=1))[0]
        assert len(bad) == 0, f"{what}: {name} differs for {len(bad)} bodies, first {bad[:5].tolist()}: {sg[bad[:2]]} vs {so[bad[:2]]}"


def _lockstep(g, o, iters, nsteps, what):
    """One GPU step, then the oracle builds its constraints, we check the two constraint SETS are
    identical, and the oracle solves in the GPU's order."""
    total = 0
    for s in range(nsteps):
        st = g.step(DT, iters)
        m = o.build(DT)
        assert st["constraints"] == m, f"{what} step {s}: {st['constraints']} constraints vs oracle {m}"
        cand, tcand = o.stats()
        # The device sweep reports exactly {j<i : tight_i overlaps fat_j}.  The reference's tree can
        # prune a leaf that merely touches the query (rounded parent boxes); never a contact.
        assert cand <= st["candidate_pairs"] == o.brute_pairs(), \
            f"{what} step {s}: candidate pairs {st['candidate_pairs']} vs tree {cand} / brute force {o.brute_pairs()}"
        if m:
            ga, gb, gf, gs, gc = g.constraints()
            oa, ob, of, osub = o.constraints(m)
            index = {k: i for i, k in enumerate(zip(oa.tolist(), ob.tolist(), of.tolist(), osub.tolist()))}
            assert len(index) == m
            try:
                perm = np.array([index[k] for k in zip(ga.tolist(), gb.tolist(), gf.tolist(), gs.tolist())], dtype=np.uint32)
            except KeyError as e:
                raise AssertionError(f"{what} step {s}: GPU constraint {e} not produced by the oracle")
            assert len(set(perm.tolist())) == m
            # colours are a proper edge colouring: within one colour no dynamic body repeats
            for c in np.unique(gc):
                sel = gc == c
                bodies = np.concatenate([ga[sel].astype(np.int64), gb[sel][gb[sel] >= 0].astype(np.int64)])
                assert len(np.unique(bodies)) == len(bodies), f"colour {c} touches a body twice"
            assert np.all(np.diff(gc.astype(np.int64)) >= 0), "rows are not colour-major"
            o.solve_order(perm, iters)
            total += m
        _assert_state_equal(g, o, f"{what} step {s}")
    return total


def test_c1_balls_lockstep_bit_exact():
    """BASELINE config 0: 512 spheres in the demo box, 10 iterations.  First 100 steps from rest
    (free fall), then the impact + settling phase where every step has hundreds of contacts."""
    bodies, terrain, iters = scenes.build_config("C1")
    g, o = _pair(bodies, terrain)
    _lockstep(g, o, iters, 100, "C1 free fall")
    g.step(DT, iters, nsteps=45); o.step(DT, iters, 45)       # both in their own order: no contacts yet
    _assert_state_equal(g, o, "C1 before impact")
    total = _lockstep(g, o, iters, 60, "C1 impact")
    assert total > 5000


def test_jittered_pile_lockstep_bit_exact():
    """A jittered (non-lattice) pile: sphere-sphere contacts in all directions, sphere-wall and
    sphere-floor contacts, fat-box refreshes."""
    shapes, mass, rest, fric, force = scenes.balls_scene(num=6, jitter=0.2, seed=3)
    shapes["p"][:, 1] -= 24.0        # start just above the floor so the pile forms quickly
    bodies = (shapes, mass, rest, fric, force)
    terrain = scenes.box_terrain(4.0, 10.0, 4.0)
    g, o = _pair(bodies, terrain)
    total = _lockstep(g, o, 20, 150, "jittered pile")
    assert total > 20000


def test_capsules_lockstep_bit_exact():
    """Capsule bodies (capsules.rs scene, shrunk): capsule-capsule, capsule-triangle with 1 and 2
    contacts per face, rotating bodies (quaternion + world inertia update)."""
    shapes, mass, rest, fric, force = scenes.capsules_scene(num=4, jitter=0.3, seed=5)
    shapes["p"][:, 1] -= 26.0
    rng = np.random.default_rng(0)
    bodies = (shapes, mass, rest, fric, force)
    terrain = scenes.box_terrain(12.0, 10.0, 12.0)
    g, o = _pair(bodies, terrain)
    om = rng.uniform(-2, 2, (len(shapes), 3)).astype(np.float32)
    v = rng.uniform(-1, 1, (len(shapes), 3)).astype(np.float32)
    g.set_velocity(0, v, om); o.set_velocity(0, v, om)
    total = _lockstep(g, o, 20, 150, "capsules")
    assert total > 1000


def test_mixed_spheres_capsules_on_heightfield():
    """All four body-pair kernels + both terrain kernels against a 2 x 20 x 20 triangle height
    field (terrain grid with many cells)."""
    s1 = scenes.balls_scene(num=5, jitter=0.2, seed=11)
    s2 = scenes.capsules_scene(num=3, jitter=0.3, seed=12)
    shapes = np.concatenate([s1[0], s2[0]])
    shapes["p"][:len(s1[0]), 1] -= 20.0
    shapes["p"][len(s1[0]):, 1] -= 20.0
    n = len(shapes)
    order = np.random.default_rng(1).permutation(n)     # interleave kinds so i/j roles mix
    shapes = shapes[order]
    bodies = (shapes, np.full(n, 1.5, np.float32), np.full(n, 0.3, np.float32), np.full(n, 0.6, np.float32),
              np.tile(np.array([0, -9.8, 0], np.float32), (n, 1)))
    terrain = scenes.heightfield_terrain(nq=20, size=40.0, y0=-10.0, amp=0.5, freq=0.3)
    g, o = _pair(bodies, terrain)
    total = _lockstep(g, o, 20, 120, "mixed/heightfield")
    assert total > 3000


def test_integrate_and_complete_motion_api():
    """RigidBodyVec::integrate / complete_motion as separate calls (physics.rs:222,262)."""
    shapes, mass, rest, fric, force = scenes.capsules_scene(num=3)
    g, o = _pair((shapes, mass, rest, fric, force), None)
    rng = np.random.default_rng(2)
    om = rng.uniform(-3, 3, (len(shapes), 3)).astype(np.float32)
    v = rng.uniform(-3, 3, (len(shapes), 3)).astype(np.float32)
    g.set_velocity(0, v, om); o.set_velocity(0, v, om)
    for _ in range(5):
        g.integrate(DT); o.integrate(DT)
        g.complete_motion(); o.complete_motion()
    _assert_state_equal(g, o, "integrate")
    assert np.array_equal(_bits(g.inv_moment()), _bits(o.inv_moment()))
    cg, co = g.colliders(), o.colliders()
    assert np.array_equal(_bits(cg["p"]), _bits(co["p"])) and np.array_equal(_bits(cg["v"]), _bits(co["v"]))


def test_solver_as_given_is_bit_identical_to_sequential():
    """MGFB_ORDER_AS_GIVEN: level-scheduled solve == the reference's sequential Gauss-Seidel over
    the list exactly as given (reference order of World::step), no permutation involved."""
    shapes, mass, rest, fric, force = scenes.balls_scene(num=6, jitter=0.2, seed=3)
    shapes["p"][:, 1] -= 24.0
    terrain = scenes.box_terrain(4.0, 10.0, 4.0)
    g, o = _pair((shapes, mass, rest, fric, force), terrain)
    _lockstep(g, o, 20, 80, "warm-up")
    for s in range(5):
        g.complete_motion(); g.integrate(DT)
        m = o.build(DT)
        assert m > 300
        d = o.manifolds(m)
        perm, imp, st = g.solve_manifolds(d["obj_a"], d["obj_b"], d["normal"], d["tangent"], d["ncontacts"], d["local_a"], d["local_b"],
                                          DT, 20, order=L.ORDER_AS_GIVEN, static_center=d["static_center"], static_friction=d["static_friction"])
        assert st["groups"] >= 1 and st["constraints"] == m
        o.solve_order(np.arange(m, dtype=np.uint32), 20)
        _assert_state_equal(g, o, f"as-given solve {s}")
        # the level schedule preserves the relative order of constraints that share a body
        pos = np.empty(m, np.int64); pos[perm] = np.arange(m)
        last = {}
        for k in range(m):
            for b in (int(d["obj_a"][k]), int(d["obj_b"][k])):
                if b >= 0:
                    if b in last:
                        assert pos[last[b]] < pos[k]
                    last[b] = k


def test_solver_multi_contact_manifolds_and_coloured_order():
    """User-built manifolds with 1..4 contacts (ContactPruner output for compound shapes),
    coloured order: the oracle replays the returned permutation."""
    rng = np.random.default_rng(5)
    n = 300
    shapes, mass, rest, fric, force = scenes.balls_scene(num=7, jitter=0.1, seed=9)
    shapes = shapes[:n]
    g, o = _pair((shapes, mass[:n], rest[:n], fric[:n], force[:n]), None)
    v = rng.uniform(-2, 2, (n, 3)).astype(np.float32); om = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    g.set_velocity(0, v, om); o.set_velocity(0, v, om)
    g.integrate(DT); o.integrate(DT)
    m = 900
    a = rng.integers(0, n, m).astype(np.int32)
    b = rng.integers(0, n, m).astype(np.int32)
    b[b == a] = (a[b == a] + 1) % n
    b[rng.random(m) < 0.15] = -1                       # some static partners
    swap = (rng.random(m) < 0.05) & (b < 0)            # ... a few with the static body in slot a
    a2 = np.where(swap, -1, a).astype(np.int32); b2 = np.where(swap, a, b).astype(np.int32)
    nrm = rng.normal(size=(m, 3)).astype(np.float32); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    t0 = np.cross(nrm, rng.normal(size=(m, 3))).astype(np.float32); t0 /= np.linalg.norm(t0, axis=1, keepdims=True)
    t1 = np.cross(nrm, t0).astype(np.float32)
    d = dict(obj_a=a2, obj_b=b2, static_center=rng.uniform(-1, 1, (m, 3)).astype(np.float32),
             static_friction=rng.uniform(0, 1, m).astype(np.float32), normal=nrm.astype(np.float32),
             tangent=np.concatenate([t0, t1], axis=1).astype(np.float32), ncontacts=rng.integers(1, 5, m).astype(np.uint32),
             local_a=rng.uniform(-0.5, 0.5, (m, 12)).astype(np.float32), local_b=rng.uniform(-0.5, 0.5, (m, 12)).astype(np.float32))
    for order in (L.ORDER_COLOURED, L.ORDER_AS_GIVEN):
        perm, imp_g, st = g.solve_manifolds(d["obj_a"], d["obj_b"], d["normal"], d["tangent"], d["ncontacts"], d["local_a"], d["local_b"],
                                            DT, 8, order=order, static_center=d["static_center"], static_friction=d["static_friction"])
        assert sorted(perm.tolist()) == list(range(m))
        imp_o = o.solve_manifolds(d, DT, 8, perm=perm)
        _assert_state_equal(g, o, f"multi-contact order={order}")
        assert np.array_equal(_bits(imp_g), _bits(imp_o))


def test_edge_cases(ctx):
    w = mgf_b200.World(device=0)
    assert w.step(DT, 20)["constraints"] == 0              # empty world
    w.add_bodies(mgf_b200.sphere((0, 5, 0), 0.5), 1.0, 0.3, 0.6, (0, -9.8, 0))
    st = w.step(DT, 20)                                     # one body, no terrain
    assert st["bodies"] == 1 and st["constraints"] == 0
    with pytest.raises(mgf_b200.MgfbError) as e:
        w.add_bodies(mgf_b200.sphere((0, 0, 0), 0.0), 1.0, 0.3, 0.6, (0, 0, 0))   # assert!(radius > 0)
    assert e.value.code == L.ERR_INVALID_ARG
    with pytest.raises(mgf_b200.MgfbError) as e:
        w.add_bodies(mgf_b200.sphere((0, 0, 0), 1.0), 0.0, 0.3, 0.6, (0, 0, 0))   # zero mass -> singular / inf
    assert e.value.code in (L.ERR_SINGULAR_INERTIA, L.ERR_INVALID_ARG)
    with pytest.raises(mgf_b200.MgfbError):
        w.state(first=5, n=10)


def test_work_list_overflow_regrows_and_stays_exact():
    """Tiny initial capacities: a dense blob overflows the pair/contact lists; the step regrows
    them, reruns the post-integration part and must still match the oracle bit for bit."""
    rng = np.random.default_rng(3)
    n = 1500
    shapes = np.zeros(n, dtype=L.SHAPE_DTYPE)
    shapes["kind"] = L.SPHERE
    shapes["p"][:, 0:3] = rng.uniform(-1.2, 1.2, (n, 3)); shapes["p"][:, 3] = 0.5     # ~ everything overlaps everything
    bodies = (shapes, np.ones(n, np.float32), np.full(n, 0.3, np.float32), np.full(n, 0.6, np.float32), np.zeros((n, 3), np.float32))
    g, o = _pair(bodies, None)
    st = g.step(DT, 2)
    assert st["overflow"] != 0 and st["candidate_pairs"] > 8 * 1500
    m = o.build(DT)
    assert st["constraints"] == m
    ga, gb, gf, gs, gc = g.constraints()
    oa, ob, of, osub = o.constraints(m)
    index = {k: i for i, k in enumerate(zip(oa.tolist(), ob.tolist(), of.tolist(), osub.tolist()))}
    perm = np.array([index[k] for k in zip(ga.tolist(), gb.tolist(), gf.tolist(), gs.tolist())], dtype=np.uint32)
    o.solve_order(perm, 2)
    _assert_state_equal(g, o, "after overflow")


def test_pipelined_step_equals_synchronous_step_bit_for_bit():
    """mgfb_step_enqueue / mgfb_step_wait (transfers of step k overlap the kernels of step k+1) must deliver, step by
    step, exactly what set_velocity + step + get_state deliver."""
    import torch
    shapes, mass, rest, fric, force = scenes.balls_scene(num=6, jitter=0.2, seed=3)
    shapes["p"][:, 1] -= 24.0
    bodies = (shapes, mass, rest, fric, force)
    terrain = scenes.box_terrain(4.0, 10.0, 4.0)
    n = len(shapes)
    a = mgf_b200.World(device=0, solver_schedule=SCHEDULE[0]); b = mgf_b200.World(device=0, solver_schedule=SCHEDULE[0])
    for w in (a, b):
        w.add_bodies(*bodies); w.set_terrain(*terrain)
        w.step(DT, 10, nsteps=60)          # into the contact-rich phase
    pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    rng = np.random.default_rng(1)
    nsteps = 12
    vin = [pin((n, 3)) for _ in range(nsteps)]; win = [pin((n, 3)) for _ in range(nsteps)]
    outs = [tuple(pin(s) for s in ((n, 3), (n, 4), (n, 3), (n, 3))) for _ in range(nsteps)]
    _, _, v0, w0 = a.state()
    for k in range(nsteps):
        vin[k][:] = v0 + rng.uniform(-0.05, 0.05, (n, 3)).astype(np.float32); win[k][:] = w0
    ref = []
    for k in range(nsteps):
        a.set_velocity(0, vin[k], win[k])
        sa = a.step(DT, 10)
        ref.append((sa["constraints"], [x.copy() for x in a.state()]))
    stats = []
    for k in range(nsteps):
        b.step_enqueue(DT, 10, vin[k], win[k], *outs[k])
        if k > 0:
            stats.append(b.step_wait())
    stats.append(b.step_wait())
    with pytest.raises(mgf_b200.MgfbError):
        b.step_wait()                       # nothing in flight
    assert sum(c for c, _ in ref) > 1000
    for k in range(nsteps):
        assert stats[k]["constraints"] == ref[k][0]
        for name, got, want in zip("x q v omega".split(), outs[k], ref[k][1]):
            assert np.array_equal(_bits(got), _bits(want)), f"step {k}: {name} differs"
    _assert_same = [np.array_equal(_bits(x), _bits(y)) for x, y in zip(a.state(), b.state())]
    assert all(_assert_same)
    # MGFB_INPUT_ADD: the inputs are added (external impulses); the synchronous twin adds on the host
    dv = pin((n, 3)); dw = pin((n, 3))
    dv[:] = rng.uniform(-0.02, 0.02, (n, 3)).astype(np.float32); dw[:] = 0.0
    _, _, va, wa = a.state()
    a.set_velocity(0, (va + dv).astype(np.float32), (wa + dw).astype(np.float32)); a.step(DT, 10)
    b.step_enqueue(DT, 10, dv, dw, *outs[0], add=True); b.step_wait()
    assert all(np.array_equal(_bits(x), _bits(y)) for x, y in zip(a.state(), b.state()))


def test_snapshot_restore_into_gpu_and_oracle_continues_bit_identically():
    """mgfb_bodies_get_fat_bounds / mgfb_bodies_set_state: a world stepped on the GPU, saved, and loaded into (a) a fresh
    GPU world and (b) the oracle continues bit-identically in all three (the benchmark's settled window relies on it)."""
    shapes, mass, rest, fric, force = scenes.balls_scene(num=6, jitter=0.2, seed=5)
    shapes["p"][:, 1] -= 24.0
    bodies = (shapes, mass, rest, fric, force)
    terrain = scenes.box_terrain(4.0, 10.0, 4.0)
    g, o = _pair(bodies, terrain)
    g.step(DT, 20, nsteps=90)
    snap = g.snapshot()
    assert np.abs(snap["omega"]).max() > 0 and np.abs(snap["colliders"]["v"]).max() > 0
    g2, _ = _pair(bodies, terrain)
    g2.restore(snap); o.restore(snap)
    for k, v in g2.snapshot().items():
        assert np.array_equal(np.ascontiguousarray(v).view(np.uint8), np.ascontiguousarray(snap[k]).view(np.uint8)), k
    for k, v in o.snapshot().items():
        assert np.array_equal(np.ascontiguousarray(v).view(np.uint8), np.ascontiguousarray(snap[k]).view(np.uint8)), k
    assert np.array_equal(_bits(g.inv_moment()), _bits(g2.inv_moment()))
    total = _lockstep(g2, o, 20, 8, "restored world")
    assert total > 1500
    g.step(DT, 20, nsteps=8)
    for a, b in zip(g.state(), g2.state()):
        assert np.array_equal(_bits(a), _bits(b))
    # error behaviour: a collider cannot change kind or radius; negative half extents are the reference's bounds.rs:125 assert
    bad = snap["colliders"].copy(); bad["p"][0, 3] = 0.7
    with pytest.raises(mgf_b200.MgfbError) as e:
        g2.set_state(0, colliders=bad)
    assert e.value.code == L.ERR_INVALID_ARG
    fat = snap["fat"].copy(); fat[3, 4] = -1.0
    with pytest.raises(mgf_b200.MgfbError) as e:
        g2.set_state(0, fat=fat)
    assert e.value.code == L.ERR_NAN_BOUNDS


def test_pipelined_overflow_regrows_requeues_and_stays_exact():
    """A pipelined step that overflows a work list (dense blob: ~500 pairs per body against 16 slots) while younger steps
    are queued behind it: mgfb_step_wait regrows, re-runs the step from after its integration and re-queues the younger
    steps whole (their ADDed velocity inputs must be applied exactly once).  Same bits as the synchronous calls."""
    import torch
    rng = np.random.default_rng(3)
    n = 1500
    shapes = np.zeros(n, dtype=L.SHAPE_DTYPE)
    shapes["kind"] = L.SPHERE
    shapes["p"][:, 0:3] = rng.uniform(-1.2, 1.2, (n, 3)); shapes["p"][:, 3] = 0.5
    bodies = (shapes, np.ones(n, np.float32), np.full(n, 0.3, np.float32), np.full(n, 0.6, np.float32), np.zeros((n, 3), np.float32))
    a = mgf_b200.World(device=0, solver_schedule=SCHEDULE[0]); b = mgf_b200.World(device=0, solver_schedule=SCHEDULE[0])
    a.add_bodies(*bodies); b.add_bodies(*bodies)
    pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    nsteps = 3
    vin = [pin((n, 3)) for _ in range(nsteps)]; win = [pin((n, 3)) for _ in range(nsteps)]
    outs = [tuple(pin(s) for s in ((n, 3), (n, 4), (n, 3), (n, 3))) for _ in range(nsteps)]
    for k in range(nsteps):
        vin[k][:] = rng.uniform(-0.05, 0.05, (n, 3)).astype(np.float32); win[k][:] = rng.uniform(-0.05, 0.05, (n, 3)).astype(np.float32)
    ref = []
    for k in range(nsteps):
        _, _, v, w = a.state()
        a.set_velocity(0, v + vin[k], w + win[k])
        sa = a.step(DT, 2)
        ref.append((sa["constraints"], [x.copy() for x in a.state()]))
    for k in range(nsteps):
        b.step_enqueue(DT, 2, vin[k], win[k], *outs[k], add=True)       # all three in flight
    stats = [b.step_wait() for _ in range(nsteps)]
    assert stats[0]["overflow"] != 0 and stats[1]["overflow"] == 0
    for k in range(nsteps):
        assert stats[k]["constraints"] == ref[k][0], (k, stats[k]["constraints"], ref[k][0])
        for got, want in zip(outs[k], ref[k][1]):
            assert np.array_equal(_bits(got), _bits(want)), f"pipelined step {k} differs from the synchronous step"
    for x, y in zip(a.state(), b.state()):
        assert np.array_equal(_bits(x), _bits(y))


def test_handover_selftest_sees_no_torn_record():
    """The 32-byte hand-over store of the dataflow solver is checked at context creation (selftest.cuh); run it harder here:
    writers overwrite records 20 000 times while readers on other SMs poll -- every record read must be whole."""
    with mgf_b200.Context(device=0) as c:
        torn, seen = c.selftest_handover(20000)
        assert torn == 0
        assert seen > 16 * 32       # the readers did watch the records change
        assert c.cfg.solver_schedule == L.SCHEDULE_DATAFLOW
