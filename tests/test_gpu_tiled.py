"""GPU parity of a world TILED across contexts (mgf_b200/csrc/tile.cuh), through the C ABI.

Every tile is its own mgfb_ctx with its own stream; ghosts and boundary velocities move through
peer-mapped device memory.  Here the tiles share ONE GPU and one process (each stepped from its
own host thread, cooperative grids capped so the persistent solvers are co-resident), which is
what the single-GPU test box can run; bench.py --gpus N runs the same code one process per GPU.

The tiled solve is equivalent to a sequential sweep in the order tiling.executed_order()
returns; the oracle (ONE untiled world holding every body) replays that order and every tile's
owned bodies must match it bit for bit, step after step."""
import threading

import numpy as np
import pytest

import oracle_lib
from mgf_b200 import scenes, tiling

pytestmark = pytest.mark.gpu
DT = np.float32(1.0 / 60.0)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _parallel(fns):
    """Run one callable per tile concurrently (the tiles wait for each other on the device)."""
    out = [None] * len(fns); err = [None] * len(fns)

    def run(k):
        try:
            out[k] = fns[k]()
        except BaseException as e:  # noqa: BLE001
            err[k] = e
    ts = [threading.Thread(target=run, args=(k,)) for k in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
        assert not t.is_alive(), "tile step did not return"
    for e in err:
        if e is not None:
            raise e
    return out


SCHEDULE = [0]


@pytest.fixture(autouse=True, params=[0, 1], ids=["dataflow", "phases"])
def _schedule(request):
    """Both solver schedules (include/mgfb.h mgfb_solver_schedule): body chains running across the tile
    boundary through peer-memory inboxes, and the barrier-phased solve with explicit exchange phases."""
    SCHEDULE[0] = request.param
    yield


def _make(bodies, terrain, ntiles, ctas):
    shapes = bodies[0]
    parts = tiling.slab_partition(tiling.shape_centres_x(shapes), ntiles)
    tiles = []
    for r in range(ntiles):
        t = tiling.TiledWorld(r, ntiles, device=0, max_cooperative_ctas=ctas, tile_timeout_ms=4000, solver_schedule=SCHEDULE[0])
        t.add_bodies(parts[r], *bodies)
        t.set_terrain(*terrain)
        tiles.append(t)
    tiling.connect_local(tiles, ghost_capacity=max(256, len(shapes) // 2))
    o = oracle_lib.OracleWorld()
    o.add_bodies(*bodies); o.set_terrain(*terrain)
    return tiles, o


def _lockstep(tiles, o, iters, nsteps, what):
    total = boundary = 0
    for s in range(nsteps):
        stats = _parallel([lambda t=t: t.step(DT, iters) for t in tiles])
        m = o.build(DT)
        assert sum(st["constraints"] for st in stats) == m, f"{what} step {s}: {[st['constraints'] for st in stats]} vs oracle {m}"
        per_rank = [t.constraints() for t in tiles]
        if m:
            ga, gb, gf, gs = tiling.executed_order(per_rank)
            oa, ob, of, osub = o.constraints(m)
            index = {k: i for i, k in enumerate(zip(oa.tolist(), ob.tolist(), of.tolist(), osub.tolist()))}
            try:
                perm = np.array([index[k] for k in zip(ga.tolist(), gb.tolist(), gf.tolist(), gs.tolist())], dtype=np.uint32)
            except KeyError as e:
                raise AssertionError(f"{what} step {s}: tiled constraint {e} not produced by the oracle")
            assert len(set(perm.tolist())) == m, "a constraint was produced by two tiles"
            o.solve_order(perm, iters)
            total += m
            boundary += sum(st["boundary_constraints"] for st in stats)
        ostate = o.state()
        for t in tiles:
            for name, sg, so in zip("x q v omega".split(), t.state(), ostate):
                so = so[t.ids]
                bad = np.nonzero((_bits(sg) != _bits(so)).any(axis=1))[0]
                assert len(bad) == 0, f"{what} step {s} tile {t.rank}: {name} differs for {len(bad)} bodies, first gid {t.ids[bad[:4]].tolist()}"
    return total, boundary


def test_two_tiles_lockstep_bit_exact():
    bodies = scenes.pile_xyz(16, 5, 5, jitter=0.01, seed=3)
    terrain = scenes.box_terrain(12.0, 10.0, 6.0)
    tiles, o = _make(bodies, terrain, 2, ctas=24)
    total, boundary = _lockstep(tiles, o, 10, 30, "2 tiles")
    assert total > 10000 and boundary > 300, (total, boundary)


def test_three_tiles_lockstep_bit_exact():
    """The middle tile has both neighbours: it sends ghosts left, receives ghosts from the right."""
    bodies = scenes.pile_xyz(30, 4, 4, jitter=0.01, seed=5)
    terrain = scenes.box_terrain(20.0, 10.0, 5.0)
    tiles, o = _make(bodies, terrain, 3, ctas=16)
    total, boundary = _lockstep(tiles, o, 8, 20, "3 tiles")
    assert total > 5000 and boundary > 200, (total, boundary)


def test_tiled_equals_untiled_constraint_set_and_multi_step_enqueue():
    """mgfb_step_n on every tile with no host synchronisation in between, against the oracle's own order-free invariants:
    same number of constraints and finite state; then one lock-step to show the tiles are still in sync."""
    bodies = scenes.pile_xyz(16, 4, 6, jitter=0.01, seed=9)
    terrain = scenes.box_terrain(12.0, 10.0, 6.0)
    tiles, o = _make(bodies, terrain, 2, ctas=24)
    _lockstep(tiles, o, 6, 3, "warm")
    # 5 steps enqueued back to back on both tiles; the oracle follows with the reference order of
    # each step, which is NOT the tiled order, so only order-free quantities are compared here
    stats = _parallel([lambda t=t: t.step(DT, 6, nsteps=5) for t in tiles])
    assert all(st["constraints"] > 0 for st in stats)
    for t in tiles:
        assert all(np.isfinite(a).all() for a in t.state())


def test_tile_too_thin_is_reported():
    """A world 4 bodies wide cut in 2: the edge bodies touch both neighbours' layers -> MGFB_ERR_TILE,
    never a silent wrong answer.  (With 3 tiles of width 2 the middle tile's bodies are ghosts on the
    left AND touch ghosts from the right.)"""
    import mgf_b200
    from mgf_b200 import _lib as L
    bodies = scenes.pile_xyz(3, 3, 3, jitter=0.0)
    terrain = scenes.box_terrain(8.0, 10.0, 4.0)
    shapes = bodies[0]
    parts = tiling.slab_partition(tiling.shape_centres_x(shapes), 3)
    tiles = []
    for r in range(3):
        t = tiling.TiledWorld(r, 3, device=0, max_cooperative_ctas=16, tile_timeout_ms=3000, solver_schedule=SCHEDULE[0])
        t.add_bodies(parts[r], *bodies); t.set_terrain(*terrain); tiles.append(t)
    tiling.connect_local(tiles, ghost_capacity=256)
    codes = [None] * 3

    def run(k):
        try:
            tiles[k].step(DT, 4); codes[k] = L.OK
        except mgf_b200.MgfbError as e:
            codes[k] = e.code
    _parallel([lambda k=k: run(k) for k in range(3)])
    assert L.ERR_TILE in codes, codes


def test_tiled_pipelined_steps_equal_synchronous_steps():
    """mgfb_step_enqueue / mgfb_step_wait on every tile (host transfers of step k overlapping the kernels of step
    k+1, the tiles meeting on the device) must leave exactly the state the synchronous steps leave."""
    import torch
    bodies = scenes.pile_xyz(16, 4, 5, jitter=0.01, seed=11)
    terrain = scenes.box_terrain(12.0, 10.0, 6.0)
    sync_tiles, _ = _make(bodies, terrain, 2, ctas=24)
    pipe_tiles, _ = _make(bodies, terrain, 2, ctas=24)
    nsteps = 6
    _parallel([lambda t=t: t.step(DT, 8, nsteps=nsteps) for t in sync_tiles])
    pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()

    def run(t):
        n = len(t.ids)
        outs = [tuple(pin(s) for s in ((n, 3), (n, 4), (n, 3), (n, 3))) for _ in range(2)]
        for k in range(nsteps):
            t.world.step_enqueue(DT, 8, None, None, *outs[k & 1])
            if k > 0:
                t.world.step_wait()
        t.world.step_wait()
        return outs[(nsteps - 1) & 1]
    last = _parallel([lambda t=t: run(t) for t in pipe_tiles])
    for ts, tp, out in zip(sync_tiles, pipe_tiles, last):
        for name, a, b, c in zip("x q v omega".split(), ts.state(), tp.state(), out):
            assert np.array_equal(_bits(a), _bits(b)), f"tile {ts.rank}: {name} differs between synchronous and pipelined steps"
            assert np.array_equal(_bits(a), _bits(c)), f"tile {ts.rank}: {name} read back by the pipeline differs"


def test_rebin_migrates_bodies_between_tiles_and_stays_bit_exact():
    """SURVEY 8e "re-bin every k steps": the top layer of the left half slides across the pile into the right half.  Every 8
    steps the world is re-tiled by the bodies' current x (tiling.rebin_local: snapshot -> new slabs -> restored state); bodies
    change owner, and the whole run stays bit-identical to the ONE untiled oracle world replaying the executed order."""
    nx, ny, nz = 16, 4, 4
    bodies = scenes.pile_xyz(nx, ny, nz, jitter=0.01, seed=13)
    terrain = scenes.box_terrain(14.0, 10.0, 6.0)
    tiles, o = _make(bodies, terrain, 2, ctas=24)
    # push: global id = (ix * ny + iy) * nz + iz ; the whole top layer (iy = ny - 1) gets v = (14, 0, 0) and slides over the pile
    gid = np.arange(nx * ny * nz)
    pushed = gid[(gid // nz) % ny == ny - 1]
    v = np.zeros((len(gid), 3), np.float32); v[pushed, 0] = 14.0
    w = np.zeros_like(v)
    o.set_velocity(0, v, w)
    for t in tiles:
        t.world.set_velocity(0, v[t.ids], w[t.ids])
    owners0 = {int(g): t.rank for t in tiles for g in t.ids}
    total = 0; moved = set()
    for chunk in range(6):
        tot, _ = _lockstep(tiles, o, 8, 8, f"rebin chunk {chunk}")
        total += tot
        tiles = tiling.rebin_local(tiles, ghost_capacity=256, min_width=2.5)
        owners = {int(g): t.rank for t in tiles for g in t.ids}
        assert sorted(owners) == sorted(owners0)                      # every body owned exactly once
        moved |= {g for g in owners if owners[g] != owners0[g]}
        for t in tiles:                                              # the restored state is the oracle's, bit for bit
            for sg, so in zip(t.state(), o.state()):
                assert np.array_equal(_bits(sg), _bits(so[t.ids]))
    assert total > 10000
    assert len(moved) >= 6 and len(moved & set(pushed.tolist())) >= 3, f"bodies that changed owner: {sorted(moved)}"
