"""GPU: mgfb_bvh_* (src/bvh.rs BVH<AABB, u32> on the device) against the oracle's incrementally built
tree and against brute force: same leaves reported by query and raytrace (as sets; the reference may
prune touching leaves, never the other way round), identical Intersections bit for bit, across inserts,
removals and slot reuse."""
import numpy as np
import pytest

import mgf_b200
import oracle_lib
from mgf_b200 import _lib as L
from test_bvh_oracle import _boxes, overlaps

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_bvh_query_matches_brute_force_and_contains_the_oracle_set(ctx):
    n = 3000
    boxes = _boxes(n, 3)
    g = mgf_b200.BVH(ctx); o = oracle_lib.OracleBVH()
    gi = g.insert(boxes, np.arange(n)); oi = [o.insert(b, v) for v, b in enumerate(boxes)]
    assert len(g) == n and sorted(gi.tolist()) == list(range(n))
    queries = _boxes(200, 4)
    alive = np.ones(n, bool)

    def check(tag):
        off, vals = g.query(queries)
        assert off[0] == 0 and off[-1] == len(vals)
        for q in range(len(queries)):
            got = vals[off[q]:off[q + 1]].tolist()
            brute = [v for v in np.nonzero(alive)[0].tolist() if overlaps(queries[q], boxes[v])]
            assert sorted(got) == brute, f"{tag}: query {q}"
            assert set(o.query(queries[q]).tolist()) <= set(got), f"{tag}: query {q} misses a leaf the reference reports"
    check("after insert")
    kill = np.arange(0, n, 3)
    g.remove(gi[kill]); [o.remove(oi[k]) for k in kill]
    alive[kill] = False
    assert len(g) == n - len(kill)
    check("after remove")
    # freed slots are reused; values are new
    extra = _boxes(500, 5)
    boxes2 = np.concatenate([boxes, extra]); alive = np.concatenate([alive, np.ones(500, bool)])
    gi2 = g.insert(extra, np.arange(n, n + 500)); [o.insert(b, n + v) for v, b in enumerate(extra)]
    assert set(gi2.tolist()) <= set(gi[kill].tolist())
    boxes = boxes2
    check("after reuse")
    b, v = g.get(gi2[7]); assert v == n + 7 and np.array_equal(b, extra[7])
    with pytest.raises(mgf_b200.MgfbError):
        g.remove(gi[kill][:1])         # freed first, so still free (slots are reused last-freed-first): no leaf there (the reference panics)


def test_bvh_raytrace_matches_the_oracle_bit_for_bit(ctx):
    n = 2000
    boxes = _boxes(n, 7)
    g = mgf_b200.BVH(ctx); o = oracle_lib.OracleBVH()
    g.insert(boxes, np.arange(n)); [o.insert(b, v) for v, b in enumerate(boxes)]
    rng = np.random.default_rng(8)
    origin = rng.uniform(-12, 12, (150, 3)).astype(np.float32)
    direction = (rng.uniform(-10, 10, (150, 3)).astype(np.float32) - origin)
    direction[::10, 1] = 0.0          # slab-parallel components
    rays = np.concatenate([origin, direction], axis=1).astype(np.float32)
    segs = np.concatenate([origin, origin + direction * np.float32(0.5)], axis=1).astype(np.float32)
    for kind, parts in ((L.RAY, rays), (L.SEGMENT, segs)):
        off, vals, hits = g.raytrace(kind, parts)
        total = 0
        for q in range(len(parts)):
            gv = vals[off[q]:off[q + 1]]; gh = hits[off[q]:off[q + 1]]
            ov, oh = o.raytrace(kind, parts[q])
            order_g = np.argsort(gv); order_o = np.argsort(ov)
            assert set(ov.tolist()) <= set(gv.tolist())
            # every leaf both report carries the identical Intersection
            common = np.isin(gv[order_g], ov)
            assert np.array_equal(gv[order_g][common], ov[order_o])
            assert np.array_equal(_bits(gh["t"][order_g][common]), _bits(oh["t"][order_o]))
            assert np.array_equal(_bits(gh["p"][order_g][common]), _bits(oh["p"][order_o]))
            # and what only the device reports is a true box hit (the reference pruned it above the leaf)
            single, hit1 = oracle_lib.intersections_batch(kind, np.repeat(parts[q][None], len(gv), 0),
                                                          np.concatenate([mgf_b200.aabb(boxes[v][:3], boxes[v][3:]) for v in gv]) if len(gv) else np.zeros(0, dtype=L.SHAPE_DTYPE)) if len(gv) else (None, np.zeros(0))
            assert hit1.all()
            total += len(gv)
        assert total > 300


def test_bvh_empty_and_capacity(ctx):
    g = mgf_b200.BVH(ctx)
    off, vals = g.query(_boxes(5, 1))
    assert off.tolist() == [0] * 6 and len(vals) == 0
    g.insert(_boxes(10, 2), np.arange(10))
    big = np.array([[0, 0, 0, 100, 100, 100]], np.float32)
    off, vals = g.query(big)
    assert sorted(vals.tolist()) == list(range(10))
    with pytest.raises(mgf_b200.MgfbError) as e:
        g.insert(np.array([[0, 0, 0, -1, 1, 1]], np.float32), [99])     # assert!(r >= 0), bounds.rs:125-127
    assert e.value.code == L.ERR_NAN_BOUNDS
    # a result set larger than the caller's arrays: MGFB_ERR_CAPACITY with the needed total
    import ctypes as C
    offsets = np.zeros(2, np.uint32); values = np.zeros(3, np.uint32); total = C.c_uint32()
    st = ctx.lib.mgfb_bvh_query_batch(g.h, L.ptr(big), 1, L.ptr(offsets), L.ptr(values), 3, C.byref(total))
    assert st == L.ERR_CAPACITY and total.value == 10


def test_remove_rejects_bad_batches_without_changing_anything(ctx):
    """A batch naming a free slot, an out-of-range slot or the same slot twice is refused BEFORE anything is removed."""
    import mgf_b200
    from mgf_b200 import _lib as L
    rng = np.random.default_rng(11)
    boxes = np.concatenate([rng.uniform(-5, 5, (64, 3)), rng.uniform(0.1, 1.0, (64, 3))], axis=1).astype(np.float32)
    t = mgf_b200.BVH(ctx)
    idx = t.insert(boxes, np.arange(64, dtype=np.uint32))
    everything = np.array([[0, 0, 0, 100, 100, 100]], np.float32)
    before = sorted(t.query(everything)[1].tolist())
    for bad in ([idx[3], idx[7], idx[3]], [idx[1], 10 ** 6], ):
        with pytest.raises(mgf_b200.MgfbError) as e:
            t.remove(np.array(bad, dtype=np.uint32))
        assert e.value.code == L.ERR_INVALID_ARG
        assert len(t) == 64 and sorted(t.query(everything)[1].tolist()) == before
    t.remove(np.array([idx[3], idx[7]], dtype=np.uint32))
    assert len(t) == 62 and sorted(t.query(everything)[1].tolist()) == [v for v in before if v not in (3, 7)]
    with pytest.raises(mgf_b200.MgfbError):
        t.remove(np.array([idx[3]], dtype=np.uint32))      # already free
    assert len(t) == 62
