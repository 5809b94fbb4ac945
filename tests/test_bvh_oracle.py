"""CPU: the oracle's BVH (src/bvh.rs restated: incremental insert, remove, balance, query, raytrace)
against brute force over the same boxes.  The reference's own BVH test (bvh.rs:514) is in oracle/kat.cpp."""
import numpy as np

import oracle_lib
from mgf_b200 import _lib as L


def _boxes(n, seed):
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(-10, 10, (n, 3)), rng.uniform(0.1, 1.5, (n, 3))], axis=1).astype(np.float32)


def overlaps(a, b):   # collision.rs:22-29, in f32
    return bool(np.all(np.abs(a[:3] - b[:3]) <= (a[3:] + b[3:])))


def test_oracle_bvh_query_is_the_brute_force_set_up_to_touching():
    boxes = _boxes(400, 1)
    t = oracle_lib.OracleBVH()
    idx = [t.insert(b, v) for v, b in enumerate(boxes)]
    for k in range(0, 400, 7):       # remove some, the tree rebalances
        t.remove(idx[k])
    alive = [v for v in range(400) if v % 7 != 0]
    for q in _boxes(60, 2):
        got = sorted(t.query(q).tolist())
        brute = [v for v in alive if overlaps(q, boxes[v])]
        assert set(got) <= set(brute) and len(got) == len(set(got))
        # the tree may only prune leaves that merely touch (rounded parent unions): strict overlaps are always found
        strict = [v for v in alive if np.all(np.abs(q[:3] - boxes[v][:3]) < (q[3:] + boxes[v][3:]) * np.float32(0.999))]
        assert set(strict) <= set(got)
