"""GPU: mgfb_compound_* (src/compound.rs:232-352 on the device) against the reference's own test (compound.rs:362-388) and,
bit for bit, against the oracle's Compound -- same contacts in the same callback order (the tree is grown like bvh.rs
grows it), same ray hits, closest points and bounds -- on seeded random compounds under random transforms."""
import numpy as np
import pytest

import mgf_b200
import oracle_lib
from compound_cases import from_arc_x_to_y, random_compound, random_particles, random_quats, random_rhs, reference_test_compound
from mgf_b200 import _lib as L
from mgf_b200 import api

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def test_reference_test_compound_on_device(ctx):
    c = mgf_b200.Compound(ctx, reference_test_compound())
    test_sphere = api.sphere((0.0, 8.0, 0.0), 1.0, v=(0.0, -1.5, 0.0))
    out, counts = c.contacts(test_sphere)
    assert counts[0] == 0
    c.set_transform((0, 0, 0), from_arc_x_to_y())
    out, counts = c.contacts(test_sphere)
    assert counts[0] >= 1
    last = out[0, counts[0] - 1]
    assert abs(last["t"] - 0.6666663) <= 1e-6 and np.allclose(last["a"], [0.0, 6.0, 0.0], atol=1e-6, rtol=1e-6)
    c.set_transform((0, 0, 0), (1, 0, 0, 0))
    rect = api.rectangle((0.0, -2.0, 0.0), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), 6.0, 6.0)
    rect["v"][0] = (0.0, 3.0, 0.0)
    assert c.contacts(rect)[1][0] >= 1


@pytest.mark.parametrize("ncomp,seed", [(1, 1), (2, 2), (3, 3), (7, 4), (24, 5), (150, 6)])
def test_compound_matches_oracle_bit_for_bit(ctx, ncomp, seed):
    rng = np.random.default_rng(seed)
    comps = random_compound(rng, ncomp, spread=3.0 + 0.05 * ncomp)
    g = mgf_b200.Compound(ctx, comps); o = oracle_lib.OracleCompound(comps)
    rhs = random_rhs(rng, 2000)
    rays, segs = random_particles(rng, 2000)
    pts = rng.uniform(-8, 8, (500, 3)).astype(np.float32)
    total = 0; hits = 0
    for disp, rot in [((0, 0, 0), (1, 0, 0, 0)), (rng.uniform(-2, 2, 3), random_quats(rng, 1)[0]), (rng.uniform(-1, 1, 3), random_quats(rng, 1)[0])]:
        g.set_transform(disp, rot); o.set_transform(disp, rot)
        for a, b in zip(g.bounds(), o.bounds()):
            assert np.array_equal(_bits(a), _bits(b))
        go, gc = g.contacts(rhs); oo, oc = o.contacts(rhs)
        assert np.array_equal(gc, oc), np.nonzero(gc != oc)[0][:5]
        bad = np.nonzero((_bits(go).reshape(len(rhs), -1) != _bits(oo).reshape(len(rhs), -1)).any(axis=1))[0]
        assert len(bad) == 0, (bad[:5], go[bad[0], :gc[bad[0]]], oo[bad[0], :oc[bad[0]]])
        total += int(gc.sum())
        for kind, parts in ((L.RAY, rays), (L.SEGMENT, segs)):
            gi, gh = g.intersections(kind, parts); oi, oh = o.intersections(kind, parts)
            assert np.array_equal(gh, oh) and np.array_equal(_bits(gi), _bits(oi))
            hits += int(gh.sum())
        assert np.array_equal(_bits(g.closest_points(pts)), _bits(o.closest_points(pts)))
    assert total > 100 and hits > 30
    g.close()


def test_compound_error_behaviour(ctx):
    with pytest.raises(mgf_b200.MgfbError) as e:
        mgf_b200.Compound(ctx, api.triangle((0, 0, 0), (1, 0, 0), (0, 1, 0)))        # not a Component
    assert e.value.code == L.ERR_INVALID_ARG
    with pytest.raises(mgf_b200.MgfbError) as e:
        mgf_b200.Compound(ctx, api.sphere((0, 0, 0), 0.0))                            # assert!(r > 0), geom.rs:300
    assert e.value.code == L.ERR_INVALID_ARG
    empty = mgf_b200.Compound(ctx, api.make_shapes(0))
    with pytest.raises(mgf_b200.MgfbError) as e:
        empty.bounds()                                                                # "BVH is empty", bvh.rs:263
    assert e.value.code == L.ERR_STATE
    assert empty.contacts(api.sphere((0, 0, 0), 1.0, v=(0, 1, 0)))[1][0] == 0
    c = mgf_b200.Compound(ctx, reference_test_compound())
    with pytest.raises(mgf_b200.MgfbError) as e:
        c.contacts(api.plane((0, 1, 0), 0.0))                                         # Plane is not BoundedBy<AABB>
    assert e.value.code == L.ERR_INVALID_ARG
