"""GPU: mgfb_gjk_batch / mgfb_separation_batch (csrc/gjk.cuh) against the reference's golden
vectors and, bit for bit, against the oracle on seeded random pairs of every shape-kind pair."""
import numpy as np
import pytest

import gjk_cases
import oracle_lib
import mgf_b200
from mgf_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_reference_gjk_vectors_on_device(ctx):
    n = gjk_cases.check_golden(lambda a, b: mgf_b200.gjk_batch(ctx, a, b), lambda a, b: mgf_b200.separation_batch(ctx, a, b))
    assert n == 9


def test_gjk_epa_batch_bit_exact_vs_oracle(ctx):
    a, b = gjk_cases.mixed_pairs(16 * 400)
    out, status, iters = mgf_b200.gjk_batch(ctx, a, b)
    oout, ohit, oiters = oracle_lib.gjk_batch(a, b)
    assert 2 not in status, "EPA polytope outgrew the device capacity on plain random shapes"
    assert np.array_equal(status, ohit)
    assert ohit.sum() > 500
    assert np.array_equal(iters, oiters)
    for f in ("a", "b", "n", "t"):
        g, o = out[f].reshape(len(a), -1), oout[f].reshape(len(a), -1)
        # bit-exact, except that a NaN only has to be a NaN (x86 and the GPU pick different quiet-NaN payloads;
        # the reference's barycentric division yields NaN on a degenerate closest face)
        bad = np.nonzero(((_bits(g) != _bits(o)) & ~(np.isnan(g) & np.isnan(o))).any(axis=1))[0]
        assert len(bad) == 0, f"{f} differs for {len(bad)} pairs, first {bad[:5].tolist()} kinds {a['kind'][bad[:5]]} x {b['kind'][bad[:5]]}"


def test_separation_batch_bit_exact_vs_oracle(ctx):
    a, b = gjk_cases.mixed_pairs(16 * 400, seed=21)
    sep, some = mgf_b200.separation_batch(ctx, a, b)
    osep, osome = oracle_lib.separation_batch(a, b)
    assert np.array_equal(some, osome) and 0 < osome.sum() < len(osome)
    assert np.array_equal(_bits(sep), _bits(osep))


def test_gjk_edge_cases(ctx):
    out, status, iters = mgf_b200.gjk_batch(ctx, np.zeros(0, L.SHAPE_DTYPE), np.zeros(0, L.SHAPE_DTYPE))
    assert len(out) == 0
    tri = mgf_b200.triangle((0, 0, 0), (1, 0, 0), (0, 1, 0))
    with pytest.raises(mgf_b200.MgfbError) as e:
        mgf_b200.gjk_batch(ctx, tri, mgf_b200.sphere((0, 0, 0), 1.0))
    assert e.value.code == L.ERR_INVALID_ARG
    # NaN input must neither hang nor crash, and ends like the oracle does
    bad = mgf_b200.sphere((float("nan"), 0, 0), 1.0)
    _, status, _ = mgf_b200.gjk_batch(ctx, bad, mgf_b200.sphere((0, 0, 0), 1.0))
    _, ostatus, _ = oracle_lib.gjk_batch(bad, mgf_b200.sphere((0, 0, 0), 1.0))
    assert int(status[0]) == int(ostatus[0])


def test_convex_mesh_pairs_bit_exact_vs_oracle(ctx):
    """ConvexMesh (mesh.rs:141-236) as a GJK / EPA shape against every other kind, both orders, and itself."""
    pool, a, b = gjk_cases.mesh_pairs(9 * 300)
    mgf_b200.convex_vertices_set(ctx, pool); oracle_lib.convex_vertices_set(pool)
    out, status, iters = mgf_b200.gjk_batch(ctx, a, b)
    oout, ohit, oiters = oracle_lib.gjk_batch(a, b)
    assert np.array_equal(status, ohit), np.nonzero(status != ohit)[0][:10]
    assert (ohit == 1).sum() > 200 and (ohit == 0).sum() > 200
    assert np.array_equal(iters, oiters)
    for f in ("a", "b", "n", "t"):
        g, o = out[f].reshape(len(a), -1), oout[f].reshape(len(a), -1)
        bad = np.nonzero(((_bits(g) != _bits(o)) & ~(np.isnan(g) & np.isnan(o))).any(axis=1))[0]
        assert len(bad) == 0, f"{f} differs for {len(bad)} pairs, first {bad[:5].tolist()} kinds {a['kind'][bad[:5]]} x {b['kind'][bad[:5]]}"
    sep, some = mgf_b200.separation_batch(ctx, a, b)
    osep, osome = oracle_lib.separation_batch(a, b)
    assert np.array_equal(some, osome) and np.array_equal(_bits(sep), _bits(osep))
    # a slice outside the pool is an error, not a wild read
    with pytest.raises(mgf_b200.MgfbError) as e:
        mgf_b200.gjk_batch(ctx, mgf_b200.convex_mesh(len(pool) - 2, 5), mgf_b200.sphere((0, 0, 0), 1.0))
    assert e.value.code == L.ERR_INVALID_ARG
