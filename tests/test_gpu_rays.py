"""GPU: mgfb_intersections_batch (ray casts, collision.rs:163-373) against the reference's own
vectors and, bit for bit, against the oracle on seeded mixed batches of every shape kind."""
import numpy as np
import pytest

import mgf_b200
import oracle_lib
import ray_cases
from mgf_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_gpu_matches_reference_ray_vectors(ctx):
    fails = ray_cases.check_kat(lambda k, p, s: mgf_b200.intersections_batch(ctx, k, p, s))
    assert not fails, "\n".join(fails)


@pytest.mark.parametrize("seed", [1, 2])
def test_gpu_rays_and_segments_bit_exact_vs_oracle(ctx, seed):
    rays, segs, shapes = ray_cases.random_queries(20000, seed)
    for kind, parts in ((L.RAY, rays), (L.SEGMENT, segs)):
        og, hg = mgf_b200.intersections_batch(ctx, kind, parts, shapes)
        oo, ho = oracle_lib.intersections_batch(kind, parts, shapes)
        assert np.array_equal(hg, ho), f"hit flags differ for {int((hg != ho).sum())} queries (first {np.nonzero(hg != ho)[0][:5].tolist()})"
        assert 0.15 < ho.mean() < 0.9, ho.mean()          # the batch exercises both outcomes
        sel = ho == 1
        nan = np.isnan(oo["t"])                               # (NaN payloads are not compared)
        assert np.array_equal(_bits(og["t"][sel & ~nan]), _bits(oo["t"][sel & ~nan]))
        assert np.array_equal(_bits(og["p"][sel & ~nan]), _bits(oo["p"][sel & ~nan]))
        for k in range(7):                                    # every shape kind produced hits
            assert ho[shapes["kind"] == k].sum() > 0, k


def test_gpu_rays_edge_cases(ctx):
    out, hit = mgf_b200.intersections_batch(ctx, L.RAY, np.zeros((0, 6), np.float32), np.zeros(0, dtype=L.SHAPE_DTYPE))
    assert len(out) == 0 and len(hit) == 0
    with pytest.raises(mgf_b200.MgfbError) as e:
        mgf_b200.intersections_batch(ctx, 7, np.zeros((1, 6), np.float32), mgf_b200.sphere((0, 0, 0), 1.0))
    assert e.value.code == L.ERR_INVALID_ARG
    bad = mgf_b200.sphere((0, 0, 0), 1.0); bad["p"][0, 3] = 0.0      # assert!(radius > 0), geom.rs:300
    with pytest.raises(mgf_b200.MgfbError):
        mgf_b200.intersections_batch(ctx, L.RAY, np.zeros((1, 6), np.float32), bad)
    # zero direction: every divide is by zero -> None or NaN exactly as on the CPU
    z = np.array([[0.5, 0.5, 0.5, 0, 0, 0]], np.float32)
    for s in (mgf_b200.aabb((0, 0, 0), (1, 1, 1)), mgf_b200.plane((0, 1, 0), 0.0)):
        og, hg = mgf_b200.intersections_batch(ctx, L.RAY, z, s); oo, ho = oracle_lib.intersections_batch(L.RAY, z, s)
        assert hg.tolist() == ho.tolist()
