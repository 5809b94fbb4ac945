"""GPU parity of the narrowphase kernels, through the C ABI (mgfb_contacts_batch):
(1) the reference's own unit-test vectors, (2) bit-exact agreement with the oracle on seeded
random and adversarial batches for every pair kind."""
import numpy as np
import pytest

import kat_check
import oracle_lib
import mgf_b200
from mgf_b200 import _lib as L

pytestmark = pytest.mark.gpu


def test_reference_vectors_on_device(ctx):
    failures = kat_check.check_cases(lambda k, r, a: mgf_b200.contacts_batch(ctx, k, r, a))
    assert not failures, "\n".join(failures)


def _rand_shapes(rng, kind, n, scale=2.0, vel=1.5):
    s = np.zeros(n, dtype=L.SHAPE_DTYPE)
    s["kind"] = kind
    p = s["p"]
    if kind == L.SPHERE:
        p[:, 0:3] = rng.uniform(-scale, scale, (n, 3)); p[:, 3] = rng.uniform(0.2, 1.2, n)
    elif kind == L.CAPSULE:
        p[:, 0:3] = rng.uniform(-scale, scale, (n, 3)); p[:, 3:6] = rng.uniform(-1.5, 1.5, (n, 3)); p[:, 6] = rng.uniform(0.2, 1.0, n)
    elif kind == L.TRIANGLE:
        c = rng.uniform(-scale, scale, (n, 3))
        for k in range(3):
            p[:, 3 * k:3 * k + 3] = c + rng.uniform(-2.0, 2.0, (n, 3))
    elif kind == L.RECTANGLE:
        p[:, 0:3] = rng.uniform(-scale, scale, (n, 3))
        u0 = rng.normal(size=(n, 3)); u0 /= np.linalg.norm(u0, axis=1, keepdims=True)
        t = rng.normal(size=(n, 3)); u1 = np.cross(u0, t); u1 /= np.linalg.norm(u1, axis=1, keepdims=True)
        p[:, 3:6] = u0; p[:, 6:9] = u1; p[:, 9:11] = rng.uniform(0.5, 2.5, (n, 2))
    elif kind == L.PLANE:
        nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        p[:, 0:3] = nrm; p[:, 3] = rng.uniform(-1, 1, n)
    s["v"] = rng.uniform(-vel, vel, (n, 3))
    return s


def _special_cases(recv, arg, rng):
    """Adversarial rows: zero velocity, coincident centres, axis-aligned / parallel capsules,
    velocity exactly along an edge, large sweeps."""
    n = len(arg)
    arg["v"][0:n // 16] = 0.0
    if arg["kind"][0] == recv["kind"][0]:
        k = slice(n // 16, n // 8)
        arg["p"][k] = recv["p"][k]                       # coincident shapes
        arg["v"][n // 16: n // 16 + n // 32] = 0.0       # ... and not moving
    if arg["kind"][0] == L.CAPSULE and recv["kind"][0] == L.CAPSULE:
        k = slice(n // 8, n // 4)
        arg["p"][k, 3:6] = recv["p"][k, 3:6] * rng.choice([-2.0, -1.0, 0.5, 1.0, 3.0], (n // 4 - n // 8, 1)).astype(np.float32)
    if arg["kind"][0] == L.CAPSULE and recv["kind"][0] == L.TRIANGLE:
        k = slice(n // 8, n // 4)                         # capsule parallel to edge ab, integer-ish coords
        recv["p"][k] = np.round(recv["p"][k])
        e = recv["p"][k, 3:6] - recv["p"][k, 0:3]
        arg["p"][k, 3:6] = e * rng.choice([-1.0, 0.5, 1.0, 2.0], (n // 4 - n // 8, 1)).astype(np.float32)
        arg["p"][k, 0:3] = np.round(arg["p"][k, 0:3])
        k2 = slice(n // 4, n // 4 + n // 8)               # capsule lying parallel to the face plane
        recv["p"][k2, 1] = 0.0; recv["p"][k2, 4] = 0.0; recv["p"][k2, 7] = 0.0
        arg["p"][k2, 4] = 0.0; arg["p"][k2, 1] = rng.uniform(0.5, 2.0, n // 8); arg["v"][k2] = [0.0, -2.5, 0.0]
    arg["v"][-n // 16:] *= 8.0
    return recv, arg


KINDS = [
    (L.SPHERE_X_MSPHERE, L.SPHERE, L.SPHERE), (L.CAPSULE_X_MSPHERE, L.CAPSULE, L.SPHERE),
    (L.SPHERE_X_MCAPSULE, L.SPHERE, L.CAPSULE), (L.CAPSULE_X_MCAPSULE, L.CAPSULE, L.CAPSULE),
    (L.PLANE_X_MSPHERE, L.PLANE, L.SPHERE), (L.PLANE_X_MCAPSULE, L.PLANE, L.CAPSULE),
    (L.TRI_X_MSPHERE, L.TRIANGLE, L.SPHERE), (L.TRI_X_MCAPSULE, L.TRIANGLE, L.CAPSULE),
    (L.RECT_X_MSPHERE, L.RECTANGLE, L.SPHERE), (L.RECT_X_MCAPSULE, L.RECTANGLE, L.CAPSULE),
]


def _assert_bit_equal(got, want, what):
    # bit-exact, except that NaN payload/sign is not part of IEEE arithmetic (x86 SSE makes
    # 0xFFC00000, the GPU 0x7FFFFFFF): degenerate inputs (zero-area triangle) must give NaN on both.
    gf = np.ascontiguousarray(got).view(np.float32); wf = np.ascontiguousarray(want).view(np.float32)
    g = gf.view(np.uint32); w = wf.view(np.uint32)
    bad = np.nonzero(((g != w) & ~(np.isnan(gf) & np.isnan(wf))).reshape(len(got), -1).any(axis=1))[0]
    assert len(bad) == 0, f"{what}: {len(bad)} rows differ, first {bad[:5].tolist()}"


@pytest.mark.parametrize("kind,rk,ak", KINDS)
def test_batch_bit_exact_vs_oracle(ctx, kind, rk, ak):
    rng = np.random.default_rng(1000 + kind)
    n = 20000
    recv = _rand_shapes(rng, rk, n); arg = _rand_shapes(rng, ak, n)
    recv["v"] = 0.0
    recv, arg = _special_cases(recv, arg, rng)
    out_g, cnt_g = mgf_b200.contacts_batch(ctx, kind, recv, arg)
    out_o, cnt_o = oracle_lib.contacts_batch(kind, recv, arg)
    assert np.array_equal(cnt_g, cnt_o), f"hit counts differ in {np.count_nonzero(cnt_g != cnt_o)} rows"
    assert cnt_o.sum() > n // 50, "test batch has too few hits to mean anything"
    mask = np.arange(2)[None, :] < cnt_o[:, None]
    _assert_bit_equal(out_g[mask], out_o[mask], f"pair kind {kind}")


@pytest.mark.parametrize("rk,ak", [(L.SPHERE, L.SPHERE), (L.CAPSULE, L.SPHERE), (L.SPHERE, L.CAPSULE), (L.CAPSULE, L.CAPSULE)])
def test_body_pairs_local_contacts_bit_exact(ctx, rk, ak):
    rng = np.random.default_rng(7 + rk * 2 + ak)
    n = 20000
    recv = _rand_shapes(rng, rk, n); arg = _rand_shapes(rng, ak, n)
    recv, arg = _special_cases(recv, arg, rng)
    og, cg, lg = mgf_b200.contacts_batch(ctx, L.MCOMP_X_MCOMP, recv, arg, want_local=True)
    oo, co, lo = oracle_lib.contacts_batch(L.MCOMP_X_MCOMP, recv, arg, want_local=True)
    assert np.array_equal(cg, co)
    mask = np.arange(2)[None, :] < co[:, None]
    _assert_bit_equal(lg[mask], lo[mask], "local contacts")


@pytest.mark.parametrize("rk", [L.SPHERE, L.CAPSULE])
def test_body_vs_terrain_triangle_bit_exact(ctx, rk):
    rng = np.random.default_rng(70 + rk)
    n = 20000
    recv = _rand_shapes(rng, rk, n); tri = _rand_shapes(rng, L.TRIANGLE, n)
    tri["v"] = rng.uniform(-1, 1, (n, 3))   # carries mesh.x
    og, cg, lg = mgf_b200.contacts_batch(ctx, L.MCOMP_X_TRI, recv, tri, want_local=True)
    oo, co, lo = oracle_lib.contacts_batch(L.MCOMP_X_TRI, recv, tri, want_local=True)
    assert np.array_equal(cg, co)
    assert co.sum() > 100
    mask = np.arange(2)[None, :] < co[:, None]
    _assert_bit_equal(lg[mask], lo[mask], "terrain local contacts")


def test_empty_batch_and_bad_kind(ctx):
    e = np.zeros(0, dtype=L.SHAPE_DTYPE)
    out, cnt = mgf_b200.contacts_batch(ctx, L.SPHERE_X_MSPHERE, e, e)
    assert len(cnt) == 0
    s = mgf_b200.sphere((0, 0, 0), 1.0)
    with pytest.raises(mgf_b200.MgfbError) as err:
        mgf_b200.contacts_batch(ctx, L.CAPSULE_X_MSPHERE, s, s)   # receiver is not a capsule
    assert err.value.code == L.ERR_INVALID_ARG


def test_contact_pruner_on_device_bit_exact_and_feeds_the_solver(ctx):
    """mgfb_manifolds_prune = ContactPruner::push + Manifold::from (manifold.rs:42-148) on the device, against the oracle on
    seeded groups that hit every branch of push; the manifolds it returns then go through mgfb_solver_solve unchanged."""
    import pruner_cases
    contacts, offsets = pruner_cases.random_groups(5000)
    g = mgf_b200.manifolds_prune(ctx, contacts, offsets); o = oracle_lib.manifolds_prune(contacts, offsets)
    assert np.array_equal(g["ncontacts"], o["ncontacts"])
    assert set(np.unique(o["ncontacts"]).tolist()) >= {0, 1, 2, 3, 4}
    for k in ("time", "normal", "tangent", "local_a", "local_b"):
        a, b = g[k].reshape(len(offsets) - 1, -1), o[k].reshape(len(offsets) - 1, -1)
        bad = np.nonzero(((a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b))).any(axis=1))[0]
        assert len(bad) == 0, (k, bad[:5], a[bad[:2]], b[bad[:2]])
