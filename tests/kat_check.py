"""Shared checker for tests/golden/reference_kat.json (the reference's own unit-test vectors)."""
import json
import os

import numpy as np

from mgf_b200 import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))
F32_EPS = np.float32(1.1920929e-07)


def load_cases():
    with open(os.path.join(HERE, "golden", "reference_kat.json")) as f:
        return json.load(f)["cases"]


def to_shape(d):
    s = np.zeros(1, dtype=L.SHAPE_DTYPE)
    s["kind"] = d["kind"]
    s["p"][0, :len(d["p"])] = np.asarray(d["p"], dtype=np.float32)
    s["v"][0] = np.asarray(d["v"], dtype=np.float32)
    return s


def relative_eq(a, b, eps):
    """approx 0.3 relative_eq with max_relative = f32::EPSILON, all in f32."""
    a = np.float32(a); b = np.float32(b)
    if a == b:
        return True
    if np.isinf(a) or np.isinf(b):
        return False
    d = np.float32(abs(np.float32(a - b)))
    if d <= np.float32(eps):
        return True
    return d <= np.float32(max(abs(a), abs(b))) * F32_EPS


def check_cases(batch_fn):
    """batch_fn(pair_kind, recv, arg) -> (contacts[n,2], counts[n]); returns list of failure strings."""
    failures = []
    for c in load_cases():
        out, counts = batch_fn(c["pair_kind"], to_shape(c["recv"]), to_shape(c["arg"]))[:2]
        cnt = int(counts[0])
        if c["count"] is not None and cnt != c["count"]:
            failures.append(f'{c["name"]} ({c["src"]}): count {cnt} != {c["count"]}')
            continue
        if c["min_count"] is not None and cnt < c["min_count"]:
            failures.append(f'{c["name"]} ({c["src"]}): count {cnt} < {c["min_count"]}')
            continue
        for e in c["expect"]:
            i = e["i"] if e["i"] >= 0 else cnt - 1   # -1 = last_contact
            got = np.atleast_1d(out[0, i][e["field"]]).astype(np.float32)
            want = np.atleast_1d(np.asarray(e["val"], dtype=np.float32))
            if e["mode"] == "eq":
                ok = bool(np.all(got == want))
            elif e["mode"] == "rel":
                ok = all(relative_eq(g, w, e["eps"]) for g, w in zip(got, want))
            elif e["mode"] == "one_minus_lt":
                ok = bool(np.float32(1.0) - got[0] < np.float32(e["val"]))
            else:
                raise ValueError(e["mode"])
            if not ok:
                failures.append(f'{c["name"]} ({c["src"]}): {e["field"]}[{i}] = {got.tolist()} want {want.tolist()} ({e["mode"]})')
    return failures
