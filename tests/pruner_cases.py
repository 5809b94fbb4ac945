"""Shared inputs of the ContactPruner tests (manifold.rs:42-148): seeded groups of LocalContacts built to exercise every branch
of push -- earlier time replaces, later time is dropped, near points merge (either the a's or the b's), far points append."""
import numpy as np

from mgf_b200 import _lib as L


def random_groups(ngroups, seed=9, max_per_group=7):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(0, max_per_group + 1, ngroups)
    sizes[0] = 0                                            # an empty pruner: Manifold::from divides 0 by 0
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    n = int(offsets[-1])
    c = np.zeros(n, dtype=L.LOCAL_CONTACT_DTYPE)
    for g in range(ngroups):
        base_t = rng.uniform(0, 1)
        centre = rng.uniform(-3, 3, 3)
        for k in range(offsets[g], offsets[g + 1]):
            r = rng.random()
            # times: mostly equal within the 1e-6 window, sometimes clearly earlier / later
            t = base_t + (rng.uniform(-9e-7, 9e-7) if r < 0.7 else rng.uniform(-0.2, 0.2))
            # points: clustered (merge, threshold^2 = 0.5) or spread (append)
            spread = 0.3 if rng.random() < 0.5 else 2.5
            a = centre + rng.normal(size=3) * spread
            b = a + rng.normal(size=3) * (0.05 if rng.random() < 0.7 else 1.5)
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            c[k]["global"]["a"] = a; c[k]["global"]["b"] = b; c[k]["global"]["n"] = nrm; c[k]["global"]["t"] = max(t, 0.0)
            c[k]["local_a"] = rng.normal(size=3) * rng.uniform(0.2, 2.0); c[k]["local_b"] = rng.normal(size=3) * rng.uniform(0.2, 2.0)
    return c, offsets


def lc(a, b, n, t, la, lb):
    c = np.zeros(1, dtype=L.LOCAL_CONTACT_DTYPE)
    c["global"]["a"] = a; c["global"]["b"] = b; c["global"]["n"] = n; c["global"]["t"] = t; c["local_a"] = la; c["local_b"] = lb
    return c
