"""world_size-2 gloo check of the tiling HOST logic (no GPU): the slab plan covers every body once,
descriptors come back in rank order, and every rank derives the same executed order."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgf_b200 import scenes, tiling  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    bodies = scenes.pile_xyz(12, 3, 3, jitter=0.01, seed=2)
    x = tiling.shape_centres_x(bodies[0])
    parts = tiling.slab_partition(x, world)
    mine = parts[rank]
    # descriptors (here: stand-ins) return in rank order on every rank
    descs = tiling.all_gather_bytes(bytes([rank]) * 896)
    assert [d[0] for d in descs] == list(range(world)) and all(len(d) == 896 for d in descs)
    # the plan covers every body exactly once and slabs are ordered along x
    got = [None] * world
    dist.all_gather_object(got, mine.tolist())
    flat = sorted(i for g in got for i in g)
    assert flat == list(range(len(x)))
    for r in range(world - 1):
        assert x[np.array(got[r])].max() <= x[np.array(got[r + 1])].min()
    assert abs(len(got[0]) - len(got[-1])) <= 1
    # executed order: interior rows of every rank first, then boundary rows, identical everywhere
    rng = np.random.default_rng(rank)
    ni, nb = 5 + rank, 2 + rank
    a = rng.integers(0, 100, ni + nb).astype(np.uint32); b = rng.integers(-1, 100, ni + nb).astype(np.int32)
    z = np.zeros(ni + nb, np.uint32)
    colour = np.concatenate([np.sort(rng.integers(0, 9, ni)), 32 + np.sort(rng.integers(0, 3, nb))]).astype(np.uint32)
    rec = (a, b, z, z, colour, ni)
    allrec = [None] * world
    dist.all_gather_object(allrec, rec)
    oa, ob, _, _ = tiling.executed_order(allrec)
    want_a = np.concatenate([r[0][:r[5]] for r in allrec] + [r[0][r[5]:] for r in allrec])
    assert np.array_equal(oa, want_a) and len(ob) == sum(len(r[0]) for r in allrec)
    digest = [None] * world
    dist.all_gather_object(digest, oa.tobytes())
    assert all(d == digest[0] for d in digest)
    dist.barrier()
    if rank == 0:
        print("TILING_CPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
