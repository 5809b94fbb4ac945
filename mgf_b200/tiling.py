"""One world tiled across GPUs: host-side driver of include/mgfb.h's tile API.

mgf itself is single-threaded; this is the scale-out the north star asks for.  Bodies are split
into slabs along x, one slab per GPU (one process per GPU under torchrun, or several contexts in
one process).  This module only PLANS (which body goes where) and WIRES (swaps the tiles'
memory descriptors between ranks through torch.distributed -- plumbing); every byte of the
per-step exchange moves GPU to GPU inside the CUDA kernels (mgf_b200/csrc/tile.cuh).
"""
import numpy as np

from . import _lib as L
from .api import World

INTERIOR_COLOURS = 32   # csrc/kernels.cuh TILE_INTERIOR_COLOURS: colours >= 32 are boundary colours


def slab_partition(x, nranks, min_width=0.0):
    """Split bodies into `nranks` slabs of (nearly) equal count by x coordinate.

    Returns a list of ascending global-id arrays, slab 0 leftmost.  Ties are broken by id, so the
    result is deterministic.  `min_width` > 0 widens slabs that would come out thinner than that (a
    tile must stay wider than its two ghost layers, mgfb.h): cuts are pushed right, bodies follow;
    a world narrower than nranks * min_width raises ValueError."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    n = len(x)
    order = np.lexsort((np.arange(n), x))
    cuts = [(n * r) // nranks for r in range(nranks + 1)]
    if min_width > 0.0 and n:
        xs = x[order]
        if xs[-1] - xs[0] < min_width * nranks:
            raise ValueError(f"the bodies span {xs[-1] - xs[0]:.3g} along x: too narrow for {nranks} tiles at least {min_width:.3g} wide")
        lo = xs[0]
        for r in range(1, nranks):
            # the cut coordinate is the first body of slab r; it must leave min_width to the left and to every slab still to come
            want = max(xs[min(cuts[r], n - 1)], lo + min_width)
            want = min(want, xs[-1] - min_width * (nranks - r))
            cuts[r] = int(np.searchsorted(xs, want, side="left"))
            lo = want
    return [np.sort(order[cuts[r]:cuts[r + 1]]).astype(np.uint32) for r in range(nranks)]


def shape_centres_x(shapes):
    """x of Shape::center for an array of mgfb_shape (sphere: c.x; capsule: a.x + d.x/2)."""
    p = shapes["p"]
    return np.where(shapes["kind"] == L.CAPSULE, p[:, 0] + 0.5 * p[:, 3], p[:, 0])


def all_gather_bytes(blob, group=None):
    """Every rank's `blob`, in rank order (torch.distributed; gloo or nccl -- host plumbing)."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, blob, group=group)
    return out


class TiledWorld:
    """Rank `rank`'s tile.  Usage (every rank):

        tw = TiledWorld(rank, nranks, device=local_rank)
        tw.add_bodies(owned_ids, shapes, mass, restitution, friction, world_force)   # GLOBAL arrays
        tw.set_terrain(verts, faces, x)
        tw.connect(all_gather_bytes)         # or connect_local([...]) for contexts of one process
        tw.step(dt, iters)
    """

    def __init__(self, rank, nranks, device=0, **cfg):
        self.rank, self.nranks = rank, nranks
        self.device, self.cfg = device, dict(cfg)
        self.world = World(device=device, **cfg)
        self.ids = np.zeros(0, np.uint32)
        self.desc = None
        self.params = None      # add_body arguments of the owned bodies (kept for rebin)
        self.terrain = None

    def add_bodies(self, owned_ids, shapes, mass, restitution, friction, world_force):
        ids = np.ascontiguousarray(owned_ids, dtype=np.uint32)
        assert len(ids) > 0, "a tile needs at least one body"
        assert np.all(np.diff(ids.astype(np.int64)) > 0), "owned ids must be ascending"
        n = len(shapes)
        pick = lambda a, shape: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float32), shape)[ids])
        self.add_owned(ids, np.ascontiguousarray(shapes[ids]), pick(mass, (n,)), pick(restitution, (n,)), pick(friction, (n,)), pick(world_force, (n, 3)))

    def add_owned(self, ids, shapes, mass, restitution, friction, world_force):
        """Like add_bodies, but the arrays hold ONLY this tile's bodies (big worlds: no rank ever
        materialises the whole scene)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        assert len(ids) == len(shapes) > 0 and np.all(np.diff(ids.astype(np.int64)) > 0)
        n = len(ids)
        f32 = lambda a, shape: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float32), shape))
        self.params = (np.ascontiguousarray(shapes, dtype=L.SHAPE_DTYPE), f32(mass, (n,)), f32(restitution, (n,)), f32(friction, (n,)), f32(world_force, (n, 3)))
        self.world.add_bodies(*self.params)
        self.world.set_gid(ids)
        self.ids = ids

    def set_terrain(self, verts, faces, x=(0.0, 0.0, 0.0)):
        self.terrain = (verts, faces, x)
        self.world.set_terrain(verts, faces, x)

    def close(self):
        self.world.ctx.close()

    def export(self, ghost_capacity=None):
        if ghost_capacity is None:
            ghost_capacity = max(4096, len(self.ids) // 2)
        self.desc = self.world.tile_export(ghost_capacity)
        return self.desc

    def connect(self, gather=all_gather_bytes, ghost_capacity=None):
        if self.desc is None and ghost_capacity is None:
            # one capacity for every tile (mgfb_tile_connect insists): the largest of the ranks' own defaults
            ghost_capacity = max(int(c) for c in gather(str(max(4096, len(self.ids) // 2)).encode()))
        descs = gather(self.desc or self.export(ghost_capacity))
        assert len(descs) == self.nranks
        self.world.tile_connect(self.rank, descs)

    def step(self, dt, iters=20, nsteps=1):
        return self.world.step(dt, iters, nsteps)

    def state(self):
        return self.world.state()

    def constraints(self):
        """(gid_a, gid_b, face, sub, colour) of the last step in solve order, and the number of
        interior rows (they come first; the rest are boundary rows)."""
        a, b, face, sub, colour = self.world.constraints()
        return a, b, face, sub, colour, int(np.count_nonzero(colour < INTERIOR_COLOURS))


def connect_local(tiles, ghost_capacity=None):
    """Wire tiles that live in ONE process (tests; several GPUs driven by one host thread each)."""
    if ghost_capacity is None:
        ghost_capacity = max(max(4096, len(t.ids) // 2) for t in tiles)   # one capacity for every tile
    descs = [t.export(ghost_capacity) for t in tiles]
    for t in tiles:
        t.world.tile_connect(t.rank, descs)


def executed_order(per_rank):
    """The sequential Gauss-Seidel order the tiled solver is equivalent to, per iteration: every
    rank's interior rows (rank by rank), then every rank's boundary rows.  `per_rank` is a list of
    TiledWorld.constraints() results; returns (a, b, face, sub) arrays in that global order."""
    parts = []
    for a, b, f, s, _, ni in per_rank:
        parts.append((a[:ni], b[:ni], f[:ni], s[:ni]))
    for a, b, f, s, _, ni in per_rank:
        parts.append((a[ni:], b[ni:], f[ni:], s[ni:]))
    return tuple(np.concatenate([p[k] for p in parts]) for k in range(4))


# ---------------------------------------------------------------- re-binning (SURVEY.md 8e "re-bin every k steps")
def _retile(parcels, rank, nranks, device, cfg, terrain, min_width):
    """Build rank `rank`'s tile of a fresh partition from every old tile's parcel = (ids, add_body arguments, snapshot)."""
    ids = np.concatenate([p[0] for p in parcels])
    order = np.argsort(ids, kind="stable")
    ids = ids[order]
    params = tuple(np.concatenate([p[1][k] for p in parcels])[order] for k in range(5))
    snap = {k: np.concatenate([p[2][k] for p in parcels])[order] for k in parcels[0][2]}
    parts = slab_partition(shape_centres_x(snap["colliders"]), nranks, min_width)    # indices into the id-sorted arrays
    mine = parts[rank]
    assert len(mine) > 0, "a tile would be left without bodies"
    t = TiledWorld(rank, nranks, device=device, **cfg)
    t.add_owned(ids[mine], *[a[mine] for a in params])
    if terrain is not None:
        t.set_terrain(*terrain)
    t.world.restore({k: v[mine] for k, v in snap.items()})
    return t


def rebin(tile, gather=all_gather_bytes, ghost_capacity=None, min_width=0.0):
    """Re-tile the world by the bodies' CURRENT x: every rank publishes its bodies (creation arguments + the state snapshot of
    mgfb_bodies_get_state / get_colliders / get_fat_bounds), the slabs are cut afresh and every rank builds a new tile holding
    what it now owns, state restored bit for bit (mgfb_bodies_set_state).  Host-side and collective (every rank calls it after
    the same step); meant for every few hundred steps.  Returns the new TiledWorld; the old one is closed."""
    parcels = gather((tile.ids, tile.params, tile.world.snapshot()))
    t = _retile(parcels, tile.rank, tile.nranks, tile.device, tile.cfg, tile.terrain, min_width)
    tile.close()
    t.connect(gather, ghost_capacity)
    return t


def rebin_local(tiles, ghost_capacity=None, min_width=0.0):
    """rebin() for tiles that live in ONE process (tests)."""
    parcels = [(t.ids, t.params, t.world.snapshot()) for t in tiles]
    new = [_retile(parcels, t.rank, t.nranks, t.device, t.cfg, t.terrain, min_width) for t in tiles]
    for t in tiles:
        t.close()
    connect_local(new, ghost_capacity)
    return new
