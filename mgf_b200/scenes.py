"""Synthetic scenes of BASELINE.json's configs, generated on the host in f32.

All generators are parameterisations of the reference demo's own loops
(mgf_demo/balls.rs:74-96, capsules.rs:67-95, terrain world.rs:118-150); randomness is one
documented 64-bit LCG so that the CUDA path and the oracle see identical bits.
"""
import numpy as np

from . import _lib as L

F = np.float32


def lcg_uniform(n, seed):
    """n floats in [0,1): x <- x*6364136223846793005 + 1442695040888963407 (mod 2^64), top 24 bits."""
    out = np.empty(n, dtype=np.float32)
    x = np.uint64(seed)
    a, c = np.uint64(6364136223846793005), np.uint64(1442695040888963407)
    with np.errstate(over="ignore"):
        for i in range(n):
            x = x * a + c
            out[i] = F(int(x >> np.uint64(40)) / float(1 << 24))
    return out


def lcg_uniform_fast(n, seed):
    """Same sequence as lcg_uniform, vectorised by jumping the LCG (used for big scenes)."""
    # x_k = A_k * x_0 + C_k with A_k, C_k built by doubling; compute blockwise.
    M = (1 << 64) - 1
    a, c = 6364136223846793005, 1442695040888963407
    out = np.empty(n, dtype=np.float32)
    x = seed & M
    # generate sequentially in Python ints but in chunks using numpy uint64 wraparound
    B = 1 << 16
    As = np.empty(B, dtype=np.uint64); Cs = np.empty(B, dtype=np.uint64)
    ak, ck = 1, 0
    for i in range(B):
        ak = (ak * a) & M
        ck = (ck * a + c) & M
        As[i] = ak; Cs[i] = ck
    pos = 0
    while pos < n:
        m = min(B, n - pos)
        with np.errstate(over="ignore"):
            xs = As[:m] * np.uint64(x) + Cs[:m]
        out[pos:pos + m] = ((xs >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(np.float32)
        x = int(xs[m - 1])
        pos += m
    return out


def box_terrain(hx=10.0, wall_h=10.0, hz=10.0, pos=(0.0, -10.0, 0.0)):
    """world.rs:118-150: open box, 8 verts / 10 faces; winding decides the solid side."""
    V = np.array([[-hx, 0, -hz], [-hx, 0, hz], [hx, 0, hz], [hx, 0, -hz],
                  [-hx, wall_h, -hz], [-hx, wall_h, hz], [hx, wall_h, hz], [hx, wall_h, -hz]], dtype=np.float32)
    Fc = np.array([[0, 1, 3], [1, 2, 3], [0, 5, 1], [0, 4, 5], [0, 3, 7], [0, 7, 4], [2, 6, 3], [3, 6, 7], [1, 5, 2], [2, 5, 6]],
                  dtype=np.uint32)
    return V, Fc, np.asarray(pos, dtype=np.float32)


def heightfield_terrain(nq=100, size=100.0, y0=-10.0, amp=0.5, freq=0.3):
    """nq x nq quads (2*nq^2 triangles), y = amp*sin(freq x)*cos(freq z) + y0, CCW seen from above
    so normals point up (world.rs:137-141)."""
    xs = np.linspace(-size / 2, size / 2, nq + 1, dtype=np.float32)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    Y = (F(amp) * np.sin(F(freq) * X) * np.cos(F(freq) * Z) + F(y0)).astype(np.float32)
    V = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(np.float32)
    idx = lambda i, j: i * (nq + 1) + j
    faces = []
    i, j = np.meshgrid(np.arange(nq), np.arange(nq), indexing="ij")
    i = i.ravel(); j = j.ravel()
    a, b, c, d = idx(i, j), idx(i, j + 1), idx(i + 1, j + 1), idx(i + 1, j)
    # (x_i,z_j)=a, (x_i,z_j+1)=b, (x_i+1,z_j+1)=c, (x_i+1,z_j)=d ; normal up: (b-a)x(c-a) = z-hat x (x+z) -> +y
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.uint32)
    return V, faces, np.zeros(3, dtype=np.float32)


def _grid_positions(num, rad, y_base=10.0):
    """balls.rs:74-92 / capsules.rs:75-91 loop order i (x), j (y), k (z)."""
    shift = F(2.5) * F(rad)
    centerx = shift * F(num) / F(2.0)
    centery = shift * F(num) / F(2.0)
    ii, jj, kk = np.meshgrid(np.arange(num), np.arange(num), np.arange(num), indexing="ij")
    ii = ii.ravel().astype(np.float32); jj = jj.ravel().astype(np.float32); kk = kk.ravel().astype(np.float32)
    x = ii * F(2.5) * F(rad) - centerx
    y = F(y_base) + jj * F(2.5) * F(rad) + centery * F(2.0)
    z = kk * F(2.5) * F(rad) - centerx
    return np.stack([x, y, z], axis=1).astype(np.float32)


def balls_scene(num=11, extra=0, jitter=0.0, seed=1, rad=0.5, demo_extra_ball=False):
    """Spheres r=rad on the demo's num^3 lattice (+ `extra` more on a top layer), optional
    LCG jitter in [-jitter, jitter] per coordinate.  Returns (shapes, mass, rest, fric, force)."""
    pos = _grid_positions(num, rad)
    if extra:
        side = int(np.ceil(np.sqrt(extra)))
        shift = F(2.5) * F(rad)
        top = pos[:, 1].max() + shift
        k = np.arange(extra)
        ex = np.stack([(k % side).astype(np.float32) * shift - shift * F(side) / F(2.0),
                       np.full(extra, top, dtype=np.float32) + (k // (side * side)).astype(np.float32) * shift,
                       ((k // side) % side).astype(np.float32) * shift - shift * F(side) / F(2.0)], axis=1).astype(np.float32)
        pos = np.concatenate([pos, ex])
    if demo_extra_ball:  # balls.rs:94-96
        pos = np.concatenate([pos, np.array([[0.0, 130.0, 0.0]], dtype=np.float32)])
    n = len(pos)
    if jitter:
        u = lcg_uniform_fast(3 * n, seed).reshape(n, 3)
        pos = (pos + (u * F(2.0) - F(1.0)) * F(jitter)).astype(np.float32)
    shapes = np.zeros(n, dtype=L.SHAPE_DTYPE)
    shapes["kind"] = L.SPHERE
    shapes["p"][:, 0:3] = pos
    shapes["p"][:, 3] = rad
    return (shapes, np.full(n, 1.0, np.float32), np.full(n, 0.3, np.float32), np.full(n, 0.6, np.float32),
            np.tile(np.array([0.0, -9.8, 0.0], np.float32), (n, 1)))


def capsules_scene(num=11, jitter=0.0, seed=1, count=None):
    """capsules.rs:67-95: Capsule a=(-0.5,0,0) d=(1,0,0) r=1 moved so its centre is on the
    lattice with rad=2.0 (spacing 5.0)."""
    pos = _grid_positions(num, 2.0)
    if count is not None:
        pos = pos[:count]
    n = len(pos)
    if jitter:
        u = lcg_uniform_fast(3 * n, seed).reshape(n, 3)
        pos = (pos + (u * F(2.0) - F(1.0)) * F(jitter)).astype(np.float32)
    shapes = np.zeros(n, dtype=L.SHAPE_DTYPE)
    shapes["kind"] = L.CAPSULE
    # set_pos: disp = p - center(); a += disp ; center = a + d*0.5 = (0,0,0) initially
    a0 = np.array([-0.5, 0.0, 0.0], np.float32)
    shapes["p"][:, 0:3] = a0 + (pos - np.zeros(3, np.float32))
    shapes["p"][:, 3:6] = np.array([1.0, 0.0, 0.0], np.float32)
    shapes["p"][:, 6] = 1.0
    return (shapes, np.full(n, 1.0, np.float32), np.full(n, 0.3, np.float32), np.full(n, 0.6, np.float32),
            np.tile(np.array([0.0, -9.8, 0.0], np.float32), (n, 1)))


def pile_scene(num=46, extra=2664, jitter=0.01, seed=1, rad=0.5, squeeze=0.98, floor_y=-10.0):
    """The C2 body set as a pile that is already in contact: the balls.rs lattice (same loop
    order, same `extra` top layer, same LCG jitter) with the spacing squeezed from 2.5*rad to
    squeeze*2*rad and the bottom layer resting on the floor, so contact-constraint work is
    present from step 0 (the free-fall version needs ~600 steps before the pile forms)."""
    shapes, mass, rest, fric, force = balls_scene(num, extra, 0.0, seed, rad)
    pos = shapes["p"][:, 0:3].astype(np.float32)
    scale = F(squeeze * 2.0 * rad) / (F(2.5) * F(rad))
    pos = (pos * scale).astype(np.float32)
    pos[:, 1] += F(floor_y + rad) - pos[:, 1].min()
    n = len(pos)
    if jitter:
        u = lcg_uniform_fast(3 * n, seed).reshape(n, 3)
        pos = (pos + (u * F(2.0) - F(1.0)) * F(jitter)).astype(np.float32)
    shapes["p"][:, 0:3] = pos
    return shapes, mass, rest, fric, force


def pile_xyz(nx, ny, nz, jitter=0.01, seed=1, rad=0.5, squeeze=0.98, floor_y=-10.0):
    """A pile already in contact on an nx x ny x nz lattice (x-major order, like balls.rs:80-92), used
    for worlds tiled along x: ids ascend with x, so a slab is a contiguous id range."""
    step = F(squeeze * 2.0 * rad)
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pos = np.stack([(ii.ravel().astype(np.float32) - F(nx - 1) / F(2.0)) * step,
                    F(floor_y + rad) + jj.ravel().astype(np.float32) * step,
                    (kk.ravel().astype(np.float32) - F(nz - 1) / F(2.0)) * step], axis=1).astype(np.float32)
    n = len(pos)
    if jitter:
        u = lcg_uniform_fast(3 * n, seed).reshape(n, 3)
        pos = (pos + (u * F(2.0) - F(1.0)) * F(jitter)).astype(np.float32)
    shapes = np.zeros(n, dtype=L.SHAPE_DTYPE)
    shapes["kind"] = L.SPHERE
    shapes["p"][:, 0:3] = pos
    shapes["p"][:, 3] = rad
    return (shapes, np.full(n, 1.0, np.float32), np.full(n, 0.3, np.float32), np.full(n, 0.6, np.float32),
            np.tile(np.array([0.0, -9.8, 0.0], np.float32), (n, 1)))


def capsule_pile(nx, ny, nz, jitter=0.02, seed=1, floor_y=-10.0):
    """capsules.rs:69-73 capsules (a = centre - (0.5,0,0), d = (1,0,0), r = 1) packed into a pile that
    is already in contact: spacing 2.9 along the axis, 1.96 across, bottom layer on the floor."""
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pos = np.stack([(ii.ravel().astype(np.float32) - F(nx - 1) / F(2.0)) * F(2.9),
                    F(floor_y + 1.0) + jj.ravel().astype(np.float32) * F(1.96),
                    (kk.ravel().astype(np.float32) - F(nz - 1) / F(2.0)) * F(1.96)], axis=1).astype(np.float32)
    n = len(pos)
    if jitter:
        u = lcg_uniform_fast(3 * n, seed).reshape(n, 3)
        pos = (pos + (u * F(2.0) - F(1.0)) * F(jitter)).astype(np.float32)
    shapes = np.zeros(n, dtype=L.SHAPE_DTYPE)
    shapes["kind"] = L.CAPSULE
    shapes["p"][:, 0:3] = pos + np.array([-0.5, 0.0, 0.0], np.float32)
    shapes["p"][:, 3:6] = np.array([1.0, 0.0, 0.0], np.float32)
    shapes["p"][:, 6] = 1.0
    return (shapes, np.full(n, 1.0, np.float32), np.full(n, 0.3, np.float32), np.full(n, 0.6, np.float32),
            np.tile(np.array([0.0, -9.8, 0.0], np.float32), (n, 1)))


def config_c3(scale=1.0):
    """BASELINE configs[2]: 50 000 capsules over a static 20 000-triangle mesh floor (Capsule x Triangle
    narrowphase).  The floor is SURVEY 8(d)'s height field y = 0.5 sin(0.3x) cos(0.3z) - 10; the capsules
    sit on it as a 50 x 20 x 50 pile so the capsule-triangle and capsule-capsule kernels work from step 0.
    `scale` < 1 shrinks the body count (parity tests on small boxes)."""
    nx, ny, nz = max(2, int(round(50 * scale))), max(2, int(round(20 * scale))), max(2, int(round(50 * scale)))
    bodies = capsule_pile(nx, ny, nz)
    size = max(2.9 * nx, 1.96 * nz) + 12.0
    shapes = bodies[0]
    shapes["p"][:, 1] += 0.6   # clear the crests of the height field
    return bodies, heightfield_terrain(nq=100, size=size, y0=-10.0, amp=0.5, freq=0.3), 20


def config_c5(scale=1.0):
    """BASELINE configs[4] restated (RigidBodyVec holds single Components and Compound x Mesh has
    no impl in the reference): a MIXED body set -- spheres r = 0.5 and capsules d = (1,0,0) r = 0.4
    interleaved -- 200 000 bodies over a 200 000-triangle height field.  The GJK half of the config is the
    separate static-pair batch (tests/gjk_cases.py)."""
    nx, ny, nz = max(2, int(round(100 * scale))), max(2, int(round(20 * scale))), max(2, int(round(100 * scale)))
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pos = np.stack([(ii.ravel().astype(np.float32) - F(nx - 1) / F(2.0)) * F(1.38),
                    F(-10.0 + 0.9) + jj.ravel().astype(np.float32) * F(0.88),
                    (kk.ravel().astype(np.float32) - F(nz - 1) / F(2.0)) * F(0.88)], axis=1).astype(np.float32)
    n = len(pos)
    u = lcg_uniform_fast(3 * n, 5).reshape(n, 3)
    pos = (pos + (u * F(2.0) - F(1.0)) * F(0.01)).astype(np.float32)
    shapes = np.zeros(n, dtype=L.SHAPE_DTYPE)
    is_cap = ((ii + jj + kk).ravel() % 2) == 1
    shapes["kind"] = np.where(is_cap, L.CAPSULE, L.SPHERE)
    shapes["p"][:, 0:3] = pos
    shapes["p"][:, 3] = 0.5
    shapes["p"][is_cap, 0] -= 0.5
    shapes["p"][is_cap, 3:6] = np.array([1.0, 0.0, 0.0], np.float32)
    shapes["p"][is_cap, 6] = 0.4
    bodies = (shapes, np.full(n, 1.0, np.float32), np.full(n, 0.3, np.float32), np.full(n, 0.6, np.float32),
              np.tile(np.array([0.0, -9.8, 0.0], np.float32), (n, 1)))
    nq = max(4, int(round(316 * scale)))   # 2 * 316^2 = 199 712 triangles
    return bodies, heightfield_terrain(nq=nq, size=max(1.38 * nx, 0.88 * nz) + 10.0, y0=-10.0, amp=0.5, freq=0.3), 20


def tiled_pile(ntiles, tile=0, nx=50, ny=40, nz=50):
    """Tile `tile` of ONE pile of ntiles*nx x ny x nz spheres in one box, split along x (weak
    scaling: nx*ny*nz = 100 000 bodies per tile by default, the C2 body count).  Tiles abut at the
    lattice spacing; jitter comes from the LCG seeded 1 + tile.  Returns the tile's bodies, their
    global ids (tile-major, ascending with x) and the terrain of the whole world."""
    shapes, mass, rest, fric, force = pile_xyz(nx, ny, nz, 0.01, 1 + tile)
    pitch = F(nx * 0.98)
    shapes["p"][:, 0] += (F(tile) - F(ntiles - 1) / F(2.0)) * pitch
    n = len(shapes)
    ids = (np.arange(n, dtype=np.uint32) + np.uint32(tile * n))
    return (shapes, mass, rest, fric, force), ids, box_terrain(0.8 * nx * ntiles, 40.0, 0.8 * nz)


CONFIGS = {
    # name: (scene kwargs, terrain kwargs, iters)
    "C1": dict(kind="balls", num=8, extra=0, jitter=0.0, box=(10.0, 10.0, 10.0), iters=10),
    "demo": dict(kind="balls", num=11, extra=0, jitter=0.0, box=(10.0, 10.0, 10.0), iters=20, demo_extra_ball=True),
    "C2": dict(kind="balls", num=46, extra=2664, jitter=0.01, box=(80.0, 40.0, 80.0), iters=20),
    "C2pile": dict(kind="pile", num=46, extra=2664, jitter=0.01, box=(80.0, 40.0, 80.0), iters=20),
    "C4tile": dict(kind="balls", num=63, extra=0, jitter=0.01, box=(80.0, 60.0, 80.0), iters=20),
}


def build_config(name):
    cfg = CONFIGS[name]
    if cfg["kind"] == "pile":
        bodies = pile_scene(cfg["num"], cfg["extra"], cfg["jitter"], 1)
        hx, wh, hz = cfg["box"]
        return bodies, box_terrain(hx, wh, hz), cfg["iters"]
    bodies = balls_scene(cfg["num"], cfg["extra"], cfg["jitter"], 1, demo_extra_ball=cfg.get("demo_extra_ball", False))
    hx, wh, hz = cfg["box"]
    terrain = box_terrain(hx, wh, hz)
    return bodies, terrain, cfg["iters"]
