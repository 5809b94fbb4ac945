"""Dev aid: a few steps of a small mixed scene + the batch APIs, to run under compute-sanitizer."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import mgf_b200
from mgf_b200 import scenes, _lib as L
s1 = scenes.pile_xyz(6, 5, 6, jitter=0.01, seed=3); s2 = scenes.capsule_pile(3, 2, 3, jitter=0.02, seed=4)
s2[0]["p"][:, 1] += 6.0
shapes = np.concatenate([s1[0], s2[0]])
n = len(shapes)
w = mgf_b200.World(device=0)
w.add_bodies(shapes, np.ones(n, np.float32), np.full(n, 0.3, np.float32), np.full(n, 0.6, np.float32), np.tile(np.array([0, -9.8, 0], np.float32), (n, 1)))
w.set_terrain(*scenes.box_terrain(6.0, 10.0, 6.0))
tot = 0; paths = {}
for _ in range(int(os.environ.get("MGFB_SAN_STEPS", "90"))):
    st = w.step(np.float32(1 / 60), 8)
    tot += st["constraints"]; paths[st["broadphase_path"]] = paths.get(st["broadphase_path"], 0) + 1
print("constraints", tot, "broadphase paths (1 coherent, 2 rebuilt, 3 sweep):", paths)
import ray_cases
rays, segs, shp = ray_cases.random_queries(500, 1)
out, hit = mgf_b200.intersections_batch(w.ctx, L.RAY, rays, shp); print("ray hits", int(hit.sum()))
b = mgf_b200.BVH(w.ctx)
rng = np.random.default_rng(0)
boxes = np.concatenate([rng.uniform(-5, 5, (300, 3)), rng.uniform(0.1, 1, (300, 3))], 1).astype(np.float32)
idx = b.insert(boxes, np.arange(300)); b.remove(idx[::3]); off, vals = b.query(boxes[:50]); print("bvh results", len(vals))
off, vals, hits = b.raytrace(L.RAY, rays[:50]); print("bvh ray results", len(vals))
