"""Development tool (GPU): how many bodies replace their stored fat box per step in each benchmark window (the size of the
incremental part of a temporally coherent broadphase)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
DT = np.float32(bench.DT)
for name in sys.argv[1:] or ["C2settled", "C2pile", "C3", "C5"]:
    g, bodies, terrain, iters, snap = bench.gpu_preroll(name)
    rows = [g.step(DT, iters) for _ in range(40)]
    r = np.array([x["fat_refreshes"] for x in rows]); p = np.array([x["candidate_pairs"] for x in rows]); c = np.array([x["constraints"] for x in rows])
    print(f"{name}: bodies {len(bodies[0])}  fat refreshes per step: mean {r.mean():.1f} max {r.max()} first10 {r[:10].tolist()}  candidate pairs {p.mean():.0f}  constraints {c.mean():.0f}", flush=True)
    g.ctx.close()
