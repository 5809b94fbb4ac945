"""Dev aid: solver time vs problem size for both schedules (not part of the bench contract)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, mgf_b200
from mgf_b200 import scenes
sizes = ((8, 0), (20, 0), (46, 2664), (80, 0)) if len(sys.argv) < 2 else [(int(a), 0) for a in sys.argv[1:]]
for num, extra in sizes:
    bodies = scenes.pile_scene(num, extra, 0.01, 1)
    for sched in (0, 1):
        g = mgf_b200.World(device=0, solver_schedule=sched); g.add_bodies(*bodies); g.set_terrain(*scenes.box_terrain(160, 40, 160))
        dt = np.float32(1 / 60)
        g.step(dt, 20, nsteps=3)
        sm = []; 
        for _ in range(5):
            st = g.step(dt, 20); sm.append(st['solve_ms'])
        ph = st['phases'] * 20
        print(num, 'df' if sched == 0 else 'ph', {k: v for k, v in st.items() if k in ('constraints', 'phases', 'step_ms', 'colouring_rounds')},
              'solve_ms %.4f' % min(sm), 'us/phase %.2f' % (1e3 * min(sm) / ph), 'Gci/s %.2f' % (st['constraints'] * 20 / min(sm) / 1e6), flush=True)
        g.ctx.close()
