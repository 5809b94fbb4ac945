import sys, time, numpy as np
sys.path.insert(0,'.')
import torch, mgf_b200
from mgf_b200 import scenes
bodies, terrain, iters = scenes.build_config("C2pile")
g = mgf_b200.World(device=0); g.add_bodies(*bodies); g.set_terrain(*terrain)
dt=np.float32(1/60)
g.step(dt, iters, nsteps=3)
for k in range(5):
    st=g.step(dt, iters); print({k:v for k,v in st.items() if k in ('constraints','groups','step_ms','solve_ms','colouring_rounds','candidate_pairs')})
