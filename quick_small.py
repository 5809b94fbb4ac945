import sys, numpy as np
sys.path.insert(0,'.')
import torch, mgf_b200
from mgf_b200 import scenes
for num, extra in ((8,0),(20,0),(46,2664)):
    bodies = scenes.pile_scene(num, extra, 0.01, 1)
    g = mgf_b200.World(device=0); g.add_bodies(*bodies); g.set_terrain(*scenes.box_terrain(80,40,80))
    dt=np.float32(1/60)
    g.step(dt, 20, nsteps=3)
    st=g.step(dt, 20)
    ph=st['groups']*20
    print(num, {k:v for k,v in st.items() if k in ('constraints','groups','step_ms','solve_ms','colouring_rounds')}, 'us/phase %.2f'%(1e3*st['solve_ms']/ph))
