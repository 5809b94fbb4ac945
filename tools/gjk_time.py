"""Development tool (GPU): wall clock of mgfb_gjk_batch on the benchmark's 1 M-pair batch (bench.py gjk_block)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gjk_cases, mgf_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
a, b = gjk_cases.mixed_pairs(n, seed=7)
ctx = mgf_b200.Context(device=0)
mgf_b200.gjk_batch(ctx, a[:4096], b[:4096])
for _ in range(2):
    t0 = time.perf_counter(); out, status, iters = mgf_b200.gjk_batch(ctx, a, b); dt = time.perf_counter() - t0
    print(f"{n} pairs: {dt * 1e3:.1f} ms = {n / dt / 1e6:.3f} M pairs/s; status {dict(zip(*[x.tolist() for x in np.unique(status, return_counts=True)]))}", flush=True)
