"""Development tool (GPU): A/B of the solver kernels on the benchmark workloads.  For each workload the scene is pre-rolled
once; every variant then restores the same snapshot, warms up and times the same steps (library CUDA events, L2 flushed).
Usage: python tools/solver_ab.py [workload ...]   (variants: env MGFB_AB="kernel:warps,..." default "1:0,3:0")"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import mgf_b200

DT = np.float32(1 / 60)


def main():
    names = sys.argv[1:] or ["C2pile", "C2settled"]
    variants = os.environ.get("MGFB_AB", "").split(",")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    nsteps = int(os.environ.get("MGFB_AB_STEPS", "12"))
    for name in names:
        g, bodies, terrain, iters, snap = bench.gpu_preroll(name)
        g.ctx.close()
        ref_state = None
        for variant in variants:
            from mgf_b200 import _lib
            _lib._lib = None   # load the named build
            os.environ["MGFB_LIB"] = os.path.join(ROOT, "mgf_b200", "lib", f"libmgfb{variant}.so")
            w = mgf_b200.World(device=0)
            w.add_bodies(*bodies); w.set_terrain(*terrain); w.restore(snap)
            w.step(DT, iters, nsteps=3)
            rows = bench.time_steps_device(w, torch, flush, DT, iters, nsteps)
            st = w.state()
            if ref_state is None:
                ref_state = st
            same = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(st, ref_state))
            ms = np.mean([r["step_ms"] for r in rows]); sm = np.mean([r["solve_ms"] for r in rows])
            print(f"{name:10s} build '{variant}': step {ms:.4f} ms  solve {sm:.4f} ms  constraints {np.mean([r['constraints'] for r in rows]):.0f}"
                  f"  colours {np.mean([r['phases'] for r in rows]):.2f}  same bits as first variant: {same}", flush=True)
            if os.environ.get("MGFB_AB_PHASES", "1") == "1" and hasattr(w, "step_profile"):
                acc = {}
                for _ in range(6):
                    flush.add_(1.0)
                    _, pr = w.step_profile(DT, iters)
                    for k, v in pr.items():
                        if isinstance(v, float):
                            acc[k] = acc.get(k, 0.0) + v / 6
                print("           phases (us): " + "  ".join(f"{k} {1e3 * v:.1f}" for k, v in acc.items() if v > 0), flush=True)
            w.ctx.close()


if __name__ == "__main__":
    main()
