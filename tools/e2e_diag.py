"""Development tool (GPU, torchrun): where does the end-to-end (pipelined, host-buffer) step time go on N tiles?
Per rank: concurrent PCIe copy bandwidth at the step's transfer sizes, then the bench.py e2e loop with host timestamps
around mgfb_step_enqueue and mgfb_step_wait, for a few pipeline depths and with the transfers switched off.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_diag.py [steps]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import bench
from mgf_b200 import _lib as L


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dt = np.float32(bench.DT)
    if world > 1:
        tw, bodies, terrain, iters = bench.build_settled_tiled(world, rank, local, "dataflow")
        g = tw.world
    else:
        g, bodies, terrain, iters, snap = bench.gpu_preroll("C2settled", local)
    n = len(bodies[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def report(tag, vals):
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        if world > 1:
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
        else:
            out = [t]
        if rank == 0:
            print(tag, " | ".join(" ".join(f"{x:.3f}" for x in o.tolist()) for o in out), flush=True)

    # ---- (a) all ranks copy at once, the step's sizes, separate streams for the two directions
    pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True)
    hin, hout = pin((n, 6)), pin((n, 13))
    din, dout = torch.empty((n, 6), device="cuda"), torch.empty((n, 13), device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for both in (False, True):
        barrier()
        t0 = time.perf_counter()
        for _ in range(50):
            with torch.cuda.stream(s1):
                hout.copy_(dout, non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    din.copy_(hin, non_blocking=True)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 50 * 1e3
        report(f"copy per step (ms) d2h 5.2MB{' + h2d 2.4MB' if both else ''}, all ranks at once:", [ms])

    # ---- (b) the e2e loop with timestamps
    lib, h = g.ctx.lib, g.ctx.h
    hv, hw = pin((n, 3)).numpy(), pin((n, 3)).numpy()
    hv[:] = 0; hw[:] = 0
    st = L.StepStats()
    g.step(dt, iters, nsteps=3)
    configs = ((3, True, True), (3, True, True), (2, True, True), (4, True, True), (3, False, True), (3, True, False), (3, False, False), (1, True, True))
    if os.environ.get("MGFB_DIAG_SHORT") == "1":
        configs = ((3, True, True), (3, True, True), (4, True, True), (3, False, False))
    if os.environ.get("MGFB_DIAG_SHORT") == "2":
        configs = ((3, True, True), (3, True, True))
    for depth, with_in, with_out in configs:
        outs = [(pin((n, 3)).numpy(), pin((n, 4)).numpy(), pin((n, 3)).numpy(), pin((n, 3)).numpy()) for _ in range(depth)]
        barrier()
        enq, wai, dev, sol = [], [], [], []
        t0 = time.perf_counter()
        waited = 0
        for k in range(steps):
            hx, hq, ov, ow = outs[k % depth]
            ta = time.perf_counter()
            g.ctx.check(lib.mgfb_step_enqueue(h, dt, iters, L.INPUT_ADD, L.ptr(hv) if with_in else None, L.ptr(hw) if with_in else None,
                                              L.ptr(hx) if with_out else None, L.ptr(hq) if with_out else None, L.ptr(ov) if with_out else None,
                                              L.ptr(ow) if with_out else None))
            tb = time.perf_counter()
            enq.append(tb - ta)
            if k - waited + 1 >= depth:
                g.ctx.check(lib.mgfb_step_wait(h, C.byref(st))); waited += 1
                wai.append(time.perf_counter() - tb); dev.append(st.step_ms); sol.append(st.solve_ms)
        while waited < steps:
            tb = time.perf_counter()
            g.ctx.check(lib.mgfb_step_wait(h, C.byref(st))); waited += 1
            wai.append(time.perf_counter() - tb); dev.append(st.step_ms); sol.append(st.solve_ms)
        barrier()
        wall = (time.perf_counter() - t0) / steps * 1e3
        e = np.array(enq) * 1e3; w = np.array(wai) * 1e3; d = np.array(dev); so = np.array(sol)
        report(f"depth {depth} in {int(with_in)} out {int(with_out)}: wall/step, enqueue mean p99, wait mean p99, device mean p99, solve mean (ms):",
               [wall, e.mean(), np.percentile(e, 99), w.mean(), np.percentile(w, 99), d.mean(), np.percentile(d, 99), so.mean()])
    # ---- per-phase device time of the synchronous step (mgfb_step_profile), every rank
    acc = {}
    for _ in range(20):
        _, pr = g.step_profile(dt, iters)
        for k, v in pr.items():
            if isinstance(v, float):
                acc[k] = acc.get(k, 0.0) + v / 20
    if rank == 0:
        print("phases:", " ".join(acc.keys()), flush=True)
    report("phase ms per rank:", [acc[k] for k in acc])
    # ---- (c) the synchronous call, device-timed, for reference
    barrier()
    t0 = time.perf_counter()
    ms = []
    sm = []
    for _ in range(steps):
        r = g.step(dt, iters); ms.append(r["step_ms"]); sm.append(r["solve_ms"])
    barrier()
    report("synchronous mgfb_step: wall/step, device mean, solve mean (ms):", [(time.perf_counter() - t0) / steps * 1e3, float(np.mean(ms)), float(np.mean(sm))])


if __name__ == "__main__":
    main()
