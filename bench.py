#!/usr/bin/env python
"""bench.py -- mgf step hot path on B200: contact-constraint iterations/s (+ narrowphase pairs/s).

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE
JSON line from rank 0.  A "step" is one World::step (mgf_demo/world.rs:227-294) of the
workload below with 20 solver iterations; `value` is whole-job constraint-iterations/s with
all inputs resident in HBM, timed on the device with CUDA events (the library's own events
on its launching stream), L2 flushed between timed steps; `e2e` is the same metric through
the public API with host buffers (pinned H2D of velocities + D2H of the full body state every
step, wall clock).  `roofline` is for the dominant kernel (k_solve), `cpu_baseline` is the
oracle (a C++ port of the reference: the Rust reference cannot be built here) on one core.

`--impl reference` times that CPU port alone on the same workload (bounded sample).

Workload "C2pile": BASELINE.json configs[1] body set (100 000 spheres r=0.5, balls.rs lattice
46^3 + 2664, LCG jitter +-0.01 seed 1, box 160x160x40, restitution 0.3, friction 0.6, g=-9.8,
dt=1/60, 20 iterations) arranged as a pile already in contact (lattice spacing squeezed to
0.98*2r, bottom layer on the floor) so that the ~295 k contact constraints per step of the
settled pile exist from step 0 in BOTH arms; the free-fall variant needs ~600 steps (10 min of
CPU time) before its first dense contact.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_CONSTRAINT_ITER = 268.0   # 216 B read + 52 B written
WORKLOAD = "C2pile"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled every ~5 ms by NVML during the timed region (the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; B200_PROFILING.md)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []          # (sm_mhz, reasons bitmask)
        self.smax = None
        self.h = None
        self._stop = False
        self.t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates by PCI order like CUDA_VISIBLE_DEVICES-less CUDA; map through the UUID to be safe
            import torch
            uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
            self.h = None
            for k in range(pynvml.nvmlDeviceGetCount()):
                hk = pynvml.nvmlDeviceGetHandleByIndex(k)
                u = pynvml.nvmlDeviceGetUUID(hk)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u or u.replace("GPU-", "") == uuid:
                    self.h = hk
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:   # noqa: BLE001
            self.err = repr(e)
            self.h = None

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                  int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.005)

    def mark(self):
        """Index of the next sample: brackets the timed region inside a longer-running sampler."""
        return len(self.rows)

    def stop(self, first=0, last=None):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: " + getattr(self, "err", "?")]}
        self._stop = True
        self.t.join(1.0)
        last = len(self.rows) if last is None else max(last, first + 1)
        rows = self.rows[first:last + 1]
        sm = sorted(r[0] for r in rows)
        mask = 0
        for r in rows:
            mask |= r[1]
        reasons = sorted(k for k, bit in self.REASONS.items() if mask & bit)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": reasons, "samples": len(sm)}


def build_scene():
    from mgf_b200 import scenes
    return scenes.build_config(WORKLOAD)


def run_reference(args, rank):
    """CPU arm: the oracle port of the reference, single thread (mgf is single-threaded and its
    Gauss-Seidel sweep is inherently serial, solver.rs:73-77)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib   # bench.py's reference arm is one of the two places allowed to execute oracle/
    if args.gpus == 1:
        bodies, terrain, iters = build_scene()
        workload = WORKLOAD
    else:   # the same one-box world the N-GPU arm tiles, whole, on one core
        from mgf_b200 import scenes
        tiles = [scenes.tiled_pile(args.gpus, t, nz=50 * args.bodies_per_gpu // 100000) for t in range(args.gpus)]
        bodies = tuple(np.concatenate([t[0][k] for t in tiles]) for k in range(5))
        terrain = tiles[0][2]; iters = 20
        workload = (f"pile of {args.gpus} x {args.bodies_per_gpu} spheres (50x40x{50 * args.bodies_per_gpu // 100000} lattice per tile, same "
                    "radius/spacing/material as C2pile), one box")
    w = oracle_lib.OracleWorld()
    w.add_bodies(*bodies); w.set_terrain(*terrain)
    dt = np.float32(1.0 / 60.0)
    budget_s = 150.0
    t_start = time.time()
    warm = 0
    for _ in range(args.warmup):
        if time.time() - t_start > budget_s * 0.3:
            break
        w.step(dt, iters); warm += 1
    sec = 0.0; ci = 0; pairs = 0; steps = 0
    for _ in range(args.steps):
        s, c, p = w.time_steps(dt, iters, 1)
        sec += s; ci += c; pairs += p; steps += 1
        if time.time() - t_start > budget_s:
            break
    val = ci / sec
    line = {
        "impl": "reference", "metric": "contact_constraint_iterations_per_second", "value": val, "unit": "constraint-iters/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "bodies": len(bodies[0]), "solver_iterations": iters, "dt": 1.0 / 60.0,
                   "constraints_per_step": ci / iters / steps, "candidate_pairs_per_step": pairs / steps},
        "narrowphase_pairs_per_second": pairs / sec,
        "cpu_baseline": {"value": val, "unit": "constraint-iters/s", "cores": 1, "kind": "port",
                         "sample": f"{steps} full World::step of {workload} after {warm} warm-up steps, C++ port of the reference "
                                   "(oracle/), g++ -O2 -ffp-contract=off, whole step timed like balls.rs:107-109; host has "
                                   f"{os.cpu_count()} cores, 1 used (reference is single-threaded)"},
        "e2e": {"value": val, "unit": "constraint-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mgf_b200", choices=["mgf_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bodies-per-gpu", type=int, default=100000,
                    help="N > 1 only: 100000 (default, 50x40x50 per tile) or 250000 (50x40x125 per tile: 8 GPUs = the 2 M-sphere C4 scene)")
    ap.add_argument("--schedule", default="dataflow", choices=["dataflow", "phases"],
                    help="solver schedule (include/mgfb.h mgfb_solver_schedule); tiled worlds always use phases")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "mgf_b200" else args.warmup
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU: mgf_b200 has no CPU fallback"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import mgf_b200

    dt = np.float32(1.0 / 60.0)
    if world == 1:
        bodies, terrain, iters = build_scene()
        g = mgf_b200.World(device=local_rank, solver_schedule=0 if args.schedule == "dataflow" else 1)
        g.add_bodies(*bodies); g.set_terrain(*terrain)
        workload = WORKLOAD
        parallelism = "single GPU"
    else:
        # ONE world of `world` x 100 000 spheres in one box, one slab per GPU (weak scaling).  Ghost
        # bodies and boundary velocities cross NVLink inside the kernels (csrc/tile.cuh); torch.distributed
        # only swaps the tiles' memory descriptors once, here.
        from mgf_b200 import scenes, tiling
        nz = 50 * args.bodies_per_gpu // 100000
        bodies, ids, terrain = scenes.tiled_pile(world, rank, nz=nz)
        iters = 20
        tw = tiling.TiledWorld(rank, world, device=local_rank, solver_schedule=0 if args.schedule == "dataflow" else 1)
        tw.add_owned(ids, *bodies); tw.set_terrain(*terrain)
        tw.connect(tiling.all_gather_bytes, ghost_capacity=32768 * max(1, nz // 50))
        g = tw.world
        workload = f"pile of {world} x {len(bodies[0])} spheres (50x40x{nz} lattice per tile, same radius/spacing/material as C2pile), one box"
        parallelism = (f"{world} slabs along x, one per GPU; ghost bodies once per step through NVLink peer memory; "
                       + ("solver: the constraint chains of boundary bodies continue on the neighbour GPU, every hand-over one 32-byte "
                          "peer store from inside the solver kernel (no exchange phase, no grid barrier)" if args.schedule == "dataflow"
                          else "boundary velocities every solver iteration through peer memory inside the barrier-phased solver kernel"))
    n = len(bodies[0])
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # 256 MB > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (the clock sampler starts here: nvidia-smi needs ~0.1 s before its first sample)
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    g.step(dt, iters, nsteps=args.warmup)
    g.totals(reset=True)
    # ---- timed: K steps, device time from the library's CUDA events, L2 flushed between steps
    barrier()
    mark0 = sampler.mark()
    step_ms = []; solve_ms = []; cons = []; pairs = []; groups = []; ghosts = []; bcons = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1.0); torch.cuda.synchronize()
        st = g.step(dt, iters)
        step_ms.append(st["step_ms"]); solve_ms.append(st["solve_ms"]); cons.append(st["constraints"])
        pairs.append(st["candidate_pairs"] + st["terrain_candidates"]); groups.append(st["phases"])
        ghosts.append(st["ghosts"]); bcons.append(st["boundary_constraints"])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop(mark0, sampler.mark())
    tot = g.totals(reset=True)
    total_ms = float(sum(step_ms)); units = float(sum(cons)) * iters; npairs = float(sum(pairs))
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); total_ms_max = t.item()
        u = torch.tensor([units, npairs], device="cuda", dtype=torch.float64); dist.all_reduce(u, op=dist.ReduceOp.SUM)
        units_all, pairs_all = u[0].item(), u[1].item()
    else:
        total_ms_max, units_all, pairs_all = total_ms, units, npairs
    value = units_all / (total_ms_max * 1e-3)

    # ---- end to end: public API with host buffers; every step: pinned H2D of that step's inputs (v, omega) and D2H of its
    # result (x, q, v, omega), through the pipelined step API (mgfb_step_enqueue / mgfb_step_wait: the transfers of step
    # k overlap the kernels of step k+1, two output buffer sets); on a tiled world every rank runs the same sequence.
    pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    hv, hw = pin((n, 3)), pin((n, 3))
    outs = [(pin((n, 3)), pin((n, 4)), pin((n, 3)), pin((n, 3))) for _ in range(2)]
    # the per-step inputs are external velocity increments (zero here, so the e2e loop steps the SAME world the
    # device-timed loop stepped)
    hv[:] = 0.0; hw[:] = 0.0
    from mgf_b200 import _lib as L
    lib, h = g.ctx.lib, g.ctx.h
    import ctypes as C
    st = L.StepStats()
    barrier()
    e2e_units = 0.0; e2e_dev_ms = 0.0
    t0 = time.perf_counter()
    for k in range(args.steps):
        hx, hq, ov, ow = outs[k & 1]
        g.ctx.check(lib.mgfb_step_enqueue(h, dt, iters, L.INPUT_ADD, L.ptr(hv), L.ptr(hw), L.ptr(hx), L.ptr(hq), L.ptr(ov), L.ptr(ow)))   # H2D + step + D2H queued
        if k > 0:
            g.ctx.check(lib.mgfb_step_wait(h, C.byref(st)))   # step k-1's state is in outs[(k-1) & 1]
            e2e_units += st.constraints * iters; e2e_dev_ms += st.step_ms
    g.ctx.check(lib.mgfb_step_wait(h, C.byref(st)))
    e2e_units += st.constraints * iters; e2e_dev_ms += st.step_ms
    e2e_api = "mgfb_step_enqueue / mgfb_step_wait (pipelined, 2 steps in flight)"
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = t.item()
        u = torch.tensor([e2e_units], device="cuda", dtype=torch.float64); dist.all_reduce(u, op=dist.ReduceOp.SUM); e2e_units = u.item()
    e2e_val = e2e_units / e2e_s
    assert np.isfinite(outs[0][0]).all() and np.isfinite(outs[(args.steps - 1) & 1][0]).all()

    # ---- roofline of the dominant kernel (k_solve), live CUDA-event durations
    peak, peak_src = measured_peak()
    solve_s = sum(solve_ms) * 1e-3
    achieved = ALGO_BYTES_PER_CONSTRAINT_ITER * units / solve_s / 1e9
    dataflow = args.schedule == "dataflow"
    roofline = {"kernel": "k_solve_df (persistent sequential-impulse solver; rows wait on per-row inboxes filled by their predecessors, no grid barrier)" if dataflow
                else "k_solve (persistent cooperative sequential-impulse solver, one grid barrier per colour)", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one k_solve_df launch of this workload, `ncu --set full`
                # (profiles/r01_v4_k_solve_df_raw.csv; ncu starts every replay pass from a flushed L2)
                "traffic": 145.2e6 if (dataflow and world == 1) else None, "traffic_unit": "bytes per launch (ncu, cold L2)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CONSTRAINT_ITER * units / len(solve_ms),
                "avg_launch_ms": sum(solve_ms) / len(solve_ms), "share_of_step": sum(solve_ms) / total_ms,
                "note": "268 B per constraint-iteration x constraints x 20 iterations per launch; the rows (~70 MB) are "
                        "L2-resident, so the kernel is bound by the latency of its 180 chain hand-overs, not by HBM bytes "
                        "(profiles/r01_summary_v4.md)"}

    # ---- CPU baseline (rank 0, N=1): bounded sample of the same workload on the oracle
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        o = oracle_lib.OracleWorld()
        o.add_bodies(*bodies); o.set_terrain(*terrain)
        o.step(dt, iters, min(args.warmup, 3))
        sec, ci, pr = o.time_steps(dt, iters, 8)
        cpu = {"value": ci / sec, "unit": "constraint-iters/s", "cores": 1, "kind": "port",
               "pairs_per_second": pr / sec, "ms_per_step": 1e3 * sec / 8,
               "sample": f"8 full World::step of {WORKLOAD} after {min(args.warmup, 3)} warm-up steps on the C++ port of the reference "
                         f"(the Rust reference cannot be built: no rustc); 1 of {os.cpu_count()} host cores (reference is single-threaded)"}

    if rank == 0:
        line = {
            "metric": "contact_constraint_iterations_per_second", "value": value, "unit": "constraint-iters/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "bodies_per_gpu": n, "solver_iterations": iters, "dt": 1.0 / 60.0,
                       "constraints_per_step": units_all / iters / args.steps, "candidate_pairs_per_step": pairs_all / args.steps,
                       "colour_groups": sum(groups) / len(groups), "l2": "flushed between timed steps (256 MB write)",
                       "parallelism": parallelism,
                       "rank0_ghosts_per_step": sum(ghosts) / len(ghosts), "rank0_boundary_constraints_per_step": sum(bcons) / len(bcons),
                       "arithmetic": "--fmad=false, IEEE div/sqrt: bit-exact vs the CPU port"},
            "narrowphase_pairs_per_second": pairs_all / (total_ms_max * 1e-3),
            "solver_only_constraint_iters_per_second": units / solve_s,
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "constraint-iters/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 52,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "api": e2e_api,
                    "device_ms_per_step": e2e_dev_ms / args.steps, "constraints_per_step": e2e_units / iters / args.steps / max(world, 1)},
            "gpu_launches": tot["kernel_launches"], "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
