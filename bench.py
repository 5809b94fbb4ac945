#!/usr/bin/env python
"""bench.py -- mgf step hot path on B200: contact-constraint iterations/s (+ narrowphase pairs/s).

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE
JSON line from rank 0.  A "step" is one World::step (mgf_demo/world.rs:227-294) of the
workload below with 20 solver iterations; `value` is whole-job constraint-iterations/s with
all inputs resident in HBM, timed on the device with CUDA events (the library's own events
on its launching stream), L2 flushed between timed steps; `e2e` is the same metric through
the public API with host buffers (pinned H2D of velocities + D2H of the full body state every
step, wall clock).  `roofline` is for the dominant kernel (the solver), `kernels` gives the
other phases of the step against their own algorithmic bytes, `cpu_baseline` is the oracle
(a C++ port of the reference: the Rust reference cannot be built here) on one core.

`--impl reference` times that CPU port alone on the same workload (bounded sample).

Workloads (`--workload`, default C2settled at N=1):
  C2settled  BASELINE.json configs[1] as BASELINE.md section 2 writes it: 100 000 spheres r=0.5 dropped from
             the balls.rs lattice (46^3 + 2664, spacing 1.25, LCG jitter +-0.01 seed 1) into the 160x160x40
             box, restitution 0.3, friction 0.6, g=-9.8, dt=1/60, 20 iterations; the timed window starts at
             step 600 (the disordered pile).  Steps 0..599 are run ONCE on the GPU; the state at step 600
             (x, q, v, omega, colliders, stored fat boxes) is a snapshot that BOTH arms load, so the GPU and
             the CPU port time the same steps 600.. on identical inputs (tests/test_gpu_configs.py lock-steps
             them bit for bit from that snapshot).
  C2pile     the same body set squeezed into a lattice pile already in contact (round 1's workload: the
             solver's best case, kept for comparison).
  C1         BASELINE configs[0]: 512 spheres in the demo box, 10 iterations, window from step 400.
  C3         BASELINE configs[2]: 50 000 capsules on a 20 000-triangle mesh floor (Capsule x Triangle).
  C5         BASELINE configs[4] restated: 200 000 mixed sphere/capsule bodies on a 200 000-triangle mesh.
At N=1 with no --workload the line also carries short runs of the other workloads under `other_workloads`.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_CONSTRAINT_ITER = 268.0   # 216 B read + 52 B written
# name: (scene, pre-roll steps run once on the GPU before the window, description)
WORKLOADS = {
    "C2settled": ("C2", 600, "C2settled: 100000 spheres dropped into the 160x160x40 box (balls.rs lattice 46^3+2664, jitter 0.01), window from step 600"),
    "C2pile": ("C2pile", 0, "C2pile: the C2 body set as a squeezed lattice pile already in contact, window from step 0"),
    "C1": ("C1", 400, "C1: 512 spheres in the demo box (balls.rs n=8), 10 iterations, window from step 400"),
    "C3": ("C3", 60, "C3: 50000 capsules on a 20000-triangle height-field floor, window from step 60"),
    "C5": ("C5", 60, "C5: 200000 mixed sphere/capsule bodies on a 199712-triangle height field, window from step 60"),
}
DEFAULT_WORKLOAD = "C2settled"
DT = 1.0 / 60.0


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


# ------------------------------------------------------------------------------------------ workloads
def build_scene(name):
    """(bodies, terrain, iters) of a workload's scene at step 0 (host-side generators, mgf_b200/scenes.py)."""
    from mgf_b200 import scenes
    scene = WORKLOADS[name][0]
    if scene == "C3":
        return scenes.config_c3()
    if scene == "C5":
        return scenes.config_c5()
    return scenes.build_config(scene)


def config_of(name, nbodies, iters):
    """The `config` object: the same keys and values in both arms."""
    return {"workload": WORKLOADS[name][2], "bodies": int(nbodies), "solver_iterations": int(iters), "dt": DT,
            "window_first_step": WORKLOADS[name][1]}


def snapshot_path(name):
    return os.path.join(tempfile.gettempdir(), f"mgfb_bench_snapshot_{name}_{WORKLOADS[name][1]}.npz")


def save_snapshot(name, snap):
    import numpy as np
    tmp = snapshot_path(name) + f".{os.getpid()}.tmp.npz"
    np.savez(tmp, **snap)
    os.replace(tmp, snapshot_path(name))


def load_snapshot(name):
    import numpy as np
    p = snapshot_path(name)
    if not os.path.exists(p):
        return None
    with np.load(p) as z:
        return {k: z[k] for k in z.files}


def gpu_preroll(name, device=0):
    """Steps 0..window_first_step-1 on the GPU (once); returns the world, positioned at the window's first step,
    its scene and the snapshot of that state."""
    import numpy as np
    import mgf_b200
    bodies, terrain, iters = build_scene(name)
    g = mgf_b200.World(device=device)
    g.add_bodies(*bodies); g.set_terrain(*terrain)
    pre = WORKLOADS[name][1]
    if pre:
        # profiling runs (ncu launch lists) skip the pre-roll's ~13 000 launches by loading the snapshot an earlier run on this box
        # cached: restoring it continues bit-identically (tests/test_gpu_step.py snapshot tests)
        snap = load_snapshot(name) if os.environ.get("MGFB_BENCH_REUSE_SNAPSHOT") == "1" else None
        if snap is not None:
            g.restore(snap)
        else:
            g.step(np.float32(DT), iters, nsteps=pre)
    return g, bodies, terrain, iters, g.snapshot()


def reference_snapshot(name):
    """The window's initial state for the CPU arm: the cached snapshot, else made by a GPU pre-roll in a CHILD process
    (this process never loads the CUDA library), else -- no GPU -- by pre-rolling on the CPU port itself."""
    if WORKLOADS[name][1] == 0:
        return None, "step 0 of the scene (no pre-roll)"
    snap = load_snapshot(name)
    if snap is not None:
        return snap, "snapshot of the GPU pre-roll (cached by an earlier bench.py run on this box)"
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--make-snapshot", name], capture_output=True, text=True)
    snap = load_snapshot(name)
    if r.returncode == 0 and snap is not None:
        return snap, "snapshot of a GPU pre-roll run in a child process (inputs only; the timed region is CPU-only)"
    return "cpu", "pre-rolled on the CPU port itself (no GPU available for the pre-roll)"


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank):
    """CPU arm: the oracle port of the reference, single thread (mgf is single-threaded and its
    Gauss-Seidel sweep is inherently serial, solver.rs:73-77)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib   # bench.py's reference arm is one of the two places allowed to execute oracle/
    dt = np.float32(DT)
    if args.gpus == 1:
        name = args.workload or DEFAULT_WORKLOAD
        bodies, terrain, iters = build_scene(name)
        config = config_of(name, len(bodies[0]), iters)
        snap, how = reference_snapshot(name)
    elif args.bodies_per_gpu == 100000:   # the same one-box world the N-GPU arm tiles, whole, on one core
        name = "C2settled"
        b1, _, iters = build_scene(name)
        one, how = reference_snapshot(name)
        if not isinstance(one, dict):
            w1 = oracle_lib.OracleWorld(); w1.add_bodies(*b1); w1.set_terrain(*build_scene(name)[1]); w1.step(dt, iters, WORKLOADS[name][1])
            one = w1.snapshot()
        bodies = tuple(np.concatenate([b1[k]] * args.gpus) for k in range(5))
        parts = [settled_tile_snapshot(one, t, args.gpus) for t in range(args.gpus)]
        snap = {k: np.concatenate([p[k] for p in parts]) for k in one}
        terrain = settled_tiles_terrain(args.gpus)
        config = settled_tiles_config(args.gpus, len(b1[0]), iters)
    else:   # --bodies-per-gpu 250000: the lattice pile of BASELINE configs[3]
        from mgf_b200 import scenes
        tiles = [scenes.tiled_pile(args.gpus, t, nz=50 * args.bodies_per_gpu // 100000) for t in range(args.gpus)]
        bodies = tuple(np.concatenate([t[0][k] for t in tiles]) for k in range(5))
        terrain = tiles[0][2]; iters = 20
        config = tiled_config(args.gpus, args.bodies_per_gpu, iters)
        snap, how = None, "step 0 of the scene (no pre-roll)"
    w = oracle_lib.OracleWorld()
    w.add_bodies(*bodies); w.set_terrain(*terrain)
    if isinstance(snap, dict):
        w.restore(snap)
    elif snap == "cpu":
        w.step(dt, iters, WORKLOADS[name][1])
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
        "config": config,
        "work": {"constraints_per_step": ci / iters / steps, "candidate_pairs_per_step": pairs / steps, "initial_state": how},
        "narrowphase_pairs_per_second": pairs / sec,
        "cpu_baseline": {"value": val, "unit": "constraint-iters/s", "cores": 1, "kind": "port",
                         "sample": f"{steps} full World::step of the workload after {warm} warm-up steps, C++ port of the reference "
                                   "(oracle/), g++ -O2 -ffp-contract=off, whole step timed like balls.rs:107-109; host has "
                                   f"{os.cpu_count()} cores, 1 used (reference is single-threaded)"},
        "e2e": {"value": val, "unit": "constraint-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


SEAM_PITCH = 160.5   # the C2 box is 160 wide: copies half a unit further apart start just out of contact across the seam


def settled_tile_snapshot(snap, tile, ntiles):
    """Tile `tile` of the N-GPU workload: the C2settled snapshot translated along x to its place in a row of `ntiles` copies
    (numpy f32 adds, the same in both arms).  The inner walls are gone, so neighbouring piles lean on each other over the seams."""
    import numpy as np
    off = np.float32((tile - (ntiles - 1) / 2.0) * SEAM_PITCH)
    out = {k: np.array(v, copy=True) for k, v in snap.items()}
    out["x"][:, 0] += off
    out["colliders"]["p"][:, 0] += off
    out["fat"][:, 0] += off
    return out


def settled_tiles_terrain(ntiles):
    from mgf_b200 import scenes
    return scenes.box_terrain(SEAM_PITCH * ntiles / 2.0, 40.0, 80.0)


def settled_tiles_config(world, n, iters):
    return {"workload": f"{world} x C2settled: {world} copies of the settled 100000-sphere pile (state of step 600) side by side in one box "
                        f"{SEAM_PITCH * world:g} x 160, one per GPU, no inner walls: neighbouring piles meet over the tile seams",
            "bodies": world * n, "solver_iterations": iters, "dt": DT, "window_first_step": 600}


def build_settled_tiled(world, rank, local_rank, schedule):
    """Every rank pre-rolls the C2 scene on its own GPU (identical bits everywhere), moves it to its slot and loads it into its tile."""
    import numpy as np
    from mgf_b200 import tiling
    g, bodies, _, iters, snap = gpu_preroll("C2settled", local_rank)
    g.ctx.close()
    if rank == 0:
        save_snapshot("C2settled", snap)
    n = len(bodies[0])
    tw = tiling.TiledWorld(rank, world, device=local_rank, solver_schedule=0 if schedule == "dataflow" else 1)
    tw.add_owned(np.arange(n, dtype=np.uint32) + np.uint32(rank * n), *bodies)
    terrain = settled_tiles_terrain(world)
    tw.set_terrain(*terrain)
    tw.world.restore(settled_tile_snapshot(snap, rank, world))
    tw.connect(tiling.all_gather_bytes, ghost_capacity=32768)
    return tw, bodies, terrain, iters


def build_tiled(world, rank, local_rank, bodies_per_gpu, schedule):
    from mgf_b200 import scenes, tiling
    nz = 50 * bodies_per_gpu // 100000
    bodies, ids, terrain = scenes.tiled_pile(world, rank, nz=nz)
    tw = tiling.TiledWorld(rank, world, device=local_rank, solver_schedule=0 if schedule == "dataflow" else 1)
    tw.add_owned(ids, *bodies); tw.set_terrain(*terrain)
    tw.connect(tiling.all_gather_bytes, ghost_capacity=32768 * max(1, nz // 50))
    return tw, bodies, terrain


def tiled_config(world, bodies_per_gpu, iters):
    nz = 50 * bodies_per_gpu // 100000
    return {"workload": f"pile of {world} x {bodies_per_gpu} spheres (50x40x{nz} lattice per tile, same radius/spacing/material as C2pile), one box",
            "bodies": world * bodies_per_gpu, "solver_iterations": iters, "dt": DT, "window_first_step": 0}


# ------------------------------------------------------------------------------------------ GPU arm helpers
def time_steps_device(g, torch, flush, dt, iters, nsteps):
    """nsteps World::steps, each timed by the library's CUDA events on its stream, L2 flushed in between."""
    rows = []
    for _ in range(nsteps):
        flush.fill_(1.0); torch.cuda.synchronize()
        rows.append(g.step(dt, iters))
    return rows


def phase_rooflines(g, torch, flush, dt, iters, nsteps, peak):
    """Per-phase device time (mgfb_step_profile: an event between the phases) against each phase's ALGORITHMIC bytes
    (SURVEY.md section 8d): integrate 272 B/body; narrowphase 64 B (SxS), 72 (SxCap / CapxS), 88 (CapxCap), 72 (TrixS),
    84 (TrixCap) per candidate pair + 64 B per LocalContact; ContactConstraint::new 2x84 + 64 read + 100 written per
    constraint; solver 268 B per constraint-iteration."""
    acc = {}; n = 0
    for _ in range(nsteps):
        flush.fill_(1.0); torch.cuda.synchronize()
        st, pr = g.step_profile(dt, iters)
        n += 1
        pb = pr["pairs"]; tp = pr["terrain_pairs"]
        bytes_ = {
            "integrate": 272.0 * st["bodies"],
            "body_grid": None, "pair_sweep": None, "colouring": None,
            # (+ 76 B per contact: the chain colouring's per-constraint setup -- key 8, links 24, group 4, two degree counters 8, the two
            # positions the key's class is read from 32 -- is done by the thread that emits the contact since round 2)
            "narrow_bodies": 64.0 * pb[0] + 72.0 * (pb[1] + pb[2]) + 88.0 * pb[3] + (64.0 + 76.0) * pr["body_contacts"],
            "terrain": 72.0 * tp[0] + 84.0 * tp[1] + (64.0 + 44.0) * pr["terrain_contacts"],
            "build_rows": (2 * 84.0 + 64.0 + 100.0) * st["constraints"],
            "solve": ALGO_BYTES_PER_CONSTRAINT_ITER * st["constraints"] * iters,
        }
        units = {"integrate": st["bodies"], "body_grid": st["bodies"], "pair_sweep": sum(pb), "narrow_bodies": sum(pb), "terrain": sum(tp),
                 "colouring": st["constraints"], "build_rows": st["constraints"], "solve": st["constraints"] * iters}
        for k in bytes_:
            a = acc.setdefault(k, {"ms": 0.0, "bytes": 0.0, "units": 0.0, "has_bytes": bytes_[k] is not None})
            a["ms"] += pr[k]; a["bytes"] += bytes_[k] or 0.0; a["units"] += units[k]
    out = {}
    unit_names = {"integrate": "bodies", "body_grid": "bodies", "pair_sweep": "candidate pairs", "narrow_bodies": "candidate pairs",
                  "terrain": "(body, face) candidates", "colouring": "constraints", "build_rows": "constraints", "solve": "constraint-iterations"}
    for k, a in acc.items():
        ms = a["ms"] / n
        row = {"ms": ms, "units_per_step": a["units"] / n, "unit": unit_names[k],
               "units_per_second": (a["units"] / n) / (ms * 1e-3) if ms > 0 else None}
        if a["has_bytes"] and ms > 0:
            gbs = a["bytes"] / n / (ms * 1e-3) / 1e9
            row.update({"algorithmic_bytes_per_step": a["bytes"] / n, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak})
        out[k] = row
    return out


def short_run(name, torch, flush, device, peak):
    """A secondary workload: pre-roll, 3 warm-up + 10 timed steps, device-timed, no CPU arm."""
    import numpy as np
    dt = np.float32(DT)
    g, bodies, terrain, iters, snap = gpu_preroll(name, device)
    g.step(dt, iters, nsteps=3)
    rows = time_steps_device(g, torch, flush, dt, iters, 10)
    ms = sum(r["step_ms"] for r in rows); sms = sum(r["solve_ms"] for r in rows)
    cons = sum(r["constraints"] for r in rows); prs = sum(r["candidate_pairs"] + r["terrain_candidates"] for r in rows)
    phases = phase_rooflines(g, torch, flush, dt, iters, 4, peak)
    out = {"config": config_of(name, len(bodies[0]), iters), "value": cons * iters / (ms * 1e-3), "unit": "constraint-iters/s",
           "ms_per_step": ms / len(rows), "solve_ms": sms / len(rows), "constraints_per_step": cons / len(rows),
           "candidate_pairs_per_step": prs / len(rows), "narrowphase_pairs_per_second": prs / (ms * 1e-3),
           "colours": sum(r["phases"] for r in rows) / len(rows),
           "solver_frac_of_hbm_peak": ALGO_BYTES_PER_CONSTRAINT_ITER * cons * iters / (sms * 1e-3) / 1e9 / peak if sms > 0 else None,
           "kernels": {k: {kk: v[kk] for kk in ("ms", "units_per_second", "unit", "frac_of_hbm_peak") if kk in v} for k, v in phases.items()}}
    g.ctx.close()
    return out


def gjk_block(device, with_cpu):
    """BASELINE configs[4]'s discrete half: 1 M static convex pairs (all 16 kind pairs of Sphere / Capsule / AABB / OBB, centres
    U(-2,2)^3, sizes U(0.3,1), LCG seed 7) through mgfb_gjk_batch (GJK + EPA, collision.rs:497-519).  Host buffers in and out:
    the call's wall clock includes both PCIe copies.  The CPU port runs a 20 k sample of the same pairs; that sample is also
    compared bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import gjk_cases   # input generator only (no oracle code)
    import mgf_b200
    n = 1 << 20
    a, b = gjk_cases.mixed_pairs(n, seed=7)
    ctx = mgf_b200.Context(device=device)
    mgf_b200.gjk_batch(ctx, a[:4096], b[:4096])
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        out, status, iters = mgf_b200.gjk_batch(ctx, a, b)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    blk = {"pairs": n, "pairs_per_second": n / best, "ms_per_batch": 1e3 * best, "contacts": int((status == 1).sum()),
           "epa_iterations_mean": float(iters[status == 1].mean()), "timing": "wall clock of mgfb_gjk_batch, host buffers (H2D + kernels + D2H), best of 3",
           "status_histogram": {int(k): int(v) for k, v in zip(*np.unique(status, return_counts=True))}}
    if with_cpu:
        import oracle_lib
        m = 20000
        t0 = time.perf_counter()
        oout, ohit, oit = oracle_lib.gjk_batch(a[:m], b[:m])
        cs = time.perf_counter() - t0
        same = bool(np.array_equal(ohit, status[:m]) and all(
            np.array_equal(np.ascontiguousarray(out[f][:m]).view(np.uint32)[~np.isnan(oout[f].reshape(m, -1)).any(axis=1)],
                           np.ascontiguousarray(oout[f]).view(np.uint32)[~np.isnan(oout[f].reshape(m, -1)).any(axis=1)]) for f in ("a", "b", "n", "t")))
        blk["cpu_port"] = {"pairs": m, "pairs_per_second": m / cs, "cores": 1, "bit_identical_on_sample": same}
    ctx.close()
    return blk


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mgf_b200", choices=["mgf_b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help=f"N = 1 only; default {DEFAULT_WORKLOAD} (plus short runs of the others under other_workloads)")
    ap.add_argument("--make-snapshot", default=None, choices=sorted(WORKLOADS), help="internal: GPU pre-roll of a workload, cached for the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true")
    ap.add_argument("--bodies-per-gpu", type=int, default=100000,
                    help="N > 1 only: 100000 (default, 50x40x50 per tile) or 250000 (50x40x125 per tile: 8 GPUs = the 2 M-sphere C4 scene)")
    ap.add_argument("--schedule", default="dataflow", choices=["dataflow", "phases"],
                    help="solver schedule (include/mgfb.h mgfb_solver_schedule)")
    args = ap.parse_args()
    if args.make_snapshot:
        g, _, _, _, snap = gpu_preroll(args.make_snapshot)
        save_snapshot(args.make_snapshot, snap)
        return
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

    dt = np.float32(DT)
    snap = None
    parity_check = None
    initial_state = "step 0 of the scene (no pre-roll)"
    if world == 1:
        name = args.workload or DEFAULT_WORKLOAD
        g, bodies, terrain, iters, snap = gpu_preroll(name, local_rank)
        if WORKLOADS[name][1]:
            save_snapshot(name, snap)
            initial_state = (f"steps 0..{WORKLOADS[name][1] - 1} run once on the GPU; the snapshot of that state (x, q, v, omega, colliders, stored fat "
                             "boxes) is what the CPU arm loads too")
        if args.schedule != "dataflow":
            g.ctx.close()
            g = mgf_b200.World(device=local_rank, solver_schedule=1)
            g.add_bodies(*bodies); g.set_terrain(*terrain); g.restore(snap)
        config = config_of(name, len(bodies[0]), iters)
        parallelism = "single GPU"
    else:
        # Before anything is timed: a small tiled world lock-stepped against ONE untiled world on the CPU port (the logic of
        # tests/mp_tiled_check.py, this process group, these GPUs): cross-GPU bit parity shown by the run that is measured.
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import mp_tiled_check
        parity_check = mp_tiled_check.check(rank, world, local_rank, 3, 0 if args.schedule == "dataflow" else 1)
        # ONE world of `world` x 100 000 spheres in one box, one slab per GPU (weak scaling).  Ghost
        # bodies and boundary velocities cross NVLink inside the kernels (csrc/tile.cuh); torch.distributed
        # only swaps the tiles' memory descriptors once, here.
        if args.bodies_per_gpu == 100000:
            # the N = 1 workload, N times: every GPU holds one settled C2 pile, the piles meet over the tile seams (weak scaling)
            tw, bodies, terrain, iters = build_settled_tiled(world, rank, local_rank, args.schedule)
            config = settled_tiles_config(world, len(bodies[0]), iters)
            initial_state = "every rank runs steps 0..599 of the C2 scene on its own GPU (identical bits) and loads the result, translated to its slot, into its tile"
        else:
            iters = 20
            tw, bodies, terrain = build_tiled(world, rank, local_rank, args.bodies_per_gpu, args.schedule)
            config = tiled_config(world, args.bodies_per_gpu, iters)
        g = tw.world
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
    t_wall0 = time.perf_counter()
    rows = time_steps_device(g, torch, flush, dt, iters, args.steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop(mark0, sampler.mark())
    tot = g.totals(reset=True)
    step_ms = [r["step_ms"] for r in rows]; solve_ms = [r["solve_ms"] for r in rows]; cons = [r["constraints"] for r in rows]
    pairs = [r["candidate_pairs"] + r["terrain_candidates"] for r in rows]; groups = [r["phases"] for r in rows]
    ghosts = [r["ghosts"] for r in rows]; bcons = [r["boundary_constraints"] for r in rows]
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
    if snap is not None:
        g.restore(snap)      # the e2e loop walks the same window again (its warm-up steps go through the pipelined calls, below)
    pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    hv, hw = pin((n, 3)), pin((n, 3))
    DEPTH = 3   # steps in flight (mgfb_step_enqueue allows 4): rides out host / PCIe jitter, which lock-stepped neighbour tiles otherwise amplify
    outs = [(pin((n, 3)), pin((n, 4)), pin((n, 3)), pin((n, 3))) for _ in range(DEPTH)]
    # the per-step inputs are external velocity increments (zero here, so the e2e loop steps the SAME world the
    # device-timed loop stepped)
    hv[:] = 0.0; hw[:] = 0.0
    from mgf_b200 import _lib as L
    lib, h = g.ctx.lib, g.ctx.h
    import ctypes as C
    st = L.StepStats()
    # warm-up THROUGH the pipelined calls: their first use creates the copy streams and allocates the staging buffers of every
    # slot (cudaMalloc / cudaMallocHost synchronise the device, with peer mappings on N GPUs for milliseconds)
    for k in range(args.warmup):
        hx, hq, ov, ow = outs[k % DEPTH]
        g.ctx.check(lib.mgfb_step_enqueue(h, dt, iters, L.INPUT_ADD, L.ptr(hv), L.ptr(hw), L.ptr(hx), L.ptr(hq), L.ptr(ov), L.ptr(ow)))
        if k + 1 >= DEPTH:
            g.ctx.check(lib.mgfb_step_wait(h, C.byref(st)))
    for k in range(min(args.warmup, DEPTH - 1)):
        g.ctx.check(lib.mgfb_step_wait(h, C.byref(st)))
    g.totals(reset=True)
    barrier()
    e2e_units = 0.0; e2e_dev_ms = 0.0
    t0 = time.perf_counter()
    waited = 0
    for k in range(args.steps):
        hx, hq, ov, ow = outs[k % DEPTH]
        g.ctx.check(lib.mgfb_step_enqueue(h, dt, iters, L.INPUT_ADD, L.ptr(hv), L.ptr(hw), L.ptr(hx), L.ptr(hq), L.ptr(ov), L.ptr(ow)))   # H2D + step + D2H queued
        if k - waited + 1 >= DEPTH:
            g.ctx.check(lib.mgfb_step_wait(h, C.byref(st)))   # the oldest queued step's state is in its buffer set (read here by a real consumer)
            e2e_units += st.constraints * iters; e2e_dev_ms += st.step_ms; waited += 1
    while waited < args.steps:
        g.ctx.check(lib.mgfb_step_wait(h, C.byref(st)))
        e2e_units += st.constraints * iters; e2e_dev_ms += st.step_ms; waited += 1
    e2e_api = f"mgfb_step_enqueue / mgfb_step_wait (pipelined, {DEPTH} steps in flight)"
    barrier()
    e2e_s = time.perf_counter() - t0
    tot_e2e = g.totals(reset=True)
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = t.item()
        u = torch.tensor([e2e_units], device="cuda", dtype=torch.float64); dist.all_reduce(u, op=dist.ReduceOp.SUM); e2e_units = u.item()
    e2e_val = e2e_units / e2e_s
    assert np.isfinite(outs[0][0]).all() and np.isfinite(outs[(args.steps - 1) % DEPTH][0]).all()

    # ---- roofline of the dominant kernel (the solver), live CUDA-event durations
    peak, peak_src = measured_peak()
    solve_s = sum(solve_ms) * 1e-3
    achieved = ALGO_BYTES_PER_CONSTRAINT_ITER * units / solve_s / 1e9
    dataflow = args.schedule == "dataflow"
    roofline = {"kernel": "k_solve_df (persistent sequential-impulse solver; rows wait on per-row inboxes filled by their predecessors, no grid barrier)" if dataflow
                else "k_solve (persistent cooperative sequential-impulse solver, one grid barrier per colour)", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one solver launch, `ncu --set full` (profiles/; ncu starts every replay
                # pass from a flushed L2); null until this round's capture of this workload is in profiles/
                "traffic": NCU_TRAFFIC.get((config["workload"].split(":")[0], dataflow and world == 1)), "traffic_unit": "bytes per launch (ncu, cold L2)",
                "l2_throughput_frac_ncu": NCU_L2_FRAC.get((config["workload"].split(":")[0], dataflow and world == 1)),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CONSTRAINT_ITER * units / len(solve_ms),
                "avg_launch_ms": sum(solve_ms) / len(solve_ms), "share_of_step": sum(solve_ms) / total_ms,
                "colours_per_iteration": sum(groups) / len(groups),
                "note": "268 B per constraint-iteration x constraints x 20 iterations per launch.  The rows are L2-resident at this size, so "
                        "`achieved` is ALGORITHMIC bytes served mostly from L2, not DRAM traffic: the kernel is bound by the latency of "
                        "its chain hand-overs (colours x iterations), see profiles/ and DESIGN.md section 4"}

    kernels = None; others = None; cpu = None
    if world == 1:
        # ---- the other phases of the step against their own algorithmic bytes (a separate, event-instrumented pass)
        if snap is not None:
            g.restore(snap); g.step(dt, iters, nsteps=args.warmup)
        kernels = phase_rooflines(g, torch, flush, dt, iters, 8, peak)
    # ---- CPU baseline (rank 0, N=1): bounded sample of the same workload on the oracle
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        o = oracle_lib.OracleWorld()
        o.add_bodies(*bodies); o.set_terrain(*terrain)
        if WORKLOADS[name][1]:
            o.restore(snap)
        ncpu = 8 if n > 10000 else 200
        o.step(dt, iters, 2)
        sec, ci, pr = o.time_steps(dt, iters, ncpu)
        cpu = {"value": ci / sec, "unit": "constraint-iters/s", "cores": 1, "kind": "port",
               "pairs_per_second": pr / sec, "ms_per_step": 1e3 * sec / ncpu,
               "sample": f"{ncpu} full World::step of the workload (from the same snapshot) after 2 warm-up steps on the C++ port of the reference "
                         f"(the Rust reference cannot be built: no rustc); 1 of {os.cpu_count()} host cores (reference is single-threaded)"}
    # ---- the north star's parity clause, measured by this very run: C1 stepped 250 times (free fall, impact, settling) in the
    # reference's own constraint order on the GPU and natively on the CPU port; max relative error of positions / velocities
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from mgf_b200 import scenes, _lib as LL
        cb, ct, ci = scenes.build_config("C1")
        pg = mgf_b200.World(device=local_rank, step_order=LL.STEP_ORDER_REFERENCE); po = oracle_lib.OracleWorld()
        for w in (pg, po):
            w.add_bodies(*cb); w.set_terrain(*ct)
        pg.step(dt, ci, nsteps=250); po.step(dt, ci, 250)
        (gx, _, gv, _), (ox, _, ov, _) = pg.state(), po.state()
        rel = lambda a, b: float((np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-30)).max())
        parity = {"scene": "C1 (512 spheres, 10 iterations), steps 0..249, mgfb_config.step_order = REFERENCE vs the CPU port's native World::step",
                  "max_rel_error_x": rel(gx, ox), "max_rel_error_v": rel(gv, ov), "bit_identical": bool(np.array_equal(gx.view(np.uint32), ox.view(np.uint32)) and
                                                                                                         np.array_equal(gv.view(np.uint32), ov.view(np.uint32))),
                  "north_star_tolerance": 1e-4}
        pg.ctx.close()
    if world == 1 and args.workload is None and not args.no_other_workloads:
        g.ctx.close()
        others = {}
        for other in WORKLOADS:
            if other != name:
                others[other] = short_run(other, torch, flush, local_rank, peak)
        others["gjk_epa_batch"] = gjk_block(local_rank, rank == 0 and not args.no_cpu_baseline)

    # ---- N = 8: BASELINE configs[3] (C4: 2 M spheres over 8 GPUs = 250 000 per GPU) as an extra block of the same run
    c4 = None
    if world == 8 and args.bodies_per_gpu == 100000:
        barrier()
        g.ctx.close()
        tw4, b4, _ = build_tiled(world, rank, local_rank, 250000, args.schedule)
        g4 = tw4.world
        barrier()
        g4.step(dt, iters, nsteps=args.warmup)
        barrier()
        r4 = time_steps_device(g4, torch, flush, dt, iters, min(args.steps, 20))
        barrier()
        ms4 = float(sum(r["step_ms"] for r in r4)); u4 = float(sum(r["constraints"] for r in r4)) * iters
        t = torch.tensor([ms4], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u = torch.tensor([u4], device="cuda", dtype=torch.float64); dist.all_reduce(u, op=dist.ReduceOp.SUM)
        c4 = {"config": tiled_config(world, 250000, iters), "value": u.item() / (t.item() * 1e-3), "unit": "constraint-iters/s", "steps": len(r4),
              "ms_per_step": t.item() / len(r4), "constraints_per_step": u.item() / iters / len(r4),
              "rank0_solve_ms": sum(r["solve_ms"] for r in r4) / len(r4), "rank0_ghosts_per_step": sum(r["ghosts"] for r in r4) / len(r4),
              "timing": "device (library CUDA events), max over ranks, L2 flushed between steps"}
        tw4.close()

    if rank == 0:
        line = {
            "metric": "contact_constraint_iterations_per_second", "value": value, "unit": "constraint-iters/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "work": {"constraints_per_step": units_all / iters / args.steps, "candidate_pairs_per_step": pairs_all / args.steps,
                     "colour_groups": sum(groups) / len(groups), "l2": "flushed between timed steps (256 MB write)",
                     # how the body pair set was found in the timed steps (mgfb_step_stats.broadphase_path; the set is the same on every path)
                     "broadphase_steps": {label: sum(1 for r in rows if r.get("broadphase_path") == code)
                                          for code, label in ((0, "grid_and_sweep"), (1, "coherent_cache"), (2, "cache_rebuilt"), (3, "sweep_chosen_by_cache"))},
                     "parallelism": parallelism, "bodies_per_gpu": n,
                     "rank0_ghosts_per_step": sum(ghosts) / len(ghosts), "rank0_boundary_constraints_per_step": sum(bcons) / len(bcons),
                     "arithmetic": "--fmad=false, IEEE div/sqrt: bit-exact vs the CPU port",
                     "initial_state": initial_state},
            "narrowphase_pairs_per_second": pairs_all / (total_ms_max * 1e-3),
            "solver_only_constraint_iters_per_second": units / solve_s,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "constraint-iters/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 52,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "api": e2e_api,
                    "device_ms_per_step": e2e_dev_ms / args.steps, "constraints_per_step": e2e_units / iters / args.steps / max(world, 1),
                    "gpu_launches": tot_e2e["kernel_launches"]},
            "gpu_launches": tot["kernel_launches"], "clocks": clocks, "wall_s_timed_region": t_wall,
            "other_workloads": others, "reference_order_parity": parity, "parity_check": parity_check, "c4": c4,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ncu `--set full` captures of one solver launch (profiles/): (workload, dataflow single GPU) -> DRAM bytes, L2 throughput fraction
# dram__bytes_read.sum + dram__bytes_write.sum and lts__throughput of ONE k_solve_df launch (ncu --set full, profiles/r02_v1_k_solve_df_raw.csv
# for C2settled, profiles/r01_v4_k_solve_df_raw.csv for C2pile)
NCU_TRAFFIC = {("C2pile", True): 145.2e6, ("C2settled", True): 38.2e6}
NCU_L2_FRAC = {("C2pile", True): 0.32, ("C2settled", True): 0.24}


if __name__ == "__main__":
    main()
