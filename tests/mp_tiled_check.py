"""Multi-process, multi-GPU parity check of the tiled world (run under torchrun, one rank per GPU;
driven by tests/test_gpu_tiled_mp.py).  Rank 0 also holds the oracle: ONE untiled world with every
body.  Every step the ranks' constraint lists are gathered, the oracle replays the executed order
(tiling.executed_order) and each rank compares its owned bodies with the oracle bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mgf_b200 import scenes, tiling  # noqa: E402


def check(rank, world, local, nsteps, schedule=0):
    """The lock-step itself; the process group (any backend able to all_gather objects) must exist.  Returns a summary dict on
    every rank (raises AssertionError on the first difference)."""
    dt = np.float32(1.0 / 60.0); iters = 10
    bodies = scenes.pile_xyz(12 * world, 5, 6, jitter=0.01, seed=11)
    terrain = scenes.box_terrain(10.0 * world, 10.0, 6.0)
    parts = tiling.slab_partition(tiling.shape_centres_x(bodies[0]), world)
    tw = tiling.TiledWorld(rank, world, device=local, tile_timeout_ms=10000, solver_schedule=schedule)
    tw.add_bodies(parts[rank], *bodies)
    tw.set_terrain(*terrain)
    tw.connect(tiling.all_gather_bytes, ghost_capacity=2048)
    o = None
    if rank == 0:
        import oracle_lib
        o = oracle_lib.OracleWorld()
        o.add_bodies(*bodies); o.set_terrain(*terrain)
    dist.barrier()
    total = boundary = 0
    for s in range(nsteps):
        st = tw.step(dt, iters)
        mine = tw.constraints()
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
        counts = [None] * world
        dist.all_gather_object(counts, (st["constraints"], st["boundary_constraints"], st["ghosts"]))
        expect = [None]
        if rank == 0:
            m = o.build(dt)
            assert sum(c[0] for c in counts) == m, f"step {s}: {counts} vs oracle {m}"
            if m:
                ga, gb, gf, gs = tiling.executed_order(per_rank)
                oa, ob, of, osub = o.constraints(m)
                index = {k: i for i, k in enumerate(zip(oa.tolist(), ob.tolist(), of.tolist(), osub.tolist()))}
                perm = np.array([index[k] for k in zip(ga.tolist(), gb.tolist(), gf.tolist(), gs.tolist())], dtype=np.uint32)
                assert len(set(perm.tolist())) == m
                o.solve_order(perm, iters)
            expect = [o.state()]
            total += m; boundary += sum(c[1] for c in counts)
        dist.broadcast_object_list(expect, src=0)
        for name, sg, so in zip("x q v omega".split(), tw.state(), expect[0]):
            so = so[tw.ids]
            bad = np.nonzero((np.ascontiguousarray(sg).view(np.uint32) != np.ascontiguousarray(so).view(np.uint32)).any(axis=1))[0]
            assert len(bad) == 0, f"rank {rank} step {s}: {name} differs for {len(bad)} bodies"
    summary = [None]
    if rank == 0:
        assert total > 0 and boundary > 0, (total, boundary)
        summary = [{"ranks": world, "steps": nsteps, "bodies": len(bodies[0]), "constraints": total, "boundary_constraints": boundary,
                    "result": "ok: every rank's bodies bit-identical to ONE untiled world on the CPU port replaying the executed order, every step"}]
    dist.broadcast_object_list(summary, src=0)
    dist.barrier()
    tw.close()
    return summary[0]


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = check(rank, world, local, nsteps)
    if rank == 0:
        print(f"MP_TILED_OK ranks={world} steps={nsteps} constraints={out['constraints']} boundary={out['boundary_constraints']}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
