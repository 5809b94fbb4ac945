"""CPU coverage of the multi-GPU path's host side: plan + wiring over gloo, world_size 2."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_is_balanced_sorted_and_deterministic():
    from mgf_b200 import tiling
    x = np.array([3.0, -1.0, 2.0, 2.0, 0.5, 9.0, -4.0], np.float32)
    parts = tiling.slab_partition(x, 3)
    assert sorted(np.concatenate(parts).tolist()) == list(range(7))
    assert [len(p) for p in parts] == [2, 2, 3]
    assert all(np.all(np.diff(p.astype(np.int64)) > 0) for p in parts)
    assert max(x[parts[0]]) <= min(x[parts[1]]) and max(x[parts[1]]) <= min(x[parts[2]])
    again = tiling.slab_partition(x, 3)
    assert all(np.array_equal(p, q) for p, q in zip(parts, again))


def test_tiling_host_logic_two_ranks_gloo():
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29633", LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_tiling_cpu.py")], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[1][-2000:] for o in outs)
    assert "TILING_CPU_OK" in outs[0][0]
