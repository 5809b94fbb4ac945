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


def test_slab_partition_min_width_widens_thin_slabs():
    import pytest
    from mgf_b200 import tiling
    x = np.concatenate([np.linspace(0.0, 1.0, 90), np.linspace(1.0, 10.0, 10)]).astype(np.float32)   # 90 % of the bodies in the first tenth
    thin = tiling.slab_partition(x, 4)
    assert x[thin[1]].max() - x[thin[1]].min() < 0.5
    wide = tiling.slab_partition(x, 4, min_width=2.0)
    assert sorted(np.concatenate(wide).tolist()) == list(range(100)) and all(len(p) > 0 for p in wide)
    firsts = [x[p].min() for p in wide]
    assert all(b - a >= 2.0 - 1e-6 for a, b in zip(firsts, firsts[1:])) and x.max() - firsts[-1] >= 2.0 - 1e-6
    assert all(x[wide[r]].max() <= x[wide[r + 1]].min() for r in range(3))
    with pytest.raises(ValueError):
        tiling.slab_partition(x, 4, min_width=3.0)


def test_tiling_host_logic_two_ranks_gloo():
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29633", LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_tiling_cpu.py")], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[1][-2000:] for o in outs)
    assert "TILING_CPU_OK" in outs[0][0]
