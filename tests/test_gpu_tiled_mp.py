"""One process per GPU (torchrun), tiles wired through cudaIpc peer mappings: needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nranks", [2, 4])
def test_tiled_world_one_process_per_gpu_bit_exact(nranks):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + nranks), os.path.join(ROOT, "tests", "mp_tiled_check.py"), "12"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    assert r.returncode == 0 and "MP_TILED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
