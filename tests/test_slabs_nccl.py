"""NCCL transport of the z-slab path: needs >= 2 GPUs on the box (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_slabs_over_nccl(libb2m):
    ndev = libb2m.b2m_device_count()
    if ndev < 2:
        pytest.skip("one GPU on this box: the NCCL transport needs two (run under gpurun --gpus 2)")
    world = 4 if ndev >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tests" / "slab_nccl_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900,
                       env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    sys.stdout.write(p.stdout[-4000:])
    assert p.returncode == 0 and "SLAB_NCCL_RESULT PASS" in p.stdout
