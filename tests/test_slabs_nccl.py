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


def test_two_devices_in_one_process(libb2m):
    """one process driving two GPUs (local slab group over devices 0 and 1, then a plain Engine on device 1): per-device
    state such as the opt-in to > 48 KB of dynamic shared memory must not be remembered process-wide"""
    if libb2m.b2m_device_count() < 2:
        pytest.skip("one GPU on this box")
    import numpy as np
    import cases
    from nii2mesh_b200 import lib
    vol, iso = cases.volumes(big=False)["gyroid96"]
    e0 = lib.Engine(0)
    v0, t0, _ = e0.meshify(vol, iso, 0, 1, 1, 1)
    e1 = lib.Engine(1)
    v1, t1, _ = e1.meshify(vol, iso, 0, 1, 1, 1)
    assert np.array_equal(t0, t1) and np.array_equal(v0.view(np.uint64), v1.view(np.uint64))
    grp = lib.LocalSlabGroup(2, devices=[0, 1])
    nz = vol.shape[0]
    gv, gt, _ = grp.meshify(vol, [0, nz // 2, nz], iso, original_mc=0, pre_smooth=1, only_largest=1, fill_bubbles=1)
    grp.close()
    assert np.array_equal(gt, t0) and np.array_equal(gv.view(np.uint64), v0.view(np.uint64))
