"""NCCL transport of the z-slab path: needs >= 2 GPUs on the box (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run_worker(libb2m, env, port, extra=(), timeout=900):
    ndev = libb2m.b2m_device_count()
    if ndev < 2:
        pytest.skip("one GPU on this box: the NCCL transport needs two (run under gpurun --gpus 2)")
    world = 4 if ndev >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "slab_nccl_worker.py"), *extra]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout,
                       env=dict(os.environ, MASTER_ADDR="127.0.0.1", **env))
    sys.stdout.write(p.stdout[-4000:])
    return p


def test_slabs_over_nccl(libb2m):
    """default transport: scalar blocks through the node-local host segment, fast (sort-free) seam merge"""
    p = _run_worker(libb2m, {}, 29533)
    assert p.returncode == 0 and "SLAB_NCCL_RESULT PASS" in p.stdout


@pytest.mark.parametrize("env", [{"B2M_SCALARS_NCCL": "1"}, {"B2M_SEAM_ENT_CAP": "2", "B2M_SEAM_PAIR_CAP": "8"}, {"B2M_SEAM_SLOW": "1"}],
                         ids=["scalars_by_nccl", "seam_block_overflow", "seam_slow_path"])
def test_slabs_over_nccl_other_paths(libb2m, env):
    """the scalar exchange as an NCCL all-gather (what ranks on different nodes use), and the sorted-list seam merge
    (forced, and reached through the overflow of a tiny seam block)"""
    p = _run_worker(libb2m, dict(env, B2M_TEST_VOLS="blobs2,bet"), 29534)
    assert p.returncode == 0 and "SLAB_NCCL_RESULT PASS" in p.stdout


def test_failed_rank_releases_its_peers(libb2m):
    """a rank that fails on its own must not leave the others inside a collective: it poisons the host
    segment, the peers' host waits notice, abort the communicator and return an error"""
    if not os.environ.get("B2M_TEST_ABORT"):
        pytest.skip("opt-in (B2M_TEST_ABORT=1): a peer that is NOT released costs the whole timeout in GPU time")
    p = _run_worker(libb2m, {}, 29535, extra=("--inject-failure",), timeout=150)
    assert p.returncode == 0 and "SLAB_NCCL_ABORT PASS" in p.stdout


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
