"""CPU tests of the host side of the z-slab path (world_size 2, gloo): partitioning, the 128-byte communicator id
carried by a torch.distributed group, gathering and assembling the per-rank mesh blocks.  No GPU, no oracle."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import slabs  # noqa: E402


def test_partition_covers_and_orders():
    for nz in (8, 9, 40, 142, 1024, 2048):
        for w in (1, 2, 3, 4, 8):
            if nz < 4 * w and w > 1:
                with pytest.raises(ValueError):
                    slabs.partition(nz, w)
                continue
            c = slabs.partition(nz, w)
            assert c[0] == 0 and c[-1] == nz and len(c) == w + 1
            sizes = [b - a for a, b in zip(c, c[1:])]
            assert min(sizes) >= (4 if w > 1 else 1) and max(sizes) - min(sizes) <= 1


def _fake_parts(world, rng):
    """a random welded mesh cut the way the library cuts it: per rank an edge block, a centroid block (all centroid
    blocks follow all edge blocks), extras on the last rank, triangles in rank order"""
    ne = rng.integers(3, 20, world)
    nc = rng.integers(0, 6, world)
    nx = 2
    nt = rng.integers(1, 30, world)
    NV, NT = int(ne.sum() + nc.sum() + nx), int(nt.sum())
    V = rng.standard_normal((NV, 3))
    T = rng.integers(0, NV, (NT, 3)).astype(np.int32)
    parts, eo, co, to = [], 0, int(ne.sum()), 0
    for r in range(world):
        x = nx if r == world - 1 else 0
        v = np.concatenate([V[eo:eo + ne[r]], V[co:co + nc[r]], V[NV - nx:NV][:x]])
        parts.append(dict(nverts=NV, ntris=NT, nv_edge=int(ne[r]), nv_cent=int(nc[r]), nv_extra=x, ntris_local=int(nt[r]),
                          v_edge_off=eo, v_cent_off=co, v_extra_off=NV - nx, tri_off=to, verts=v, tris=T[to:to + nt[r]]))
        eo += int(ne[r]); co += int(nc[r]); to += int(nt[r])
    return parts, V, T


def test_assemble_blocks():
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 8):
        parts, V, T = _fake_parts(world, rng)
        v, t = slabs.assemble(parts)
        assert np.array_equal(v, V) and np.array_equal(t, T)
    parts, V, T = _fake_parts(3, rng)
    parts[1]["tri_off"] += 1  # overlap / gap must be detected
    with pytest.raises(ValueError):
        slabs.assemble(parts)


WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["B2M_ROOT"]); sys.path.insert(0, os.path.join(os.environ["B2M_ROOT"], "tests"))
import torch.distributed as dist
from nii2mesh_b200 import slabs
import test_slabs_host as T
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
ident = slabs.broadcast_id(dist, rank, lambda: bytes(range(128)))
assert ident == bytes(range(128)), "id broadcast"
cuts = slabs.partition(40, world)
parts, V, Tr = T._fake_parts(world, np.random.default_rng(11))   # same seed on every rank
mine = parts[rank]
class R:  # what lib.SlabResult exposes
    pass
r = R(); r.r = R(); r.r.nverts, r.r.ntris = mine["nverts"], mine["ntris"]
for k in ("nv_edge", "nv_cent", "nv_extra", "ntris_local", "v_edge_off", "v_cent_off", "v_extra_off", "tri_off"):
    setattr(r, k, mine[k])
got = slabs.gather_parts(dist, rank, world, r, mine["verts"], mine["tris"])
if rank == 0:
    v, t = slabs.assemble(got)
    assert np.array_equal(v, V) and np.array_equal(t, Tr)
    print("HOST_SLABS_OK", cuts)
dist.barrier()
dist.destroy_process_group()
'''


def test_world2_gloo_id_gather_assemble(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(w)]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300,
                       env=dict(os.environ, B2M_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1"))
    assert p.returncode == 0 and "HOST_SLABS_OK" in p.stdout, p.stdout[-3000:]
