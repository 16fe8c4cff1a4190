#!/usr/bin/env python3
"""torchrun worker of tests/test_slabs_nccl.py: N ranks (one GPU each) mesh ONE volume as N z-slabs over the
NCCL transport of libb2m (b2m_comm_create_nccl); rank 0 assembles the blocks and compares them, bit for bit,
with the single-GPU path on the same volume.  torch.distributed only carries the NCCL id and the results."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def inject_failure(eng, comm, dist, rank, world, vols):
    """the last rank hands in a slab that does not fit (usage error, detected locally); every other rank must come back
    with an error instead of waiting for ever in the first collective"""
    import torch
    from nii2mesh_b200 import lib
    vol, iso = vols["blobs2"]
    nz = vol.shape[0]
    cuts = [(i * nz) // world for i in range(world + 1)]
    lo, hi = cuts[rank], cuts[rank + 1]
    # a good call first: NCCL sets its peer connections up lazily, inside the first send / receive, where no host wait of
    # ours can see the poison flag
    eng.meshify_slab_host(comm, vol[lo:hi], vol.shape, lo, iso, original_mc=0, pre_smooth=1, only_largest=1, fill_bubbles=1)
    if rank == world - 1:
        hi -= 1            # one plane short of the volume: rejected by the argument check of this rank only
    failed = False
    try:
        eng.meshify_slab_host(comm, vol[lo:hi], vol.shape, lo, iso, original_mc=0, pre_smooth=1, only_largest=1, fill_bubbles=1)
    except lib.B2MError as ex:
        failed = True
        print(f"rank {rank}: {ex}", flush=True)
    flag = torch.tensor([0 if failed else 1], device="cuda")
    dist.all_reduce(flag)
    eng.lib.b2m_comm_destroy(comm)
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB_NCCL_ABORT", "PASS" if int(flag.item()) == 0 else "FAIL: some rank did not fail", flush=True)
    sys.exit(0)


def main():
    import torch
    import torch.distributed as dist
    from nii2mesh_b200 import lib, slabs
    import cases
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = lib.Engine(local)
    comm = slabs.nccl_comm_from_torch(eng, dist, rank, world)
    vols = cases.volumes()
    failures = 0
    if "--inject-failure" in sys.argv:
        return inject_failure(eng, comm, dist, rank, world, vols)
    names = os.environ.get("B2M_TEST_VOLS", "blobs2,gyroid160,sphere64,bet").split(",")
    for name in names:
        vol, iso = vols[name]
        cuts = slabs.partition(vol.shape[0], world)
        for backend, omc, ps, ol, fb in cases.flag_sets(name):
            flags = dict(original_mc=omc, pre_smooth=ps, only_largest=ol, fill_bubbles=fb, backend=backend)
            r, v, t = eng.meshify_slab_host(comm, vol[cuts[rank]:cuts[rank + 1]], vol.shape, cuts[rank], iso, **flags)
            parts = slabs.gather_parts(dist, rank, world, r, v, t)
            if rank == 0:
                V, T = slabs.assemble(parts)
                sv, st, sr = eng.meshify(vol, iso, omc, ps, ol, fb, backend)
                ok = (np.array_equal(T, st) and np.array_equal(V.view(np.uint64), sv.view(np.uint64)))
                print(f"{name} b{backend} o{omc} p{ps} l{ol} f{fb}: {len(V)} verts {len(T)} tris {'OK' if ok else 'MISMATCH'}", flush=True)
                failures += not ok
    flag = torch.tensor([failures], device="cuda")
    dist.broadcast(flag, 0)
    eng.lib.b2m_comm_destroy(comm)
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB_NCCL_RESULT", "PASS" if failures == 0 else f"FAIL {failures}", flush=True)
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
