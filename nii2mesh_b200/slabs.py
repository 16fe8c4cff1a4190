"""Host-side helpers of the z-slab path: how a volume is cut, how the library's NCCL communicator
is bootstrapped from an existing torch.distributed group (plumbing only: 128 bytes), how the per-rank blocks of
a slab run are put together.  No CUDA here: everything in this file runs (and is tested) on CPU ranks too."""
import numpy as np


def partition(nz, world, min_planes=4):
    """z boundaries [0, ..., nz] of `world` contiguous slabs, as even as possible; raster order = rank order"""
    if world < 1 or nz < world * (min_planes if world > 1 else 1):
        raise ValueError(f"cannot cut {nz} planes into {world} slabs of >= {min_planes} planes")
    cuts = [(i * nz) // world for i in range(world + 1)]
    return cuts


def broadcast_id(dist, rank, make_id, device=None):
    """rank 0 calls make_id() -> 128 bytes; every rank returns the same bytes (works on gloo and nccl groups)"""
    import torch
    t = torch.zeros(128, dtype=torch.uint8, device=device or "cpu")
    if rank == 0:
        raw = bytes(make_id())
        if len(raw) != 128:
            raise ValueError("an NCCL unique id is 128 bytes")
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def nccl_comm_from_torch(eng, dist, rank, world):
    """the library's own NCCL communicator (b2m_comm_create_nccl), id carried by the torch group"""
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if dev == "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    return eng.nccl_comm(broadcast_id(dist, rank, eng.nccl_unique_id, dev), rank, world)


def part_of(r, verts, tris):
    """plain-python description of one rank's blocks (picklable: SlabResult is a ctypes struct)"""
    return dict(nverts=r.r.nverts, ntris=r.r.ntris, nv_edge=r.nv_edge, nv_cent=r.nv_cent, nv_extra=r.nv_extra,
                ntris_local=r.ntris_local, v_edge_off=r.v_edge_off, v_cent_off=r.v_cent_off, v_extra_off=r.v_extra_off,
                tri_off=r.tri_off, verts=verts, tris=tris)


def gather_parts(dist, rank, world, r, verts, tris):
    """all ranks' parts on rank 0 (None elsewhere)"""
    out = [None] * world if rank == 0 else None
    dist.gather_object(part_of(r, verts, tris), out, dst=0)
    return out


def assemble(parts):
    """[part dict] of all ranks -> (verts[nverts,3] f64, tris[ntris,3] i32) of the whole volume.  Checks that the
    blocks tile the arrays exactly (no gap, no overlap)."""
    nv, nt = parts[0]["nverts"], parts[0]["ntris"]
    V = np.empty((nv, 3), np.float64)
    T = np.empty((nt, 3), np.int32)
    seen_v = np.zeros(nv, np.uint8)
    seen_t = np.zeros(nt, np.uint8)
    for p in parts:
        if (p["nverts"], p["ntris"]) != (nv, nt):
            raise ValueError("ranks disagree on the global counts")
        ne, nc, nx = p["nv_edge"], p["nv_cent"], p["nv_extra"]
        v = p["verts"]
        for off, a, b in ((p["v_edge_off"], 0, ne), (p["v_cent_off"], ne, ne + nc), (p["v_extra_off"], ne + nc, ne + nc + nx)):
            V[off:off + (b - a)] = v[a:b]
            seen_v[off:off + (b - a)] += 1
        T[p["tri_off"]:p["tri_off"] + p["ntris_local"]] = p["tris"]
        seen_t[p["tri_off"]:p["tri_off"] + p["ntris_local"]] += 1
    if not (np.all(seen_v == 1) and np.all(seen_t == 1)):
        raise ValueError("slab blocks do not tile the assembled mesh")
    return V, T
