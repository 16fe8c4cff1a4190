#!/usr/bin/env python3
"""bench.py — Gvoxels/s of the meshify() hot path (smooth + CC + MC + weld) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 1024] [--volume 2048] [--impl reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): synthetic "gyroid + bumps"
float32 volume (128^3 period), Lewiner MC33, -p 1 -l 1 -b 1, isolevel 0.  A step is one whole meshify() pass.
  N = 1 : G1024, 1024^3 voxels on one B200.
  N > 1 : ONE volume of N * 2^30 voxels cut into N z-slabs, one rank (process, GPU) per slab, NCCL over NVLink
          for the halo planes / seam lists (b2m_meshify_slab): 2048x1024x1024, 2048x2048x1024 and, at N = 8,
          2048^3 (BASELINE configs[4]).  Per-GPU work is fixed (2^30 voxels): weak scaling.

  --volume V : STRONG scaling instead: ONE V^3 volume (BASELINE configs[4] as written: 2048^3) cut into N z-slabs whatever
          N is (a slab holds up to 2^32 voxels, so 2048^3 runs on 2, 4 and 8 GPUs), known answer asserted.
  N > 1 also runs a parity pre-pass (config.parity): a G256 volume meshed as N slabs over NCCL, assembled on rank 0 and
          compared with rank 0's single-GPU mesh and with the digest recorded from the unmodified reference.

  value : volume resident in HBM, mesh left in HBM; K steps between one CUDA-event pair on the library's
          stream (b2m_timer_start/stop), barrier + synchronize on both sides, max over ranks.
  e2e   : the same through the reference-facing C entry point — meshify() (include/meshify.h) at N = 1,
          b2m_meshify_slab_host() at N > 1 — with a pinned HOST volume in and malloc()'d HOST mesh out, copies
          inside the timed region.
  roofline : the dominant kernel of the step, from per-launch CUDA-event pairs recorded live during the timed
          steps (b2m_set_profile); algorithmic bytes per launch are in kernel_bytes() below,
          peak = MEASURED_PEAKS.json hbm_gbs (fallback 6650 GB/s).
  cpu_baseline : the unmodified reference (oracle/_ref, built by oracle/build_ref.sh) on a bounded sample of the
          same generator and flags; meshify() is single-threaded, so one volume per host core runs concurrently.
  --impl reference : only that CPU reference, each step = one G256 volume per host core.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLAGS = dict(original_mc=0, pre_smooth=1, only_largest=1, fill_bubbles=1, backend=0)
ISO = 0.0
METRIC = "Gvoxels/s meshify (smooth+CC+MC+weld)"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed ncu capture of this
    round's kernels (profiles/r2_ncu_full_summary.csv: G1024, same flags, `ncu --set full` of tools/profile_step.py;
    a number taken under ncu can only be a committed one), or None"""
    import csv
    p = ROOT / "profiles" / "r2_ncu_full_summary.csv"
    if not p.exists():
        return None
    try:
        rows = list(csv.reader(p.open()))
        hdr = rows[0]
        ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(rows[1][ir], 1e9)
        best = None
        for r in rows[2:]:
            if ("k_" + kernel) in r[ik]:
                t = (float(r[ir]) + float(r[iw])) * unit
                best = t if best is None else max(best, t)   # several launches (6-/18-connectivity): the larger
        return best
    except Exception:  # noqa: BLE001
        return None


def kernel_bytes(name, n, nv, nt, nwords):
    """ALGORITHMIC bytes one launch of kernel `name` must move: n voxels f32 (per GPU),
    nwords 32-voxel bit words, nv/nt mesh vertices/triangles (per GPU)."""
    table = {
        "smooth3": 8 * n,                        # R V + W V
        "minmax": 4 * n,
        "threshold": 4 * n + nwords * 4,         # R V + W bits
        "mc_classify": 4 * n,                    # R V (composed on the fly)
        "mc_emit": 24 * nv + 12 * nt,            # surface term S
        "tri_degen": 12 * nt + nt // 4,           # R indices + W keep bits/counts (positions only for the few near-corner ones)
        "tri_finish": 24 * nt,                   # R indices + W renumbered, compacted indices
        "cc_local": nwords * 4, "cc_border": nwords * 4, "cc_flatten": nwords * 4, "cc_best": nwords * 4,
        "cc_select": nwords * 8, "dilate_bbox": nwords * 12,
    }
    return table.get(name)


def known_counts(gshape):
    """PRE-weld vertex / triangle counts of the reference for a G volume of gshape = (nz, ny, nx), all multiples of 128:
    exact polynomial in the tiles per axis (a, b, c), fitted on reference runs of 10 small boxes and checked on 6 held-out
    ones (tools/make_golden_big.py -> tests/golden/golden_big.json "boxlaw"); for cubes it is SURVEY.md 8e's cubic law."""
    if any(d % 128 for d in gshape):
        return None
    c, b, a = (d // 128 for d in gshape)
    f = lambda k3, k2, k1: k3 * a * b * c + k2 * (a * b + b * c + c * a) + k1 * (a + b + c)  # noqa: E731
    return f(79416, 15152, -470), f(158848, 30304, -944)


def workload_string(world, n, gshape, strong=False):
    """config.workload: the same string on the b2m arm and on the reference arm"""
    if world == 1:
        return f"G{n} gyroid+bumps {n}^3 f32, Lewiner MC33 -p1 -l1 -b1 iso 0 (BASELINE configs[2])"
    N = gshape[0] * gshape[1] * gshape[2] // world
    return (f"gyroid+bumps {gshape[2]}x{gshape[1]}x{gshape[0]} f32 ({world} x {N} voxels), Lewiner MC33 "
            f"-p1 -l1 -b1 iso 0, ONE volume in {world} z-slabs (BASELINE configs[4]" +
            (" as written: the same cube at every GPU count)" if strong else " family; 2048^3 at 8 GPUs)"))


def slab_volume(world, n):
    """global (nz, ny, nx) of the N-GPU workload: N * n^3 voxels, doubling x, then y, then z"""
    nz = ny = nx = n
    k = world
    for ax in ("x", "y", "z") * 4:
        if k <= 1:
            break
        if ax == "x":
            nx *= 2
        elif ax == "y":
            ny *= 2
        else:
            nz *= 2
        k //= 2
    assert nx * ny * nz == world * n ** 3, "--gpus must be a power of two"
    return nz, ny, nx


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons of one GPU sampled DURING the timed region: NVML in-process (nvidia_ml_py;
    a sample every 20 ms from the first millisecond), `nvidia-smi -lms 100` as the fallback (its start-up alone can
    outlast a short timed region on an 8-GPU box)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc, self.stop_flag, self.nvml, self.max_mhz = index, [], None, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:  # noqa: BLE001
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = (pynvml, h)
        except Exception:  # noqa: BLE001
            self.nvml = None

    def run(self):
        if self.nvml:
            nv, h = self.nvml
            while not self.stop_flag:
                try:
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:  # noqa: BLE001
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), nv.nvmlDeviceGetPowerUsage(h) / 1000.0, int(mask)))
                except Exception:  # noqa: BLE001
                    break
                time.sleep(0.02)
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                c = [x.strip() for x in line.split(",")]
                if len(c) >= 9:
                    mask = sum(bit for (name, bit), col in zip(self.REASONS, (8, 5, 7, 6)) if c[col].lower().startswith("active"))
                    self.max_mhz = float(c[2])
                    self.rows.append((float(c[1]), float(c[3]), mask))
        except Exception:  # noqa: BLE001
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2.0)
        rows = list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        sm = sorted(r[0] for r in rows)
        mask = 0
        for r in rows:
            mask |= r[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(n for n, bit in self.REASONS if mask & bit),
                "samples": len(rows), "power_w_max": max(r[1] for r in rows), "source": "nvml" if self.nvml else "nvidia-smi"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip() + f" ({os.cpu_count()} logical cores)"
    except Exception:  # noqa: BLE001
        pass
    return f"unknown ({os.cpu_count()} logical cores)"


def ref_meshify_time(n, steps, warmup, threads):
    """time the unmodified reference (oracle/_ref; falls back to the oracle port) on `threads` concurrent
    G<n> volumes per step (meshify() itself is single-threaded; ctypes releases the GIL)"""
    from nii2mesh_b200 import synth
    import oracle
    vol = synth.gyroid(n)
    if oracle.ref_available("lewiner"):
        R, kind = oracle.Ref("lewiner"), "reference"
        fn = lambda: R.meshify(vol, ISO, 0, 1, 1, 1)  # noqa: E731
    else:
        O, kind = oracle.Oracle(), "port"
        fn = lambda: O.meshify(vol, ISO, 0, 1, 1, 1, 0)  # noqa: E731
    last = [None] * threads

    def one_step():
        def work(i):
            last[i] = fn()
        th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        for t in th:
            t.start()
        for t in th:
            t.join()
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    r = last[0]
    assert all(x["rc"] == 0 for x in last)
    return dt, kind, len(r["verts"]), len(r["tris"])


def host_threads():
    try:
        c = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        c = os.cpu_count() or 1
    return max(1, min(c, 32))


def run_reference(args, rank, emit):
    if rank != 0:
        return
    n, T = 256, host_threads()
    dt, kind, nv, nt = ref_meshify_time(n, args.steps, min(args.warmup, 1), T)
    val = T * n ** 3 / dt / 1e9
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gvoxels/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_string(max(1, args.gpus), args.size, slab_volume(max(1, args.gpus), args.size)),
                      "sample": f"each step = {T} G{n} volumes of the same generator and flags, one per host core (bounded sample "
                                f"of the workload: the reference's per-voxel rate is flat in the volume size, BASELINE.md section 3)"},
           "cpu_baseline": {"value": val, "unit": "Gvoxels/s", "cores": T, "per_core_mvox_s": n ** 3 / dt / 1e6, "kind": kind, "cpu": cpu_model(),
                            "sample": f"{T} x G{n} ({n}^3 voxels) per step, {nv} verts {nt} tris each; meshify() is "
                                      f"single-threaded, one volume per core"},
           "e2e": {"value": val, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


def verify_slabs(eng, comm, dist, rank, world, tile):
    """parity pre-pass of every multi-GPU run: G256 (BASELINE config 3 family, Lewiner -p1 -l1 -b1) meshed as `world`
    z-slabs over NCCL, the blocks gathered and assembled on rank 0, against (a) rank 0's single-GPU mesh of the same
    volume, array for array, and (b) the topology digest recorded from the unmodified reference"""
    from nii2mesh_b200 import slabs
    sys.path.insert(0, str(ROOT))
    n = 256
    nzl = n // world
    d = eng.tiled_volume(tile, (nzl, n, n), z_offset=rank * nzl)
    try:
        sr = eng.meshify_slab(comm, d, (n, n, n), rank * nzl, ISO, **FLAGS)
        v, t = eng.fetch_slab(sr)
    finally:
        d.free()
    parts = slabs.gather_parts(dist, rank, world, sr, v, t)
    if rank != 0:
        return None
    V, T = slabs.assemble(parts)
    dw = eng.tiled_volume(tile, (n, n, n))
    try:
        sv, st, r1 = eng.meshify_device(dw, ISO, **FLAGS)
    finally:
        dw.free()
    same = bool(np.array_equal(T, st) and np.array_equal(V.view(np.uint64), sv.view(np.uint64)))
    out = {"volume": "G256 in %d z-slabs over NCCL, assembled" % world, "nverts": int(len(V)), "ntris": int(len(T)),
           "equals_single_gpu_arrays": same}
    try:
        from oracle.canon import topology_digest
        g = json.loads((ROOT / "tests" / "golden" / "golden_big.json").read_text())["gyroid"]["256"]
        out["equals_reference_digest"] = bool(topology_digest(V, T)[2] == g["digest"] and (len(V), len(T)) == (g["nverts"], g["ntris"]))
    except Exception as ex:  # noqa: BLE001
        out["equals_reference_digest"] = None
        out["note"] = f"reference digest not checked: {ex}"
    assert same and out["equals_reference_digest"] is not False, f"slab parity pre-pass failed: {out}"
    return out


def main():
    # ONE JSON line on stdout: anything else a library prints there (NCCL's version banner, ...) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=1024, help="cube edge of the per-GPU volume (multiple of 128)")
    ap.add_argument("--impl", default="b2m")
    ap.add_argument("--volume", type=int, default=0, help="strong scaling: ONE cube of this edge (e.g. 2048) cut into --gpus z-slabs")
    ap.add_argument("--no-verify", action="store_true", help="skip the N > 1 parity pass")
    ap.add_argument("--verify-first", action="store_true", help="run the parity pass before the timed steps instead of after them")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, emit)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from nii2mesh_b200 import lib, synth
    eng = lib.Engine(local)
    n = args.size
    strong = args.volume > 0
    if strong:
        assert world > 1 and args.volume % world == 0, "--volume needs --gpus > 1 slabs that divide it"
        gshape = (args.volume,) * 3
    else:
        gshape = slab_volume(world, n)      # (nz, ny, nx) of the whole volume
    GN = gshape[0] * gshape[1] * gshape[2]
    nzl = gshape[0] // world
    z0 = rank * nzl
    sshape = (nzl, gshape[1], gshape[2])    # this rank's slab
    N = nzl * gshape[1] * gshape[2]
    tile = synth.gyroid_tile(128)
    dvol = eng.tiled_volume(tile, sshape, z_offset=z0)
    comm = None
    if world > 1:
        # torch.distributed is plumbing only: it carries the 128-byte NCCL id of the library's own communicator
        from nii2mesh_b200 import slabs
        comm = slabs.nccl_comm_from_torch(eng, dist, rank, world)

    parity = None
    if world > 1 and not args.no_verify and args.verify_first:
        parity = verify_slabs(eng, comm, dist, rank, world, tile)

    def step():
        if world == 1:
            return eng.meshify_device(dvol, ISO, fetch=False, **FLAGS)[2]
        sr = eng.meshify_slab(comm, dvol, gshape, z0, ISO, **FLAGS)
        step.last = sr
        return sr.r
    step.last = None

    # ---- device-resident throughput -------------------------------------------------------------
    for _ in range(args.warmup):
        r = step()
    eng.set_profile(True)
    ktot, kcnt = {}, {}
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    eng.timer_start()
    launches = 0
    stage = np.zeros(8)
    per_step_k = []
    for _ in range(args.steps):
        r = step()
        launches += r.launches
        stage += np.array(list(r.ms))
    ms_total = eng.timer_stop()
    # per-kernel CUDA-event times of the LAST timed step (the library keeps one call's event pairs; reading 77 of them
    # through ctypes after every step cost 0.4 ms of idle GPU per step inside the timed region)
    per_step_k.append(eng.kernel_times())
    barrier()
    clocks = sampler.stop()
    eng.set_profile(False)
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = GN / (ms_step * 1e-3) / 1e9
    for ks in per_step_k:
        for name, ms in ks:
            ktot[name] = ktot.get(name, 0.0) + ms
            kcnt[name] = kcnt.get(name, 0) + 1
    if world > 1 and not args.no_verify and not args.verify_first:
        parity = verify_slabs(eng, comm, dist, rank, world, tile)
    nv, nt = r.nverts, r.ntris                      # global counts
    if world == 1:
        lnv, lnt = nv, nt
    else:
        sr = step.last
        lnv, lnt = sr.nv_edge + sr.nv_cent + sr.nv_extra, sr.ntris_local
    nwords = nzl * gshape[1] * ((gshape[2] + 31) // 32)
    peak, peak_src = peaks()
    top = max(ktot, key=ktot.get)
    top_ms = ktot[top] / kcnt[top]
    kb = kernel_bytes(top, N, lnv, lnt, nwords)
    achieved = kb / (top_ms * 1e-3) / 1e9 if kb else None
    # whole-step algorithmic traffic (60 B/voxel for -p1 -l1 -b1, plus the surface term)
    step_bytes = 60 * GN + 24 * nv + 12 * nt
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": ncu_traffic(top) if (world == 1 and n == 1024) else None,
                "traffic_source": "profiles/r2_ncu_full_summary.csv (ncu --set full of the same kernels, G1024, committed with this code)",
                "peak_source": peak_src,
                "kernel_ms": top_ms, "kernel_share_of_step": ktot[top] / ms_step,
                "algorithmic_bytes_per_launch": kb, "scope": "rank 0's GPU" if world > 1 else "the GPU",
                "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                               "frac": step_bytes / (ms_step * 1e-3) / 1e9 / (peak * world)},
                "kernels_ms_per_step": {k: round(v, 4) for k, v in sorted(ktot.items(), key=lambda kv: -kv[1])},
                "kernels_sampled": "CUDA-event pairs of the last timed step"}

    # ---- end to end through the reference-facing entry point with host buffers ---------------------
    e2e = None
    if not args.no_e2e:
        L = eng.lib
        hp = C.c_void_p()
        eng._chk(L.b2m_host_alloc(C.byref(hp), N * 4))
        hvol = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=sshape)
        rz, ry, rx = nzl // 128, gshape[1] // 128, gshape[2] // 128
        if (rz * 128, ry * 128, rx * 128) == sshape and z0 % 128 == 0:
            hvol.reshape(rz, 128, ry, 128, rx, 128)[...] = tile[None, :, None, :, None, :]
        else:
            hvol[...] = dvol.to_host()
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        os.environ["B2M_DEVICE"] = str(local)
        dvol.free()
        if world == 1:
            L.meshify.argtypes = [C.c_void_p, C.POINTER(C.c_short), C.c_int, C.c_float, C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_bool, C.c_bool,
                                  C.c_bool, C.c_bool]
            dim = (C.c_short * 3)(n, n, n)

            def one():
                pt, pp, cnt, cnv = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
                rc = L.meshify(hp, dim, 0, ISO, C.byref(pt), C.byref(pp), C.byref(cnt), C.byref(cnv), True, True, True, False)
                assert rc == 0 and (cnv.value, cnt.value) == (nv, nt)
                libc.free(pp)
                libc.free(pt)
                return nv * 24 + nt * 12
            api = "meshify() (include/meshify.h), pinned host volume in, malloc'd host mesh out"
        else:
            o2 = lib.Opts(ISO, 0, 1, 1, 1, 0, 0)
            gd = (C.c_int64 * 3)(gshape[2], gshape[1], gshape[0])

            def one():
                pv, pt = C.c_void_p(), C.c_void_p()
                sr2 = lib.SlabResult()
                eng._chk(L.b2m_meshify_slab_host(eng.ctx, comm, hp, gd, z0, nzl, C.byref(o2), C.byref(pv), C.byref(pt),
                                                 C.byref(sr2)))
                assert (sr2.r.nverts, sr2.r.ntris) == (nv, nt)
                libc.free(pv)
                libc.free(pt)
                one.last = (sr2.r.h2d_ms, sr2.r.ms[7], sr2.r.d2h_ms)
                return int(sr2.r.d2h_bytes)  # bytes that crossed PCIe (Lewiner vertices travel as f32 and are widened on the host)
            api = ("b2m_meshify_slab_host() (include/b2m.h), one z-slab per rank: pinned host planes in, malloc'd host "
                   "mesh blocks out")
        ke = max(1, min(args.steps, 5))
        for _ in range(2):
            one()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            d2h = one()
        barrier()
        dt = max_over_ranks((time.perf_counter() - t0) / ke)
        breakdown = None
        if world > 1:
            h, dv, dd = one.last
            tt = torch.tensor([h, dv, dd], dtype=torch.float64, device="cuda")
            allr = [torch.zeros_like(tt) for _ in range(world)]
            dist.all_gather(allr, tt)
            per_rank = [[round(float(x), 1) for x in a.tolist()] for a in allr]
            breakdown = {"h2d_ms": round(max_over_ranks(h), 2), "device_ms": round(max_over_ranks(dv), 2),
                         "d2h_ms": round(max_over_ranks(dd), 2), "note": "max over ranks, last timed call; device_ms includes "
                         "waiting in the collectives for ranks whose H2D finished later",
                         "per_rank_h2d_device_d2h_ms": per_rank}
        if world == 1:
            # two untimed calls through b2m_meshify_host (what meshify() wraps) for the copy/compute breakdown
            r2 = lib.Result()
            pv, pt = C.c_void_p(), C.c_void_p()
            o2 = lib.Opts(ISO, 0, 1, 1, 1, 0, 0)
            for _ in range(2):  # the second call is the steady state (output blocks pre-faulted from the previous totals)
                eng._chk(L.b2m_meshify_host(eng.ctx, hp, (C.c_int64 * 3)(n, n, n), C.byref(o2), C.byref(pv), C.byref(pt), C.byref(r2)))
                libc.free(pv)
                libc.free(pt)
            breakdown = {"h2d_ms": round(r2.h2d_ms, 2), "device_ms": round(r2.ms[7], 2), "d2h_ms": round(r2.d2h_ms, 2)}
            d2h = int(r2.d2h_bytes)  # bytes that crossed PCIe: Lewiner vertices travel as f32 (exact) and are widened on the host
        pageable = None
        if world == 1:
            # the drop-in caller hands meshify() a malloc()'d volume (src/nii2mesh.c:141): the same call on pageable memory
            # (the library stages it through pinned buffers with its copy pool)
            pg = np.empty(sshape, np.float32)
            pg[...] = hvol
            ppg = C.c_void_p(pg.ctypes.data)

            def one_pg():
                pt, pp, cnt, cnv = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
                rc = L.meshify(ppg, dim, 0, ISO, C.byref(pt), C.byref(pp), C.byref(cnt), C.byref(cnv), True, True, True, False)
                assert rc == 0 and (cnv.value, cnt.value) == (nv, nt)
                libc.free(pp)
                libc.free(pt)
            one_pg()
            t0 = time.perf_counter()
            for _ in range(3):
                one_pg()
            dtp = (time.perf_counter() - t0) / 3
            pageable = {"value": GN / dtp / 1e9, "ms_per_step": dtp * 1e3, "steps": 3,
                        "note": "same meshify() call, volume in pageable (numpy / malloc) memory"}
            del pg
        L.b2m_get_copy_threads.restype = C.c_int
        e2e = {"value": GN / dt / 1e9, "breakdown": breakdown, "pageable_input": pageable, "copy_threads": int(L.b2m_get_copy_threads()),
               "unit": "Gvoxels/s", "h2d_bytes_per_step": N * 4,
               "d2h_bytes_per_step": d2h, "d2h_delivered_bytes_per_step": (nv * 24 + nt * 12) if world == 1 else None, "bytes_scope": "per rank" if world > 1 else "whole job",
               "ms_per_step": dt * 1e3, "steps": ke, "api": api}
        L.b2m_host_free(hp)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sn, T = 256, host_threads()
        dt, kind, cnv, cnt = ref_meshify_time(sn, 1, 0, T)
        cpu = {"value": T * sn ** 3 / dt / 1e9, "unit": "Gvoxels/s", "cores": T, "per_core_mvox_s": sn ** 3 / dt / 1e6, "kind": kind, "cpu": cpu_model(),
               "sample": f"{T} x G{sn} ({sn}^3 voxels, same generator and flags) concurrently, one per host core, "
                         f"{dt:.1f} s, {cnv} verts {cnt} tris each; meshify() itself is single-threaded"}
    # known answer of the G family: the reference's PRE-weld counts are exactly cubic in the number
    # of 128-voxel tiles per axis; they pin the full-size runs no CPU oracle can reach (2048^3: 336 902 112 / 673 869 568)
    known = None
    exp = known_counts(gshape)
    if exp:
        known = {"pre_nverts": exp[0], "pre_ntris": exp[1], "match": (r.pre_nverts, r.pre_ntris) == exp,
                 "source": "reference's pre-weld count law of the G family (exact fit on reference runs, tests/golden/golden_big.json)"}
        assert known["match"], f"pre-weld counts {r.pre_nverts}/{r.pre_ntris} differ from the reference's {exp}"
    if rank == 0:
        workload = workload_string(world, n, gshape, strong)
        if world == 1:
            par = "single GPU"
        else:
            par = f"{world} z-slabs of {nzl} planes, one rank per GPU; NCCL halo/seam exchange (b2m_meshify_slab)"
        out = {"metric": METRIC, "value": value, "unit": "Gvoxels/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": workload, "voxels_per_gpu": N, "voxels": GN, "parallelism": par, "parity": parity,
                          "l2": "inputs (4 B/voxel volume) larger than the 126 MB L2; no flush needed",
                          "mesh": {"nverts": nv, "ntris": nt, "pre_nverts": r.pre_nverts, "pre_ntris": r.pre_ntris},
                          "known_answer": known},
               "stage_ms": {k: round(float(v) / args.steps, 4) for k, v in zip(lib.STAGES, stage)},
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        emit(out)
    if comm is not None:
        eng.lib.b2m_comm_destroy(comm)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
