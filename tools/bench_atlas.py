#!/usr/bin/env python3
"""BASELINE configs[3]: data/D99_atlas_v2.0_right.nii.gz, one mesh per label (365 non-empty labels), -p 1 -b 0, Lewiner.
Times b2m_atlas_scan + b2m_meshify_label_device over all labels (volume resident on the device, meshes left on the
device), with 1..T host threads (one b2m_ctx / stream each: per-label work is launch-latency bound).  The reference's
cost for the same job is one whole-volume meshify() per label (measured on a sample of labels with oracle/_ref)."""
import json
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import lib, synth  # noqa: E402


def main():
    vol, _ = synth.load_nifti(ROOT / "tests" / "golden" / "D99_atlas_v2.0_right.nii.gz")
    out = {"volume": list(vol.shape), "voxels": int(vol.size)}
    for T in (1, 2, 4, 8):
        engs = [lib.Engine(0) for _ in range(T)]
        d = engs[0].upload(vol)
        for rep in range(2):  # first repetition warms the workspaces
            t0 = time.perf_counter()
            infos = [i for i in engs[0].atlas_scan(d)[1:] if i.nvox > 0]
            t1 = time.perf_counter()
            tot = [0, 0]
            lock = threading.Lock()

            def work(k):
                nv = nt = 0
                for info in infos[k::T]:
                    _, _, r = engs[k].meshify_label(d, info, 0.5, 0, 1, 0, 0, fetch=False)
                    nv += r.nverts
                    nt += r.ntris
                with lock:
                    tot[0] += nv
                    tot[1] += nt
            th = [threading.Thread(target=work, args=(k,)) for k in range(T)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            t2 = time.perf_counter()
        out[f"threads_{T}"] = {"labels": len(infos), "scan_ms": round((t1 - t0) * 1e3, 2), "mesh_ms": round((t2 - t1) * 1e3, 2),
                               "total_verts": tot[0], "total_tris": tot[1]}
        d.free()
        for e in engs:
            e.close()
    try:
        import oracle
        if oracle.ref_available("lewiner"):
            R = oracle.Ref("lewiner")
            ts = []
            for lab in (1, 164, 301):
                b = ((vol > np.float32(lab - 0.5)) & (vol < np.float32(lab + 0.5))).astype(np.float32)
                t0 = time.perf_counter()
                R.meshify(b, 0.5, 0, 1, 0, 0)
                ts.append(time.perf_counter() - t0)
            out["reference_cpu"] = {"s_per_label": round(float(np.mean(ts)), 3), "labels_sampled": 3,
                                    "estimate_s_all_labels_1core": round(float(np.mean(ts)) * 365, 1)}
    except Exception as ex:  # noqa: BLE001
        out["reference_cpu"] = {"error": str(ex)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
