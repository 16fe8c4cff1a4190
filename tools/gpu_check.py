#!/usr/bin/env python3
"""Development diagnostic (run on a GPU box): stage-by-stage parity of libb2m against the oracle,
printing as much as possible per run.  The pytest suite in tests/ is the judged version of this."""
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import lib, synth  # noqa: E402
from oracle import Oracle, Ref, ref_available  # noqa: E402
from oracle.canon import assert_same_mesh  # noqa: E402

QUICK = "--quick" in sys.argv
O = Oracle()
E = lib.Engine(0)
fails = []


def step(name, fn):
    t0 = time.time()
    try:
        msg = fn()
        print(f"[ok]   {name} ({time.time() - t0:.2f}s) {msg or ''}", flush=True)
    except Exception as e:  # noqa: BLE001
        fails.append(name)
        print(f"[FAIL] {name}: {type(e).__name__}: {e}", flush=True)
        tb = traceback.format_exc().splitlines()
        print("\n".join(tb[-6:]), flush=True)


def vols():
    out = {
        "sphere24": (synth.noisy_sphere(24), 0.0),
        "sphere40": (synth.noisy_sphere(40), 0.0),
        "blobs": (synth.random_blobs((30, 37, 41), seed=3), 0.2),
        "blobs2": (synth.random_blobs((33, 64, 70), seed=5, smooth=1), 0.1),
        "gyroid96": (synth.gyroid(96, P=32), 0.0),
    }
    if not QUICK:
        out["sphere64"] = (synth.noisy_sphere(64), 0.0)
        out["gyroid160"] = (synth.gyroid(160, P=64), 0.0)
        bet = ROOT / "tests" / "golden" / "bet.nii.gz"
        if bet.exists():
            out["bet"] = (synth.load_nifti(bet)[0], 67.729)
    return out


V = vols()


def check_smooth():
    for name, (v, _) in V.items():
        a = E.smooth(v)
        b = O.smooth(v)
        nd = int((a.view(np.uint32) != b.view(np.uint32)).sum())
        assert nd == 0, f"{name}: {nd} voxels differ, max abs {np.abs(a - b).max()}"
    return f"{len(V)} volumes bit-exact"


def check_front():
    n = 0
    for name, (v, iso) in V.items():
        for ps, ol, fb in ((0, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1), (0, 1, 1)):
            a = E.front(v, iso, ps, ol, fb)
            b = O.front(v, iso, ps, ol, fb)
            tag = f"{name} p{ps} l{ol} b{fb}"
            assert a["lo"] == b["lo"] and a["hi"] == b["hi"], f"{tag}: bbox {a['lo']} {a['hi']} vs {b['lo']} {b['hi']}"
            assert a["iso"] == b["iso"] and a["mn"] == b["mn"] and a["mx"] == b["mx"], f"{tag}: range/iso"
            if ol or fb:
                dm = int(((a["mask"] != 0) != (b["mask"] != 0)).sum())
                assert dm == 0, f"{tag}: mask differs in {dm} voxels ({int((a['mask']!=0).sum())} vs {int((b['mask']!=0).sum())})"
            nd = int((a["img"].view(np.uint32) != b["img"].view(np.uint32)).sum())
            assert nd == 0, f"{tag}: composed volume differs in {nd} voxels"
            n += 1
    return f"{n} cases bit-exact"


def check_mc():
    n = 0
    for name, (v, iso) in V.items():
        f = O.front(v, iso, 1, 1, 0)
        for omc in (0, 1):
            ov, ot = O.mc(f["img"], f["lo"], f["hi"], f["iso"], omc, 0)
            gv, gt, r = E.mc(f["img"], f["lo"], f["hi"], f["iso"], omc, 0)
            tag = f"{name} lewiner o{omc}"
            assert gv.shape == ov.shape and gt.shape == ot.shape, f"{tag}: shapes {gv.shape} {gt.shape} vs {ov.shape} {ot.shape}"
            assert np.array_equal(gt, ot), f"{tag}: {int((gt != ot).any(axis=1).sum())} triangles differ"
            assert np.array_equal(gv, ov), f"{tag}: {int((gv != ov).any(axis=1).sum())} vertices differ"
            n += 1
        # raw (unsmoothed) volume: exercises the ambiguous MC33 cases
        f = O.front(v, iso, 0, 0, 0)
        ov, ot = O.mc(f["img"], f["lo"], f["hi"], f["iso"], 0, 0)
        gv, gt, r = E.mc(f["img"], f["lo"], f["hi"], f["iso"], 0, 0)
        assert gv.shape == ov.shape and gt.shape == ot.shape, f"{name} raw: shapes {gv.shape} {gt.shape} vs {ov.shape} {ot.shape}"
        assert np.array_equal(gt, ot), f"{name} raw: {int((gt != ot).any(axis=1).sum())} triangles differ"
        assert np.array_equal(gv, ov), f"{name} raw: {int((gv != ov).any(axis=1).sum())} vertices differ"
        n += 1
    return f"{n} cases identical arrays (order included)"


def check_weld_hook():
    n = 0
    for name, (v, iso) in V.items():
        if v.size > 300000:
            continue
        f = O.front(v, iso, 0, 0, 0)
        for backend in (0, 1):
            ov, ot = O.mc(f["img"], f["lo"], f["hi"], f["iso"], 0, backend)
            wv, wt = O.weld(ov, ot)
            wt = O.degenerate(wv, wt)
            gv, gt = E.weld(ov, ot)
            assert len(gv) == len(wv), f"{name} b{backend}: nverts {len(gv)} vs {len(wv)}"
            assert_same_mesh(gv, gt, wv, wt)
            n += 1
    return f"{n} meshes"


def check_full():
    n = 0
    for name, (v, iso) in V.items():
        for backend, omc in ((0, 0), (0, 1), (1, 0)):
            for ps, ol, fb in ((0, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1)):
                tag = f"{name} backend{backend} o{omc} p{ps} l{ol} b{fb}"
                try:
                    o = O.meshify(v, iso, omc, ps, ol, fb, backend)
                    gv, gt, r = E.meshify(v, iso, omc, ps, ol, fb, backend)
                    assert o["rc"] == 0
                    assert len(gv) == len(o["verts"]) and len(gt) == len(o["tris"]), \
                        f"counts {len(gv)}/{len(gt)} vs {len(o['verts'])}/{len(o['tris'])} (pre {r.pre_nverts}/{r.pre_ntris} vs {o['pre_nv']}/{o['pre_nt']})"
                    assert_same_mesh(gv, gt, o["verts"], o["tris"])
                    n += 1
                except Exception as e:  # noqa: BLE001
                    fails.append(tag)
                    print(f"   [FAIL] {tag}: {type(e).__name__}: {e}", flush=True)
    return f"{n} meshes identical after canonical sorting"


def check_ref_bet():
    if "bet" not in V or not ref_available():
        return "skipped"
    v, iso = V["bet"]
    R = Ref("lewiner")
    r = R.meshify(v, iso, 0, 1, 1, 0)
    gv, gt, res = E.meshify(v, iso, 0, 1, 1, 0, 0)
    assert_same_mesh(gv, gt, r["verts"], r["tris"])
    return f"bet vs compiled reference: {len(gv)} verts {len(gt)} tris, times {res.times()}"


def timing():
    out = []
    for n in ((128, 256) if QUICK else (128, 256, 512)):
        v = synth.gyroid(n)
        d = E.upload(v)
        for _ in range(2):
            _, _, r = E.meshify_device(d, 0.0, 0, 1, 1, 1, 0, fetch=False)
        t = r.times()
        out.append(f"\n      G{n}: {n**3 / t['total'] / 1e6:.2f} Gvox/s total {t['total']:.2f} ms " +
                   " ".join(f"{k}={x:.2f}" for k, x in t.items() if k != 'total') +
                   f" nv={r.nverts} nt={r.ntris} pre={r.pre_nverts}/{r.pre_ntris} launches={r.launches}")
        d.free()
    return "".join(out)


step("smooth", check_smooth)
step("front", check_front)
step("mc", check_mc)
step("weld-hook", check_weld_hook)
step("full", check_full)
step("bet-vs-ref", check_ref_bet)
step("timing", timing)
print("FAILS:", fails)
sys.exit(1 if fails else 0)
