import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from nii2mesh_b200 import lib
from oracle import Oracle
import cases
E = lib.Engine(0); O = Oracle()
vol, iso = cases.volumes(big=False)["sphere40"]
for ps, ol in ((1, 1), (0, 0)):
    f = O.front(vol, iso, ps, ol, 0)
    print(ps, ol, f["lo"], f["hi"], vol.shape, flush=True)
    for omc in (0, 1):
        gv, gt, _ = E.mc(f["img"], f["lo"], f["hi"], f["iso"], omc, 0)
        print(" lew", omc, len(gv), len(gt), flush=True)
    gv, gt, _ = E.mc(f["img"], f["lo"], f["hi"], f["iso"], 0, 1)
    print(" classic", len(gv), len(gt), flush=True)
    cv, ct = E.weld(gv, gt)
    print(" weld", len(cv), len(ct), flush=True)
