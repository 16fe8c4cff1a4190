import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from nii2mesh_b200 import lib
from oracle import Oracle
import cases
E = lib.Engine(0); O = Oracle()
for name, (vol, iso) in cases.volumes(big=False).items():
    a = E.smooth(vol); b = O.smooth(vol)
    print(name, int((a.view(np.uint32) != b.view(np.uint32)).sum()), flush=True)
# unsafe inputs: denormals, -0, huge, inf
rng = np.random.default_rng(0)
v = rng.standard_normal((20, 33, 40)).astype(np.float32)
v[3:6, 4:9, 5:30] = 1e-42; v[7, 7, 7] = -0.0; v[10:12, 10:20, 3:9] = 3e38; v[15, 5, 5] = 1e-39; v[2, 2, 2:12] = 0.0
a = E.smooth(v); b = O.smooth(v)
print("unsafe mix", int((a.view(np.uint32) != b.view(np.uint32)).sum()), flush=True)
