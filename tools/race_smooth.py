import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from nii2mesh_b200 import lib, synth
eng = lib.Engine(0)
for shape in ((40, 50, 140), (17, 20, 33), (64, 64, 64)):
    v = synth.random_blobs(shape, seed=3)
    a = eng.smooth(v)
    print(shape, float(a.mean()))
print("DONE")
