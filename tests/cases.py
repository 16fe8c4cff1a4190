"""Named deterministic parity inputs shared by tools/make_golden.py and the tests."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
from nii2mesh_b200 import synth  # noqa: E402

LEWINER, CLASSIC = 0, 1


def volumes(big=True):
    """name -> (volume[z,y,x] f32, isolevel).  All finish in seconds on the CPU oracle."""
    out = {
        "sphere24": (synth.noisy_sphere(24), 0.0),
        "sphere40": (synth.noisy_sphere(40), 0.0),
        "blobs": (synth.random_blobs((30, 37, 41), seed=3), 0.2),          # odd dims, many clusters + bubbles
        "blobs_w33": (synth.random_blobs((17, 20, 33), seed=7), 0.1),       # nx one past a bit word
        "blobs_w31": (synth.random_blobs((19, 23, 31), seed=8), 0.1),       # nx one short of a bit word
        "blobs2": (synth.random_blobs((33, 64, 70), seed=5, smooth=1), 0.1),
        "thin4": (synth.random_blobs((9, 11, 4), seed=9, smooth=1), 0.0),   # a dim < 5: the smooth is skipped
        "gyroid96": (synth.gyroid(96, P=32), 0.0),
        "isoreset": (synth.noisy_sphere(24), 1.0e6),                        # isolevel out of range -> reset to mid
    }
    if big:
        out["sphere64"] = (synth.noisy_sphere(64), 0.0)
        out["gyroid160"] = (synth.gyroid(160, P=64), 0.0)
        bet = ROOT / "tests" / "golden" / "bet.nii.gz"
        out["bet"] = (synth.load_nifti(bet)[0], 67.729)                      # Otsu "medium" isolevel (BASELINE config 1)
    return out


def flag_sets(name=""):
    """(backend, originalMC, preSmooth, onlyLargest, fillBubbles)"""
    fs = [(LEWINER, 0, 0, 0, 0), (LEWINER, 0, 1, 1, 0), (LEWINER, 0, 1, 1, 1), (LEWINER, 0, 1, 0, 1),
          (LEWINER, 0, 0, 1, 1), (LEWINER, 1, 0, 0, 0), (LEWINER, 1, 1, 1, 0),
          (CLASSIC, 0, 0, 0, 0), (CLASSIC, 0, 1, 1, 0), (CLASSIC, 0, 1, 1, 1)]
    return fs


def flat_volume():
    return np.full((8, 9, 10), 3.0, np.float32)
