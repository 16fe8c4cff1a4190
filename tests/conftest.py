"""pytest configuration: the `gpu` marker, shared fixtures (oracle, compiled reference, engine)."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_lewiner():
    from oracle import Ref, ref_available, build
    build()
    if not ref_available("lewiner"):
        pytest.skip("compiled reference (oracle/_ref) not present")
    return Ref("lewiner")


@pytest.fixture(scope="session")
def ref_classic():
    from oracle import Ref, ref_available, build
    build()
    if not ref_available("classic"):
        pytest.skip("compiled reference (oracle/_ref) not present")
    return Ref("classic")


@pytest.fixture(scope="session")
def libb2m():
    """the built C-ABI library (python -m nii2mesh_b200.build); building is part of the CPU check"""
    from nii2mesh_b200 import build as b
    b.build()
    from nii2mesh_b200 import lib
    return lib.load()


@pytest.fixture(scope="session")
def eng(libb2m):
    """a libb2m engine on cuda:0; GPU tests FAIL (not skip) if the device or library is missing"""
    from nii2mesh_b200 import lib
    return lib.Engine(0)


@pytest.fixture(scope="session")
def bet():
    from nii2mesh_b200 import synth
    vol, hdr = synth.load_nifti(GOLDEN / "bet.nii.gz")
    return vol, hdr


def bits_differ(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return int((a != b).sum())
