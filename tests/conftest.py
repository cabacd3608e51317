import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
# golden files that hold a point cloud and the reference's outputs for it
POINT_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                     if f.endswith(".npz") and not f.startswith(("forward_", "retrieval", "quantization", "intensity", "ctor_", "interp_", "keyframe", "widths")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
