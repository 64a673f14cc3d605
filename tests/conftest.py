import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the PyRayT reference tree (build container only)")


def load_case(name):
    """(FlatScene, rays (13,N), reference frame (15,rows), generation_limit) of a committed golden case."""
    from pyrayt_b200.scene import FlatScene

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    scene = FlatScene.from_json(open(os.path.join(GOLDEN_DIR, name + ".scene.json")).read())
    return scene, z["rays"], z["frame"], int(z["generation_limit"])


def assert_frames_match(got, want, rtol=1e-9, what=""):
    """Parity bar of BASELINE.json: generation / id / surface bit-exact, positions and
    directions within `rtol` relative (absolute floor 1e-12 for values near zero)."""
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    for col in (0, 4, 5):
        assert np.array_equal(got[col], want[col]), f"{what}: integer column {col} differs"
    np.testing.assert_allclose(got, want, rtol=rtol, atol=1e-12, equal_nan=True, err_msg=what)


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return 0
