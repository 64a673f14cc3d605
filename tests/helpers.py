"""Shared test helpers: golden-case loader and the parity assertion."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


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
