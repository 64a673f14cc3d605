"""Shared test helpers: golden-case loader and the parity assertion."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
_ALL_NPZ = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
GOLDEN_CASES = [c for c in _ALL_NPZ if not c.startswith("render_")]   # trace frames (make_golden.py)
RENDER_CASES = [c for c in _ALL_NPZ if c.startswith("render_")]       # renderer hit images (make_render_golden.py)


def load_render_case(name):
    """(FlatScene, camera rays (2,4,N), reference distance (N,), surface (N,), canvas (v,h,4), (h, v))."""
    from pyrayt_b200.scene import FlatScene

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    scene = FlatScene.from_json(open(os.path.join(GOLDEN_DIR, name + ".scene.json")).read())
    return scene, z["rays"], z["distance"], z["surface"], z["canvas"], tuple(int(x) for x in z["resolution"])


def load_case(name):
    """(FlatScene, rays (13,N), reference frame (15,rows), generation_limit) of a committed golden case."""
    from pyrayt_b200.scene import FlatScene

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    scene = FlatScene.from_json(open(os.path.join(GOLDEN_DIR, name + ".scene.json")).read())
    return scene, z["rays"], z["frame"], int(z["generation_limit"])


def assert_frames_match(got, want, rtol=1e-9, what=""):
    """Parity bar of BASELINE.json: generation / id / surface bit-exact; hit positions within
    `rtol` relative to the size of the position vector (floor 1, i.e. scene units), directions
    (unit vectors) within `rtol` absolute, the remaining float columns within `rtol` relative."""
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    for col in (0, 4, 5):
        assert np.array_equal(got[col], want[col]), f"{what}: integer column {col} differs"
    np.testing.assert_allclose(got[1:4], want[1:4], rtol=rtol, atol=0, equal_nan=True, err_msg=what)
    for lo in (6, 9):  # start point x0,y0,z0 and hit point x1,y1,z1
        scale = np.maximum(1.0, np.linalg.norm(want[lo:lo + 3], axis=0))
        err = np.abs(got[lo:lo + 3] - want[lo:lo + 3]) / scale
        assert np.array_equal(np.isnan(got[lo:lo + 3]), np.isnan(want[lo:lo + 3])), what
        assert np.nanmax(err, initial=0.0) <= rtol, f"{what}: position error {np.nanmax(err):.3e}"
    terr = np.abs(got[12:15] - want[12:15])
    assert np.array_equal(np.isnan(got[12:15]), np.isnan(want[12:15])), what
    assert np.nanmax(terr, initial=0.0) <= rtol, f"{what}: direction error {np.nanmax(terr):.3e}"
