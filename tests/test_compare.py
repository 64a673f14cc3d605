"""pyrayt_b200.compare.frame_agreement: the statistic that states the FP32 fast mode's contract."""
import numpy as np
import torch

from oracle import oracle
from pyrayt_b200 import compare
from tests.helpers import load_case


def _frames():
    scene, rays, _, gl = load_case("config4_stack")
    rays = np.ascontiguousarray(rays[:, :256])
    frame, _ = oracle.trace(scene, rays, gl)
    return torch.from_numpy(frame), rays.shape[1]


def test_identical_frames_agree_completely():
    f, n = _frames()
    rep = compare.frame_agreement(f, f.clone(), 0, n)
    assert rep["rays_with_different_ids"] == 0 and rep["rays_beyond_tolerance"] == 0
    assert rep["rays_agreeing"] == n and rep["rows_compared"] == f.shape[1]
    assert rep["id_columns_equal_on_compared_rows"] and rep["max_error_on_agreeing_rays"] == 0.0
    assert rep["rows_by_error"]["<=1e-7"] == rep["rows_by_error"]["all"] == f.shape[1]


def test_small_and_large_deviations_are_told_apart():
    f, n = _frames()
    g = f.clone()
    scale = float(f[6:12].abs().max())
    g[9] += 2e-6 * scale  # every hit point off by 2e-6 of the scene scale: inside the 1e-5 tolerance
    rep = compare.frame_agreement(f, g, 0, n)
    assert rep["rays_beyond_tolerance"] == 0 and 1e-6 < rep["max_error_on_agreeing_rays"] < 1e-5
    assert rep["rows_by_error"]["<=1e-6"] == 0 and rep["rows_by_error"]["<=1e-5"] == f.shape[1]
    ray = int(f[4, 5])
    row = int(torch.nonzero(f[4] == ray)[2])
    g[13, row] += 1e-3  # one tilt of one ray far off
    rep = compare.frame_agreement(f, g, 0, n)
    assert rep["rays_beyond_tolerance"] == 1 and rep["rays_agreeing"] == n - 1
    assert rep["max_error_on_agreeing_rays"] < 1e-5  # the offending ray is not among the agreeing ones


def test_another_surface_or_a_missing_row_is_a_different_path():
    f, n = _frames()
    g = f.clone()
    row = 40
    g[5, row] += 1  # another surface id in one row
    rep = compare.frame_agreement(f, g, 0, n)
    assert rep["rays_with_different_ids"] == 1
    last = int(torch.nonzero(f[4] == f[4, 7])[-1])  # drop the last row of one ray
    keep = torch.ones(f.shape[1], dtype=torch.bool)
    keep[last] = False
    rep = compare.frame_agreement(f, f[:, keep], 0, n)
    assert rep["rays_with_different_ids"] == 1 and rep["rows_b"] == f.shape[1] - 1
    assert rep["rows_compared"] == f.shape[1] - int((f[4] == f[4, 7]).sum())
