"""Lean device -> host transfer of a frame: host-side column rebuild (CPU) and the whole path (GPU)."""
import ctypes

import numpy as np
import pytest

from pyrayt_b200 import _lib
from tests.helpers import GOLDEN_CASES, load_case


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_host_expand_rebuilds_the_five_columns(name):
    """prt_host_expand_frame is host code: checked here against the reference's golden frames."""
    lib = _lib.load()
    scene, rays, frame, gl = load_case(name)
    rows = frame.shape[1]
    rays = np.ascontiguousarray(rays)
    idx = (frame[4] - rays[12, 0]).astype(np.int64)
    assert np.array_equal(rays[12, idx], frame[4])
    packed = (idx.astype(np.uint64) << np.uint64(24)) | (frame[5].astype(np.int64) + 1).astype(np.uint64)
    goff = np.zeros(gl + 1, dtype=np.int64)
    goff[1:] = np.cumsum(np.bincount(frame[0].astype(int), minlength=gl)[:gl])
    for threads in (1, 3):
        out = np.ascontiguousarray(frame.copy())
        out[[0, 1, 2, 4, 5]] = -777.0
        rc = lib.prt_host_expand_frame(_p(packed), rows, _p(goff), gl, _p(rays[8]), _p(rays[9]), _p(rays[10]),
                                       _p(rays[12]), _p(out), rows, threads)
        assert rc == 0
        assert np.array_equal(out, frame, equal_nan=True)
    bad = goff.copy()
    bad[-1] += 1
    out = np.ascontiguousarray(frame.copy())
    assert lib.prt_host_expand_frame(_p(packed), rows, _p(bad), gl, _p(rays[8]), _p(rays[9]), _p(rays[10]),
                                     _p(rays[12]), _p(out), rows, 1) != 0 or rows == 0


@pytest.mark.gpu
def test_lean_transfer_equals_full_copy(cuda_device):
    import torch

    import pyrayt_b200
    from oracle import oracle

    for name in ("config4_stack", "config5_cavity", "thick_lens_zoo"):
        scene, rays, _, gl = load_case(name)
        eng = pyrayt_b200.Engine(scene, device=0)
        d = torch.from_numpy(np.ascontiguousarray(rays)).cuda()
        want, _ = oracle.trace(scene, rays, gl)
        for method in ("single", "wavefront"):
            full = eng.trace(d, generation_limit=gl, to_host=True, lean=False, method=method).frame.numpy().copy()
            lean = eng.trace(d, generation_limit=gl, to_host=True, lean=True, method=method).frame.numpy().copy()
            assert np.array_equal(full, want, equal_nan=True) and np.array_equal(lean, want, equal_nan=True), name
            h_rays = torch.from_numpy(np.ascontiguousarray(rays)).pin_memory()
            lean2 = eng.trace(d, generation_limit=gl, to_host=True, lean=True, host_rays=h_rays, method=method)
            assert np.array_equal(lean2.frame.numpy(), want, equal_nan=True), name
    # rows the host could not rebuild: ids that are not consecutive, fractional intensity is fine, huge surface ids
    scene, rays, _, gl = load_case("config4_stack")
    eng = pyrayt_b200.Engine(scene, device=0)
    odd = np.ascontiguousarray(rays).copy()
    odd[12] = odd[12] * 3.0 + 0.5  # neither consecutive nor integral
    odd[9] = np.linspace(1.0, 2.0, odd.shape[1])
    want, _ = oracle.trace(scene, odd, gl)
    got = eng.trace(torch.from_numpy(odd).cuda(), generation_limit=gl, to_host=True, lean=True)
    assert np.array_equal(got.frame.numpy(), want, equal_nan=True)  # fell back to copying every column
    shifted = np.ascontiguousarray(rays).copy()
    shifted[12] += 1000.0  # consecutive ids with an offset (a rank's share of a sharded trace)
    shifted[8] = 5.0       # non-zero input generation: kept in the generation-0 rows
    want, _ = oracle.trace(scene, shifted, gl)
    got = eng.trace(torch.from_numpy(shifted).cuda(), generation_limit=gl, to_host=True, lean=True)
    assert np.array_equal(got.frame.numpy(), want, equal_nan=True)


@pytest.mark.gpu
def test_lean_transfer_verifies_the_generation_column(cuda_device):
    """A device frame whose generation column is not what the host would rebuild from the row offsets
    (caller-edited here) must be copied in full, not silently 'corrected'."""
    import torch

    import pyrayt_b200

    scene, rays, _, gl = load_case("config4_stack")
    eng = pyrayt_b200.Engine(scene, device=0)
    d = torch.from_numpy(np.ascontiguousarray(rays)).cuda()
    res = eng.trace(d, generation_limit=gl)
    goff = np.concatenate(([0], np.cumsum(res.gen_counts)))
    ok = eng._frame_to_host(res.frame, res.rows, goff, d, lean=True).numpy().copy()
    assert eng.last_transfer == "lean" and np.array_equal(ok, res.frame.cpu().numpy(), equal_nan=True)
    edited = res.frame.clone()
    edited[0, res.rows // 2] = 77.0
    got = eng._frame_to_host(edited, res.rows, goff, d, lean=True).numpy()
    assert eng.last_transfer == "full"
    assert np.array_equal(got, edited.cpu().numpy(), equal_nan=True)
