"""Small end-to-end exercise of every kernel, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python scripts/sanitize_run.py
    compute-sanitizer --tool racecheck python scripts/sanitize_run.py
    compute-sanitizer --tool synccheck python scripts/sanitize_run.py

Golden cases (all record modes, staging overflow + retry, the wavefront driver, the captured small-trace
sequence, the lean host transfer, the diagnose variant, the FP32 fast mode), a 973-leaf scene read from global
memory, sources, component.intersect, nearest / render hits, the ordering kernels
and the frame read-outs; every result is compared with the oracle.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import pyrayt_b200  # noqa: E402
from oracle import oracle, sources_np  # noqa: E402
from pyrayt_b200 import analytics, workloads  # noqa: E402
from tests.helpers import GOLDEN_CASES, RENDER_CASES, load_case, load_render_case  # noqa: E402

checked = 0
for name in GOLDEN_CASES:
    scene, rays, _, gl = load_case(name)
    rays = np.ascontiguousarray(rays[:, :700])
    eng = pyrayt_b200.Engine(scene, 0)
    d = torch.from_numpy(rays).cuda()
    want, ctr = oracle.trace(scene, rays, gl)
    res = eng.trace(d, generation_limit=gl)
    assert np.array_equal(res.frame.cpu().numpy(), want, equal_nan=True), name
    res = eng.trace(d, generation_limit=gl, capacity=64, to_host=True)  # overflow, then the exact retry
    assert np.array_equal(res.frame.numpy(), want, equal_nan=True), name
    wave = eng.trace_wavefront(d, generation_limit=gl, capacity=50)  # overflow, then the exact retry
    assert np.array_equal(wave.frame.cpu().numpy(), want, equal_nan=True), name
    lean = eng.trace(d, generation_limit=gl, to_host=True, lean=True)  # pack kernel + host rebuild
    assert np.array_equal(lean.frame.numpy(), want, equal_nan=True), name
    if eng.small_fits(rays.shape[1], min(gl, 16)):
        sub, _ = oracle.trace(scene, rays, min(gl, 16))
        eng.small_ray_buffer(rays.shape[1]).copy_(d)
        for _ in range(3):  # eager, captured, replayed
            small = eng.trace_small(rays.shape[1], generation_limit=min(gl, 16))
            assert np.array_equal(small.frame.numpy(), sub, equal_nan=True), name
    sid = int(scene.leaf_sid[-1])
    res = eng.trace(d, generation_limit=gl, record="surface", detector_sid=sid)
    assert np.array_equal(res.frame.cpu().numpy(), want[:, want[5] == sid], equal_nan=True), name
    assert eng.trace(d, generation_limit=gl, record="none").counters["generations"] == ctr["generations"]
    diag = eng.trace(d, generation_limit=gl, diagnose=True)  # PRT_FLAG_DIAGNOSE variant
    odiag = oracle.diagnose(scene, rays, gl)
    assert (diag.counters["grazing_rays"], diag.counters["seam_rays"]) == (odiag["grazing_rays"], odiag["seam_rays"]), name
    f32 = eng.trace(d, generation_limit=gl, precision="fp32", to_host=True)  # the FP32 fast mode: trace + ordering + host transfer
    assert abs(f32.rows - want.shape[1]) <= max(2, want.shape[1] // 100), name
    eng.trace(d, generation_limit=gl, precision="fp32", record="none")
    full = eng.trace(d, generation_limit=gl)
    if full.rows:
        analytics.spot_stats(full, max(1, rays.shape[1] // 3), 3, surface=sid)
        analytics.focus_table(full, generation=float(want[0].max()))
    r8 = np.zeros((8, rays.shape[1]))
    r8[0:3], r8[3], r8[4:7] = rays[0:3], 1, rays[4:7]
    for renderer in (False, True):
        t, s, nrm = eng.nearest_hit(torch.from_numpy(r8).cuda(), normals=True, renderer=renderer)
        ot, osid, onrm = (oracle.render_hit if renderer else oracle.nearest)(scene, r8)
        assert np.array_equal(t.cpu().numpy(), ot) and np.array_equal(s.cpu().numpy(), osid), name
    for c in range(scene.n_components):
        hits, sids = eng.intersect(c, torch.from_numpy(r8.reshape(2, 4, -1)).cuda())
        oh, os_ = oracle.intersect(scene, c, r8.reshape(2, 4, -1))
        hits, sids = hits.cpu().numpy(), sids.cpu().numpy()
        assert np.array_equal(hits, oh, equal_nan=True) and np.array_equal(sids, os_), name
    eng.close()
    checked += 1
# a scene too large for shared memory: the trace kernel reads it in place (GLOBAL variant)
from tests import scene_util as su  # noqa: E402

scene, centres = su.lenslet_array(18, 18)
rays = su.lenslet_rays(centres, 2)
eng = pyrayt_b200.Engine(scene, 0)
res = eng.trace(torch.from_numpy(rays).cuda(), generation_limit=8)
want, _ = oracle.trace(scene, rays, 8, threads=4)
assert np.array_equal(res.frame.cpu().numpy(), want, equal_nan=True), "lenslet array"
eng.trace(torch.from_numpy(rays[:, :256].copy()).cuda(), generation_limit=8, diagnose=True, record="none")
eng.close()
checked += 1
for name in RENDER_CASES:
    scene, rays, dist, surf, _, _ = load_render_case(name)
    eng = pyrayt_b200.Engine(scene, 0)
    t, s, _ = eng.nearest_hit(torch.from_numpy(np.ascontiguousarray(rays[..., :1500])).cuda(), renderer=True)
    assert np.array_equal(s.cpu().numpy(), surf[:1500]), name
    checked += 1
for wl in workloads.WORKLOADS.values():
    if hasattr(wl.source, "templates"):  # the reference's own Source classes: whole sources, sin / cos within 2 ulp
        k = 90 * len(wl.source.templates)
        got = wl.source.generate(k, device=0).cpu().numpy()
        np.testing.assert_allclose(got, sources_np.from_source(wl.source, k), rtol=1e-12, atol=1e-15, err_msg=wl.name)
    else:
        got = wl.source.generate(1000, device=0, first_index=77).cpu().numpy()
        assert np.array_equal(got, sources_np.from_source(wl.source, 1000, first_index=77)), wl.name
    checked += 1
torch.cuda.synchronize()
print(f"sanitize_run: {checked} groups checked against the oracle")
