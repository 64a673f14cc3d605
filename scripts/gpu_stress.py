"""Random-scene parity stress on a GPU: many seeded scenes (every primitive, operation, material, left-deep and
right-nested trees, tight / huge / too-small boxes, scaled poses) traced through the C ABI and compared bit for
bit with the oracle; the FP32 fast mode is compared within its tolerance on the scenes it supports, and the
diagnose counters with the oracle's.

    python scripts/gpu_stress.py [n_scenes] [rays_per_scene]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import pyrayt_b200  # noqa: E402
from oracle import oracle  # noqa: E402
from pyrayt_b200 import compare  # noqa: E402
from tests import scene_util as su  # noqa: E402

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
threads = os.cpu_count() or 1
t0 = time.time()
rows = generic = fp32_scenes = fp32_rays = fp32_off = diag_rays = 0
for seed in range(10_000, 10_000 + n_scenes):
    scene, rays = su.random_scene_and_rays(seed, n_rays=n_rays)
    eng = pyrayt_b200.Engine(scene, device=0)
    d = torch.from_numpy(rays).cuda()
    res = eng.trace(d, generation_limit=16)
    want, octr = oracle.trace(scene, rays, 16, threads=threads)
    got = res.frame.cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True), f"seed {seed}: FP64 frame differs from the oracle"
    assert res.counters["generations"] == octr["generations"], seed
    rows += got.shape[1]
    if seed % 10 == 0:  # the diagnose variant (five searches per generation) on a tenth of the scenes
        sub = np.ascontiguousarray(rays[:, :256])
        dg = eng.trace(torch.from_numpy(sub).cuda(), generation_limit=16, diagnose=True, record="none")
        od = oracle.diagnose(scene, sub, 16, threads=threads)
        assert (dg.counters["grazing_rays"], dg.counters["seam_rays"]) == (od["grazing_rays"], od["seam_rays"]), seed
        diag_rays += dg.counters["grazing_rays"] + dg.counters["seam_rays"]
    f32 = eng.trace(d, generation_limit=16, precision="fp32")
    rep = compare.frame_agreement(res.frame, f32.frame, 0, n_rays)
    assert rep["id_columns_equal_on_compared_rows"], seed
    generic += int(any(scene.comp_node_begin[c + 1] - scene.comp_node_begin[c] == 5 and
                       scene.node_kind[scene.comp_node_begin[c] + 3] != 0 for c in range(scene.n_components)))
    fp32_scenes += 1
    fp32_rays += n_rays
    fp32_off += rep["rays_with_different_ids"] + rep.get("rays_beyond_tolerance", 0)
    eng.close()
print(f"gpu_stress: {n_scenes} random scenes x {n_rays} rays, {rows} rows: FP64 frames bit-equal to the oracle; "
      f"diagnose counters equal on {n_scenes // 10} scenes ({diag_rays} flagged rays); FP32 mode on {fp32_scenes} scenes "
      f"({generic} of them with right-nested trees: the single-precision interpreter): {fp32_off} of {fp32_rays} rays off their FP64 path or beyond 1e-5 "
      f"({fp32_off / max(fp32_rays, 1):.2e}); {time.time() - t0:.0f} s")
