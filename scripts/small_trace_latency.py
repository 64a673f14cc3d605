"""Latency of many small traces (the optimiser-loop use, lens_design.ipynb cells 28-33)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import pyrayt_b200  # noqa: E402
from tests import fakes, scene_util as su  # noqa: E402

glass = fakes.BasicRefractor(1.5)
lens = fakes.CSG(fakes.Surface(fakes.Sphere(2.0), glass, su.translate(1.9, 0, 0)),
                 fakes.Surface(fakes.Sphere(2.0), glass, su.translate(-1.9, 0, 0)), 2, (-0.1, 0.1, -1, 1, -1, 1))
det = fakes.Surface(fakes.Plane(4, 4), fakes._AbsorbingMaterial(), su.translate(3, 0, 0) @ su.rot_y(90))
src = fakes.ConeOfRays(np.radians(3.0), world=su.translate(-3, 0, 0))
tracer = pyrayt_b200.RayTracer(src, [lens, det], rays_per_source=21, generation_limit=10)
tracer.trace()
for label, move in (("static scene", False), ("moving detector", True)):
    t0 = time.perf_counter()
    for k in range(200):
        if move:
            det.move_x(1e-3)
        df = tracer.trace()
    dt = (time.perf_counter() - t0) / 200
    print(f"RayTracer.trace() 21 rays, {label}: {dt * 1e3:.3f} ms per call, frame {df.shape}")
eng = tracer._engine
rays = torch.from_numpy(np.asarray(src.generate_rays(21))).cuda()
t0 = time.perf_counter()
for k in range(200):
    res = eng.trace(rays, generation_limit=10, to_host=True)
print(f"Engine.trace() 21 rays to host: {(time.perf_counter() - t0) / 200 * 1e3:.3f} ms per call")
eng.small_ray_buffer(21).copy_(rays)
for label, graph in (("eager launches", False), ("CUDA graph replay", True)):
    for k in range(3):
        eng.trace_small(21, generation_limit=10, use_graph=graph)
    t0 = time.perf_counter()
    for k in range(500):
        res = eng.trace_small(21, generation_limit=10, use_graph=graph)
    print(f"Engine.trace_small() 21 rays to host, {label}: {(time.perf_counter() - t0) / 500 * 1e3:.3f} ms per call")
