"""Renderers' nearest-hit loop on the GPU.

``EdgeRender`` and ``ShadedRenderer`` (tinygfx/g3d/renderers.py:11-249) run, per camera pixel, the
same loop as ``RayTracer._st_propagate`` -- with one difference that is part of their output: the hit
distance and surface are read from the *unfiltered* hit array, so a pixel whose hits on a component
are all behind the camera picks up that component's first, negative, hit (renderers.py:79-84).
``prt_render_hit`` reproduces exactly that; ``prt_nearest_hit`` is the tracer's variant.

Only PROPAGATE is on the hot path: INTERACT (edge detection / ``surface.shade``, Gooch shading), the
canvas and everything else stay the reference's own host code.  ``install()`` swaps just
``_st_propagate`` of the reference's own classes, which is all ``tinygfx.g3d.renderers.draw`` and
``RayTracer.show`` need; ``camera_nearest`` is the same launch for callers without the reference.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine
from .scene import flatten


def _propagate(rays, components, device: int = 0, normals: bool = False, renderer: bool = True,
               engine: Engine = None):
    import torch

    rays = np.ascontiguousarray(np.asarray(rays, dtype=np.float64))
    eng = engine if engine is not None else Engine(flatten(components), device)
    t, sid, nrm = eng.nearest_hit(torch.from_numpy(rays).to(torch.device("cuda", eng.device)), normals=normals,
                                  renderer=renderer)
    return t.cpu().numpy(), sid.cpu().numpy(), (nrm.cpu().numpy() if normals else None)


def camera_nearest(camera, components, device: int = 0, normals: bool = True, engine: Engine = None,
                   renderer: bool = False):
    """Nearest surface per camera pixel.

    camera: anything with ``generate_rays() -> (2,4,N)`` and ``get_resolution() -> (h, v)``
    (tinygfx.g3d.OrthographicCamera, world_objects.py:499-537).  Returns a dict of (v, h) images:
    ``distance`` (+inf = background), ``surface`` (id, -1 = background) and ``normal`` (v, h, 3).
    renderer=False reports only hits in front of the camera (the ray tracer's rule); renderer=True is
    the reference renderers' rule (see the module docstring).
    """
    h, v = camera.get_resolution()
    t, sid, nrm = _propagate(camera.generate_rays(), components, device, normals, renderer, engine)
    out = {"distance": t.reshape(v, h), "surface": sid.reshape(v, h)}
    if normals:
        out["normal"] = np.moveaxis(nrm.reshape(3, v, h), 0, -1)
    return out


def install() -> None:
    """Route PROPAGATE of the reference's own renderers (``tinygfx.g3d.renderers.EdgeRender`` /
    ``ShadedRenderer``, hence ``draw()`` and ``RayTracer.show()``) through ``prt_render_hit``."""
    from tinygfx.g3d import renderers

    def _st_propagate(self):
        self._hit_distances, self._hit_surfaces, _ = _propagate(self._rays, self._shapes)
        self._state = self.States.INTERACT

    renderers.EdgeRender._st_propagate = _st_propagate
    renderers.ShadedRenderer._st_propagate = _st_propagate
