"""Renderers' nearest-hit loop on the GPU.

``EdgeRender`` and ``ShadedRenderer`` (tinygfx/g3d/renderers.py:11-249) run, per camera pixel, the
same loop as ``RayTracer._st_propagate`` -- with one difference that is part of their output: the hit
distance and surface are read from the *unfiltered* hit array, so a pixel whose hits on a component
are all behind the camera picks up that component's first, negative, hit (renderers.py:79-84).
``prt_render_hit`` reproduces exactly that; ``prt_nearest_hit`` is the tracer's variant.

The classes below mirror the reference's two renderers (same constructor, ``render()`` and result
layout): PROPAGATE is one kernel launch, INTERACT (edge detection / ``surface.shade``, Gooch shading)
stays the reference's host code.  ``install()`` swaps just ``_st_propagate`` of the reference's own
classes, which is all ``tinygfx.g3d.renderers.draw`` and ``RayTracer.show`` need.
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


def edge_canvas(surface_image: np.ndarray) -> np.ndarray:
    """EdgeRender._st_interact (renderers.py:96-118): RGBA canvas with the surface boundaries drawn."""
    from scipy import ndimage

    hit = np.asarray(surface_image)
    h_diffs = np.abs(np.diff(hit, axis=-1, prepend=-1))
    v_diffs = np.abs(np.diff(hit, axis=0, prepend=-1))
    edges = ndimage.binary_dilation(h_diffs + v_diffs, ndimage.generate_binary_structure(2, 2),
                                    iterations=np.maximum(1, int(np.max(hit.shape) / 300)))
    canvas = np.zeros((*hit.shape, 4), dtype=float)
    canvas[..., :] = np.logical_not(edges)[..., np.newaxis]
    canvas[..., 3] = edges
    return canvas


class EdgeRender(object):
    """tinygfx.g3d.renderers.EdgeRender (renderers.py:11-126) with PROPAGATE on the GPU."""

    ray_offset_value = 1e-6

    def __init__(self, camera, surfaces, device: int = 0):
        self._camera = camera
        self._shapes = surfaces if hasattr(surfaces, "__iter__") else (surfaces,)
        self._device = device
        self._simulation_complete = False
        self._results = None

    def reset(self):
        self._simulation_complete = False
        self._results = None

    def render(self):
        self.reset()
        self._rays = self._camera.generate_rays()
        self._hit_distances, self._hit_surfaces, _ = _propagate(self._rays, self._shapes, self._device)
        hit_matrix = self._hit_surfaces.reshape(self._camera.get_resolution()[-1], -1)
        self._results = edge_canvas(hit_matrix)
        self._simulation_complete = True
        return self._results

    def get_results(self):
        return self._results


class ShadedRenderer(object):
    """tinygfx.g3d.renderers.ShadedRenderer (renderers.py:129-249): nearest hits on the GPU, then every
    surface's own ``shade(rays, distances, light_positions=...)`` on the host, as the reference does."""

    def __init__(self, camera, shapes, light_position, device: int = 0):
        self._light = np.asarray(light_position)
        self._camera = camera
        self._shapes = shapes if hasattr(shapes, "__iter__") else (shapes,)
        self._device = device
        self._surface_lut = tuple()
        for shape in self._shapes:
            self._surface_lut += tuple(shape.surface_ids)
        self._simulation_complete = False
        self._results = None

    def reset(self):
        self._simulation_complete = False
        self._results = None

    def render(self):
        self.reset()
        self._rays = self._camera.generate_rays()
        self._hit_distances, self._hit_surfaces, _ = _propagate(self._rays, self._shapes, self._device)
        canvas = np.zeros((4, self._rays.shape[-1]))
        for sid, surface in self._surface_lut:
            mask = self._hit_surfaces == sid
            if np.any(mask):
                canvas[:, mask] = surface.shade(self._rays[..., mask], self._hit_distances[mask],
                                                light_positions=self._light)
        self._results = canvas.T.reshape(*self._camera.get_resolution()[::-1], 4)
        self._simulation_complete = True
        return self._results


def install() -> None:
    """Route PROPAGATE of the reference's own renderers (``tinygfx.g3d.renderers.EdgeRender`` /
    ``ShadedRenderer``, hence ``draw()`` and ``RayTracer.show()``) through ``prt_render_hit``."""
    from tinygfx.g3d import renderers

    def _st_propagate(self):
        self._hit_distances, self._hit_surfaces, _ = _propagate(self._rays, self._shapes)
        self._state = self.States.INTERACT

    renderers.EdgeRender._st_propagate = _st_propagate
    renderers.ShadedRenderer._st_propagate = _st_propagate
