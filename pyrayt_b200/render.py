"""Renderer glue: the per-pixel nearest-hit loop of tinygfx's renderers on the GPU.

``EdgeRender`` and ``ShadedRenderer`` (tinygfx/g3d/renderers.py:72-94,:188-210) run the same loop as
``RayTracer._st_propagate`` over the rays of an ``OrthographicCamera``; here that loop is one launch
of ``prt_nearest_hit``.  Drawing (matplotlib, Gooch shading) stays the reference's host code.

One deliberate difference: for a pixel whose component hits are all behind the camera the reference
renderers pick up a *negative* distance (they index the unfiltered hit array, renderers.py:82); this
path, like the ray tracer itself, reports a miss.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine
from .scene import flatten


def camera_nearest(camera, components, device: int = 0, normals: bool = True, engine: Engine = None):
    """Nearest surface per camera pixel.

    camera: anything with ``generate_rays() -> (2,4,N)`` and ``get_resolution() -> (h, v)``
    (tinygfx.g3d.OrthographicCamera, world_objects.py:499-537).  Returns a dict of (v, h) images:
    ``distance`` (+inf = background), ``surface`` (id, -1 = background) and ``normal`` (v, h, 3).
    """
    import torch

    rays = np.ascontiguousarray(np.asarray(camera.generate_rays(), dtype=np.float64))
    h, v = camera.get_resolution()
    eng = engine if engine is not None else Engine(flatten(components), device)
    t, sid, nrm = eng.nearest_hit(torch.from_numpy(rays).to(torch.device("cuda", eng.device)), normals=normals)
    out = {"distance": t.cpu().numpy().reshape(v, h), "surface": sid.cpu().numpy().reshape(v, h)}
    if normals:
        out["normal"] = np.moveaxis(nrm.cpu().numpy().reshape(3, v, h), 0, -1)
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
