"""Duck-typed stand-ins for the reference's scene objects (the GPU box has no PyRayT tree).

They expose exactly the attributes pyrayt_b200.scene.flatten reads from tinygfx /
pyrayt objects, so the drop-in RayTracer can be exercised without the reference.
"""
import itertools

import numpy as np

_ids = itertools.count(1000)


class TracableMaterial:
    def trace(self, surface, ray_set):
        raise NotImplementedError


class _AbsorbingMaterial(TracableMaterial):
    pass


class _ReflectingMaterial(TracableMaterial):
    pass


class BasicRefractor(TracableMaterial):
    def __init__(self, n):
        self._refractive_index = n


class SellmeierRefractor(TracableMaterial):
    def __init__(self, b1=0, b2=0, b3=0, c1=0, c2=0, c3=0):
        self.b1, self.b2, self.b3, self.c1, self.c2, self.c3 = b1, b2, b3, c1, c2, c3


class Gooch:  # no trace(): untraceable
    pass


class Sphere:
    def __init__(self, r):
        self._radius = r


class Plane:
    def __init__(self, w, l):
        self._width, self._length = w, l


class Cylinder:
    def __init__(self, r, lo, hi):
        self._radius, self._h_min, self._h_max, self._capped = r, lo, hi, True


class Cube:
    def __init__(self, lo, hi):
        self.axis_spans = np.sort(np.vstack((lo, hi)), axis=0).T


class Surface:
    def __init__(self, primitive, material, world=None):
        self._surface_primitive = primitive
        self.material = material
        self._world = np.eye(4) if world is None else np.asarray(world, dtype=float)
        self._normal_scale = 1
        self._id = next(_ids)

    def _get_object_transform(self):
        return np.linalg.inv(self._world)

    def get_id(self):
        return self._id

    def move_x(self, dx):
        t = np.eye(4)
        t[0, 3] = dx
        self._world = t @ self._world
        return self


class Operation:
    def __init__(self, value):
        self.value = value


class CSG:
    def __init__(self, left, right, op, spans):
        self._l_child, self._r_child, self._operation = left, right, Operation(op)
        self._aobb = Cube(np.asarray(spans, dtype=float).reshape(3, 2)[:, 0], np.asarray(spans, dtype=float).reshape(3, 2)[:, 1])
        if op == 3:
            right._normal_scale = -1


class ArraySource:
    """Source plugin: generate_rays(n) -> (13, n) RaySet array (pyrayt/components.py:481-496)."""

    def __init__(self, rays):
        self._rays = np.asarray(rays, dtype=float)

    def generate_rays(self, n):
        return self._rays[:, :n].copy()


class _RefSource:
    """Stand-in for the reference's deterministic sources: same class names and attributes."""

    _kind = None
    _attr = None

    def __init__(self, value, wavelength=0.633, world=None):
        setattr(self, self._attr, value)
        self._wavelength = wavelength
        self._world_coordinate_transform = np.eye(4) if world is None else np.asarray(world, dtype=float)

    def generate_rays(self, n):
        from oracle import sources_np

        p = [getattr(self, self._attr), self._wavelength, float(n), 0.0] + list(self._world_coordinate_transform[:3].reshape(12))
        return sources_np.reference_source(self._kind, p, n)


class LineOfRays(_RefSource):
    _kind, _attr = 10, "_spacing"


class CircleOfRays(_RefSource):
    _kind, _attr = 11, "_diameter"


class ConeOfRays(_RefSource):
    _kind, _attr = 12, "_angle"


class WedgeOfRays(_RefSource):
    _kind, _attr = 13, "_angle"


class Lamp:
    """Stand-in for pyrayt.components.Lamp (:616-654): attributes only; its own generate_rays draws from
    NumPy's global stream like the reference's."""

    def __init__(self, width, length, max_angle=90.0, wavelength=0.633, world=None):
        self._width, self._length, self._max_angle, self._wavelength = width, length, max_angle * np.pi / 180, wavelength
        self._world_coordinate_transform = np.eye(4) if world is None else np.asarray(world, dtype=float)

    def generate_rays(self, n):
        uv = np.random.random_sample((2, n))
        theta, phi = np.arccos(1 - uv[0] * (1 - np.cos(self._max_angle))), uv[1] * 2 * np.pi
        rays = np.zeros((2, 4, n))
        rays[0, 3] = 1
        rays[0, 1] = self._width * (np.random.random_sample(n) - 0.5)
        rays[0, 2] = self._length * (np.random.random_sample(n) - 0.5)
        rays[1, 0], rays[1, 1], rays[1, 2] = np.cos(theta), np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi)
        rays = np.matmul(self._world_coordinate_transform, rays)
        rays[1] /= np.linalg.norm(rays[1], axis=0)
        out = np.zeros((13, n))
        out[:8] = rays.reshape(8, n)
        out[9], out[10], out[11], out[12] = 100.0 * np.cos(theta), self._wavelength, 1.0, np.arange(n)
        return out


class OrthographicCamera:
    """Stand-in for tinygfx.g3d.OrthographicCamera (world_objects.py:499-537): rays along +x."""

    def __init__(self, h_pixels, h_width, aspect, world=None):
        self._h, self._w, self._vw, self._v = h_pixels, h_width, aspect * h_width, int(aspect * h_pixels)
        self._world = np.eye(4) if world is None else np.asarray(world, dtype=float)

    def get_resolution(self):
        return (self._h, self._v)

    def generate_rays(self):
        hs = np.linspace(self._w / 2, -self._w / 2, self._h)
        vs = np.linspace(self._vw / 2, -self._vw / 2, self._v)
        rays = np.zeros((2, 4, self._h * self._v))
        rays[0, 3] = 1
        ys, zs = np.meshgrid(hs, vs)
        rays[0, 1], rays[0, 2], rays[1, 0] = ys.reshape(-1), zs.reshape(-1), 1
        rays = np.matmul(self._world, rays)
        rays[1] /= np.linalg.norm(rays[1], axis=0)
        return rays
