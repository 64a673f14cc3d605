"""The BASELINE.json configs as concrete inputs.

Scenes are built from the reference's own component factories in the build
container (tests/golden/make_golden.py) and shipped flattened under
``pyrayt_b200/data/*.scene.json`` because the GPU box has no PyRayT tree; the
sources are the seeded synthetic sources of ``pyrayt_b200.sources``.
"""
from __future__ import annotations

import os
import math
from dataclasses import dataclass
from typing import Optional

from . import sources
from .scene import FlatScene

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


@dataclass
class Workload:
    name: str
    scene_file: str
    source: object  # sources.SyntheticSource or sources.ReferenceSourceSet
    n_rays: int
    generation_limit: int
    description: str

    def scene(self) -> FlatScene:
        with open(os.path.join(_DATA, self.scene_file)) as fh:
            return FlatScene.from_json(fh.read())


def _lensmakers(r1, r2, n_lens, thickness):
    # examples/convex_collimator.py:4-17
    return 1 / ((n_lens - 1) * (1 / r1 - 1 / r2 + (n_lens - 1) * thickness / (n_lens * r1 * r2)))


# config 1: ConeOfRays(cone_angle=6).move_x(-focus) (examples/convex_collimator.py:31)
CONFIG1_SOURCE = sources.ReferenceSourceSet([
    (12, 6.0 * math.pi / 180.0, 0.633, (1.0, 0.0, 0.0, -_lensmakers(2, -2, 1.5, 0.25),
                                    0.0, 1.0, 0.0, 0.0,
                                    0.0, 0.0, 1.0, 0.0))])


def _config3_world():
    # LineOfRays(...).move_x(-0.5).rotate_y(-3): world = R_y(-3 deg) @ T(-0.5, 0, 0)
    # (examples/chromatic_dispersion.py:18-23; tinygfx/g3d/world_objects.py rotate_y / move_x)
    a = -3.0 * math.pi / 180.0  # the reference's own degree conversion (world_objects.py:67-69)
    c, s = math.cos(a), math.sin(a)
    return (c, 0.0, s, c * -0.5,
            0.0, 1.0, 0.0, 0.0,
            -s, 0.0, c, -s * -0.5)


# config 3: 11 LineOfRays(spacing=0.1, wavelength=linspace(0.44, 0.75, 11)[k])
CONFIG3_WAVELENGTHS = tuple(0.44 + k * ((0.75 - 0.44) / 10) for k in range(10)) + (0.75,)  # np.linspace's formula
CONFIG3_SOURCE = sources.ReferenceSourceSet([(10, 0.1, wl, _config3_world()) for wl in CONFIG3_WAVELENGTHS])

CONFIG2_SOURCE = sources.solid_angle_cone(seed=1, apex=(-2.04, 0.0, 0.0), half_angle_deg=10.0, wavelength=0.633)
CONFIG4_SOURCE = sources.field_fan(seed=4, x_start=-10.0, radius=10.0, field_deg=(0.0, 2.0, 5.0),
                                   wavelengths=(0.486, 0.588, 0.656))
CONFIG5_SOURCE = sources.lambertian_cone(seed=5, apex=(0.0, 0.0, 0.0), half_angle_deg=20.0, wavelength=0.633)

WORKLOADS = {
    "config1": Workload("config1", "config1_collimator.scene.json", CONFIG1_SOURCE, 50, 100,
                        "examples/convex_collimator.py as shipped: biconvex lens + baffle, ConeOfRays(6), 50 rays"),
    "config2": Workload("config2", "config2_tutorial.scene.json", CONFIG2_SOURCE, 100_000, 100,
                        "tutorial condenser lens + aperture stop + detector, 100k rays"),
    "config3": Workload("config3", "config3_prism.scene.json", CONFIG3_SOURCE, 11 * 95_326, 10,
                        "examples/chromatic_dispersion.py prism + baffle, 11 LineOfRays sources x 95,326 rays "
                        "(1,048,586 rays), Sellmeier BK7"),
    "config4": Workload("config4", "config4_stack.scene.json", CONFIG4_SOURCE, 1 << 24, 64,
                        "synthetic 10-element spherical-lens CSG stack with 2 stops + detector (35 leaves), 2^24 rays"),
    "config5": Workload("config5", "config5_cavity.scene.json", CONFIG5_SOURCE, 1 << 25, 32,
                        "parabolic mirror + BK7 TIR light pipe + two cuboid mirrors + detector (6 leaves), 2^25 rays per GPU"),
}
