"""The BASELINE.json configs as concrete inputs.

Scenes are built from the reference's own component factories in the build
container (tests/golden/make_golden.py) and shipped flattened under
``pyrayt_b200/data/*.scene.json`` because the GPU box has no PyRayT tree; the
sources are the seeded synthetic sources of ``pyrayt_b200.sources``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

from . import sources
from .scene import FlatScene

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


@dataclass
class Workload:
    name: str
    scene_file: str
    source: Optional[sources.SyntheticSource]
    n_rays: int
    generation_limit: int
    description: str

    def scene(self) -> FlatScene:
        with open(os.path.join(_DATA, self.scene_file)) as fh:
            return FlatScene.from_json(fh.read())


CONFIG2_SOURCE = sources.solid_angle_cone(seed=1, apex=(-2.04, 0.0, 0.0), half_angle_deg=10.0, wavelength=0.633)
CONFIG4_SOURCE = sources.field_fan(seed=4, x_start=-10.0, radius=10.0, field_deg=(0.0, 2.0, 5.0),
                                   wavelengths=(0.486, 0.588, 0.656))
CONFIG5_SOURCE = sources.lambertian_cone(seed=5, apex=(0.0, 0.0, 0.0), half_angle_deg=20.0, wavelength=0.633)

WORKLOADS = {
    "config2": Workload("config2", "config2_tutorial.scene.json", CONFIG2_SOURCE, 100_000, 100,
                        "tutorial condenser lens + aperture stop + detector, 100k rays"),
    "config4": Workload("config4", "config4_stack.scene.json", CONFIG4_SOURCE, 1 << 24, 64,
                        "synthetic 10-element spherical-lens CSG stack with 2 stops + detector (35 leaves), 2^24 rays"),
    "config5": Workload("config5", "config5_cavity.scene.json", CONFIG5_SOURCE, 1 << 25, 32,
                        "parabolic mirror + BK7 TIR light pipe + two cuboid mirrors + detector (6 leaves), 2^25 rays per GPU"),
}
