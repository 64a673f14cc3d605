"""Algorithmic work of one trace: the numerators of the roofline fractions.

Flops count every + - * / sqrt compare/select/abs of the *reference's* arithmetic as 1
(integer work free), derived line by line from the cited reference code:

    world->object transform of one ray          33   (world_objects.py:367-369, affine part)
    Sphere / Cylinder / Plane / Cube / Paraboloid test   34 / 62 / 45 / 55 / 62
                                                     (primitives.py:252-271,:664-712,:451-492,:528-578,:333-398)
    CSG node with m output slots                55 + 2m   (world AABB csg.py:126-128 + merge :36-61,:147-149)
    nearest hit of a component with m slots     2m + 1    (_pyrayt.py:380-386)
    interaction: glass / mirror / absorber segment  144 / 87 / 15
                                                     (_pyrayt.py:404-407,:449,:177; world_objects.py:409-418;
                                                      operations.py:105-107,:125-162; materials.py:140-145)

Bytes are the API's own input and output: 104 B per ray in (13 float64,
_pyrayt.py:21) and 120 B per frame row out (15 float64, _pyrayt.py:154-165).
"""
from __future__ import annotations

import numpy as np

from .scene import FlatScene, NODE_LEAF

TRANSFORM_FLOPS = 33
PRIM_FLOPS = {1: 34, 2: 62, 3: 45, 4: 55, 5: 62}  # Sphere, Paraboloid, Plane, Cube, Cylinder
GLASS_SEGMENT, MIRROR_SEGMENT, ABSORBER_SEGMENT = 144, 87, 15
RAY_BYTES_IN = 104
ROW_BYTES_OUT = 120


def flops_per_generation(scene: FlatScene) -> int:
    """Work of one ray in one generation: every leaf test + CSG overhead + nearest-hit."""
    total = sum(TRANSFORM_FLOPS + PRIM_FLOPS[int(t)] for t in scene.leaf_type)
    for c in range(scene.n_components):
        b, e = int(scene.comp_node_begin[c]), int(scene.comp_node_begin[c + 1])
        stack = []
        for k in scene.node_kind[b:e]:
            if k == NODE_LEAF:
                stack.append(2)
            else:
                r, l = stack.pop(), stack.pop()
                stack.append(l + r)
                total += 55 + 2 * (l + r)
        total += 2 * stack[0] + 1
    return int(total)


def algorithmic_flops(scene: FlatScene, counters: dict) -> int:
    seg = counters["segments"]
    ab = counters.get("absorber_segments", 0)
    mi = counters.get("mirror_segments", 0)
    gl = seg - ab - mi
    return int(counters["generations"] * flops_per_generation(scene)
               + gl * GLASS_SEGMENT + mi * MIRROR_SEGMENT + ab * ABSORBER_SEGMENT)


def algorithmic_bytes(n_rays: int, rows: int) -> int:
    return int(n_rays * RAY_BYTES_IN + rows * ROW_BYTES_OUT)
