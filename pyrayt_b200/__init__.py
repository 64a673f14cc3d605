"""pyrayt_b200 -- B200-native ray-propagation hot path behind PyRayT's RayTracer API.

Host side is Python (like the reference); the generation loop, CSG hit-interval
merging, nearest-hit selection, material interaction and record logging run in
hand-written sm_100a CUDA kernels reached through a C ABI (include/pyrayt_b200.h).
There is no CPU fallback: without the built CUDA library every entry point raises.
"""
from ._lib import FRAME_COLUMNS, PrtError, load as load_library
from .scene import FlatScene, SceneError, flatten
from .engine import Engine, TraceResult
from .tracer import RayTracer, UntraceableSurfaceError, install
from . import analytics, render, sources

__all__ = [
    "RayTracer", "Engine", "TraceResult", "FlatScene", "flatten", "SceneError", "PrtError",
    "UntraceableSurfaceError", "install", "sources", "render", "analytics", "FRAME_COLUMNS", "load_library",
]
