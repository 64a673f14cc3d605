"""ctypes binding of ``libpyrayt_b200.so`` (the C ABI in include/pyrayt_b200.h).

There is deliberately no fallback: if the CUDA library is missing or will not
load, every entry point raises.  Build it with ``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C pyrayt_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYRAYT_B200_LIB") or os.path.join(_HERE, "libpyrayt_b200.so")  # override: experiments only

ABI_VERSION = 3
FRAME_COLS = 15
STAGE_COLS = 9  # doubles per staged record (PRT_STAGE_COLS)
RAY_ROWS = 13
RECORD_ALL, RECORD_SURFACE, RECORD_NONE = 0, 1, 2
SELECT_ALL, SELECT_SURFACE, SELECT_GENERATION = 0, 1, 2
SPOT_COLS, SPOT_CENTER_COLS, SPOT_MAX_GROUPS = 16, 4, 256

FRAME_COLUMNS = (
    "generation", "intensity", "wavelength", "index", "id", "surface",
    "x0", "y0", "z0", "x1", "y1", "z1", "x_tilt", "y_tilt", "z_tilt",
)  # pyrayt/_pyrayt.py:15,:154-165

COUNTER_FIELDS = (
    "rays", "generations", "segments", "rows_reserved", "rows_dropped", "tie_rays",
    "untraceable_hits", "bad_w", "nan_rays", "limit_rays", "absorber_segments", "mirror_segments",
    "grazing_rays", "seam_rays",  # PRT_FLAG_DIAGNOSE only
)
FLAG_DIAGNOSE = 1  # PRT_FLAG_DIAGNOSE
FLAG_FP32 = 2  # PRT_FLAG_FP32: the optional single-precision fast mode
LAYOUT_FP32_RECORDS = 0x100  # prt_gather_frame: the staged records were written by an FP32 trace
STAGE_COLS_FP32 = 5
COUNTER_WORDS = 16

# every symbol include/pyrayt_b200.h declares
EXPORTS = (
    "prt_abi_version", "prt_last_error", "prt_tile_rays", "prt_scene_create", "prt_scene_destroy",
    "prt_scene_n_leaves", "prt_trace", "prt_scan_runs", "prt_gather_frame", "prt_intersect",
    "prt_generate_source", "prt_fp64_probe", "prt_nearest_hit", "prt_scene_update",
    "prt_render_hit", "prt_wave_tile", "prt_trace_wavefront", "prt_frame_pack", "prt_host_expand_frame", "prt_spot_moments", "prt_spot_centers", "prt_axis_table_blocks", "prt_axis_table",
)


class PrtError(RuntimeError):
    pass


class PrtParams(ctypes.Structure):
    _fields_ = [
        ("generation_limit", ctypes.c_int32),
        ("record_mode", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("ray_offset", ctypes.c_double),
        ("detector_sid", ctypes.c_int64),
    ]


class PrtRecords(ctypes.Structure):
    _fields_ = [
        ("d_stage", ctypes.c_void_p),
        ("capacity", ctypes.c_int64),
        ("d_run_start", ctypes.c_void_p),
        ("d_run_count", ctypes.c_void_p),
        ("d_run_base", ctypes.c_void_p),
        ("n_tiles", ctypes.c_int64),
    ]


class PrtWaveWorkspace(ctypes.Structure):
    _fields_ = [
        ("d_state", ctypes.c_void_p),
        ("d_flag", ctypes.c_void_p),
        ("d_hit_t", ctypes.c_void_p),
        ("d_hit_leaf", ctypes.c_void_p),
        ("d_tile_count", ctypes.c_void_p),
        ("d_tile_base", ctypes.c_void_p),
        ("d_alive", ctypes.c_void_p),
        ("n_tiles", ctypes.c_int64),
    ]


class PrtSourceDesc(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("seed", ctypes.c_uint64),
        ("origin", ctypes.c_double * 3),
        ("p", ctypes.c_double * 16),
    ]


_lib = None


def load():
    """Load the CUDA library or fail loudly (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PrtError(
            f"{LIB_PATH} is missing: the B200 trace kernels are not built and pyrayt_b200 has no CPU "
            "fallback.  Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc)."
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.prt_abi_version.restype = ctypes.c_int
    lib.prt_last_error.restype = ctypes.c_char_p
    lib.prt_tile_rays.restype = ctypes.c_int
    lib.prt_scene_create.restype = ctypes.c_int
    lib.prt_scene_create.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp)]
    lib.prt_scene_destroy.restype = None
    lib.prt_scene_destroy.argtypes = [vp]
    lib.prt_scene_n_leaves.restype = ctypes.c_int
    lib.prt_scene_n_leaves.argtypes = [vp]
    lib.prt_trace.restype = ctypes.c_int
    lib.prt_trace.argtypes = [vp, ctypes.POINTER(PrtParams), vp, i64, i64, ctypes.POINTER(PrtRecords), vp, vp]
    lib.prt_wave_tile.restype = ctypes.c_int
    lib.prt_trace_wavefront.restype = ctypes.c_int
    lib.prt_trace_wavefront.argtypes = [vp, ctypes.POINTER(PrtParams), vp, i64, i64, ctypes.POINTER(PrtWaveWorkspace),
                                        vp, i64, i64, vp, vp, vp, vp]
    lib.prt_scan_runs.restype = ctypes.c_int
    lib.prt_scan_runs.argtypes = [ctypes.POINTER(PrtRecords), i32, vp, vp]
    lib.prt_gather_frame.restype = ctypes.c_int
    lib.prt_gather_frame.argtypes = [vp, ctypes.POINTER(PrtRecords), vp, i64, i64, i32, vp, vp, i64, i64, i32, vp]
    lib.prt_intersect.restype = ctypes.c_int
    lib.prt_intersect.argtypes = [vp, i32, vp, i64, vp, vp, ctypes.POINTER(i32), vp]
    lib.prt_generate_source.restype = ctypes.c_int
    lib.prt_generate_source.argtypes = [ctypes.POINTER(PrtSourceDesc), vp, i64, i64, i64, vp]
    lib.prt_nearest_hit.restype = ctypes.c_int
    lib.prt_nearest_hit.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.prt_render_hit.restype = ctypes.c_int
    lib.prt_render_hit.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.prt_scene_update.restype = ctypes.c_int
    lib.prt_scene_update.argtypes = [vp, vp, vp]
    f64 = ctypes.c_double
    lib.prt_spot_moments.restype = ctypes.c_int
    lib.prt_spot_moments.argtypes = [vp, i64, i64, i32, f64, i64, i32, vp, vp, i32, vp]
    lib.prt_spot_centers.restype = ctypes.c_int
    lib.prt_spot_centers.argtypes = [vp, vp, i32, vp, vp]
    lib.prt_axis_table_blocks.restype = i64
    lib.prt_axis_table_blocks.argtypes = [i64]
    lib.prt_axis_table.restype = ctypes.c_int
    lib.prt_axis_table.argtypes = [vp, i64, i64, i32, f64, i64, i64, vp, vp, vp, vp, i64, i64, vp]
    lib.prt_frame_pack.restype = ctypes.c_int
    lib.prt_frame_pack.argtypes = [vp, i64, i64, vp, i64, i64, vp, i32, vp, vp, vp]
    lib.prt_host_expand_frame.restype = ctypes.c_int
    lib.prt_host_expand_frame.argtypes = [vp, i64, vp, i32, vp, vp, vp, vp, vp, i64, i32]
    lib.prt_fp64_probe.restype = ctypes.c_int
    lib.prt_fp64_probe.argtypes = [vp, i32, i32, vp]
    if lib.prt_abi_version() != ABI_VERSION:
        raise PrtError(f"ABI mismatch: library {lib.prt_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().prt_last_error()
        raise PrtError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
