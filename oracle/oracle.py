"""ctypes binding of ``oracle/libprt_oracle.so`` (the C restatement).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs, never by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FRAME_COLS = 15
RAY_ROWS = 13
_dp = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libprt_oracle.so")
    src = os.path.join(_HERE, "trace_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "pyrayt_b200.h")
    stale = (not os.path.exists(so)) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libprt_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.prt_oracle_trace.restype = ctypes.c_int64
        _LIB.prt_oracle_trace.argtypes = [
            ctypes.c_void_p, _dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_int,
            _dp, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint64),
        ]
        _LIB.prt_oracle_trace_diagnose.restype = ctypes.c_int64
        _LIB.prt_oracle_trace_diagnose.argtypes = [
            ctypes.c_void_p, _dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_int,
            ctypes.POINTER(ctypes.c_uint64),
        ]
        _LIB.prt_oracle_intersect.restype = ctypes.c_int
        _LIB.prt_oracle_intersect.argtypes = [
            ctypes.c_void_p, ctypes.c_int, _dp, ctypes.c_int64, _dp, ctypes.POINTER(ctypes.c_int64),
        ]
        _LIB.prt_oracle_refract.restype = ctypes.c_double
        _LIB.prt_oracle_refract.argtypes = [_dp, _dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, _dp]
        _LIB.prt_oracle_index_at.restype = ctypes.c_double
        _LIB.prt_oracle_index_at.argtypes = [ctypes.c_int, _dp, ctypes.c_double]
        _LIB.prt_oracle_reflect.argtypes = [_dp, _dp, _dp]
        _LIB.prt_oracle_prim_intersect.argtypes = [ctypes.c_int, _dp, _dp, _dp, _dp]
        _LIB.prt_oracle_prim_normal.argtypes = [ctypes.c_int, _dp, _dp, _dp]
        _LIB.prt_oracle_array_csg.argtypes = [_dp, ctypes.c_int, _dp, ctypes.c_int, ctypes.c_int, _dp]
        _LIB.prt_oracle_world_normal.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
    return _LIB


def _p(a):
    return a.ctypes.data_as(_dp)


COUNTER_NAMES = ("rays", "generations", "segments", "untraceable_hits", "nan_rays", "limit_rays")


def trace(scene, rays: np.ndarray, generation_limit: int = 10, ray_offset: float = 1e-6, threads: int = 1,
          record: bool = True):
    """RayTracer.trace() on a FlatScene and a (13,N) float64 RaySet array.

    Returns (frame (15, rows) float64 column-major, counters dict)."""
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    assert rays.ndim == 2 and rays.shape[0] == RAY_ROWS
    n = rays.shape[1]
    desc = scene.as_desc()
    ctr = (ctypes.c_uint64 * 6)()
    L = lib()
    rows = L.prt_oracle_trace(ctypes.byref(desc), _p(rays), n, n, generation_limit, ray_offset, threads, None, 0, ctr)
    if rows < 0:
        raise RuntimeError(f"oracle trace failed: {rows}")
    counters = dict(zip(COUNTER_NAMES, [int(x) for x in ctr]))
    if not record:
        return None, counters
    frame = np.empty((FRAME_COLS, max(rows, 1)), dtype=np.float64)
    rows2 = L.prt_oracle_trace(ctypes.byref(desc), _p(rays), n, n, generation_limit, ray_offset, threads,
                               _p(frame), frame.shape[1], ctr)
    assert rows2 == rows
    return frame[:, :rows], counters


def diagnose(scene, rays: np.ndarray, generation_limit: int = 10, ray_offset: float = 1e-6, threads: int = 1):
    """The PRT_FLAG_DIAGNOSE counters (include/pyrayt_b200.h): the trace's counters plus ``grazing_rays`` and
    ``seam_rays`` -- rays whose nearest-hit answer changes when the origin moves by 1e-9."""
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    n = rays.shape[1]
    desc = scene.as_desc()
    ctr = (ctypes.c_uint64 * 8)()
    rows = lib().prt_oracle_trace_diagnose(ctypes.byref(desc), _p(rays), n, n, generation_limit, ray_offset, threads,
                                           ctr)
    if rows < 0:
        raise RuntimeError(f"oracle trace failed: {rows}")
    return dict(zip(COUNTER_NAMES + ("grazing_rays", "seam_rays"), [int(x) for x in ctr]))


def trace_timed(scene, rays: np.ndarray, generation_limit: int, ray_offset: float, threads: int, frame: np.ndarray):
    """Single full pass writing into a preallocated (15, cap) frame; returns rows (bench CPU baseline)."""
    desc = scene.as_desc()
    ctr = (ctypes.c_uint64 * 6)()
    n = rays.shape[1]
    rows = lib().prt_oracle_trace(ctypes.byref(desc), _p(rays), n, n, generation_limit, ray_offset, threads,
                                  _p(frame), frame.shape[1], ctr)
    if rows < 0:
        raise RuntimeError(f"oracle trace failed: {rows}")
    return rows, dict(zip(COUNTER_NAMES, [int(x) for x in ctr]))


def intersect(scene, component: int, rays: np.ndarray):
    """component.intersect(rays (2,4,N)) -> (hits (m,N), surface ids (m,N))."""
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(8, -1)
    n = rays.shape[1]
    m = scene.component_slots(component)
    hits = np.empty((m, n), dtype=np.float64)
    sids = np.empty((m, n), dtype=np.int64)
    desc = scene.as_desc()
    rc = lib().prt_oracle_intersect(ctypes.byref(desc), component, _p(rays), n, _p(hits),
                                    sids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    if rc:
        raise RuntimeError("oracle intersect failed")
    return hits, sids


def prim_intersect(ptype: int, params, origin, direction):
    prm = np.zeros(6); prm[: len(params)] = params
    o = np.asarray(origin, dtype=np.float64); d = np.asarray(direction, dtype=np.float64)
    out = np.empty(2)
    lib().prt_oracle_prim_intersect(ptype, _p(prm), _p(o), _p(d), _p(out))
    return out


def prim_normal(ptype: int, params, point):
    prm = np.zeros(6); prm[: len(params)] = params
    q = np.asarray(point, dtype=np.float64)
    out = np.empty(3)
    lib().prt_oracle_prim_normal(ptype, _p(prm), _p(q), _p(out))
    return out


def array_csg(a1, a2, op: int):
    a1 = np.asarray(a1, dtype=np.float64); a2 = np.asarray(a2, dtype=np.float64)
    out = np.empty(len(a1) + len(a2))
    lib().prt_oracle_array_csg(_p(a1), len(a1), _p(a2), len(a2), op, _p(out))
    return out


def reflect(v, n):
    v = np.asarray(v, dtype=np.float64); n = np.asarray(n, dtype=np.float64)
    out = np.empty(3)
    lib().prt_oracle_reflect(_p(v), _p(n), _p(out))
    return out


def refract(v, n, n1, n2, n_global=1.0):
    v = np.asarray(v, dtype=np.float64); n = np.asarray(n, dtype=np.float64)
    out = np.empty(3)
    idx = lib().prt_oracle_refract(_p(v), _p(n), n1, n2, n_global, _p(out))
    return out, idx


def index_at(mat: int, matp, wavelength: float) -> float:
    mp = np.zeros(6); mp[: len(matp)] = matp
    return lib().prt_oracle_index_at(mat, _p(mp), wavelength)


def world_normal(scene, leaf: int, point):
    desc = scene.as_desc()
    q = np.asarray(point, dtype=np.float64)
    out = np.empty(3)
    lib().prt_oracle_world_normal(ctypes.byref(desc), leaf, _p(q), _p(out))
    return out


def render_hit(scene, rays: np.ndarray):
    """EdgeRender / ShadedRenderer._st_propagate (tinygfx/g3d/renderers.py:72-94): like nearest(), but a
    pixel with no positive hit reports slot 0 of the unfiltered hit array (possibly a negative distance)."""
    return nearest(scene, rays, _entry="prt_oracle_render_hit")


def nearest(scene, rays: np.ndarray, _entry="prt_oracle_nearest"):
    """_st_propagate alone: rays (2,4,N) -> (distance (N,), surface id (N,), world normals (3,N))."""
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(8, -1)
    n = rays.shape[1]
    t = np.empty(n)
    sid = np.empty(n, dtype=np.int64)
    nrm = np.empty((3, n))
    desc = scene.as_desc()
    L = lib()
    fn = getattr(L, _entry)
    fn.restype = ctypes.c_int
    rc = fn(ctypes.byref(desc), _p(rays), ctypes.c_int64(n), _p(t),
                              sid.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _p(nrm))
    if rc:
        raise RuntimeError("oracle nearest failed")
    return t, sid, nrm
