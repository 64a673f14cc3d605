"""ctypes binding of tests/emul/libprt_emul.so: the kernel's per-ray code run on the host (debug/test only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = ctypes.POINTER(ctypes.c_double)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libprt_emul.so")
        srcs = [os.path.join(_HERE, "prt_emul.cpp")] + [
            os.path.join(_HERE, "..", "..", "pyrayt_b200", "csrc", f)
            for f in ("prt_device.cuh", "prt_device_f32.cuh", "prt_encode.h", "prt_scene.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so,
                                   srcs[0]])
        _LIB = ctypes.CDLL(so)
        _LIB.prt_emul_trace.restype = ctypes.c_longlong
    return _LIB


def trace(scene, rays, generation_limit, ray_offset=1e-6, diagnose=False):
    """Returns (frame (15, rows) ordered (generation, input order), counters); diagnose = PRT_FLAG_DIAGNOSE."""
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    n = rays.shape[1]
    cap = max(1, n * generation_limit)
    rows = np.empty((cap, 15))
    nrows = np.zeros(max(n, 1), dtype=np.int32)
    ctr = (ctypes.c_ulonglong * 10)()
    desc = scene.as_desc()
    lib().prt_emul_set_diagnose(ctypes.c_int(1 if diagnose else 0))
    total = lib().prt_emul_trace(ctypes.byref(desc), rays.ctypes.data_as(_dp), ctypes.c_longlong(n),
                                 ctypes.c_longlong(n), ctypes.c_int(generation_limit), ctypes.c_double(ray_offset),
                                 rows.ctypes.data_as(_dp), ctypes.c_longlong(cap),
                                 nrows.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctr)
    assert total >= 0
    rows = rows[:total]
    # ray-major -> (generation round, input order)
    rnd = np.concatenate([np.arange(k) for k in nrows[:n]]) if total else np.zeros(0, dtype=np.int64)
    order = np.argsort(rnd, kind="stable")
    lib().prt_emul_set_diagnose(ctypes.c_int(0))
    names = ("rays", "generations", "segments", "tie_rays", "untraceable_hits", "nan_rays", "limit_rays",
             "grazing_rays", "seam_rays")
    return rows[order].T.copy(), dict(zip(names, [int(x) for x in ctr[:9]]))


def trace_f32(scene, rays, generation_limit, ray_offset=1e-6):
    """The FP32 fast mode's per-ray code on the host: frame (15, rows) ordered (generation, input order)."""
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    n = rays.shape[1]
    cap = max(1, n * generation_limit)
    rows = np.empty((cap, 15))
    nrows = np.zeros(max(n, 1), dtype=np.int32)
    desc = scene.as_desc()
    L = lib()
    L.prt_emul_trace_f32.restype = ctypes.c_longlong
    total = L.prt_emul_trace_f32(ctypes.byref(desc), rays.ctypes.data_as(_dp), ctypes.c_longlong(n),
                                 ctypes.c_longlong(n), ctypes.c_int(generation_limit), ctypes.c_double(ray_offset),
                                 rows.ctypes.data_as(_dp), ctypes.c_longlong(cap),
                                 nrows.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert total >= 0
    rows = rows[:total]
    rnd = np.concatenate([np.arange(k) for k in nrows[:n]]) if total else np.zeros(0, dtype=np.int64)
    return rows[np.argsort(rnd, kind="stable")].T.copy()


def intersect(scene, component, rays):
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(8, -1)
    n = rays.shape[1]
    m = scene.component_slots(component)
    hits = np.empty((m, n))
    sids = np.empty((m, n), dtype=np.int64)
    desc = scene.as_desc()
    rc = lib().prt_emul_intersect(ctypes.byref(desc), ctypes.c_int(component), rays.ctypes.data_as(_dp),
                                  ctypes.c_longlong(n), hits.ctypes.data_as(_dp),
                                  sids.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    assert rc == 0
    return hits, sids


def prune_flags(scene):
    """Per component: 1 = bounding-box pruning proven safe, 0 = disabled, -1 = bare leaf."""
    flags = np.zeros(scene.n_components, dtype=np.int32)
    desc = scene.as_desc()
    rc = lib().prt_emul_prune_flags(ctypes.byref(desc), flags.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert rc == 0
    return flags


def selfcheck_left_deep():
    """(disagreements, cases) of the closed-form left-deep evaluation vs the streaming merge."""
    cases = ctypes.c_longlong()
    L = lib()
    L.prt_emul_selfcheck_left_deep.restype = ctypes.c_longlong
    bad = L.prt_emul_selfcheck_left_deep(ctypes.byref(cases))
    return int(bad), int(cases.value)
