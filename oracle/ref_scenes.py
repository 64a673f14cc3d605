"""The BASELINE.json configs built from the UNMODIFIED reference's own component factories, and the
reference's ``trace()`` run on a fixed RaySet.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (needs the reference: oracle/ref_shim.py).  Used by
tests/golden/make_golden.py (fixtures), bench.py's NumPy arm (``cpu_baseline.kind = "reference"``) and
the live parity tests.  The flattened copies of these scenes ship as pyrayt_b200/data/*.scene.json.
"""
import os
import time

import numpy as np

from . import ref_shim


def _mods():
    pyrayt = ref_shim.load()
    import pyrayt.components as pc
    import pyrayt.materials as matl
    import tinygfx.g3d as cg

    return pyrayt, pc, matl, cg


def fixed_source(rays):
    """A reference ``Source`` whose generate_rays() returns the given (13, N) array
    (pyrayt/components.py:481-496: the Source interface)."""
    pyrayt, pc, _, _ = _mods()

    class FixedSource(pc.Source):
        def __init__(self, fixed):
            super().__init__()
            self._fixed = np.array(fixed, dtype=np.float64)

        def _local_ray_generation(self, n):
            rs = pyrayt.RaySet(self._fixed.shape[1])
            rs[:] = self._fixed
            return rs

    return FixedSource(rays)


def reference_trace(rays, components, generation_limit, stable=True):
    """RayTracer(FixedSource(rays), components).trace() -> (15, rows) float64 (pyrayt/_pyrayt.py:329-339)."""
    pyrayt, _, _, _ = _mods()
    tracer = pyrayt.RayTracer(fixed_source(rays), components)
    tracer.set_rays_per_source(rays.shape[1])
    tracer.set_generation_limit(generation_limit)
    if stable:
        with ref_shim.stable_argsort(), np.errstate(all="ignore"):
            df = tracer.trace()
    else:
        with np.errstate(all="ignore"):
            df = tracer.trace()
    return df.to_numpy(dtype=np.float64).T.copy() if len(df) else np.zeros((15, 0))


# ---------------------------------------------------------------- scenes

def config1_scene():
    """examples/convex_collimator.py:23-37"""
    _, pc, _, _ = _mods()
    return [pc.biconvex_lens(2, 2, 0.25, aperture=1), pc.baffle((1, 1)).move_x(1)]


def config2_scene():
    """docs/source/tutorial.rst lens + stop + detector"""
    _, pc, _, _ = _mods()
    return [pc.biconvex_lens(2, 2, 0.25, aperture=1), pc.aperture((1, 1), 0.6).move_x(0.5),
            pc.baffle((1, 1)).move_x(1)]


def config3_scene():
    """examples/chromatic_dispersion.py:10-15"""
    _, pc, _, _ = _mods()
    return [pc.equilateral_prism(1, 1).move_x(0.25), pc.baffle((1, 1)).rotate_y(90).move(1, 0, -0.5)]


def config4_scene():
    """10-element spherical-lens stack with two stops and a detector (35 leaves)."""
    _, pc, matl, _ = _mods()
    comps = []
    for i in range(10):
        if i % 2 == 0:
            lens = pc.thick_lens(60, -60, 4, aperture=25.4, material=matl.glass["BK7"])
        else:
            lens = pc.thick_lens(-80, 80, 2, aperture=25.4, material=matl.glass["SF5" if i % 4 == 1 else "SF2"])
        comps.append(lens.move_x(10 * i))
    comps.append(pc.aperture((25.4, 25.4), 12.0).move_x(35))
    comps.append(pc.aperture((25.4, 25.4), 12.0).move_x(75))
    comps.append(pc.baffle((25.4, 25.4)).move_x(100))
    return comps


def config5_scene():
    """Multi-bounce paraboloid / TIR light-pipe / cuboid-mirror scene.

    A point source at the focus of ``parabolic_mirror(50, 5, aperture=40)`` (focus at the origin)
    emits towards -x; the dish returns a collimated beam along +x.  ``m1`` (cuboid mirror, tilted
    0.3 deg) sends it back to the dish, which focuses it through the origin into the end face of a
    BK7 ``Cuboid.from_sides(200, 10, 10)`` light pipe whose axis is tilted 20 deg to the beam: rays
    zig-zag down the pipe by total internal reflection (about 9 glass interactions per ray), leave
    through the far face, and a second cuboid mirror ``m2`` folds them onto the detector baffle.
    Part of the outgoing beam also crosses the pipe sideways (two refractions).  Measured with the
    reference: 12.4 rows per ray on average, 51 % of the rays end on the detector.
    """
    _, pc, matl, cg = _mods()
    c20, s20 = np.cos(np.radians(20.0)), np.sin(np.radians(20.0))
    parab = pc.parabolic_mirror(50, 5, aperture=40)
    pipe = cg.Cuboid.from_sides(200, 10, 10, material=matl.glass["BK7"]).rotate_z(20).move(106 * c20, 106 * s20, 0)
    m1 = pc.plane_mirror(2, aperture=(60, 60)).rotate_z(0.3).move(120, 0, 0)
    m2 = pc.plane_mirror(2, aperture=(60, 60)).rotate_z(-35).move(241 * c20, 241 * s20, 0)
    det = pc.baffle((80, 80)).rotate_z(90).move(226, 160, 0)
    return [parab, pipe, m1, m2, det]


SCENES = {"config1": config1_scene, "config2": config2_scene, "config3": config3_scene,
          "config4": config4_scene, "config5": config5_scene}


# ---------------------------------------------------------------- timing the NumPy path

def _trace_chunk(args):
    """One worker of the multi-process NumPy arm: builds the scene itself (live objects do not pickle
    cheaply and surface ids are process-local), traces its ray range, returns (rows, seconds)."""
    name, rays, generation_limit = args
    comps = SCENES[name]()
    t0 = time.perf_counter()
    frame = reference_trace(rays, comps, generation_limit, stable=True)
    return frame.shape[1], time.perf_counter() - t0


def time_reference(name, rays, generation_limit, processes=1):
    """Wall-clock of the unmodified reference tracing `rays` through config `name`.

    processes = 1: RayTracer.trace() as a user runs it (NumPy's elementwise kernels are single-threaded, so this
    is one core).  processes > 1: the rays are split into contiguous index ranges, one process per range
    (valid because rays never interact; a sharded trace re-sorted by (generation, id) is the monolithic
    trace), wall-clock over the whole pool including the per-process scene construction.
    Returns {"seconds", "rows", "rays", "processes"}."""
    n = rays.shape[1]
    if processes <= 1:
        rows, sec = _trace_chunk((name, rays, generation_limit))
        return {"seconds": sec, "rows": rows, "rays": n, "processes": 1}
    import multiprocessing as mp

    bounds = [(n * k) // processes for k in range(processes + 1)]
    jobs = [(name, np.ascontiguousarray(rays[:, bounds[k]:bounds[k + 1]]), generation_limit)
            for k in range(processes) if bounds[k + 1] > bounds[k]]
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(jobs)) as pool:
        parts = pool.map(_trace_chunk, jobs)
    sec = time.perf_counter() - t0
    return {"seconds": sec, "rows": int(sum(p[0] for p in parts)), "rays": n, "processes": len(jobs)}


def argsort_mismatch(name, rays, generation_limit):
    """rows of the reference's frame that change between the stable argsort (pinned numpy
    1.20 behaviour, the parity contract) and this numpy's default (unstable for small lanes on AVX-512).
    Returns {"rows_stable", "rows_default", "rays_differing", "rows_differing"}."""
    comps = SCENES[name]()
    a = reference_trace(rays, comps, generation_limit, stable=True)
    b = reference_trace(rays, comps, generation_limit, stable=False)

    def per_ray(frame):
        out = {}
        for col in frame.T:
            out.setdefault(col[4], []).append(col.tobytes())
        return out

    ra, rb = per_ray(a), per_ray(b)
    ids = set(ra) | set(rb)
    bad_rays = [i for i in ids if ra.get(i) != rb.get(i)]
    rows_diff = sum(len(set(ra.get(i, [])) ^ set(rb.get(i, []))) for i in bad_rays)
    return {"rows_stable": int(a.shape[1]), "rows_default": int(b.shape[1]), "rays_differing": len(bad_rays),
            "rows_differing": int(rows_diff), "rays": int(rays.shape[1]), "numpy": np.__version__}
