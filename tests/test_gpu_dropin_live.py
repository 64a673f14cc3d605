"""The real drop-in on a B200: pyrayt_b200.install() + the reference's own example scripts, UNCHANGED, with
the CUDA engine underneath, compared with the unmodified reference in the same process.

Needs a GPU *and* the reference.  The reference reaches the GPU box as the git-ignored install under
baseline/_ref (oracle/stage_reference.py; it travels with the repo snapshot); /root/reference is never
read there.  Surface ids are process-global creation-order integers (tinygfx/g3d/world_objects.py:26-31),
so the comparison re-traces the *same* tracer object with the reference's original trace() instead of
building the scene twice.

Bar (BASELINE.json north_star): generation / id / surface columns bit-equal, every float within 1e-9
relative; checked with the sources generated on the device (N1, the default) and by their own host code.
"""
import os
import runpy

import numpy as np
import pytest

from oracle import ref_shim

pytestmark = [pytest.mark.gpu, pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.available(), reason="PyRayT reference not present (baseline/_ref)")]

RTOL, ATOL = 1e-9, 1e-12


@pytest.fixture()
def live(cuda_device):
    """Reference imported, trace() routed through the CUDA engine; every trace is checked against the
    reference's own trace() of the same tracer object.  Yields (pyrayt, log of compared traces)."""
    pyrayt = ref_shim.load()
    import pyrayt_b200

    ref_trace = pyrayt.RayTracer.trace
    ref_show = pyrayt.RayTracer.show
    log = []
    pyrayt_b200.install()
    b200_trace = pyrayt.RayTracer.trace
    assert b200_trace is not ref_trace

    def checked_trace(self):
        got = b200_trace(self)
        engine = self._b200._engine
        assert type(engine) is pyrayt_b200.Engine, "the drop-in must run the CUDA engine"
        got_np = got.to_numpy(dtype=np.float64).copy()
        with ref_shim.stable_argsort(), np.errstate(all="ignore"):
            want = ref_trace(self)
        want_np = want.to_numpy(dtype=np.float64)
        assert list(got.columns) == list(want.columns)
        assert got_np.shape == want_np.shape, (got_np.shape, want_np.shape)
        cols = list(got.columns)
        for name in ("generation", "id", "surface"):
            k = cols.index(name)
            assert np.array_equal(got_np[:, k], want_np[:, k]), f"{name} column differs"
        np.testing.assert_allclose(got_np, want_np, rtol=RTOL, atol=ATOL)
        log.append({"rows": got_np.shape[0], "rays": int(self._b200.last_result.counters["rays"]),
                    "launches": self._b200.last_result.launches,
                    "max_abs_err": float(np.max(np.abs(got_np - want_np))) if got_np.size else 0.0})
        # leave the B200 frame in place, as the user would see it
        self._frame.data = got
        return got

    pyrayt.RayTracer.trace = checked_trace
    pyrayt.RayTracer.show = lambda self, *a, **k: None  # no matplotlib in this image; show() is not the path
    try:
        yield pyrayt, log
    finally:
        pyrayt.RayTracer.trace = ref_trace
        pyrayt.RayTracer.show = ref_show
        for attr in ("_b200_engine_factory", "_b200_device_sources"):
            if hasattr(pyrayt.RayTracer, attr):
                delattr(pyrayt.RayTracer, attr)


@pytest.mark.parametrize("device_sources", [True, False])
@pytest.mark.parametrize("script", ["convex_collimator.py", "chromatic_dispersion.py"])
def test_reference_examples_unchanged(live, script, device_sources):
    pyrayt, log = live
    pyrayt.RayTracer._b200_device_sources = device_sources
    path = os.path.join(ref_shim.examples_dir(), script)
    runpy.run_path(path, run_name="__main__")  # the script as shipped: builds the scene, trace(), show()
    assert len(log) == 1 and log[0]["rows"] > 0 and log[0]["launches"] >= 1
    expect = {"convex_collimator.py": (50, 150), "chromatic_dispersion.py": (11, 33)}[script]
    assert (log[0]["rays"], log[0]["rows"]) == expect


@pytest.mark.parametrize("device_sources", [True, False])
def test_tutorial_scene_and_moving_components(live, device_sources):
    """docs/source/tutorial.rst: lens + stop + detector, ConeOfRays(10) at -2.04; then the scene is moved
    between traces (the tracer holds references), traced again, and a second source is added."""
    pyrayt, log = live
    pyrayt.RayTracer._b200_device_sources = device_sources
    pc = pyrayt.components
    lens = pc.biconvex_lens(2, 2, 0.25, aperture=1)
    stop = pc.aperture((1, 1), 0.6).move_x(0.5)
    det = pc.baffle((1, 1)).move_x(1)
    source = pc.ConeOfRays(cone_angle=10).move_x(-2.04)
    tracer = pyrayt.RayTracer(sources=source, components=[lens, stop, det])
    tracer.set_rays_per_source(1000)
    tracer.set_generation_limit(100)
    df = tracer.trace()
    assert tracer.get_results() is df and df.shape[1] == 15
    tracer.calculate_source_ids()
    assert "source_id" in tracer.get_results().columns
    det.move_x(0.5)
    lens.rotate_z(2.0)
    tracer.trace()
    # all four deterministic source classes + a Lamp (host-generated: global np.random, so both traces
    # of one comparison must see the same stream -> reseed inside generate_rays)
    many = [pc.LineOfRays(0.4, wavelength=0.5).move_x(-1), pc.CircleOfRays(0.5).move_x(-1.5),
            pc.WedgeOfRays(12.0).move_x(-2.04), pc.ConeOfRays(4).move_x(-2.04).rotate_z(1.0)]
    tracer2 = pyrayt.RayTracer(many, [lens, stop, det], rays_per_source=257, generation_limit=50)
    tracer2.trace()
    assert len(log) == 3 and all(entry["rows"] > 0 for entry in log)


def test_large_live_trace_takes_the_general_path(live):
    """> 4096 rays: not the captured small-trace path but trace() + lean/full host transfer."""
    pyrayt, log = live
    pc = pyrayt.components
    lens = pc.thick_lens(60, -60, 4, aperture=25.4, material=pyrayt.materials.glass["BK7"])
    lens2 = pc.thick_lens(-80, 80, 2, aperture=25.4, material=pyrayt.materials.glass["SF5"]).move_x(10)
    det = pc.baffle((25.4, 25.4)).move_x(60)
    srcs = [pc.CircleOfRays(d).move_x(-10) for d in (4.0, 8.0, 12.0)]
    tracer = pyrayt.RayTracer(srcs, [lens, lens2, det], rays_per_source=4000, generation_limit=20)
    tracer.trace()
    assert log[0]["rays"] == 12000 and log[0]["rows"] >= 5 * 12000 - 10
