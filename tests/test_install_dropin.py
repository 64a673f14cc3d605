"""install(): pyrayt.RayTracer.trace re-routed through pyrayt_b200 with LIVE reference objects.

The build container has the reference but no GPU, the GPU box has a GPU but no reference, so the
glue (flatten live objects -> rays -> engine -> pandas frame) is exercised here with an engine
stand-in that answers through the oracle; the CUDA engine itself is covered by the -m gpu suite.
"""
import numpy as np
import pytest

from oracle import oracle, ref_shim

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.available(), reason="PyRayT reference tree not present")]


class OracleEngine:
    """Same interface as pyrayt_b200.Engine.trace / update_scene, computed by the oracle (tests only)."""

    def __init__(self, scene, device=0):
        import torch

        self.scene, self.ray_device = scene, torch.device("cpu")

    def update_scene(self, scene):
        self.scene = scene

    def close(self):
        pass

    def trace(self, d_rays, generation_limit=10, ray_offset=1e-6, record="all", to_host=False, **kw):
        import torch

        from pyrayt_b200.engine import TraceResult

        frame, ctr = oracle.trace(self.scene, d_rays.numpy(), generation_limit, ray_offset)
        counters = dict(ctr, bad_w=0, rows_dropped=0, tie_rays=0)
        return TraceResult(torch.from_numpy(np.ascontiguousarray(frame)), frame.shape[1], counters, None, 0,
                           self.scene.n_leaves)


def test_install_reroutes_trace_and_matches_reference(monkeypatch):
    pyrayt = ref_shim.load()
    import pyrayt.components as pc

    import pyrayt_b200

    orig_trace = pyrayt.RayTracer.trace
    try:
        lens = pc.biconvex_lens(2, 2, 0.25, aperture=1)
        stop = pc.aperture((1, 1), 0.6).move_x(0.5)
        baffle = pc.baffle((1, 1)).move_x(1)
        sources = [pc.ConeOfRays(6).move_x(-2.04), pc.LineOfRays(0.3, wavelength=0.5).move_x(-1)]
        tracer = pyrayt.RayTracer(sources, [lens, stop, baffle], rays_per_source=50, generation_limit=100)
        with ref_shim.stable_argsort():
            want = tracer.trace().copy()
        pyrayt.RayTracer._b200_engine_factory = OracleEngine
        pyrayt.RayTracer._b200_device_sources = False  # no GPU here: the sources' own host code makes the rays
        pyrayt_b200.install()
        assert pyrayt.RayTracer.trace is not orig_trace
        tracer2 = pyrayt.RayTracer(sources, [lens, stop, baffle], rays_per_source=50, generation_limit=100)
        got = tracer2.trace()
        assert list(got.columns) == list(want.columns) and got.shape == want.shape == (300, 15)
        assert all(dt == np.float64 for dt in got.dtypes)
        for col in ("generation", "id", "surface"):
            assert np.array_equal(got[col].to_numpy(), want[col].to_numpy())
        np.testing.assert_allclose(got.to_numpy(), want.to_numpy(), rtol=1e-9, atol=1e-12)
        assert tracer2.get_results() is got
        # components are held by reference: move the detector, trace again
        baffle.move_x(0.5)
        with ref_shim.stable_argsort():
            pyrayt.RayTracer.trace = orig_trace
            want2 = pyrayt.RayTracer(sources, [lens, stop, baffle], rays_per_source=50, generation_limit=100).trace()
        pyrayt_b200.install()
        got2 = tracer2.trace()
        np.testing.assert_allclose(got2.to_numpy(), want2.to_numpy(), rtol=1e-9, atol=1e-12)
    finally:
        pyrayt.RayTracer.trace = orig_trace
        for attr in ("_b200_engine_factory", "_b200_device_sources"):
            if hasattr(pyrayt.RayTracer, attr):
                delattr(pyrayt.RayTracer, attr)
