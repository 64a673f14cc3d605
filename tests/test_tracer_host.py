"""Host logic of the drop-in RayTracer (pyrayt/_pyrayt.py:211-354) without a GPU: the engine is replaced
by a stand-in that answers through the oracle (tests only; the product always uses the CUDA engine)."""
import numpy as np
import pytest

from oracle import oracle
from tests import fakes, scene_util as su


class OracleEngine:
    """Engine.trace / update_scene with the oracle behind them (record modes included)."""

    def __init__(self, scene, device=0):
        import torch

        self.scene, self.ray_device, self.updates = scene, torch.device("cpu"), 0

    def update_scene(self, scene):
        self.scene, self.updates = scene, self.updates + 1

    def trace(self, d_rays, generation_limit=10, ray_offset=1e-6, record="all", detector_sid=-1, to_host=False, **kw):
        import torch

        from pyrayt_b200.engine import TraceResult

        frame, ctr = oracle.trace(self.scene, d_rays.numpy(), generation_limit, ray_offset)
        if record == "surface":
            frame = np.ascontiguousarray(frame[:, frame[5] == detector_sid])
        bad_w = int(np.sum((d_rays.numpy()[3] != 1.0) | (d_rays.numpy()[7] != 0.0)))
        counters = dict(ctr, bad_w=bad_w, rows_dropped=0, tie_rays=0)
        return TraceResult(torch.from_numpy(np.ascontiguousarray(frame)), frame.shape[1], counters, None, 0,
                           self.scene.n_leaves)


def _bench():
    glass = fakes.BasicRefractor(1.5)
    lens = fakes.CSG(fakes.Surface(fakes.Sphere(2.0), glass, su.translate(1.9, 0, 0)),
                     fakes.Surface(fakes.Sphere(2.0), glass, su.translate(-1.9, 0, 0)), 2, (-0.1, 0.1, -1, 1, -1, 1))
    det = fakes.Surface(fakes.Plane(4, 4), fakes._AbsorbingMaterial(), su.translate(3, 0, 0) @ su.rot_y(90))
    ys = np.linspace(-0.3, 0.3, 21)
    rays = su.make_rays(np.stack([np.full(21, -3.0), ys, np.zeros(21)], 1), np.tile([1.0, 0, 0], (21, 1)))
    return lens, det, rays


def _tracer(sources, components, **kw):
    import pyrayt_b200

    t = pyrayt_b200.RayTracer(sources, components, **kw)
    t._engine_factory = OracleEngine
    t.device_sources = False
    return t


def test_public_surface_and_frame_layout():
    import pandas as pd

    import pyrayt_b200

    lens, det, rays = _bench()
    tracer = _tracer(fakes.ArraySource(rays), [lens, det], rays_per_source=21, generation_limit=10)
    empty = tracer.get_results()  # before any trace: the reference's empty 0 x 15 float32 frame
    assert empty.shape == (0, 15) and all(dt == np.float32 for dt in empty.dtypes)
    df = tracer.trace()
    assert isinstance(df, pd.DataFrame) and list(df.columns) == list(pyrayt_b200.FRAME_COLUMNS)
    assert all(dt == np.float64 for dt in df.dtypes) and df.shape == (63, 15) and isinstance(df.index, pd.RangeIndex)
    want, _ = oracle.trace(pyrayt_b200.flatten([lens, det]), rays, 10)
    assert np.array_equal(df.to_numpy().T, want)
    assert tracer.get_results() is df
    tracer.calculate_source_ids()
    assert list(tracer.get_results()["source_id"].unique()) == [0]
    tracer.reset()
    assert tracer.get_results().shape == (0, 15)
    tracer.set_rays_per_source(5)
    tracer.set_generation_limit(1)
    assert tracer.get_rays_per_source() == 5 and tracer.get_generation_limit() == 1
    assert tracer.trace()["generation"].max() == 0


def test_scalars_are_wrapped_and_components_are_held_by_reference():
    lens, det, rays = _bench()
    tracer = _tracer(fakes.ArraySource(rays), det, rays_per_source=21)  # single component, single source
    assert tracer.trace().shape == (21, 15)
    tracer.load_components([lens, det])
    a = tracer.trace()
    det.move_x(1.0)  # the scene is re-flattened at every trace; the engine is updated in place, not rebuilt
    b = tracer.trace()
    assert np.allclose(a["x1"][a["generation"] == 2], 3.0) and np.allclose(b["x1"][b["generation"] == 2], 4.0)
    assert isinstance(tracer._engine, OracleEngine) and tracer._engine.updates >= 2


def test_several_sources_get_consecutive_ids_and_source_ids():
    lens, det, rays = _bench()
    s1, s2 = fakes.ArraySource(rays[:, :10].copy()), fakes.ArraySource(rays[:, 10:20].copy())
    tracer = _tracer([s1, s2], [lens, det], rays_per_source=10, generation_limit=10)
    df = tracer.trace()
    assert sorted(df["id"][df["generation"] == 0]) == list(range(20))
    tracer.calculate_source_ids()
    got = tracer.get_results()
    assert set(got["source_id"][got["id"] < 10]) == {0} and set(got["source_id"][got["id"] >= 10]) == {1}


def test_record_surface_extension_and_errors():
    import pyrayt_b200

    lens, det, rays = _bench()
    tracer = _tracer(fakes.ArraySource(rays), [lens, det], rays_per_source=21, generation_limit=10)
    full = tracer.trace()
    only = tracer.trace(record_surface=det)
    assert np.array_equal(only.to_numpy(), full[full["surface"] == det.get_id()].to_numpy())
    assert np.array_equal(tracer.trace(record_surface=det.get_id()).to_numpy(), only.to_numpy())
    bad = rays.copy()
    bad[3, 0] = 0.5  # homogeneous w of a position must be 1
    with pytest.raises(ValueError, match="homogeneous"):
        _tracer(fakes.ArraySource(bad), [lens, det], rays_per_source=21).trace()
    # a ray that ends on a surface without trace() raises like the reference (AttributeError)
    gooch = fakes.Surface(fakes.Sphere(1.0), fakes.Gooch(), su.translate(0, 0, 13))
    away = su.make_rays(np.array([[0.0, 0.0, 0.0]]), np.array([[0.0, 0.0, 1.0]]))
    with pytest.raises(AttributeError):
        _tracer(fakes.ArraySource(away), [gooch], rays_per_source=1).trace()
    assert issubclass(pyrayt_b200.UntraceableSurfaceError, AttributeError)


def test_no_rays_gives_the_empty_frame():
    lens, det, rays = _bench()
    miss = su.make_rays(np.array([[0.0, 50.0, 0.0]]), np.array([[0.0, 1.0, 0.0]]))  # hits nothing: no row at all
    df = _tracer(fakes.ArraySource(miss), [lens, det], rays_per_source=1).trace()
    assert df.shape == (0, 15) and all(dt == np.float32 for dt in df.dtypes)
