"""Renderers' nearest-hit loop: oracle against frames of the unmodified reference
renderers on CPU, the CUDA kernel against both on the GPU."""
import numpy as np
import pytest

from tests.helpers import RENDER_CASES, load_render_case


def _assert_hits_match(t, sid, want_t, want_sid, what):
    assert np.array_equal(sid, want_sid), f"{what}: surfaces differ"
    fin = np.isfinite(want_t)
    assert np.array_equal(np.isfinite(t), fin), what
    err = np.abs(t[fin] - want_t[fin]) / np.maximum(1.0, np.abs(want_t[fin]))
    assert np.max(err, initial=0.0) <= 1e-9, f"{what}: distance error {np.max(err):.3e}"  # BASELINE.json tolerance


def test_render_fixtures_exist():
    assert len(RENDER_CASES) >= 3


@pytest.mark.parametrize("name", RENDER_CASES)
def test_oracle_render_hit_matches_reference_renderer(name):
    from oracle import oracle

    scene, rays, dist, surf, canvas, (h, v) = load_render_case(name)
    t, sid, _ = oracle.render_hit(scene, rays)
    _assert_hits_match(t, sid, dist, surf, name)
    if name == "render_inside_view":
        # components behind the camera: the renderers keep their (negative) first hit, the tracer does not
        assert np.sum(dist < 0) > 1000
        t2, sid2, _ = oracle.nearest(scene, rays)
        assert np.all(t2[dist < 0] > 0) and np.array_equal(sid2[dist > 0], surf[dist > 0])


@pytest.mark.reference
def test_install_swaps_propagate_of_the_reference_renderers(monkeypatch):
    """Glue test with live reference objects: the oracle stands in for the kernel (no GPU here)."""
    from oracle import oracle, ref_shim

    if not ref_shim.available():
        pytest.skip("PyRayT reference tree not present")
    ref_shim.load()
    import pyrayt.components as pc
    from tinygfx.g3d import renderers
    from tinygfx.g3d.world_objects import OrthographicCamera

    from pyrayt_b200 import render
    from pyrayt_b200.scene import flatten

    comps = [pc.biconvex_lens(2, 2, 0.25, aperture=1), pc.baffle((1, 1)).move_x(1)]
    cam = OrthographicCamera(64, 3.0, 0.75)
    cam.move(0.4, 0.1, 0.0)  # inside the system: negative distances occur
    with ref_shim.stable_argsort(), np.errstate(all="ignore"):
        want_edge = renderers.EdgeRender(cam, comps).render()
        want_shaded = renderers.ShadedRenderer(cam, comps, light_position=np.array([3.0, 3.0, 9.0, 1.0])).render()

    def fake_propagate(rays, components, device=0, normals=False, renderer=True, engine=None):
        t, sid, nrm = oracle.render_hit(flatten(components), np.asarray(rays))
        return t, sid, nrm

    monkeypatch.setattr(render, "_propagate", fake_propagate)
    orig = renderers.EdgeRender._st_propagate, renderers.ShadedRenderer._st_propagate
    try:
        render.install()
        assert renderers.EdgeRender._st_propagate is not orig[0]
        with np.errstate(all="ignore"):
            got_edge = renderers.EdgeRender(cam, comps).render()
            got_shaded = renderers.ShadedRenderer(cam, comps, light_position=np.array([3.0, 3.0, 9.0, 1.0])).render()
    finally:
        renderers.EdgeRender._st_propagate, renderers.ShadedRenderer._st_propagate = orig
    assert np.array_equal(got_edge, want_edge)
    np.testing.assert_allclose(got_shaded, want_shaded, rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_hit_kernel_matches_reference_and_oracle(name, cuda_device):
    import torch

    import pyrayt_b200
    from oracle import oracle

    scene, rays, dist, surf, canvas, (h, v) = load_render_case(name)
    eng = pyrayt_b200.Engine(scene, device=0)
    d = torch.from_numpy(np.ascontiguousarray(rays)).cuda()
    t, sid, nrm = eng.nearest_hit(d, normals=True, renderer=True)
    t, sid, nrm = t.cpu().numpy(), sid.cpu().numpy(), nrm.cpu().numpy()
    _assert_hits_match(t, sid, dist, surf, name)
    ot, osid, onrm = oracle.render_hit(scene, rays)
    assert np.array_equal(t, ot) and np.array_equal(sid, osid)  # same roundings as the oracle: same bits
    assert np.array_equal(nrm, onrm, equal_nan=True)
    # the tracer's variant on the same rays
    t2, sid2, _ = eng.nearest_hit(d, renderer=False)
    ot2, osid2, _ = oracle.nearest(scene, rays)
    assert np.array_equal(t2.cpu().numpy(), ot2) and np.array_equal(sid2.cpu().numpy(), osid2)


@pytest.mark.gpu
def test_camera_nearest_on_fakes(cuda_device):
    """camera_nearest with duck-typed scene objects (no reference needed): both hit rules, image layout."""
    import pyrayt_b200
    from oracle import oracle
    from tests import fakes, scene_util as su

    glass = fakes.BasicRefractor(1.5)
    lens = fakes.CSG(fakes.Surface(fakes.Sphere(2.0), glass, su.translate(1.9, 0, 0)),
                     fakes.Surface(fakes.Sphere(2.0), glass, su.translate(-1.9, 0, 0)), 2, (-0.1, 0.1, -1, 1, -1, 1))
    ball = fakes.Surface(fakes.Sphere(0.3), fakes._ReflectingMaterial(), su.translate(1.0, 0.5, 0.2))
    behind = fakes.Surface(fakes.Sphere(0.4), fakes._ReflectingMaterial(), su.translate(-7.0, 0.0, 0.0))
    cam = fakes.OrthographicCamera(96, 2.4, 0.75, world=su.translate(-5, 0, 0))
    comps = [lens, ball, behind]
    img = pyrayt_b200.render.camera_nearest(cam, comps, renderer=True)
    t, sid, nrm = oracle.render_hit(pyrayt_b200.flatten(comps), cam.generate_rays())
    assert img["distance"].shape == img["surface"].shape == (72, 96) and img["normal"].shape == (72, 96, 3)
    assert np.array_equal(img["distance"].ravel(), t) and np.array_equal(img["surface"].ravel(), sid)
    assert np.sum(t < 0) > 0 and behind.get_id() in set(sid.tolist())
    front = pyrayt_b200.render.camera_nearest(cam, comps, renderer=False, normals=False)
    t2, sid2, _ = oracle.nearest(pyrayt_b200.flatten(comps), cam.generate_rays())
    assert np.array_equal(front["distance"].ravel(), t2) and np.array_equal(front["surface"].ravel(), sid2)
    assert behind.get_id() not in set(sid2.tolist())
