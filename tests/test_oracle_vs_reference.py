"""The oracle against the LIVE reference on fresh random inputs (build container only)."""
import numpy as np
import pytest

from oracle import oracle, ref_shim
from pyrayt_b200.scene import flatten

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.available(), reason="PyRayT reference tree not present")]


@pytest.fixture(scope="module")
def ref():
    pyrayt = ref_shim.load()
    import pyrayt.components as pc
    import pyrayt.materials as matl
    import tinygfx.g3d as cg

    return pyrayt, pc, matl, cg


def _reference_trace(ref, rays, comps, gl):
    pyrayt, pc, _, _ = ref

    class Fixed(pc.Source):
        def _local_ray_generation(self, n):
            rs = pyrayt.RaySet(rays.shape[1])
            rs[:] = rays
            return rs

    tr = pyrayt.RayTracer(Fixed(), comps)
    tr.set_rays_per_source(rays.shape[1])
    tr.set_generation_limit(gl)
    with ref_shim.stable_argsort(), np.errstate(all="ignore"):
        df = tr.trace()
    return df.to_numpy(dtype=np.float64).T if len(df) else np.zeros((15, 0))


def _random_rays(rng, n, span=6.0):
    r = np.zeros((13, n))
    r[0:3] = rng.uniform(-span, span, (3, n))
    v = rng.normal(size=(3, n))
    r[4:7] = v / np.linalg.norm(v, axis=0)
    r[3], r[9], r[11] = 1, 100, 1
    r[10] = rng.uniform(0.45, 0.75, n)
    r[12] = np.arange(n)
    return r


@pytest.mark.parametrize("seed", range(6))
def test_random_optical_bench(ref, seed):
    pyrayt, pc, matl, cg = ref
    rng = np.random.default_rng(seed)
    glasses = [matl.glass["ideal"], matl.glass["BK7"], matl.glass["SF5"], matl.glass["SF2"]]
    comps = []
    for k in range(4):
        g = glasses[int(rng.integers(0, 4))]
        kind = int(rng.integers(0, 6))
        x = -4 + 2.5 * k
        if kind == 0:
            c = pc.thick_lens(rng.uniform(3, 8), -rng.uniform(3, 8), rng.uniform(0.3, 0.8), aperture=2.0, material=g)
        elif kind == 1:
            c = pc.thick_lens(-rng.uniform(3, 8), rng.uniform(3, 8), rng.uniform(0.2, 0.5), aperture=2.0, material=g)
        elif kind == 2:
            c = pc.biconvex_lens(rng.uniform(3, 6), rng.uniform(3, 6), 0.5, aperture=2.0, material=g)
        elif kind == 3:
            c = pc.equilateral_prism(1.5, 1.5, material=g)
        elif kind == 4:
            c = pc.spherical_mirror(rng.uniform(5, 12), 0.5, aperture=2.0)
        else:
            c = pc.parabolic_mirror(rng.uniform(2, 5), 0.5, aperture=2.0)
        comps.append(c.move_x(x).rotate_z(rng.uniform(-10, 10)))
    comps.append(pc.aperture((3, 3), 1.2).move_x(5.5))
    comps += [pc.baffle((14, 14)).move_x(7), pc.baffle((14, 14)).move_x(-7)]
    rays = _random_rays(rng, 600)
    want = _reference_trace(ref, rays, comps, 12)
    got, ctr = oracle.trace(flatten(comps), rays, 12)
    assert got.shape == want.shape
    assert np.array_equal(got[[0, 4, 5]], want[[0, 4, 5]])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12, equal_nan=True)


def test_component_intersect_interface(ref):
    """component.intersect(rays) -> (hits, surface ids) incl. the ids of +inf slots and culled rays."""
    pyrayt, pc, matl, cg = ref
    rng = np.random.default_rng(3)
    comps = [pc.thick_lens(-6, 6, 0.4, aperture=2.0), pc.equilateral_prism(1, 1).move_x(2),
             pc.aperture((2, 2), 0.8).move_x(-2), pc.baffle((3, 3)).move_x(4)]
    rays13 = _random_rays(rng, 800, span=3.0)
    rays = np.zeros((2, 4, 800))
    rays[0, :3], rays[0, 3], rays[1, :3] = rays13[0:3], 1, rays13[4:7]
    scene = flatten(comps)
    for c, comp in enumerate(comps):
        with ref_shim.stable_argsort(), np.errstate(all="ignore"):
            rh, rs = comp.intersect(rays)
        oh, osid = oracle.intersect(scene, c, rays)
        np.testing.assert_allclose(oh, rh, rtol=1e-9, atol=1e-12)
        assert np.array_equal(osid, rs)
