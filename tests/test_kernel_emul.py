"""The CUDA kernels' per-ray code (prt_device.cuh) executed on the host against the oracle.

The build container has no GPU; tests/emul compiles the very functions the kernels
call (PRT_HD) for the host so the streaming CSG merge, the preorder program with
bounding-box skips and the interaction step are checked here, bit for bit.  The
`-m gpu` suite repeats the comparison on the real device.
"""
import numpy as np
import pytest

from oracle import oracle
from tests import scene_util as su
from tests.emul import emul
from tests.helpers import GOLDEN_CASES, load_case


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_emulated_kernel_equals_oracle_on_golden(name):
    scene, rays, _, gl = load_case(name)
    want, octr = oracle.trace(scene, rays, gl)
    got, ectr = emul.trace(scene, rays, gl)
    assert np.array_equal(got, want, equal_nan=True)
    for k in ("generations", "segments", "limit_rays", "nan_rays", "untraceable_hits"):
        assert ectr[k] == octr[k], k


@pytest.mark.parametrize("seed", range(24))
def test_emulated_kernel_equals_oracle_on_random_scenes(seed):
    scene, rays = su.random_scene_and_rays(seed, n_rays=384)
    want, octr = oracle.trace(scene, rays, 16)
    got, ectr = emul.trace(scene, rays, 16)
    assert np.array_equal(got, want, equal_nan=True)
    assert ectr["generations"] == octr["generations"]


@pytest.mark.parametrize("seed", range(8))
def test_component_intersect_equals_oracle(seed):
    """component.intersect: every slot -- hits, +inf padding and the surface ids both carry -- as the
    reference returns them (the emulation also cross-checks the kernel's streaming merge against the
    literal lists on every finite slot)."""
    scene, rays = su.random_scene_and_rays(100 + seed, n_rays=256)
    r = np.zeros((8, rays.shape[1]))
    r[0:3], r[3], r[4:7] = rays[0:3], 1, rays[4:7]
    for c in range(scene.n_components):
        oh, osid = oracle.intersect(scene, c, r)
        eh, esid = emul.intersect(scene, c, r)
        assert np.array_equal(eh, oh)
        assert np.array_equal(esid, osid)


def test_closed_form_left_deep_merge_equals_streaming_merge():
    """All sorted pairs over {-inf,-2,-1,1,2,3,+inf} for A, B, C and all 9 operation pairs."""
    bad, cases = emul.selfcheck_left_deep()
    assert cases == 9 * 28 ** 3 and bad == 0


def test_random_scene_stress():
    """600 more random scenes (random / tight / shrunk boxes, non-uniform scales, every primitive and
    material): frames bit-equal to the oracle, every component.intersect slot equal."""
    for seed in range(1000, 1600):
        scene, rays = su.random_scene_and_rays(seed, n_rays=256)
        want, octr = oracle.trace(scene, rays, 16)
        got, ectr = emul.trace(scene, rays, 16)
        assert np.array_equal(got, want, equal_nan=True), seed
        assert ectr["generations"] == octr["generations"], seed
        if seed % 6 == 0:
            r = np.zeros((8, 64))
            r[0:3], r[3], r[4:7] = rays[0:3, :64], 1, rays[4:7, :64]
            for c in range(scene.n_components):
                oh, osid = oracle.intersect(scene, c, r)
                eh, esid = emul.intersect(scene, c, r)
                assert np.array_equal(eh, oh) and np.array_equal(esid, osid), (seed, c)
