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


@pytest.mark.parametrize("traversal", ["ordered", "list"])
def test_both_traversals_of_the_nearest_hit_search(traversal, monkeypatch):
    """nearest_hit walks the boxed components in ray order (large scenes) or in list order (small ones);
    the encoder picks by scene size.  Both are forced here on every golden case and on random scenes
    (unboxed components, generic trees, ties) and must give the oracle's frame bit for bit."""
    monkeypatch.setenv("PRT_EMUL_TRAVERSAL", traversal)
    for name in GOLDEN_CASES:
        scene, rays, _, gl = load_case(name)
        want, _ = oracle.trace(scene, rays, gl)
        got, _ = emul.trace(scene, rays, gl)
        assert np.array_equal(got, want, equal_nan=True), name
    for seed in range(2000, 2120):
        scene, rays = su.random_scene_and_rays(seed, n_rays=128)
        want, _ = oracle.trace(scene, rays, 12)
        got, _ = emul.trace(scene, rays, 12)
        assert np.array_equal(got, want, equal_nan=True), seed


def test_closed_form_left_deep_merge_equals_streaming_merge():
    """All sorted pairs over {-inf,-2,-1,1,2,3,+inf} for A, B, C and all 9 operation pairs, plus the
    missed-inner-box case for every C."""
    bad, cases = emul.selfcheck_left_deep()
    assert cases == 9 * (28 ** 3 + 28) and bad == 0


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


def test_equal_distances_go_to_the_earlier_component_in_any_visiting_order():
    """_st_propagate keeps the first component among equal distances (strict `<`, pyrayt/_pyrayt.py:384).
    The kernel visits boxed components in the order the ray meets their boxes, so the rule is restated as
    (smallest distance, then lowest component index): coincident surfaces in both list orders, with other
    components before, between and behind them, hit from both sides and obliquely."""
    def plane(x, sid, mat=su.MAT_ABSORBER):
        return su.Leaf(su.PLANE, [4.0, 4.0], mat=mat, world=su.translate(x, 0, 0) @ su.rot_y(90), sid=sid)

    def lens(x, sid):
        a = su.Leaf(su.SPHERE, [2.0], mat=su.MAT_GLASS_CONST, matp=[1.5], world=su.translate(x + 1.9, 0, 0), sid=sid)
        b = su.Leaf(su.SPHERE, [2.0], mat=su.MAT_GLASS_CONST, matp=[1.5], world=su.translate(x - 1.9, 0, 0), sid=sid + 1)
        return su.intersect(a, b, aabb=(x - 0.1, x + 0.1, -0.7, 0.7, -0.7, 0.7))

    origins = [[-3, 0.1, 0.05], [-3, -0.2, 0.1], [6, 0.1, 0.0], [6, 0.3, -0.2], [-3, 0.0, 0.0]]
    dirs = [[1, 0, 0], [0.999, 0.02, 0.01], [-1, 0, 0], [-0.99, -0.05, 0.03], [1, 0, 0]]
    dirs = [np.asarray(d) / np.linalg.norm(d) for d in dirs]
    rays = su.make_rays(origins, dirs)
    for first, second in ((31, 32), (32, 31)):
        comps = [plane(5.0, 40), plane(2.0, first), lens(0.0, 50), plane(2.0, second), plane(-4.0, 44, su.MAT_MIRROR)]
        scene = su.build(comps)
        want, _ = oracle.trace(scene, rays, 6)
        got, _ = emul.trace(scene, rays, 6)
        assert np.array_equal(got, want, equal_nan=True)
        ends = want[5][want[0] == want[0].max()]
        assert first in set(want[5]) and second not in set(want[5]), (first, second, ends)


def test_diagnose_counters_equal_the_oracle():
    """PRT_FLAG_DIAGNOSE (rays within 1e-9 of grazing or CSG seams, counted by displacing the origin): the
    kernel's per-ray code and the oracle agree ray by ray on a crafted case and in total on random scenes."""
    scene, rays, expected = su.grazing_and_seam_case()
    for i, (gz, sm) in enumerate(expected):
        one = np.ascontiguousarray(rays[:, i:i + 1])
        o = oracle.diagnose(scene, one, 4)
        _, e = emul.trace(scene, one, 4, diagnose=True)
        assert (o["grazing_rays"], o["seam_rays"]) == (gz, sm), i
        assert (e["grazing_rays"], e["seam_rays"]) == (gz, sm), i
    for seed in range(3000, 3012):
        scene, rays = su.random_scene_and_rays(seed, n_rays=256)
        o = oracle.diagnose(scene, rays, 12)
        _, e = emul.trace(scene, rays, 12, diagnose=True)
        assert (o["grazing_rays"], o["seam_rays"]) == (e["grazing_rays"], e["seam_rays"]), seed
        assert o["generations"] == e["generations"]


FP32_CASES = list(GOLDEN_CASES)


@pytest.mark.parametrize("name", FP32_CASES)
def test_fp32_fast_mode_code_within_its_tolerance(name):
    """The FP32 fast mode's per-ray code (prt_device_f32.cuh, run on the host) against the FP64 oracle: at most
    0.5 % of the rays of these edge-case-heavy sets take a different path, and on all others every id column is
    equal, positions agree to 1e-5 of the scene scale, unit tilts and the index to 1e-5."""
    import torch

    from pyrayt_b200 import compare

    scene, rays, _, gl = load_case(name)
    want, _ = oracle.trace(scene, rays, gl)
    got = emul.trace_f32(scene, rays, gl)
    assert got is not None
    first = int(rays[12].min())
    assert np.array_equal(np.sort(rays[12]), np.arange(first, first + rays.shape[1]))
    rep = compare.frame_agreement(torch.from_numpy(want), torch.from_numpy(got), first, rays.shape[1])
    # (nested_csg: random rays through glass balls; refraction near the critical angle amplifies any rounding,
    #  1 % of its rays leave an interaction more than 1e-5 off)
    allowed = 0.02 if name == "nested_csg" else 0.005
    assert rep["rays_with_different_ids"] + rep["rays_beyond_tolerance"] <= max(1, int(allowed * rays.shape[1])), rep
    assert rep["id_columns_equal_on_compared_rows"], rep
    assert rep["max_error_on_agreeing_rays"] <= 1e-5, rep


@pytest.mark.parametrize("seed", range(4000, 4016))
def test_fp32_fast_mode_on_random_scenes_with_generic_trees(seed):
    """Random scenes (right-nested trees run the single-precision interpreter; overlapping solids, scaled poses,
    boxes that are not bounds): rays off their FP64 path stay a small minority and every other row is within
    the tolerance."""
    import torch

    from pyrayt_b200 import compare

    scene, rays = su.random_scene_and_rays(seed, n_rays=384)
    want, _ = oracle.trace(scene, rays, 12)
    got = emul.trace_f32(scene, rays, 12)
    rep = compare.frame_agreement(torch.from_numpy(want), torch.from_numpy(got), 0, rays.shape[1])
    assert rep["rays_with_different_ids"] + rep.get("rays_beyond_tolerance", 0) <= 0.03 * rays.shape[1], rep
    assert rep["id_columns_equal_on_compared_rows"], rep


def test_lenslet_array_of_973_leaves():
    """A scene far beyond the reference's examples: 18 x 18 lenslets (972 leaves in 324 left-deep components)
    and a detector.  16-bit leaf indices, the ray-ordered traversal with bisection over 324 boxes."""
    scene, centres = su.lenslet_array(18, 18)
    assert scene.n_leaves == 973 and scene.n_components == 325
    rays = su.lenslet_rays(centres, 2)
    want, octr = oracle.trace(scene, rays, 8, threads=4)
    got, ectr = emul.trace(scene, rays, 8)
    assert want.shape[1] > 2 * rays.shape[1]
    assert np.array_equal(got, want, equal_nan=True)
    assert ectr["generations"] == octr["generations"]


def test_fp32_ordered_and_list_walks_agree(monkeypatch):
    """The FP32 mode's ray-ordered traversal is a reordering of the same per-component evaluations: its frame
    equals the list-order walk's bit for bit (golden cases and the 973-leaf lenslet array)."""
    cases = [load_case(n)[:2] + (load_case(n)[3],) for n in FP32_CASES]
    scene, centres = su.lenslet_array(18, 18)
    cases.append((scene, su.lenslet_rays(centres, 1), 8))
    for scene, rays, gl in cases:
        monkeypatch.delenv("PRT_EMUL_F32_LIST", raising=False)
        monkeypatch.setenv("PRT_EMUL_F32_ORDERED", "1")
        ordered = emul.trace_f32(scene, rays, gl)
        monkeypatch.delenv("PRT_EMUL_F32_ORDERED", raising=False)
        monkeypatch.setenv("PRT_EMUL_F32_LIST", "1")
        listed = emul.trace_f32(scene, rays, gl)
        assert np.array_equal(ordered, listed, equal_nan=True)
