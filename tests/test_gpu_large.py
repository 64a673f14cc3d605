"""GPU parity at BASELINE sizes: oracle comparison where the CPU finishes in seconds,
size-independent properties at the full ray counts."""
import os

import numpy as np
import pytest

from tests.helpers import assert_frames_match, load_case

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


@pytest.fixture(scope="module")
def torch_mod(cuda_device):
    import torch

    return torch


def _engine(name):
    import pyrayt_b200
    from pyrayt_b200 import workloads

    wl = workloads.WORKLOADS[name]
    return wl, pyrayt_b200.Engine(wl.scene(), device=0)


@pytest.mark.parametrize("name,n", [("config2", 100_000), ("config4", 1 << 17), ("config5", 1 << 16)])
def test_workload_matches_oracle_bit_for_bit(name, n, torch_mod):
    from oracle import oracle, sources_np

    wl, eng = _engine(name)
    d_rays = wl.source.generate(n, device=0)
    h_rays = sources_np.from_source(wl.source, n)
    assert np.array_equal(d_rays.cpu().numpy(), h_rays), "device source differs from its NumPy restatement"
    res = eng.trace(d_rays, generation_limit=wl.generation_limit, to_host=True)
    want, octr = oracle.trace(wl.scene(), h_rays, wl.generation_limit, threads=THREADS)
    got = res.frame.numpy()
    assert_frames_match(got, want, what=name)
    assert np.array_equal(got, want, equal_nan=True)
    assert res.counters["generations"] == octr["generations"]
    assert res.ray_surface_tests == octr["generations"] * wl.scene().n_leaves


def test_config3_prism_one_million_rays(torch_mod):
    """examples/chromatic_dispersion.py geometry, 11 wavelengths x 95,326 rays (= 1,048,586)."""
    import pyrayt_b200
    from oracle import oracle

    scene, rays_small, _, gl = load_case("config3_prism")
    per = 95_326
    lam = np.linspace(0.44, 0.75, 11)
    # LineOfRays(spacing=0.1).move_x(-0.5).rotate_y(-3): rebuild from the golden rays' two end points per source
    blocks = []
    for k in range(11):
        src = rays_small[:, k * 96:(k + 1) * 96]
        t = np.linspace(0.0, 1.0, per)
        blk = np.repeat(src[:, :1], per, axis=1)
        blk[0:3] = src[0:3, :1] + (src[0:3, -1:] - src[0:3, :1]) * t
        blk[10] = lam[k]
        blocks.append(blk)
    rays = np.ascontiguousarray(np.hstack(blocks))
    rays[12] = np.arange(rays.shape[1])
    assert rays.shape[1] == 1_048_586
    eng = pyrayt_b200.Engine(scene, device=0)
    res = eng.trace(torch_mod.from_numpy(rays).cuda(), generation_limit=gl, to_host=True)
    want, _ = oracle.trace(scene, rays, gl, threads=THREADS)
    assert np.array_equal(res.frame.numpy(), want, equal_nan=True)
    # dispersion: shorter wavelengths leave the prism at a steeper angle
    last = res.frame.numpy()[:, res.frame.numpy()[0] == 2]
    tilt = [last[14, last[2] == w].mean() for w in lam]
    assert np.all(np.diff(tilt) > 0) or np.all(np.diff(tilt) < 0)


def test_config4_full_size_properties(torch_mod):
    """2^24 rays: row order, monotone survival, per-ray generation prefixes, strided oracle sample."""
    from oracle import oracle, sources_np

    torch = torch_mod
    wl, eng = _engine("config4")
    n = wl.n_rays
    d_rays = wl.source.generate(n, device=0)
    res = eng.trace(d_rays, generation_limit=wl.generation_limit)
    f = res.frame
    assert res.rows == f.shape[1] == res.counters["segments"] and res.counters["rows_dropped"] == 0
    gen, rid = f[0], f[4]
    # rows ordered by (generation, id): the key must be strictly increasing
    key = gen * float(1 << 26) + rid
    assert bool((key[1:] > key[:-1]).all())
    # survivors only shrink, and generation g rows are exactly gen_counts[g]
    gc = res.gen_counts
    assert np.all(np.diff(gc) <= 0) and gc.sum() == res.rows
    cnt = torch.bincount(gen.to(torch.int64), minlength=len(gc)).cpu().numpy()
    assert np.array_equal(cnt[: len(gc)], gc)
    # every ray's rows are generations 0..k-1 with no gaps: sum of generations == k(k-1)/2 per ray
    k = torch.bincount(rid.to(torch.int64), minlength=n)
    gs = torch.zeros(n, dtype=torch.float64, device=f.device).index_add_(0, rid.to(torch.int64), gen)
    assert bool((gs == (k * (k - 1) / 2).to(torch.float64)).all())
    # strided sample of 2048 rays against the oracle, bit for bit
    idx = np.arange(0, n, n // 2048)
    h = np.hstack([sources_np.from_source(wl.source, 1, first_index=int(i)) for i in idx])
    want, _ = oracle.trace(wl.scene(), h, wl.generation_limit, threads=THREADS)
    sel = torch.isin(rid, torch.from_numpy(idx.astype(np.float64)).to(f.device))
    got = f[:, sel].cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)
    del f, res
    eng.release_workspace()
    torch.cuda.empty_cache()


def test_config5_full_size_properties_wavefront(torch_mod):
    """BASELINE config 5 at its per-GPU size (2^25 rays, paraboloid + cuboid mirrors + TIR pipe), traced by
    the wavefront driver (rows written in place): order, per-ray generation prefixes, strided oracle sample,
    and the same frame as the single-kernel driver on the first 2^20 rays."""
    from oracle import oracle, sources_np

    torch = torch_mod
    wl, eng = _engine("config5")
    n = wl.n_rays
    d_rays = wl.source.generate(n, device=0)
    res = eng.trace_wavefront(d_rays, generation_limit=wl.generation_limit)
    f = res.frame
    assert res.rows == f.shape[1] == res.counters["segments"] and res.counters["rows_dropped"] == 0
    assert res.rows / n >= 8.0  # the scene must average >= 8 segments per ray
    gen, rid = f[0], f[4]
    for lo in range(0, res.rows - 1, 1 << 27):  # (chunked: the key tensor of 4 x 10^8 rows is 3 GB)
        hi = min(res.rows, lo + (1 << 27) + 1)
        key = gen[lo:hi] * float(1 << 26) + rid[lo:hi]
        assert bool((key[1:] > key[:-1]).all())
        del key
    gc = res.gen_counts
    assert np.all(np.diff(gc) <= 0) and gc.sum() == res.rows
    k = torch.bincount(rid.to(torch.int64), minlength=n)
    gs = torch.zeros(n, dtype=torch.float64, device=f.device).index_add_(0, rid.to(torch.int64), gen)
    assert bool((gs == (k * (k - 1) / 2).to(torch.float64)).all())
    del k, gs
    idx = np.arange(0, n, n // 1024)
    h = np.hstack([sources_np.from_source(wl.source, 1, first_index=int(i)) for i in idx])
    want, _ = oracle.trace(wl.scene(), h, wl.generation_limit, threads=THREADS)
    sel = torch.isin(rid, torch.from_numpy(idx.astype(np.float64)).to(f.device))
    assert np.array_equal(f[:, sel].cpu().numpy(), want, equal_nan=True)
    del sel, f, res
    torch.cuda.empty_cache()
    m = 1 << 20
    a = eng.trace(d_rays[:, :m].contiguous(), generation_limit=wl.generation_limit, method="single")
    b = eng.trace(d_rays[:, :m].contiguous(), generation_limit=wl.generation_limit, method="wavefront")
    assert a.rows == b.rows and bool((a.frame == b.frame).all())
    assert all(a.counters[key] == b.counters[key] for key in ("generations", "segments", "mirror_segments",
                                                              "absorber_segments", "limit_rays", "tie_rays"))
    del a, b
    eng.release_workspace()
    torch.cuda.empty_cache()


def test_sharded_ranges_reassemble_the_monolithic_frame(torch_mod):
    """tracing ray-index ranges separately and placing the blocks by the all-gathered
    counts gives the single-GPU frame bit for bit."""
    from pyrayt_b200 import dist as pdist

    wl, eng = _engine("config4")
    n = 1 << 18
    d_rays = wl.source.generate(n, device=0)
    whole = eng.trace(d_rays, generation_limit=wl.generation_limit, to_host=True)
    parts, counts = [], []
    for r in range(4):
        b, e = pdist.shard_range(n, r, 4)
        shard = wl.source.generate(e - b, device=0, first_index=b)
        assert torch_mod.equal(shard, d_rays[:, b:e])
        res = eng.trace(shard, generation_limit=wl.generation_limit, to_host=True)
        parts.append(res.frame.numpy())
        counts.append(res.gen_counts)
    glob = pdist.assemble_global_frame(parts, np.stack(counts))
    assert np.array_equal(glob, whole.frame.numpy(), equal_nan=True)


def test_record_modes_overflow_retry_and_zero_copy(torch_mod):
    wl, eng = _engine("config4")
    n = 1 << 16
    d_rays = wl.source.generate(n, device=0)
    G = wl.generation_limit
    full = eng.trace(d_rays, generation_limit=G, to_host=True)
    f = full.frame.numpy()
    det = int(wl.scene().leaf_sid[-1])
    only = eng.trace(d_rays, generation_limit=G, record="surface", detector_sid=det, to_host=True)
    assert np.array_equal(only.frame.numpy(), f[:, f[5] == det])
    none = eng.trace(d_rays, generation_limit=G, record="none")
    for k in ("rays", "generations", "segments", "limit_rays", "absorber_segments"):
        assert none.counters[k] == full.counters[k]
    tiny = eng.trace(d_rays, generation_limit=G, capacity=1000, to_host=True)  # forces the overflow retry
    assert np.array_equal(tiny.frame.numpy(), f)
    zc = eng.trace(d_rays, generation_limit=G, to_host=True, zero_copy=True)
    assert np.array_equal(zc.frame.numpy(), f)


def test_component_intersect_matches_oracle(torch_mod):
    import pyrayt_b200
    from oracle import oracle
    from tests import scene_util as su

    for seed in range(4):
        scene, rays = su.random_scene_and_rays(200 + seed, n_rays=1000)
        eng = pyrayt_b200.Engine(scene, device=0)
        r = np.zeros((8, rays.shape[1]))
        r[0:3], r[3], r[4:7] = rays[0:3], 1, rays[4:7]
        d = torch_mod.from_numpy(r).cuda()
        for c in range(scene.n_components):
            h, s = eng.intersect(c, d)
            oh, osid = oracle.intersect(scene, c, r)
            assert np.array_equal(h.cpu().numpy(), oh)
            assert np.array_equal(s.cpu().numpy(), osid)  # the ids the +inf slots carry included


def test_random_scenes_match_oracle(torch_mod):
    import pyrayt_b200
    from oracle import oracle
    from tests import scene_util as su

    for seed in range(40):
        scene, rays = su.random_scene_and_rays(seed, n_rays=4096)
        eng = pyrayt_b200.Engine(scene, device=0)
        res = eng.trace(torch_mod.from_numpy(rays).cuda(), generation_limit=16, to_host=True)
        want, _ = oracle.trace(scene, rays, 16, threads=THREADS)
        assert np.array_equal(res.frame.numpy(), want, equal_nan=True), seed


def test_drop_in_raytracer_api(torch_mod):
    """pyrayt.RayTracer's public surface (pyrayt/_pyrayt.py:211-354) on duck-typed components."""
    import pandas as pd

    import pyrayt_b200
    from oracle import oracle
    from tests import fakes, scene_util as su

    glass = fakes.BasicRefractor(1.5)
    lens = fakes.CSG(fakes.Surface(fakes.Sphere(2.0), glass, su.translate(1.9, 0, 0)),
                     fakes.Surface(fakes.Sphere(2.0), glass, su.translate(-1.9, 0, 0)), 2, (-0.1, 0.1, -1, 1, -1, 1))
    det = fakes.Surface(fakes.Plane(4, 4), fakes._AbsorbingMaterial(), su.translate(3, 0, 0) @ su.rot_y(90))
    ys = np.linspace(-0.3, 0.3, 21)
    rays = su.make_rays(np.stack([np.full(21, -3.0), ys, np.zeros(21)], 1), np.tile([1.0, 0, 0], (21, 1)))
    tracer = pyrayt_b200.RayTracer(fakes.ArraySource(rays), [lens, det], rays_per_source=21, generation_limit=10)
    df = tracer.trace()
    assert isinstance(df, pd.DataFrame) and list(df.columns) == list(pyrayt_b200.FRAME_COLUMNS)
    assert all(dt == np.float64 for dt in df.dtypes) and df.shape == (63, 15)
    assert isinstance(df.index, pd.RangeIndex)
    want, _ = oracle.trace(pyrayt_b200.flatten([lens, det]), rays, 10)
    assert np.array_equal(df.to_numpy().T, want)
    assert tracer.get_results() is df
    tracer.calculate_source_ids()
    assert "source_id" in tracer.get_results().columns
    # components are held by reference: moving one changes the next trace
    det.move_x(1.0)
    df2 = tracer.trace()
    assert np.allclose(df2["x1"][df2["generation"] == 2], 4.0)
    # extension: record only the rows that end on one surface
    only = tracer.trace(record_surface=det)
    full = tracer.trace()
    assert np.array_equal(only.to_numpy(), full[full["surface"] == det.get_id()].to_numpy())
    # setters / getters
    tracer.set_rays_per_source(5)
    tracer.set_generation_limit(1)
    assert tracer.get_rays_per_source() == 5 and tracer.get_generation_limit() == 1
    assert tracer.trace().shape == (5, 15)
    # extensions of the drop-in: the optional FP32 fast mode and the diagnose counters
    tracer.set_rays_per_source(21)
    tracer.set_generation_limit(10)
    ref = tracer.trace().to_numpy()
    tracer.precision = "fp32"
    fast = tracer.trace().to_numpy()
    assert fast.shape == ref.shape and np.array_equal(fast[:, [0, 4, 5]], ref[:, [0, 4, 5]])
    assert np.allclose(fast, ref, rtol=0, atol=1e-5 * 4.0)
    tracer.precision = "fp64"
    tracer.diagnose = True
    assert np.array_equal(tracer.trace().to_numpy(), ref)
    assert tracer.last_result.counters["grazing_rays"] == 0 and tracer.last_result.counters["seam_rays"] == 0
    tracer.diagnose = False
    tracer.set_rays_per_source(5)
    tracer.set_generation_limit(1)
    # a trace with no hit returns the reference's empty float32 frame
    away = su.make_rays([[0, 0, 10.0]], [[0, 0, 1.0]])
    empty = pyrayt_b200.RayTracer(fakes.ArraySource(away), [det], rays_per_source=1).trace()
    assert empty.shape == (0, 15) and all(dt == np.float32 for dt in empty.dtypes)
    # hitting a surface without a traceable material raises like the reference (AttributeError)
    bad = fakes.Surface(fakes.Sphere(1.0), fakes.Gooch(), su.translate(0, 0, 13))
    with pytest.raises(AttributeError):
        pyrayt_b200.RayTracer(fakes.ArraySource(away), [bad], rays_per_source=1).trace()


def test_nearest_hit_and_scene_update(torch_mod):
    """prt_nearest_hit (= _st_propagate, the renderers' per-pixel loop) and prt_scene_update."""
    import pyrayt_b200
    from oracle import oracle
    from tests import scene_util as su

    scene, rays13, _, _ = load_case("thick_lens_zoo")
    eng = pyrayt_b200.Engine(scene, device=0)
    rng = np.random.default_rng(5)
    n = 20000
    r = np.zeros((8, n))
    r[0:3] = rng.uniform(-2, 16, (3, n)) * np.array([[1.0], [0.15], [0.15]])
    d = rng.normal(size=(3, n)) * np.array([[1.0], [0.3], [0.3]])
    r[3], r[4:7] = 1, d / np.linalg.norm(d, axis=0)
    t, sid, nrm = eng.nearest_hit(torch_mod.from_numpy(r).cuda(), normals=True)
    ot, osid, onrm = oracle.nearest(scene, r)
    assert np.array_equal(t.cpu().numpy(), ot) and np.array_equal(sid.cpu().numpy(), osid)
    assert np.array_equal(nrm.cpu().numpy(), onrm, equal_nan=True)
    assert (osid >= 0).mean() > 0.3
    # update in place: same topology, different matrices -> same results as a fresh engine
    for seed in (1, 2):
        s2, rays = su.random_scene_and_rays(seed, n_rays=2048)
        eng.update_scene(s2)
        got = eng.trace(torch_mod.from_numpy(rays).cuda(), generation_limit=12, to_host=True).frame.numpy()
        want, _ = oracle.trace(s2, rays, 12)
        assert np.array_equal(got, want, equal_nan=True)


def test_camera_nearest_image_matches_oracle(torch_mod):
    """N3: the renderers' per-pixel nearest-hit loop (tinygfx/g3d/renderers.py:72-94) as one kernel."""
    import pyrayt_b200
    from oracle import oracle
    from tests import fakes, scene_util as su

    glass = fakes.BasicRefractor(1.5)
    lens = fakes.CSG(fakes.Surface(fakes.Sphere(2.0), glass, su.translate(1.9, 0, 0)),
                     fakes.Surface(fakes.Sphere(2.0), glass, su.translate(-1.9, 0, 0)), 2, (-0.1, 0.1, -1, 1, -1, 1))
    ball = fakes.Surface(fakes.Sphere(0.3), fakes._ReflectingMaterial(), su.translate(1.0, 0.5, 0.2))
    cam = fakes.OrthographicCamera(96, 2.4, 0.75, world=su.translate(-5, 0, 0))
    img = pyrayt_b200.render.camera_nearest(cam, [lens, ball])
    assert img["distance"].shape == img["surface"].shape == (72, 96) and img["normal"].shape == (72, 96, 3)
    t, sid, nrm = oracle.nearest(pyrayt_b200.flatten([lens, ball]), cam.generate_rays())
    assert np.array_equal(img["distance"].ravel(), t) and np.array_equal(img["surface"].ravel(), sid)
    assert np.array_equal(np.moveaxis(img["normal"], -1, 0).reshape(3, -1), nrm, equal_nan=True)
    assert {-1, ball.get_id(), lens._l_child.get_id()} <= set(np.unique(sid).tolist())


def test_trace_small_replays_match_trace(torch_mod):
    """N4: the captured launch sequence for small ray sets gives the frame of the ordinary path, trace
    after trace, also when the scene is re-encoded in place between replays."""
    import pyrayt_b200
    from oracle import oracle
    from tests import scene_util as su

    for name in ("config1_collimator", "thick_lens_zoo", "facing_mirrors", "config5_cavity"):
        scene, rays, _, gl = load_case(name)
        gl = min(gl, 24)
        rays = np.ascontiguousarray(rays[:, :300])
        n = rays.shape[1]
        eng = pyrayt_b200.Engine(scene, device=0)
        want, octr = oracle.trace(scene, rays, gl)
        eng.small_ray_buffer(n).copy_(torch_mod.from_numpy(rays))
        for rep in range(4):  # eager, capture + replay, replay, replay
            res = eng.trace_small(n, generation_limit=gl)
            assert res.rows == want.shape[1] and np.array_equal(res.frame.numpy(), want, equal_nan=True), (name, rep)
            assert res.counters["segments"] == octr["segments"] and res.counters["rows_dropped"] == 0
            assert np.array_equal(res.gen_counts, np.bincount(want[0].astype(int), minlength=gl)[:gl])
        sid = int(scene.leaf_sid[-1])
        for rep in range(3):
            res = eng.trace_small(n, generation_limit=gl, record="surface", detector_sid=sid)
            assert np.array_equal(res.frame.numpy(), want[:, want[5] == sid], equal_nan=True)
        eager = eng.trace_small(n, generation_limit=gl, use_graph=False)
        assert np.array_equal(eager.frame.numpy(), want, equal_nan=True)
    # in-place scene updates between replays: same structure (graph kept), then another structure (re-captured)
    s1, rays = su.random_scene_and_rays(11, n_rays=200)
    eng = pyrayt_b200.Engine(s1, device=0)
    eng.small_ray_buffer(200).copy_(torch_mod.from_numpy(rays))
    for rep in range(3):
        eng.trace_small(200, generation_limit=12)
    moved = pyrayt_b200.FlatScene.from_json(s1.to_json())
    moved.leaf_obj = moved.leaf_obj.copy()
    moved.leaf_obj.reshape(-1, 4, 4)[:, 0, 3] += 0.05  # shift every surface in object space
    eng.update_scene(moved)
    got = eng.trace_small(200, generation_limit=12)
    want, _ = oracle.trace(moved, rays, 12)
    assert np.array_equal(got.frame.numpy(), want, equal_nan=True)
    s2, rays2 = su.random_scene_and_rays(12, n_rays=200)
    eng.update_scene(s2)
    eng.small_ray_buffer(200).copy_(torch_mod.from_numpy(rays2))
    for rep in range(3):
        got = eng.trace_small(200, generation_limit=12)
        want, _ = oracle.trace(s2, rays2, 12)
        assert np.array_equal(got.frame.numpy(), want, equal_nan=True)
    with pytest.raises(pyrayt_b200.PrtError):
        eng.trace_small(5000, generation_limit=10)


def test_wavefront_trace_is_bit_identical_to_single_kernel_trace(torch_mod):
    """prt_trace_wavefront writes rows straight to their (generation, id) positions: same frame, same
    counters as prt_trace + scan + gather and as the oracle; overflow retry and record modes included."""
    import pyrayt_b200
    from oracle import oracle
    from tests import scene_util as su
    from tests.helpers import GOLDEN_CASES

    for name in GOLDEN_CASES:
        scene, rays, _, gl = load_case(name)
        eng = pyrayt_b200.Engine(scene, device=0)
        d = torch_mod.from_numpy(np.ascontiguousarray(rays)).cuda()
        want, octr = oracle.trace(scene, rays, gl)
        ref = eng.trace(d, generation_limit=gl)
        res = eng.trace_wavefront(d, generation_limit=gl)
        assert res.rows == want.shape[1], name
        assert np.array_equal(res.frame.cpu().numpy(), want, equal_nan=True), name
        for k in ("rays", "generations", "segments", "tie_rays", "untraceable_hits", "bad_w", "nan_rays",
                  "limit_rays", "absorber_segments", "mirror_segments"):
            assert res.counters[k] == ref.counters[k], (name, k)
        assert np.array_equal(res.gen_counts, ref.gen_counts)
        via = eng.trace(d, generation_limit=gl, method="wavefront", to_host=True)
        assert np.array_equal(via.frame.numpy(), want, equal_nan=True), name
        tiny = eng.trace_wavefront(d, generation_limit=gl, capacity=7, to_host=True)  # overflow -> exact retry
        assert np.array_equal(tiny.frame.numpy(), want, equal_nan=True), name
        sid = int(scene.leaf_sid[-1])
        det = eng.trace_wavefront(d, generation_limit=gl, record="surface", detector_sid=sid)
        assert np.array_equal(det.frame.cpu().numpy(), want[:, want[5] == sid], equal_nan=True), name
    for seed in range(6):
        scene, rays = su.random_scene_and_rays(300 + seed, n_rays=3000)
        eng = pyrayt_b200.Engine(scene, device=0)
        d = torch_mod.from_numpy(rays).cuda()
        want, _ = oracle.trace(scene, rays, 14)
        ev = []
        res = eng.trace_wavefront(d, generation_limit=14, nearest_events=ev)
        assert np.array_equal(res.frame.cpu().numpy(), want, equal_nan=True), seed
        assert len(ev) == 15 and all(a.elapsed_time(b) >= 0 for a, b in ev)
    # edge sizes
    scene, rays, _, gl = load_case("thick_lens_zoo")
    eng = pyrayt_b200.Engine(scene, device=0)
    for n in (0, 1, 255, 256, 257):
        sub = np.ascontiguousarray(rays[:, :n])
        want, _ = oracle.trace(scene, sub, gl)
        res = eng.trace_wavefront(torch_mod.from_numpy(sub).cuda(), generation_limit=gl)
        assert res.frame.shape == (15, want.shape[1])
        assert np.array_equal(res.frame.cpu().numpy(), want, equal_nan=True), n


@pytest.mark.parametrize("name,n,max_bad_fraction", [("config2", 1 << 20, 1e-3), ("config3", 95326 * 11, 1e-4),
                                                     ("config4", 1 << 20, 1e-4), ("config5", 1 << 20, 2e-2)])
def test_fp32_fast_mode_on_the_workloads(name, n, max_bad_fraction, torch_mod):
    """The optional FP32 fast mode against the FP64 frame of the same rays at a million rays per workload:
    positions within 1e-5 of the scene scale, tilts / index within 1e-5, id columns equal, on all rays but the
    few that pass within the tolerance of an edge (config 5's sixteen-bounce light pipe amplifies those) or of
    one of the reference's own decision thresholds: in config 2 the condenser collimates the beam to ~1e-7 rad,
    where binomial_root's isclose(b, 0) (atol 1e-8, operations.py:45-52) decides whether the stop's cylinder is
    "missed" -- a threshold below single-precision resolution, 4.5e-4 of the rays sit on the other side of it."""
    from pyrayt_b200 import compare

    wl, eng = _engine(name)
    d_rays = wl.source.generate(n, device=0)
    f64 = eng.trace(d_rays, generation_limit=wl.generation_limit)
    f32 = eng.trace(d_rays, generation_limit=wl.generation_limit, precision="fp32")
    first = int(d_rays[12, 0].item())
    rep = compare.frame_agreement(f64.frame, f32.frame, first, n)
    assert rep["rays_with_different_ids"] + rep["rays_beyond_tolerance"] <= max_bad_fraction * n, rep
    assert rep["id_columns_equal_on_compared_rows"], rep
    assert rep["max_error_on_agreeing_rays"] <= 1e-5, rep


def test_fp32_fast_mode_full_size_properties(torch_mod):
    """The FP32 fast mode at the bench size (config 4, 2^24 rays): the size-independent properties of a frame --
    (generation, id) order, monotone survival, per-ray generation prefixes with no gaps, metadata columns that
    are copies of the input rays -- and a strided sample against the FP64 oracle within the mode's tolerance."""
    from oracle import oracle, sources_np
    from pyrayt_b200 import compare

    torch = torch_mod
    wl, eng = _engine("config4")
    n = wl.n_rays
    d_rays = wl.source.generate(n, device=0)
    res = eng.trace(d_rays, generation_limit=wl.generation_limit, precision="fp32")
    f = res.frame
    assert res.rows == f.shape[1] == res.counters["segments"] and res.counters["rows_dropped"] == 0
    gen, rid = f[0], f[4]
    key = gen * float(1 << 26) + rid
    assert bool((key[1:] > key[:-1]).all())
    gc = res.gen_counts
    assert np.all(np.diff(gc) <= 0) and gc.sum() == res.rows
    ids = rid.to(torch.int64)
    k = torch.bincount(ids, minlength=n)
    gs = torch.zeros(n, dtype=torch.float64, device=f.device).index_add_(0, ids, gen)
    assert bool((gs == (k * (k - 1) / 2).to(torch.float64)).all())
    # intensity / wavelength are bit-copies of the ray's input values, the surface column holds scene ids
    assert bool(torch.equal(f[1], d_rays[9][ids])) and bool(torch.equal(f[2], d_rays[10][ids]))
    sids = torch.from_numpy(wl.scene().leaf_sid.astype(np.float64)).to(f.device)
    assert bool(torch.isin(f[5], sids).all())
    # unit tilt columns, hit point = start + distance * tilt direction (single-precision accuracy)
    nrm = (f[12] ** 2 + f[13] ** 2 + f[14] ** 2).sqrt()
    assert float((nrm - 1).abs().max()) < 1e-6
    idx = np.arange(0, n, n // 2048)
    h = np.hstack([sources_np.from_source(wl.source, 1, first_index=int(i)) for i in idx])
    want, _ = oracle.trace(wl.scene(), h, wl.generation_limit, threads=THREADS)
    sel = torch.isin(rid, torch.from_numpy(idx.astype(np.float64)).to(f.device))
    got = f[:, sel].clone()
    # renumber the sample's ids 0..2047 so that the agreement helper can index them
    lut = torch.full((n,), -1, dtype=torch.float64, device=f.device)
    lut[torch.from_numpy(idx).to(f.device)] = torch.arange(len(idx), dtype=torch.float64, device=f.device)
    got[4] = lut[got[4].to(torch.int64)]
    w = torch.from_numpy(want).to(f.device)
    w[4] = lut[w[4].to(torch.int64)]
    rep = compare.frame_agreement(w, got, 0, len(idx))
    assert rep["rays_with_different_ids"] + rep["rays_beyond_tolerance"] <= 2, rep
    assert rep["id_columns_equal_on_compared_rows"] and rep["max_error_on_agreeing_rays"] <= 1e-5, rep
    del f, res
    eng.release_workspace()
    torch.cuda.empty_cache()


def test_lenslet_array_of_973_leaves_is_read_from_global_memory(torch_mod):
    """A 973-leaf scene (18 x 18 lenslets + detector): its encoded form (several hundred KB) does not fit a
    block's shared memory, so the trace kernel reads it in place through L1 / L2 (trace_kernel<.., GLOBAL>).
    Frame bit-equal to the oracle, counters-only and diagnosing traces agree, component.intersect and the
    nearest / render hits read it in place too; the FP32 mode and the wavefront driver (shared memory only)
    refuse it with PRT_ERR_LIMIT."""
    import pyrayt_b200
    from oracle import oracle
    from tests import scene_util as su

    torch = torch_mod
    scene, centres = su.lenslet_array(18, 18)
    rays = su.lenslet_rays(centres, 64)  # 20,736 rays
    eng = pyrayt_b200.Engine(scene, device=0)
    d = torch.from_numpy(rays).cuda()
    res = eng.trace(d, generation_limit=8, to_host=True)
    want, octr = oracle.trace(scene, rays, 8, threads=THREADS)
    assert np.array_equal(res.frame.numpy(), want, equal_nan=True)
    assert res.counters["generations"] == octr["generations"]
    none = eng.trace(d, generation_limit=8, record="none")
    assert none.counters["segments"] == res.rows
    sub = np.ascontiguousarray(rays[:, :2048])
    diag = eng.trace(torch.from_numpy(sub).cuda(), generation_limit=8, diagnose=True)
    o = oracle.diagnose(scene, sub, 8, threads=THREADS)
    assert (diag.counters["grazing_rays"], diag.counters["seam_rays"]) == (o["grazing_rays"], o["seam_rays"])
    # the plugin entry points read the large scene in place too: component.intersect, nearest / render hits
    r8 = np.zeros((8, 512))
    r8[0:3], r8[3], r8[4:7] = rays[0:3, :512], 1, rays[4:7, :512]
    for renderer in (False, True):
        t, s, nrm = eng.nearest_hit(torch.from_numpy(r8).cuda(), normals=True, renderer=renderer)
        ot, osid, onrm = (oracle.render_hit if renderer else oracle.nearest)(scene, r8)
        assert np.array_equal(t.cpu().numpy(), ot) and np.array_equal(s.cpu().numpy(), osid)
    for c in (0, 161, 324):
        hits, sids = eng.intersect(c, torch.from_numpy(r8.reshape(2, 4, -1)).cuda())
        oh, os_ = oracle.intersect(scene, c, r8.reshape(2, 4, -1))
        assert np.array_equal(hits.cpu().numpy(), oh, equal_nan=True) and np.array_equal(sids.cpu().numpy(), os_)
    with pytest.raises(pyrayt_b200.PrtError, match="too large"):
        eng.trace(d, generation_limit=8, precision="fp32")
    with pytest.raises(pyrayt_b200.PrtError, match="too large"):
        eng.trace_wavefront(d, generation_limit=8)
