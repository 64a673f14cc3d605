"""Frame read-outs: the notebook's pandas reductions, oracle on CPU, kernels on the GPU."""
import numpy as np
import pytest

from tests.helpers import load_case

COLUMNS = ("generation", "intensity", "wavelength", "index", "id", "surface", "x0", "y0", "z0", "x1", "y1", "z1",
           "x_tilt", "y_tilt", "z_tilt")

# sums of ~1e5 terms in a different order than NumPy's pairwise sum: relative to the magnitude of the data
RTOL = 1e-9


def _df(frame):
    import pandas as pd

    return pd.DataFrame(np.ascontiguousarray(frame.T), columns=COLUMNS)


def _last_surface(frame):
    """The surface most rows of the last populated generation end on (the imager / detector)."""
    g = frame[0].max()
    s = frame[5][frame[0] == g]
    vals, counts = np.unique(s, return_counts=True)
    return float(vals[np.argmax(counts)]), float(g)


# ------------------------------------------------------------------ CPU: the oracle against the notebook, literally

def test_oracle_focus_table_is_the_notebook_computation():
    from oracle import analytics_np

    _, rays, frame, _ = load_case("config1_collimator")
    results = _df(frame)
    gmax = np.max(results["generation"])
    # lens_design.ipynb cell 12, statement by statement
    imager_rays = results.loc[results["generation"] == gmax]
    intercept = -imager_rays["x_tilt"] * imager_rays["y0"] / imager_rays["y_tilt"] + imager_rays["x0"]
    radii = results.loc[np.logical_and(results["generation"] == 0, results["id"].isin(imager_rays["id"]))]["y0"]
    got = analytics_np.focus_table(results, generation=gmax)
    assert np.array_equal(np.asarray(got["focus"]), np.asarray(intercept), equal_nan=True)
    assert np.array_equal(np.asarray(got["radius"]), np.asarray(radii))  # ids ascend within a generation
    assert np.array_equal(np.asarray(got["wavelength"]), np.asarray(imager_rays["wavelength"]))


def test_oracle_spot_stats_by_hand():
    from oracle import analytics_np

    _, rays, frame, _ = load_case("config2_tutorial")
    results = _df(frame)
    sid, _ = _last_surface(frame)
    n = rays.shape[1]
    per = (n + 3) // 4
    stats = analytics_np.spot_stats(results, per, 4, surface=sid)
    rows = frame[:, frame[5] == sid]
    for g in range(4):
        sub = rows[:, (rows[4] / per).astype(int) == g]
        assert stats.loc[g, "n"] == sub.shape[1]
        if sub.shape[1]:
            assert stats.loc[g, "y_mean"] == np.mean(sub[10])
            cy, cz = np.mean(sub[10]), np.mean(sub[11])
            assert np.isclose(stats.loc[g, "rms_radius"], np.sqrt(np.mean((sub[10] - cy) ** 2 + (sub[11] - cz) ** 2)),
                              rtol=1e-14)


def test_no_cpu_path():
    import torch

    from pyrayt_b200 import PrtError, analytics

    with pytest.raises(PrtError, match="no CPU path"):
        analytics.spot_stats(torch.zeros((15, 4), dtype=torch.float64), 1, 1)
    with pytest.raises(PrtError, match="no CPU path"):
        analytics.focus_table(np.zeros((15, 4)))


# ------------------------------------------------------------------ GPU parity

def _assert_stats_close(got, want, scale):
    assert list(got["n"]) == list(want["n"])
    assert list(got["n_focus"]) == list(want["n_focus"])
    for col in want.columns:
        if col in ("n", "n_focus"):
            continue
        g, w = np.asarray(got[col], dtype=float), np.asarray(want[col], dtype=float)
        assert np.array_equal(np.isnan(g), np.isnan(w)), col
        if col.endswith(("_min", "_max")):
            assert np.array_equal(g, w, equal_nan=True), col  # order-independent: exact
            continue
        ref = scale * scale if col in ("yz_cov", "sin_tilt_msd") else scale
        err = np.nanmax(np.abs(g - w), initial=0.0)
        assert err <= RTOL * ref, f"{col}: {err:.3e} (scale {ref:.3e})"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["config1_collimator", "config2_tutorial", "config3_prism", "config4_stack",
                                  "thick_lens_zoo", "facing_mirrors"])
def test_spot_stats_and_focus_table_match_oracle(name, cuda_device):
    import torch

    import pyrayt_b200
    from oracle import analytics_np
    from pyrayt_b200 import analytics

    scene, rays, _, gl = load_case(name)
    eng = pyrayt_b200.Engine(scene, device=0)
    res = eng.trace(torch.from_numpy(np.ascontiguousarray(rays)).cuda(), generation_limit=gl)
    assert res.frame.is_cuda
    frame = res.frame.cpu().numpy()
    results = _df(frame)
    sid, gmax = _last_surface(frame)
    n = rays.shape[1]
    finite = frame[9:12][np.isfinite(frame[9:12])]
    scale = max(1.0, float(np.max(np.abs(finite)))) if finite.size else 1.0
    for groups in (1, 3, 7):
        per = (n + groups - 1) // groups
        for sel in (dict(surface=sid), dict(generation=gmax), dict(generation=0.0), dict()):
            picked = analytics_np._rows(results, **sel)
            if not np.all(np.isfinite(np.asarray(picked[["y1", "z1"]]))):
                continue  # a ray that left the scene has x1 = inf: moments of such a selection are inf - inf
            want = analytics_np.spot_stats(results, per, groups, tilt_center=0.125, **sel)
            got = analytics.spot_stats(res, per, groups, tilt_center=0.125, **sel)
            f_all = np.asarray(analytics_np.focus(picked), dtype=float)
            fscale = np.max(np.abs(f_all[np.isfinite(f_all)]), initial=1.0)
            _assert_stats_close(got.drop(columns=["focus_mean", "focus_std"]),
                                want.drop(columns=["focus_mean", "focus_std"]), scale)
            for col in ("focus_mean", "focus_std"):
                g, w = np.asarray(got[col], dtype=float), np.asarray(want[col], dtype=float)
                ok = np.isfinite(w)
                assert np.array_equal(np.isfinite(g), ok), col
                assert np.all(np.abs(g[ok] - w[ok]) <= RTOL * max(1.0, fscale) * 10), col
    for sel in (dict(surface=sid), dict(generation=gmax), dict(generation=0.0), dict()):
        want = analytics_np.focus_table(results, **sel)
        got = analytics.focus_table(res, **sel)
        assert len(got) == len(want)
        for col in ("id", "radius", "focus", "wavelength"):  # IEEE element-wise arithmetic: same bits
            assert np.array_equal(np.asarray(got[col]), np.asarray(want[col]), equal_nan=True), (sel, col)


@pytest.mark.gpu
def test_focus_table_when_rays_miss_in_generation_zero(cuda_device):
    """Rays that hit nothing in generation 0 have no row, so the head of the frame is compacted and a
    later ray's generation-0 row is not at (id - first_id): the radius must still be its own y0
    (the notebook's results.loc[(generation == 0) & id.isin(...)]['y0'])."""
    import torch

    import pyrayt_b200
    from oracle import analytics_np
    from pyrayt_b200 import analytics

    scene, rays, _, gl = load_case("config1_collimator")
    rays = np.tile(rays, (1, 8))
    n = rays.shape[1]
    rays[12] = 1000 + np.arange(n)  # ids need not start at 0 either
    miss = (np.arange(n) % 3 == 0) | (np.arange(n) < 7)
    rays[4, miss] *= -1.0  # pointing away from the lens: no hit, no row
    eng = pyrayt_b200.Engine(scene, device=0)
    res = eng.trace(torch.from_numpy(np.ascontiguousarray(rays)).cuda(), generation_limit=gl)
    frame = res.frame.cpu().numpy()
    assert 0 < int(res.gen_counts[0]) == n - int(miss.sum())
    results = _df(frame)
    sid, gmax = _last_surface(frame)
    for sel in (dict(surface=sid), dict(generation=gmax), dict(generation=0.0), dict()):
        want = analytics_np.focus_table(results, **sel)
        for first_id in (1000, 0):  # the hint may be right, or useless
            got = analytics.focus_table(res, first_id=first_id, **sel)
            assert len(got) == len(want) > 0
            assert not np.any(np.isnan(np.asarray(got["radius"])))
            for col in ("id", "radius", "focus", "wavelength"):
                assert np.array_equal(np.asarray(got[col]), np.asarray(want[col]), equal_nan=True), (sel, col)


@pytest.mark.gpu
def test_analytics_edge_cases(cuda_device):
    import torch

    import pyrayt_b200
    from pyrayt_b200 import analytics

    empty = torch.empty((15, 0), dtype=torch.float64, device="cuda")
    st = analytics.spot_stats(empty, 10, 3, surface=5)
    assert list(st["n"]) == [0, 0, 0] and np.all(np.isnan(st["y_mean"]))
    assert len(analytics.focus_table(empty, generation=0)) == 0
    with pytest.raises(pyrayt_b200.PrtError, match="n_groups"):
        analytics.spot_stats(empty, 10, 257)
    with pytest.raises(pyrayt_b200.PrtError, match="rays_per_group"):
        analytics.spot_stats(empty, 0, 2)
    # rows whose group is beyond n_groups are ignored; a selection nothing matches gives zero counts
    scene, rays, _, gl = load_case("config1_collimator")
    eng = pyrayt_b200.Engine(scene, device=0)
    res = eng.trace(torch.from_numpy(np.ascontiguousarray(rays)).cuda(), generation_limit=gl)
    st = analytics.spot_stats(res, 10, 2, generation=0)
    assert list(st["n"]) == [10, 10]
    assert list(analytics.spot_stats(res, 10, 2, surface=-12345.0)["n"]) == [0, 0]
    assert len(analytics.focus_table(res, surface=-12345.0)) == 0


@pytest.mark.gpu
def test_spot_stats_large_frame(cuda_device):
    """2^20 rays of the config-4 stack: every block / group path of the reduction, against NumPy."""
    import torch

    import pyrayt_b200
    from pyrayt_b200 import analytics, workloads

    wl = workloads.WORKLOADS["config4"]
    n = 1 << 20
    eng = pyrayt_b200.Engine(wl.scene(), 0)
    res = eng.trace(wl.source.generate(n, device=0), generation_limit=wl.generation_limit)
    frame = res.frame.cpu().numpy()
    sid, _ = _last_surface(frame)
    groups, per = 16, n // 16
    got = analytics.spot_stats(res, per, groups, surface=sid)
    rows = frame[:, frame[5] == sid]
    gid = (rows[4] / per).astype(int)
    for g in range(groups):
        sub = rows[:, gid == g]
        assert got.loc[g, "n"] == sub.shape[1]
        y, z = sub[10], sub[11]
        assert abs(got.loc[g, "y_mean"] - y.mean()) <= RTOL * 25.4
        assert abs(got.loc[g, "z_mean"] - z.mean()) <= RTOL * 25.4
        want_rms = np.sqrt(np.mean((y - y.mean()) ** 2 + (z - z.mean()) ** 2))
        assert abs(got.loc[g, "rms_radius"] - want_rms) <= RTOL * 25.4
        assert got.loc[g, "y_min"] == y.min() and got.loc[g, "z_max"] == z.max()
    tab = analytics.focus_table(res, surface=sid, to_host=False)
    assert tab.shape == (4, rows.shape[1])
    assert np.array_equal(tab[0].cpu().numpy(), rows[4])
    assert np.array_equal(tab[1].cpu().numpy(), frame[7][rows[4].astype(np.int64)])
