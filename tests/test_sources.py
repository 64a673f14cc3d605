"""The seeded synthetic sources: NumPy restatement properties (CPU) and device equality (GPU)."""
import numpy as np
import pytest

from oracle import sources_np
from pyrayt_b200 import workloads


def test_uniforms_are_in_range_and_reproducible():
    i = np.arange(100000, dtype=np.uint64)
    u = sources_np.u01(7, i, 0)
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3
    assert np.array_equal(u, sources_np.u01(7, i, 0))
    assert not np.array_equal(u, sources_np.u01(8, i, 0))
    assert abs(np.corrcoef(u, sources_np.u01(7, i, 1))[0, 1]) < 0.02


def test_config4_fan_law():
    src = workloads.CONFIG4_SOURCE
    r = sources_np.from_source(src, 30000)
    assert np.all(r[0] == -10.0) and np.all(r[3] == 1) and np.all(r[7] == 0) and np.all(r[8] == 0)
    assert np.all(np.hypot(r[1], r[2]) <= 10.0 + 1e-12)
    assert np.allclose(np.hypot(r[4], r[5]), 1.0) and np.all(r[6] == 0)
    i = np.arange(30000)
    for f, ang in enumerate((0.0, 2.0, 5.0)):
        assert np.allclose(np.degrees(np.arctan2(r[5], r[4]))[i % 3 == f], ang)
    for w, lam in enumerate((0.486, 0.588, 0.656)):
        assert np.all(r[10][(i // 3) % 3 == w] == lam)
    assert np.array_equal(r[12], i)
    # a shard generated with first_index equals the slice of the whole
    part = sources_np.from_source(src, 1000, first_index=12345)
    assert np.array_equal(part, r[:, 12345:13345])


def test_cone_and_lambertian_directions_are_unit_and_inside_the_cone():
    r = sources_np.from_source(workloads.CONFIG2_SOURCE, 20000)
    assert np.allclose(np.linalg.norm(r[4:7], axis=0), 1.0, atol=1e-15)
    assert r[4].min() >= np.cos(np.radians(10.0)) - 1e-15
    # uniform in solid angle: cos(theta) is uniform on [cos(10 deg), 1]
    assert abs(r[4].mean() - (1 + np.cos(np.radians(10.0))) / 2) < 1e-4
    q = sources_np.from_source(workloads.CONFIG5_SOURCE, 20000)
    assert np.allclose(np.linalg.norm(q[4:7], axis=0), 1.0, atol=1e-15)
    assert np.all(q[4] < 0) and np.hypot(q[5], q[6]).max() <= np.sin(np.radians(20.0)) + 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["config2", "config4", "config5"])
def test_device_source_equals_numpy_restatement(name, cuda_device):
    src = workloads.WORKLOADS[name].source
    n = 1 << 20
    dev = src.generate(n, device=0, first_index=3 * n).cpu().numpy()
    assert np.array_equal(dev, sources_np.from_source(src, n, first_index=3 * n))
