"""The oracle against frames produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import oracle
from tests.helpers import GOLDEN_CASES, assert_frames_match, load_case


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_reference_frame(name):
    scene, rays, ref_frame, gl = load_case(name)
    frame, ctr = oracle.trace(scene, rays, gl)
    assert_frames_match(frame, ref_frame, what=name)
    assert ctr["segments"] == ref_frame.shape[1]
    assert ctr["untraceable_hits"] == 0


def test_config1_integration_pins():
    """int_test_ray_plane_intersection.py:48-54: 150 rows, generation-2 rays land on x = 1."""
    scene, rays, ref_frame, gl = load_case("config1_collimator")
    frame, _ = oracle.trace(scene, rays, gl)
    assert frame.shape[1] == 150
    np.testing.assert_allclose(frame[9, frame[0] == 2], 1.0)


def test_multithreaded_oracle_is_identical():
    scene, rays, _, gl = load_case("config4_stack")
    f1, c1 = oracle.trace(scene, rays, gl, threads=1)
    f4, c4 = oracle.trace(scene, rays, gl, threads=4)
    assert np.array_equal(f1, f4) and c1 == c4


def test_sharded_trace_is_bit_identical():
    """ray-range shards re-sorted by (generation, id) equal the monolithic frame."""
    scene, rays, _, gl = load_case("thick_lens_zoo")
    whole, _ = oracle.trace(scene, rays, gl)
    parts = [oracle.trace(scene, np.ascontiguousarray(rays[:, a:b]), gl)[0] for a, b in ((0, 300), (300, 701), (701, 1024))]
    cat = np.hstack(parts)
    order = np.lexsort((cat[4], cat[0]))
    assert np.array_equal(cat[:, order], whole)
