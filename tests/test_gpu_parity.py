"""GPU parity: the CUDA path through the C ABI against the oracle and the reference's golden frames."""
import numpy as np
import pytest

from tests.helpers import GOLDEN_CASES, assert_frames_match, load_case

pytestmark = pytest.mark.gpu


def _trace(scene, rays, gl, **kw):
    import torch

    import pyrayt_b200

    eng = pyrayt_b200.Engine(scene, device=0)
    res = eng.trace(torch.from_numpy(np.ascontiguousarray(rays)).cuda(), generation_limit=gl, **kw)
    return eng, res


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_case_matches_reference_and_oracle(name, cuda_device):
    from oracle import oracle

    scene, rays, ref_frame, gl = load_case(name)
    eng, res = _trace(scene, rays, gl, to_host=True)
    got = res.frame.numpy()
    want, octr = oracle.trace(scene, rays, gl)
    assert_frames_match(got, ref_frame, what=f"{name} vs reference golden")
    # the kernel and the oracle use the same roundings (no FMA, IEEE div/sqrt): identical bits
    assert np.array_equal(got, want, equal_nan=True), f"{name}: kernel differs from oracle"
    assert res.counters["segments"] == octr["segments"] == got.shape[1]
    assert res.counters["generations"] == octr["generations"]
    assert res.counters["limit_rays"] == octr["limit_rays"]
    assert res.counters["rows_dropped"] == 0 and res.counters["bad_w"] == 0
