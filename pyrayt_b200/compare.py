"""Agreement of two results frames of the same rays, computed where the frames are (device tensors).

Used to state the FP32 fast mode's contract with numbers (bench.py, tests): which rays took a different
path (another surface sequence or row count) and, over the rays that took the same path, the largest
deviation of positions (relative to the scene scale), unit tilts and refractive index.
"""
from __future__ import annotations


def _ray_signature(frame, first_id: int, n_rays: int):
    """Per ray: an order-independent 64-bit signature of its (generation, surface) rows and their number."""
    import torch

    ids = (frame[4] - float(first_id)).to(torch.int64)
    gen = frame[0].to(torch.int64)
    sid = frame[5].to(torch.int64)
    # rows of one ray have distinct generations, so a sum of per-row hashes identifies the sequence
    h = (gen * 0x9E3779B1 + 1) * ((sid + 2) * 0x85EBCA77 + 0x165667B1)
    h = (h ^ (h >> 29)) + (1 << 40)  # the added constant counts the rows
    sig = torch.zeros(n_rays, dtype=torch.int64, device=frame.device)
    sig.index_add_(0, ids, h)
    return sig, ids


def frame_agreement(frame_a, frame_b, first_id: int, n_rays: int, tol: float = 1e-5) -> dict:
    """frame_a / frame_b: (15, rows) float64 frames in (generation, id) order of rays first_id ..
    first_id + n_rays - 1 (ids consecutive).  A ray *agrees* when both frames give it the same (generation,
    surface) rows and every one of those rows is within `tol`: positions within tol x the scene scale (largest
    |coordinate| of frame_a), unit tilt and refractive index within tol.  Returns counts as Python numbers:
    rays_with_different_ids (another surface sequence or row count), rays_beyond_tolerance (same ids, some row
    further off -- e.g. another face of a one-id cuboid), and how the rows of the same-id rays spread over
    error decades."""
    import torch

    sig_a, ids_a = _ray_signature(frame_a, first_id, n_rays)
    sig_b, ids_b = _ray_signature(frame_b, first_id, n_rays)
    bad = sig_a != sig_b
    n_bad = int(bad.sum())
    keep_a = ~bad[ids_a]
    keep_b = ~bad[ids_b]
    rows_same = int(keep_a.sum())
    assert rows_same == int(keep_b.sum())
    out = {"rays": n_rays, "tolerance": tol, "rays_with_different_ids": n_bad, "rows_a": int(frame_a.shape[1]),
           "rows_b": int(frame_b.shape[1]), "rows_compared": rows_same}
    if rows_same == 0:
        return out
    scale = 1.0
    for c in range(6, 12):
        scale = max(scale, float(frame_a[c][keep_a].abs().max()))
    err = torch.zeros(rows_same, dtype=torch.float64, device=frame_a.device)  # worst relative error of each row
    for c in range(6, 12):
        err = torch.maximum(err, (frame_a[c][keep_a] - frame_b[c][keep_b]).abs() / scale)
    for c in (3, 12, 13, 14):
        err = torch.maximum(err, (frame_a[c][keep_a] - frame_b[c][keep_b]).abs())
    err = torch.nan_to_num(err, nan=0.0)  # NaN tilt columns (dead directions) are NaN in both frames
    ids_same = ids_a[keep_a]
    off = torch.zeros(n_rays, dtype=torch.bool, device=frame_a.device)
    off[ids_same[err > tol]] = True
    exact = all(bool(torch.equal(frame_a[c][keep_a], frame_b[c][keep_b])) for c in (0, 1, 2, 4, 5))
    within = ~off[ids_same]
    out.update({"scene_scale": scale, "rays_beyond_tolerance": int(off.sum()),
                "rays_agreeing": n_rays - n_bad - int(off.sum()),
                "id_columns_equal_on_compared_rows": exact,
                "max_error_on_agreeing_rays": float(err[within].max()) if bool(within.any()) else 0.0,
                "rows_by_error": {"<=1e-7": int((err <= 1e-7).sum()), "<=1e-6": int((err <= 1e-6).sum()),
                                  "<=1e-5": int((err <= 1e-5).sum()), "<=1e-4": int((err <= 1e-4).sum()),
                                  "all": rows_same}})
    return out
