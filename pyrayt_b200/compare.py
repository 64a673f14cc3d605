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


def frame_agreement(frame_a, frame_b, first_id: int, n_rays: int) -> dict:
    """frame_a / frame_b: (15, rows) float64 frames in (generation, id) order of rays first_id ..
    first_id + n_rays - 1 (ids consecutive).  Returns counts and maxima as Python numbers."""
    import torch

    sig_a, ids_a = _ray_signature(frame_a, first_id, n_rays)
    sig_b, ids_b = _ray_signature(frame_b, first_id, n_rays)
    bad = sig_a != sig_b
    n_bad = int(bad.sum())
    keep_a = ~bad[ids_a]
    keep_b = ~bad[ids_b]
    rows_same = int(keep_a.sum())
    assert rows_same == int(keep_b.sum())
    out = {"rays": n_rays, "rays_with_a_different_path": n_bad, "rows_a": int(frame_a.shape[1]),
           "rows_b": int(frame_b.shape[1]), "rows_compared": rows_same}
    if rows_same == 0:
        return out
    scale = 1.0
    for c in range(6, 12):
        scale = max(scale, float(frame_a[c][keep_a].abs().max()))
    pos = max(float((frame_a[c][keep_a] - frame_b[c][keep_b]).abs().max()) for c in range(6, 12))
    tilt = max(float((frame_a[c][keep_a] - frame_b[c][keep_b]).abs().max()) for c in range(12, 15))
    idx = float((frame_a[3][keep_a] - frame_b[3][keep_b]).abs().max())
    exact = all(bool(torch.equal(frame_a[c][keep_a], frame_b[c][keep_b])) for c in (0, 1, 2, 4, 5))
    out.update({"scene_scale": scale, "max_position_error": pos, "max_position_error_rel_scale": pos / scale,
                "max_tilt_error": tilt, "max_index_error": idx, "id_columns_equal_on_compared_rows": exact})
    return out
