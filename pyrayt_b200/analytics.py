"""Read-out of a results frame on the GPU.

What users of the reference read off the 15-column frame in pandas
(examples/lens_design.ipynb): the spot of every source on the imager
(cells 11, 19, 38), the focus of each imager ray against its launch radius or
wavelength (cells 12-16) and the coma metric (cell 20).  Here the frame stays on
the device it was traced on (``Engine.trace(..., to_host=False)``) and the
reductions run in the kernels of ``csrc/prt_analytics.cu``; only the per-source
table comes back.  Like everything else in this package there is no CPU path.

Sharded traces: the additive moments are all-reduced over ``torch.distributed``
between the two passes, so every rank gets the statistics of the whole ray set.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib
from .dist import reduce_spot_sums

SPOT_COLUMNS = ("n", "y_mean", "z_mean", "y_std", "z_std", "yz_cov", "rms_radius", "y_min", "y_max", "z_min",
                "z_max", "n_focus", "focus_mean", "focus_std", "y_tilt_mean", "y_tilt_std", "sin_tilt_msd")


def _select(surface, generation):
    if surface is not None and generation is not None:
        raise ValueError("select rows by surface or by generation, not both")
    if surface is not None:
        sid = surface.get_id() if hasattr(surface, "get_id") else surface
        return _lib.SELECT_SURFACE, float(sid)
    if generation is not None:
        return _lib.SELECT_GENERATION, float(generation)
    return _lib.SELECT_ALL, 0.0


def _device_frame(frame):
    import torch

    frame = getattr(frame, "frame", frame)  # a TraceResult
    if not isinstance(frame, torch.Tensor) or not frame.is_cuda:
        raise _lib.PrtError("analytics run on the device frame (Engine.trace(..., to_host=False)); "
                            "there is no CPU path")
    if frame.dtype != torch.float64 or frame.dim() != 2 or frame.shape[0] != _lib.FRAME_COLS:
        raise _lib.PrtError("frame must be a (15, rows) float64 tensor")
    if frame.shape[1] and frame.stride(1) != 1:
        raise _lib.PrtError("frame columns must be contiguous")
    return frame


def _stream(frame):
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream(frame.device).cuda_stream)


def spot_moments(frame, rays_per_group: int, n_groups: int, surface=None, generation=None, center=None):
    """Raw per-group moments (n_groups, 16) on the device; see prt_spot_moments in include/pyrayt_b200.h."""
    import torch

    frame = _device_frame(frame)
    lib = _lib.load()
    select, value = _select(surface, generation)
    rows = int(frame.shape[1])
    out = torch.empty((int(n_groups), _lib.SPOT_COLS), dtype=torch.float64, device=frame.device)
    with torch.cuda.device(frame.device):
        _lib.check(lib.prt_spot_moments(frame.data_ptr() if rows else None, rows, int(frame.stride(0)) if rows else 0,
                                        select, value, int(rays_per_group), int(n_groups),
                                        center.data_ptr() if center is not None else None, out.data_ptr(), 0,
                                        _stream(frame)), "prt_spot_moments")
    return out


def spot_stats(frame, rays_per_group: int, n_groups: int, surface=None, generation=None,
               tilt_center: Optional[float] = None):
    """Per-source spot statistics of the selected rows -> pandas.DataFrame indexed by source_id.

    ``source_id = floor(id / rays_per_group)`` as ``RayTracer.calculate_source_ids``
    (pyrayt/_pyrayt.py:316-327).  Standard deviations are population ones (``np.std``);
    ``rms_radius`` is about the spot centroid; ``sin_tilt_msd`` is
    ``mean((sin(y_tilt) - tilt_center)**2)`` (lens_design.ipynb cell 20, ``tilt_center = sin(angle)``;
    about 0 when not given).
    """
    import pandas as pd
    import torch

    frame = _device_frame(frame)
    lib = _lib.load()
    g = int(n_groups)
    first = spot_moments(frame, rays_per_group, g, surface, generation)
    reduce_spot_sums(first)
    center = torch.empty((g, _lib.SPOT_CENTER_COLS), dtype=torch.float64, device=frame.device)
    with torch.cuda.device(frame.device):
        _lib.check(lib.prt_spot_centers(first.data_ptr(), None, g, center.data_ptr(), _stream(frame)),
                   "prt_spot_centers")
    tilt_mean = center[:, 3].cpu().numpy().copy()
    if tilt_center is not None:
        center[:, 3] = float(tilt_center)
    second = spot_moments(frame, rays_per_group, g, surface, generation, center=center)
    reduce_spot_sums(second)
    s = second.cpu().numpy()
    c = center.cpu().numpy()
    n, nf = s[:, 0], s[:, 10]
    with np.errstate(invalid="ignore", divide="ignore"):
        ry, rz, rf = s[:, 1] / n, s[:, 2] / n, s[:, 11] / nf  # residual means (about 0: the centres are the means)
        var_y = s[:, 3] / n - ry * ry
        var_z = s[:, 4] / n - rz * rz
        cov = s[:, 5] / n - ry * rz
        var_f = s[:, 12] / nf - rf * rf
        rt = s[:, 13] / n  # pass 2 may have been centred on tilt_center instead of the mean
        var_t = s[:, 14] / n - rt * rt
        table = {
            "n": n.astype(np.int64),
            "y_mean": c[:, 0] + ry, "z_mean": c[:, 1] + rz,
            "y_std": np.sqrt(np.maximum(var_y, 0.0)), "z_std": np.sqrt(np.maximum(var_z, 0.0)),
            "yz_cov": cov,
            "rms_radius": np.sqrt(np.maximum(var_y + var_z, 0.0)),
            "y_min": np.where(n > 0, s[:, 6], np.nan), "y_max": np.where(n > 0, s[:, 7], np.nan),
            "z_min": np.where(n > 0, s[:, 8], np.nan), "z_max": np.where(n > 0, s[:, 9], np.nan),
            "n_focus": nf.astype(np.int64),
            "focus_mean": c[:, 2] + rf, "focus_std": np.sqrt(np.maximum(var_f, 0.0)),
            "y_tilt_mean": tilt_mean, "y_tilt_std": np.sqrt(np.maximum(var_t, 0.0)),
            "sin_tilt_msd": s[:, 15] / n,
        }
    df = pd.DataFrame(table, columns=list(SPOT_COLUMNS))
    df.index.name = "source_id"
    return df


def focus_table(frame, surface=None, generation=None, first_id: int = 0, gen0_rows: Optional[int] = None,
                to_host: bool = True):
    """The table of lens_design.ipynb cells 12 and 15: one row per selected frame row, in frame order,
    with ``id``, ``radius`` (y0 of that ray's generation-0 row), ``focus`` (= -x_tilt*y0/y_tilt + x0)
    and ``wavelength``.  ``gen0_rows``: rows of generation 0 at the head of the frame (default: found
    from the frame).  Returns a pandas.DataFrame, or the (4, k) device tensor when ``to_host=False``.
    """
    import pandas as pd
    import torch

    if gen0_rows is None and getattr(frame, "gen_counts", None) is not None and len(frame.gen_counts):
        gen0_rows = int(frame.gen_counts[0])  # a TraceResult knows its rows per generation
    frame = _device_frame(frame)
    lib = _lib.load()
    select, value = _select(surface, generation)
    rows = int(frame.shape[1])
    dev = frame.device
    if gen0_rows is None:
        gen0_rows = int((frame[0] == 0.0).sum().item()) if rows else 0
    nb = int(lib.prt_axis_table_blocks(rows))
    count = torch.empty(max(nb, 1), dtype=torch.int32, device=dev)
    base = torch.empty(max(nb, 1), dtype=torch.int64, device=dev)
    total = torch.zeros(2, dtype=torch.int64, device=dev)
    # the number of selected rows is only known after the count pass: size the table for one row per ray
    # (all a generation or a detector selection can hold) and redo the call in the rare case it was short
    cap = min(rows, max(int(gen0_rows), 4096))
    while True:
        table = torch.empty((4, max(cap, 1)), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.prt_axis_table(frame.data_ptr() if rows else None, rows,
                                          int(frame.stride(0)) if rows else 0, select, value, int(first_id),
                                          int(gen0_rows), count.data_ptr(), base.data_ptr(), total.data_ptr(),
                                          table.data_ptr(), int(table.stride(0)), cap, _stream(frame)),
                       "prt_axis_table")
        k = int(total[1].item())
        if k <= cap:
            break
        cap = k
    out = table[:, :k]
    if not to_host:
        return out
    h = out.cpu().numpy()
    return pd.DataFrame({"id": h[0], "radius": h[1], "focus": h[2], "wavelength": h[3]})
