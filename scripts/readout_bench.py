"""Device timing of the frame read-out kernels on a full-size config-4 frame.

    python scripts/readout_bench.py [rays]
Prints one JSON line: rows, selected rows, ms and GB/s (algorithmic bytes: the selection column for every
row + the 7 columns a selected row contributes) of one prt_spot_moments pass and of prt_axis_table.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pyrayt_b200  # noqa: E402
from pyrayt_b200 import analytics, workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
wl = workloads.WORKLOADS["config4"]
scene = wl.scene()
eng = pyrayt_b200.Engine(scene, 0)
res = eng.trace(wl.source.generate(n, device=0), generation_limit=wl.generation_limit)
eng.release_workspace()
det = float(scene.leaf_sid[-1])
groups, per = 9, (n + 8) // 9


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[1:])[len(ts[1:]) // 2], out


t_mom, sums = timed(lambda: analytics.spot_moments(res, per, groups, surface=det))
t_tab, tab = timed(lambda: analytics.focus_table(res, surface=det, to_host=False))
t_all, stats = timed(lambda: analytics.spot_stats(res, per, groups, surface=det))
sel = int(sums[:, 0].sum().item())
b_mom = 8 * res.rows + 56 * sel
b_tab = 2 * 8 * res.rows + sel * (7 * 8 + 4 * 8)
print(json.dumps({
    "rays": n, "rows": res.rows, "selected_rows": sel,
    "spot_moments": {"ms": round(t_mom, 3), "GB/s": round(b_mom / t_mom / 1e6, 1), "bytes": b_mom},
    "axis_table": {"ms": round(t_tab, 3), "GB/s": round(b_tab / t_tab / 1e6, 1), "bytes": b_tab},
    "spot_stats_end_to_end_ms": round(t_all, 3),
    "frame_GB_not_copied": round(res.rows * 120 / 1e9, 2),
    "spot": json.loads(stats[["n", "y_mean", "z_mean", "rms_radius", "focus_mean"]].to_json(orient="split"))["data"][:3],
}))
