"""Quick device-only timing of the trace kernel (experiments; not the bench contract).

    PYRAYT_B200_LIB=pyrayt_b200/variants/lib_x.so python scripts/kbench.py [workload] [rays]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pyrayt_b200  # noqa: E402
from pyrayt_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
wl = workloads.WORKLOADS[name]
eng = pyrayt_b200.Engine(wl.scene(), 0)
PREC = os.environ.get("KBENCH_PRECISION", "fp64")
if PREC != "fp64":  # every trace below runs the fast mode
    _trace = eng.trace
    eng.trace = lambda *a, **k: _trace(*a, precision=PREC, **k)
rays = wl.source.generate(n, device=0)
if os.environ.get("KBENCH_ONLY") == "k1":  # profiling runs: three recording traces and nothing else
    for it in range(3):
        res = eng.trace(rays, generation_limit=wl.generation_limit)
        torch.cuda.synchronize()
        del res
    sys.exit(0)
for mode in (("all", "none") if os.environ.get("KBENCH_ONLY") != "wave" else ()):
    ts = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "all":
            res = eng.trace(rays, generation_limit=wl.generation_limit, k1_events=(e0, e1))
        else:
            res = eng.trace(rays, generation_limit=wl.generation_limit, record="none")
            e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[1:])[len(ts[1:]) // 2]
    print(f"{os.environ.get('PYRAYT_B200_LIB', 'default'):40s} {name} n={n} record={mode:4s} K1 {t:8.3f} ms  "
          f"{n / t / 1e3:8.2f} Mrays/s  rows {res.rows} gens {res.counters['generations']}", flush=True)
ts = []
for it in range(5 if os.environ.get("KBENCH_WAVE") else 0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev = []
    e0.record()
    res = eng.trace_wavefront(rays, generation_limit=wl.generation_limit, nearest_events=ev)
    e1.record()
    torch.cuda.synchronize()
    ts.append((e0.elapsed_time(e1), sum(a.elapsed_time(b) for a, b in ev)))
    del res
if ts:
    t, ta = sorted(ts[1:])[len(ts[1:]) // 2]
    print(f"{os.environ.get('PYRAYT_B200_LIB', 'default'):40s} {name} n={n} wavefront whole step {t:8.3f} ms  "
          f"{n / t / 1e3:8.2f} Mrays/s  (nearest-hit kernels {ta:8.3f} ms)", flush=True)
ts = []
for it in range(4 if os.environ.get("KBENCH_ONLY") != "wave" else 0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = eng.trace(rays, generation_limit=wl.generation_limit)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
    del res
t = sorted(ts[1:])[len(ts[1:]) // 2] if ts else float("nan")
print(f"{os.environ.get('PYRAYT_B200_LIB', 'default'):40s} {name} n={n} single-kernel whole step (trace+scan+gather) {t:8.3f} ms  "
      f"{n / t / 1e3:8.2f} Mrays/s", flush=True)
