"""Turn the artefacts of the final measurement run (gpurun_out/) into the tracked files under profiles/."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1_final"

d = json.loads(open(os.path.join(G, f"bench_{tag}.json")).read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "K1", d["roofline"]["kernel_ms"], "tests/s", d["ray_surface_tests_per_s"])
e = d["e2e"]
print("e2e", e["value"], e["ms_per_step"], e.get("transfer_ms_per_step_measured"), "readout", e["with_device_readout"]["value"])
print("cpu", d["cpu_baseline"]["value"], "fp64 frac", d["roofline_fp64"]["frac"], "hbm frac", d["roofline"]["frac"])
for f in (f"bench_{tag}.json", f"bench_{tag}_reference.json", f"launches_{tag}.csv"):
    shutil.copy(os.path.join(G, f), os.path.join(P, f))

rows = list(csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v, u = float(r[mv].replace(",", "")), r[mu]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    a = acc.setdefault(r[kn], [0.0, 0])
    a[0] += ms
    a[1] += 1
tot = sum(a[0] for a in acc.values())
out = [f"{ms:10.3f} ms  n={n:3d}  {100 * ms / tot:5.1f}%  {k[:70]}" for k, (ms, n) in sorted(acc.items(), key=lambda kv: -kv[1][0])]
open(os.path.join(P, f"launches_{tag}_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:3]))

summary = os.path.join(P, f"trace_kernel_{tag}_ncu_summary.json")
subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), os.path.join(G, f"prof_trace_{tag}.ncu-rep"),
                summary, "ncu --set full --clock-control none, trace_kernel<true,false>, config4, 16,777,216 rays "
                         "(full bench size), steady-state launch (-s 2), round-1 final kernel"],
               capture_output=True)
s = json.load(open(summary))
r, w = float(s["dram__bytes_read.sum"]["value"]) * 1e9, float(s["dram__bytes_write.sum"]["value"]) * 1e9
t = {"workload": "config4", "rays": 16777216, "dram_bytes_per_launch": r + w, "dram_bytes_read": r, "dram_bytes_write": w,
     "source": f"profiles/trace_kernel_{tag}_ncu_summary.json (ncu --set full --clock-control none, one steady-state "
               "launch of trace_kernel<true,false> at the bench size)",
     "algorithmic_bytes_per_launch": 38386204544,
     "fp64_pipe_active_pct": float(s["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]["value"]),
     "issue_active_pct": float(s["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"]),
     "kernel_ms_under_ncu": float(s["gpu__time_duration.sum"]["value"])}
json.dump(t, open(os.path.join(P, "trace_kernel_traffic.json"), "w"), indent=1)
print({k: t[k] for k in ("dram_bytes_per_launch", "fp64_pipe_active_pct", "issue_active_pct", "kernel_ms_under_ncu")})
print("thread instr per warp instr", s.get("smsp__thread_inst_executed_per_inst_executed.ratio"), "inst", s.get("smsp__inst_executed.sum"))
