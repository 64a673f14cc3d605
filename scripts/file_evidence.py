"""File the artefacts of scripts/gpu_run_final.sh <tag> (gpurun_out/) under profiles/r2/ and print the key numbers.

    python scripts/file_evidence.py r2k
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles", "r2")
tag = sys.argv[1]


def line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


for f in ([f"bench_{tag}_config{c}_1gpu.json" for c in range(1, 6)] +
          [f"bench_{tag}_reference_arm.json", f"pytest_gpu_{tag}.log", f"host_{tag}.txt", f"launches_{tag}.csv",
           f"sanitizer_{tag}.txt"]):
    shutil.copy(os.path.join(G, f), os.path.join(P, f))
shutil.copy(os.path.join(G, f"bench_{tag}_config4_1gpu.json"), os.path.join(ROOT, "profiles", "bench_r2_final.json"))
for kind, what in (("trace", "trace_kernel<true,false>, round-2 final FP64 kernel"), ("gather", "gather_kernel<0>"),
                   ("trace_f32", "trace_kernel_f32<true,true> (4 blocks/SM, ray-ordered walk)")):
    out = os.path.join(P, f"{kind}_kernel_{tag}_ncu_summary.json".replace("trace_f32_kernel", "trace_kernel_f32"))
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"),
                    os.path.join(G, f"prof_{kind}_{tag}_config4.ncu-rep"), out,
                    f"ncu --set full --clock-control none, {what}, config4, 2^24 rays (bench size), steady-state launch"],
                   capture_output=True)
s = json.load(open(os.path.join(P, f"trace_kernel_{tag}_ncu_summary.json")))
r, w = float(s["dram__bytes_read.sum"]["value"]) * 1e9, float(s["dram__bytes_write.sum"]["value"]) * 1e9
t = {"workload": "config4", "rays": 16777216, "dram_bytes_per_launch": r + w, "dram_bytes_read": r, "dram_bytes_write": w,
     "source": f"profiles/r2/trace_kernel_{tag}_ncu_summary.json (ncu --set full --clock-control none, one steady-state "
               "launch of trace_kernel<true,false> at the bench size, round-2 final kernel; K1 writes 72-byte staged "
               "records, the ordering pass expands them to the 120-byte rows)",
     "algorithmic_bytes_per_launch": 38386204544,
     "fp64_pipe_active_pct": float(s["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]["value"]),
     "issue_active_pct": float(s["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"]),
     "kernel_ms_under_ncu": float(s["gpu__time_duration.sum"]["value"])}
json.dump(t, open(os.path.join(ROOT, "profiles", "trace_kernel_traffic.json"), "w"), indent=1)

rows = list(csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))))
hi = [i for i, r_ in enumerate(rows) if "Kernel Name" in r_][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
acc, seq = collections.OrderedDict(), []
for r_ in rows[hi + 1:]:
    if len(r_) <= mv:
        continue
    v, u = float(r_[mv].replace(",", "")), r_[mu]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    a = acc.setdefault(r_[kn], [0.0, 0])
    a[0] += ms
    a[1] += 1
    seq.append((r_[kn], ms))
tot = sum(a[0] for a in acc.values())
tr = [ms for k, ms in seq if "trace_kernel<1, 0, 0, 0>" in k]
ga = [ms for k, ms in seq if "gather_kernel<0>" in k]
sc = [ms for k, ms in seq if "scan_runs" in k or "gen_offsets" in k]
share = (sum(tr) / len(tr)) / (sum(tr) / len(tr) + sum(ga) / len(ga) + sum(sc) / max(len(ga), 1))
out = [f"{ms:10.3f} ms  n={n:3d}  {100 * ms / tot:5.1f}%  {k[:90]}" for k, (ms, n) in sorted(acc.items(), key=lambda kv: -kv[1][0])]
open(os.path.join(P, f"launches_{tag}_summary.txt"), "w").write(
    "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e\n"
    "(first 400 launches: source generation, warm-up and timed FP64 steps, the diagnose trace, read-outs, the FP32 leg;\n"
    " per-launch times are cold-cache and serialised)\n"
    f"FP64 step under ncu: trace {sum(tr) / len(tr):.2f} ms, ordering {sum(ga) / len(ga):.2f} ms, scans "
    f"{sum(sc) / max(len(ga), 1):.2f} ms -> K1 share {share:.3f}\n" + "\n".join(out) + "\n")

print(open(os.path.join(P, f"pytest_gpu_{tag}.log")).read().strip().splitlines()[-2])
print(open(os.path.join(P, f"sanitizer_{tag}.txt")).read().strip())
for c in range(1, 6):
    d = line(os.path.join(P, f"bench_{tag}_config{c}_1gpu.json"))
    f = d.get("fp32_mode") or {}
    ag = f.get("agreement_with_fp64") or {}
    e = d["e2e"]
    print(f"config{c}: {d['value'] / 1e6:9.2f} M rays/s  step {d['ms_per_step']:8.3f} ms  K1 {d['roofline']['kernel_ms']:7.3f}  "
          f"hbm frac {d['roofline']['frac']:.3f}  tests/s {d['ray_surface_tests_per_s'] / 1e9:6.1f} G  "
          f"e2e {e['value'] / 1e6:7.2f} M ({e['ms_per_step']:.1f} ms, frac of host floor {e.get('frac_of_host_peak', 0):.2f}, "
          f"d2h {e['host_peak']['d2h_gbs_per_rank']:.1f} GB/s)  readout {e['with_device_readout']['value'] / 1e6:.1f} M")
    print(f"         fp32 {f.get('value', 0) / 1e6:9.2f} M  step {f.get('ms_per_step', 0):8.3f}  K1 {f.get('kernel_ms', 0):7.3f}  "
          f"different ids {ag.get('rays_with_different_ids')} beyond tol {ag.get('rays_beyond_tolerance')} of {ag.get('rays')} "
          f"max err {ag.get('max_error_on_agreeing_rays')}")
    cb = d["cpu_baseline"]
    nd = d["near_degenerate"]
    print(f"         numpy 1 core {cb['value']:.0f} rays/s, port {cb['port']['value']:.0f} rays/s ({cb['port']['cores']} threads); "
          f"near-degenerate: grazing {nd['grazing_rays']} seam {nd['seam_rays']} tie {nd['tie_rays']} of {nd['rays']}; "
          f"argsort mismatch rows {d['argsort_mismatch']['rows_differing']}; clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
d = line(os.path.join(P, f"bench_{tag}_reference_arm.json"))
print(f"reference arm: {d['value']:.0f} rays/s ({d['cpu_baseline']['cores']} processes), port {d['cpu_baseline']['port']['value']:.0f}")
print("K1 share of the FP64 step under ncu: %.3f; traffic file:" % share, {k: t[k] for k in ("dram_bytes_per_launch", "fp64_pipe_active_pct", "issue_active_pct", "kernel_ms_under_ncu")})
for k in ("gather_kernel", "trace_kernel_f32"):
    s2 = json.load(open(os.path.join(P, f"{k}_{tag}_ncu_summary.json")))
    print(k, {m.split(".")[0]: s2[m]["value"] for m in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                        "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                                        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                                                        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
