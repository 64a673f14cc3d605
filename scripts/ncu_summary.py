"""Summarise an .ncu-rep (first kernel) into a small JSON: python scripts/ncu_summary.py rep.ncu-rep out.json 'note'"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
    'smsp__sass_average_branch_targets_threads_uniform.pct', 'sm__inst_executed.avg.per_cycle_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
]


def main():
    rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"_what": note, "_kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            d[h] = {"value": v, "unit": u}
    json.dump(d, open(out, "w"), indent=1)
    for k in ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
              'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
              'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
              'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_write.sum', 'launch__registers_per_thread',
              'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum'):
        print(k, d.get(k))
    st = {k: float(v["value"]) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled")}
    for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  stall {k.split('stalled_')[1].split('_per_')[0]:28s} {v:.3f}")


if __name__ == "__main__":
    main()
