"""bench.py -- traced rays/s and ray-surface tests/s of the PyRayT hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--workload config4] [--rays N] [--impl reference]

One "step" is one full trace of the workload's ray batch: the persistent trace kernel
(every generation, CSG merging, nearest hit, material interaction, record append) followed by
the (generation, id) ordering of the records into the 15-column frame.  With N > 1 (launched by
torchrun, one rank per GPU) every rank traces its own ray-index range of an N x larger source
("weak" scaling); the only collective is the all-gather of per-generation row counts.

value  = rays/s with the rays already resident in HBM and the frame left in HBM.
e2e    = the same through the host-buffer API: rays in pinned host memory are copied H2D,
         traced, ordered, and the whole frame is copied D2H, all inside the timed region.
roofline / roofline_fp64 / cpu_baseline: see DESIGN.md "Measurement".

--scaling weak (default): every rank traces --rays rays (N x larger source); --scaling strong: the
workload's ray count is split over the ranks by ray-index range (config 4: "16M rays at 1/2/4/8 GPUs").

--impl reference times the reference's own CPU path on this box's host cores, on a bounded sample of
the same workload: the UNMODIFIED NumPy reference (baseline/_ref, staged by oracle/stage_reference.py;
one process per core over ray-index ranges) when it is present, with the multithreaded C restatement
(oracle/, "port") timed beside it; the port alone otherwise.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "traced rays/s"
UNIT = "rays/s"
CPU_SAMPLE_RAYS = 1 << 18          # oracle port (C restatement), all host threads
NUMPY_SAMPLE_RAYS_1CORE = 1 << 14  # unmodified NumPy reference as a user runs it (one core), ~10 s
NUMPY_SAMPLE_RAYS_PER_PROC = 1 << 12  # reference arm: rays per worker process and step
MISMATCH_SAMPLE_RAYS = 1 << 12     # stable-vs-default argsort report
FP32_COMPARE_RAYS = 1 << 22        # FP32 fast mode vs FP64 frame agreement, untimed
DIAGNOSE_SAMPLE_RAYS = 1 << 22     # grazing / seam ray count (PRT_FLAG_DIAGNOSE), untimed


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config4")
    ap.add_argument("--rays", type=int, default=0,
                    help="rays per GPU (weak) / in total (strong); default: the workload's")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-numpy", action="store_true", help="skip the NumPy-reference legs even if baseline/_ref exists")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the FP32 fast-mode leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--zero-copy", action="store_true", help="e2e: gather kernel writes straight into pinned host memory")
    ap.add_argument("--full-copy", action="store_true",
                    help="e2e: copy all 15 columns over PCIe instead of rebuilding 5 of them on the host")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled (NVML, every 20 ms) during the timed region."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread, self.err = index, [], False, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # NVML missing: report it, never fail the bench
            self.nv, self.err = None, repr(exc)

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].strip().isdigit():
                return int(ids[index])
        return index

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception as exc:
                self.err = repr(exc)
                return
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "error": self.err}
        load = [s for s in self.samples if not (s[1] & 0x1)] or self.samples  # drop "GPU idle" samples
        mhz = sorted(s[0] for s in load)
        bits = 0
        for _, r in load:
            bits |= r
        return {"sm_mhz": float(mhz[len(mhz) // 2]), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(n for b, n in self.REASONS.items() if bits & b), "samples": len(load)}


def host_rays(workload, n, first=0, total=None):
    """The workload's rays [first, first + n) as a host array (NumPy restatement of the device sources)."""
    from oracle import sources_np

    return sources_np.from_source(workload.source, n, first, total)


def port_sample_size(workload, n):
    """Rays the port traces per step: a ReferenceSourceSet needs a multiple of its source count."""
    k = len(getattr(workload.source, "templates", (None,)))
    return max(k, (min(CPU_SAMPLE_RAYS, n) // k) * k)


def cpu_reference_run(workload, n_sample, threads, steps, warmup):
    """Times the oracle port (CPU restatement of the reference) on the first n_sample rays."""
    import numpy as np

    from oracle import oracle

    scene = workload.scene()
    rays = host_rays(workload, n_sample)
    cap = n_sample * min(workload.generation_limit, 40)
    frame = np.empty((15, cap))
    times, rows, ctr = [], 0, {}
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        rows, ctr = oracle.trace_timed(scene, rays, workload.generation_limit, 1e-6, threads, frame)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return {"rays_per_s": n_sample / mean, "tests_per_s": ctr["generations"] * scene.n_leaves / mean,
            "rows": rows, "ms_per_step": mean * 1e3, "counters": ctr, "steps": steps, "warmup": warmup}


def numpy_reference_available(args):
    if getattr(args, "no_numpy", False):
        return False
    try:
        from oracle import ref_shim

        return ref_shim.available()
    except Exception:
        return False


def numpy_reference_run(workload, n_sample, processes, steps, warmup):
    """Times the UNMODIFIED NumPy reference (RayTracer.trace() through a FixedSource holding the same
    seeded rays) on the first n_sample rays; processes > 1 = one process per contiguous ray range."""
    from oracle import ref_scenes

    k = len(getattr(workload.source, "templates", (None,)))
    n_sample = max(k, (n_sample // k) * k)
    rays = host_rays(workload, n_sample)
    times, rows = [], 0
    for it in range(warmup + steps):
        r = ref_scenes.time_reference(workload.name, rays, workload.generation_limit, processes)
        rows = r["rows"]
        if it >= warmup:
            times.append(r["seconds"])
    mean = sum(times) / len(times)
    return {"rays_per_s": n_sample / mean, "rows": rows, "ms_per_step": mean * 1e3, "rays": n_sample,
            "processes": r["processes"], "steps": steps, "warmup": warmup}


def argsort_mismatch_report(workload, n):
    """SURVEY 9-Q3 / north star: rows of the reference's own frame that differ between the stable argsort
    (the parity contract; pinned numpy 1.20 behaviour) and this numpy's default argsort."""
    from oracle import ref_scenes

    k = len(getattr(workload.source, "templates", (None,)))
    n = max(k, (min(MISMATCH_SAMPLE_RAYS, n) // k) * k)
    rep = ref_scenes.argsort_mismatch(workload.name, host_rays(workload, n), workload.generation_limit)
    rep["what"] = ("unmodified reference, same seeded rays, np.argsort kind='stable' (the contract the kernel "
                   "matches bit for bit) vs this numpy's default argsort")
    return rep


def run_reference(args):
    """The reference arm: the reference's CPU implementation of the path on this box's host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from pyrayt_b200 import workloads

    wl = workloads.WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n_total = args.rays or wl.n_rays
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    port_n = port_sample_size(wl, n_total)
    line_cfg = {"workload": wl.name, "description": wl.description, "generation_limit": wl.generation_limit}
    if numpy_reference_available(args):
        n = min(n_total, NUMPY_SAMPLE_RAYS_PER_PROC * cores)
        r = numpy_reference_run(wl, n, cores, steps, warmup)
        port = cpu_reference_run(wl, port_n, cores, 2, 1)
        kind = "reference"
        sample = (f"first {r['rays']} rays of {wl.name} ({wl.description}) through the UNMODIFIED NumPy reference "
                  f"(pyrayt.RayTracer.trace via a FixedSource, stable argsort), {r['processes']} processes over "
                  f"contiguous ray ranges, wall clock incl. per-process scene construction and the pandas frame")
        extra = {"port": {"value": port["rays_per_s"], "unit": UNIT, "cores": cores, "kind": "port",
                          "sample": f"first {port_n} rays, multithreaded C restatement (oracle/), full frame written",
                          "steps": port["steps"], "warmup": port["warmup"]}}
        used = r["processes"]
        tests_per_s = None
    else:
        r = cpu_reference_run(wl, port_n, cores, steps, warmup)
        kind = "port"
        sample = (f"first {port_n} rays of {wl.name} ({wl.description}), multithreaded C restatement of the "
                  "reference (oracle/), full frame written on the host; the NumPy reference is not staged "
                  "(baseline/_ref missing)")
        extra = {}
        used = cores
        tests_per_s = r["tests_per_s"]
        r["rays"] = port_n
    line_cfg["rays_per_step"] = r["rays"]
    line = {
        "impl": "reference", "metric": METRIC, "value": r["rays_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": line_cfg,
        "ray_surface_tests_per_s": tests_per_s,
        "cpu_baseline": dict({"value": r["rays_per_s"], "unit": UNIT, "cores": used, "kind": kind,
                              "sample": sample}, **extra),
        "e2e": {"value": r["rays_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def fp64_peak_tflops(torch, lib, device):
    """Measured FP64 FMA rate of this GPU from our own probe kernel (flops = threads*iters*16)."""
    import ctypes

    scratch = torch.zeros(8, dtype=torch.float64, device=device)
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    blocks, iters = sms * 16, 1 << 15
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.prt_fp64_probe(scratch.data_ptr(), blocks, iters, stream)
        assert rc == 0
        e1.record()
        e1.synchronize()
        best = max(best, blocks * 256 * iters * 16 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import pyrayt_b200
    from pyrayt_b200 import _lib, roofline, workloads
    from pyrayt_b200 import dist as pdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_node = pdist.bind_to_gpu_numa_node(local_rank)  # before any pinned allocation

    wl = workloads.WORKLOADS[args.workload]
    G = wl.generation_limit
    scene = wl.scene()
    engine = pyrayt_b200.Engine(scene, device=local_rank)
    engine.host_threads = max(2, (os.cpu_count() or 16) // world)  # ranks share the host's cores
    # ray-index range of this rank (ids stay global): weak = every rank brings its own --rays rays,
    # strong = the workload's rays split over the ranks
    n_total = (args.rays or wl.n_rays) * (world if args.scaling == "weak" else 1)
    first, last = pdist.shard_range(n_total, rank, world)
    n = last - first
    if isinstance(wl.source, pyrayt_b200.sources.ReferenceSourceSet):
        d_rays = wl.source.generate(n, device=local_rank, first_index=first, total=n_total)
    else:
        d_rays = wl.source.generate(n, device=local_rank, first_index=first)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ------------------------------------------------------------------ device-resident steps
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    k1_ms, step_ms = [], []
    res = None

    def device_step(timed):
        nonlocal res
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        res = engine.trace(d_rays, generation_limit=G, record="all", k1_events=(e0, e1) if timed else None)
        if world > 1:  # C1: the only data-path collective -- per-generation row counts of every rank
            pdist.exchange_counts(res.gen_counts, device=device)
        e2.record()
        e2.synchronize()
        if timed:
            k1_ms.append(e0.elapsed_time(e1))
            step_ms.append(e0.elapsed_time(e2))

    for _ in range(max(args.warmup, 3)):
        device_step(False)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_begin = ev()
    t_end = ev()
    t_begin.record()
    for _ in range(args.steps):
        device_step(True)
    t_end.record()
    barrier()
    clocks = sampler.stop()
    total_ms = max_over_ranks(t_begin.elapsed_time(t_end))
    rows = res.rows
    counters = res.counters
    launches_per_step = res.launches
    ms_per_step = total_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)
    tests_per_s = sum_over_ranks(res.ray_surface_tests) / (ms_per_step * 1e-3)
    rows_total = int(sum_over_ranks(rows))
    k1 = sum(k1_ms) / len(k1_ms)

    # ------------------------------------------------------------------ rays excluded from the id contract
    # (north star: "rays within 1e-9 of grazing or CSG seams ... are counted and reported"): one untimed
    # PRT_FLAG_DIAGNOSE trace of the first DIAGNOSE_SAMPLE_RAYS rays of this rank (5 searches per generation)
    nd = min(n, DIAGNOSE_SAMPLE_RAYS)
    dres = engine.trace(d_rays[:, :nd], generation_limit=G, record="none", diagnose=True)
    near = {"rays": nd, "grazing_rays": dres.counters["grazing_rays"], "seam_rays": dres.counters["seam_rays"],
            "tie_rays": dres.counters["tie_rays"],
            "what": "PRT_FLAG_DIAGNOSE on this rank's first rays: rays for which the nearest-hit answer of some "
                    "generation changes (hit <-> miss: grazing; other surface: seam) when the origin is displaced "
                    "by 1e-9 x max(1, |origin|) perpendicular to the direction; tie_rays: equal finite CSG keys"}

    # ------------------------------------------------------------------ roofline of the trace kernel
    hbm_peak, peak_src = measured_peaks()
    abytes = roofline.algorithmic_bytes(n, rows)
    aflops = roofline.algorithmic_flops(scene, counters)
    fp64_peak = fp64_peak_tflops(torch, engine.lib, device)
    roof = {"bound": "hbm", "achieved": abytes / (k1 * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": abytes / (k1 * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src,
            "kernel": "trace_kernel<true>", "kernel_ms": k1, "algorithmic_bytes_per_launch": abytes,
            "binding_resource": "fp64 pipe (see roofline_fp64); the HBM fraction is reported per the bench contract"}
    # Reference-equivalent work: the flops the reference's algorithm spends for the same answers (every leaf of
    # every component, every generation).  The kernel skips most of it (proven-box pruning, ray-ordered
    # traversal), so this rate can exceed the pipe's peak: it says how much reference work a second of K1
    # replaces, not how busy the pipe is -- the executed-work counters (ncu) are printed beside it.
    roof64 = {"bound": "fp64", "reference_equivalent_tflops": aflops / (k1 * 1e-3) / 1e12, "peak": fp64_peak,
              "unit": "TFLOP/s", "reference_equivalent_over_peak": aflops / (k1 * 1e-3) / 1e12 / fp64_peak,
              "peak_source": "measured: prt_fp64_probe (DFMA, 2 flops each) on this GPU",
              "reference_equivalent_flops_per_launch": aflops,
              "note": "reference-equivalent flops (+,-,*,/,sqrt,compare = 1 each, SURVEY 8(d)) of the tests the "
                      "kernel ANSWERED; most are answered by pruning without being executed, so the ratio to the "
                      "pipe's peak is not a utilisation and may exceed 1"}
    traffic_file = os.path.join(ROOT, "profiles", "trace_kernel_traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as fh:
            t = json.load(fh)
        if t.get("workload") == wl.name and t.get("rays") == n:
            # STORED values: one `ncu --set full` capture of this kernel on this workload made by the builder
            # (profiles/), not counters of this run -- a bench number is never taken under a profiler
            roof["traffic"] = t.get("dram_bytes_per_launch")
            roof["traffic_source"] = "stored: " + str(t.get("source"))
            # what the counters say the pipes actually did (pruning answers tests without executing them,
            # so the algorithmic fraction above can exceed it)
            roof64["executed_work_stored_ncu"] = {
                "fp64_pipe_active_pct": t.get("fp64_pipe_active_pct"), "issue_active_pct": t.get("issue_active_pct"),
                "kernel_ms_under_ncu": t.get("kernel_ms"), "source": "stored: " + str(t.get("source"))}

    # ------------------------------------------------------------------ read-out on the device frame (N2)
    # outside the timed steps: what a user reads off the frame (per-field spot on the detector) without
    # copying the frame to the host; all-reduced over the ranks between its two passes
    from pyrayt_b200 import analytics

    det = float(scene.leaf_sid[-1])
    per_group = (n_total + 8) // 9
    analytics.spot_stats(res, per_group, 9, surface=det)
    r0, r1 = ev(), ev()
    r0.record()
    spot = analytics.spot_stats(res, per_group, 9, surface=det)
    r1.record()
    r1.synchronize()
    readout = {"what": "analytics.spot_stats: 9 ray-index groups, rows ending on the detector, two passes "
                       "(prt_spot_moments x2 + prt_spot_centers) incl. the small D2H of the table",
               "ms": max_over_ranks(r0.elapsed_time(r1)), "rows_scanned_per_gpu": rows,
               "detector_rows": int(spot["n"].sum()), "rms_radius_group0": float(spot["rms_radius"][0]),
               "frame_bytes_left_on_device_per_gpu": rows * 120}

    # ------------------------------------------------------------------ end to end through host buffers
    e2e = None
    res = None  # drop the device frame of the last resident step before the host-buffer run
    torch.cuda.empty_cache()
    # ------------------------------------------------------------------ the optional FP32 fast mode (north star)
    fp32 = None
    if not args.no_fp32:
        try:
            fp32 = run_fp32_mode(torch, engine, d_rays, n, n_total, G, args.steps, ev, barrier, max_over_ranks)
        except _lib.PrtError as exc:  # the library is the judge of what the mode supports; never fail the bench for it
            fp32 = {"unsupported": str(exc)}
        torch.cuda.empty_cache()
    if not args.no_e2e:
        e2e = run_e2e(args, torch, engine, d_rays, n, n_total, G, rows, world, barrier, max_over_ranks)

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1)
    cpu, mismatch = None, None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        ns = port_sample_size(wl, n)
        r = cpu_reference_run(wl, ns, threads, 2, 1)
        port = {"value": r["rays_per_s"], "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"first {ns} rays of {wl.name}, oracle port (multithreaded C restatement), full frame written",
                "ray_surface_tests_per_s": r["tests_per_s"]}
        if numpy_reference_available(args):
            # the metric's own baseline: the unmodified NumPy reference as a user runs it (one process;
            # NumPy's elementwise kernels are single-threaded), same seeded rays
            q = numpy_reference_run(wl, min(NUMPY_SAMPLE_RAYS_1CORE, n), 1, 1, 0)
            cpu = {"value": q["rays_per_s"], "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"first {q['rays']} rays of {wl.name} through the UNMODIFIED NumPy reference "
                             "(pyrayt.RayTracer.trace, FixedSource with the same seeded rays, stable argsort), "
                             "one process, wall clock of one trace() incl. the pandas frame",
                   "rows": q["rows"], "port": port}
            mismatch = argsort_mismatch_report(wl, n)
        else:
            cpu = port

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "description": wl.description, "rays_per_gpu": n, "rays_total": n_total,
                       "generation_limit": G, "leaves": scene.n_leaves, "rows_per_ray": rows / max(n, 1),
                       "parallelism": f"ray-range x{world}", "numa_node_rank0": numa_node, "l2": "inputs larger than L2 (rays + staging >> 126 MB)"},
            "ray_surface_tests_per_s": tests_per_s, "segments_per_s": rows_total / (ms_per_step * 1e-3),
            "roofline": roof, "roofline_fp64": roof64, "cpu_baseline": cpu, "e2e": e2e, "readout": readout,
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
            "counters": {k: counters[k] for k in ("rays", "generations", "segments", "tie_rays", "rows_dropped")},
            "argsort_mismatch": mismatch, "near_degenerate": near, "fp32_mode": fp32,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_fp32_mode(torch, engine, d_rays, n, n_total, G, steps, ev, barrier, max_over_ranks):
    """PRT_FLAG_FP32: the same step in single precision (device-resident, like `value`), and its agreement with
    the FP64 frame on this rank's first FP32_COMPARE_RAYS rays."""
    from pyrayt_b200 import compare

    k1, res = [], None
    for it in range(3 + steps):
        res = None  # one frame at a time
        e0, e1 = ev(), ev()
        e0.record()
        res = engine.trace(d_rays, generation_limit=G, record="all", precision="fp32", k1_events=(e0, e1))
        torch.cuda.synchronize()
        if it >= 3:
            k1.append(e0.elapsed_time(e1))
    rows32 = res.rows
    res = None
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(steps):
        res = None
        res = engine.trace(d_rays, generation_limit=G, record="all", precision="fp32")
    t1.record()
    barrier()
    ms = max_over_ranks(t0.elapsed_time(t1)) / steps
    res = None
    nc = min(n, FP32_COMPARE_RAYS)
    sub = d_rays[:, :nc]
    a = engine.trace(sub, generation_limit=G)
    b = engine.trace(sub, generation_limit=G, precision="fp32")
    rep = compare.frame_agreement(a.frame, b.frame, int(sub[12, 0].item()), nc)
    return {"value": n_total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "kernel_ms": sum(k1) / len(k1),
            "dtype": "f32", "rows": rows32,
            "contract": "positions within 1e-5 of the scene scale, unit tilt and index within 1e-5, ids equal, "
                        "except rays passing within that distance of an edge (counted below)",
            "agreement_with_fp64": rep}


def run_e2e(args, torch, engine, d_rays, n, n_total, G, rows, world, barrier, max_over_ranks):
    """Host buffers in, host frame out: H2D of the rays and D2H of the frame inside the timed region."""
    import psutil

    h_rays = torch.empty(d_rays.shape, dtype=torch.float64, pin_memory=True)
    h_rays.copy_(d_rays)
    frame_bytes = rows * 15 * 8
    avail = psutil.virtual_memory().available
    if frame_bytes * world * 1.25 > avail:
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": int(h_rays.numel() * 8),
                "d2h_bytes_per_step": int(frame_bytes), "skipped": "host RAM too small for the pinned frame(s)"}
    h_frame = torch.empty((15, rows), dtype=torch.float64, pin_memory=True)
    dev_in = torch.empty_like(d_rays)
    steps = max(1, min(args.steps, 3))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step():
        dev_in.copy_(h_rays, non_blocking=True)
        r = engine.trace(dev_in, generation_limit=G, record="all", to_host=True, host_frame=h_frame,
                         zero_copy=args.zero_copy, host_rays=h_rays, lean=False if args.full_copy else "auto")
        assert r.rows == rows
        return r

    step()
    step()  # (the engine times one transfer of each kind, lean and full, and keeps the faster)
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    barrier()
    ms = max_over_ranks(t0.elapsed_time(t1)) / steps
    chk = float(h_frame[5, :1024].sum())  # touch the host result
    # the host-side ceiling of this path on this box: plain pinned copies, every rank at the same time
    from pyrayt_b200 import dist as pdist

    peak = pdist.measure_host_copy_peak(d_rays.device, 2 << 30, world)
    lean = engine.last_transfer == "lean"  # what the timed steps used (chosen by the engine's own timing)
    out = {"value": n_total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "h2d_bytes_per_step": int(h_rays.numel() * 8),
           "d2h_bytes_per_step": int(rows * 11 * 8 + 8) if lean else int(frame_bytes),
           "host_frame_bytes_per_step": int(frame_bytes),
           "transfer": ("lean: 10 columns + one packed word per row cross the bus, generation / intensity / "
                        "wavelength / id / surface are rebuilt on the host from the rays (after the device "
                        "verified every row)") if lean else
                       ("all 15 columns copied" + ("" if args.full_copy or args.zero_copy else
                                                   " (faster on this host than rebuilding 5 columns from the rays: "
                                                   "the engine timed both)")),
           "transfer_ms_per_step_measured": {k: v * rows for k, v in engine._xfer_ms_per_row.items()},
           "host_checksum": chk}
    moved = out["d2h_bytes_per_step"] + out["h2d_bytes_per_step"]  # per rank; copies are serial within a rank
    floor_ms = (out["d2h_bytes_per_step"] / peak["d2h_gbs_per_rank"] + out["h2d_bytes_per_step"] / peak["h2d_gbs_per_rank"]) / 1e6
    out["host_peak"] = dict(peak, what="measured in this run: pinned D2H / H2D copy rate with all ranks copying at once")
    out["host_floor_ms_per_step"] = floor_ms  # the bytes this path moves, at the measured copy rates, nothing else
    out["frac_of_host_peak"] = floor_ms / ms
    out["bus_bytes_per_step"] = int(moved)
    del h_frame

    # the same call chain when the user reads per-field spot statistics instead of the whole frame:
    # host rays in, trace, read-out kernels on the device frame, a 9 x 17 table out
    from pyrayt_b200 import analytics

    det = float(engine.scene.leaf_sid[-1])
    per_group = (n_total + 8) // 9

    def step_readout():
        dev_in.copy_(h_rays, non_blocking=True)
        r = engine.trace(dev_in, generation_limit=G, record="all")
        return analytics.spot_stats(r, per_group, 9, surface=det)

    step_readout()
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(steps):
        table = step_readout()
    t1.record()
    barrier()
    ms2 = max_over_ranks(t0.elapsed_time(t1)) / steps
    out["with_device_readout"] = {
        "what": "host rays in -> trace -> analytics.spot_stats on the device frame -> 9 x 17 table out "
                "(the frame is not copied)",
        "value": n_total / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2,
        "d2h_bytes_per_step": int(table.shape[0] * table.shape[1] * 8)}
    return out


if __name__ == "__main__":
    main()
