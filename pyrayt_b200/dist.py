"""Multi-GPU plumbing: ray-index-range sharding, one process per GPU (torch.distributed).

Rays never interact, so the trace shards by contiguous ray-index
range with every rank holding the whole (tiny) scene.  The data path needs exactly
one exchange: the per-generation row counts of every rank (C1), from which each rank
knows where its rows sit in the global (generation, id)-ordered frame.  Detector rows
(C2) can be gathered on request.  Works with NCCL on GPUs and with gloo on CPU tensors
(the host logic is tested with gloo, world_size 2).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ray-index range [begin, end) of `rank`; ranges tile [0, n_total) in rank order."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def exchange_counts(gen_counts: np.ndarray, device=None) -> np.ndarray:
    """C1: all-gather the per-generation row counts. Returns (world, G) int64."""
    import torch
    import torch.distributed as dist

    local = torch.as_tensor(np.ascontiguousarray(gen_counts, dtype=np.int64))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.numpy()[None, :]
    if device is not None:
        local = local.to(device)
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local)
    return torch.stack(out).cpu().numpy()


def global_row_offsets(all_counts: np.ndarray) -> np.ndarray:
    """First row, in the global (generation, id)-ordered frame, of every (rank, generation) block.

    Global order is generation-major; within a generation ranks follow in rank order because
    ray-index ranges are assigned in rank order (ids ascending).  Returns (world, G) int64.
    """
    all_counts = np.asarray(all_counts, dtype=np.int64)
    per_gen = all_counts.sum(axis=0)
    gen_start = np.concatenate(([0], np.cumsum(per_gen)[:-1]))
    within = np.cumsum(all_counts, axis=0) - all_counts
    return gen_start[None, :] + within


def assemble_global_frame(frames: List[np.ndarray], all_counts: np.ndarray) -> np.ndarray:
    """Place every rank's (15, rows_r) frame into the global frame (host-side reference of the layout)."""
    offs = global_row_offsets(all_counts)
    total = int(np.asarray(all_counts).sum())
    out = np.empty((15, total))
    for r, f in enumerate(frames):
        local_start = np.concatenate(([0], np.cumsum(all_counts[r])[:-1]))
        for g in range(all_counts.shape[1]):
            c = int(all_counts[r, g])
            if c:
                out[:, offs[r, g]: offs[r, g] + c] = f[:, local_start[g]: local_start[g] + c]
    return out


def gather_rows(local_rows, device=None):
    """C2: all-gather a variable number of (15, k_r) rows (e.g. the detector rows) to every rank.

    Returns the list of per-rank tensors in rank order."""
    import torch
    import torch.distributed as dist

    t = local_rows if isinstance(local_rows, torch.Tensor) else torch.as_tensor(local_rows)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [t]
    if device is not None:
        t = t.to(device)
    world = dist.get_world_size()
    k = torch.tensor([t.shape[1]], dtype=torch.int64, device=t.device)
    ks = [torch.empty_like(k) for _ in range(world)]
    dist.all_gather(ks, k)
    kmax = int(max(int(x.item()) for x in ks))
    pad = torch.zeros((t.shape[0], kmax), dtype=t.dtype, device=t.device)
    pad[:, : t.shape[1]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return [o[:, : int(kk.item())] for o, kk in zip(outs, ks)]


def reduce_spot_sums(sums) -> None:
    """Combine per-rank moment tables (n_groups, 16) of prt_spot_moments in place: the sums add,
    columns 6..9 (min / max of y1, z1) take the min / max.  No-op without a process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    lo = sums[:, [6, 8]].contiguous()
    hi = sums[:, [7, 9]].contiguous()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    sums[:, [6, 8]] = lo
    sums[:, [7, 9]] = hi


def detector_summary(frame, detector_sid: int, device=None) -> Optional[dict]:
    """Spot statistics of the rows that ended on `detector_sid`, over the ray sets of every rank:
    count, centroid, RMS radius (the moment kernels of analytics.py, all-reduced between their passes)."""
    from . import analytics

    st = analytics.spot_stats(frame, rays_per_group=1 << 62, n_groups=1, surface=float(detector_sid))
    n = int(st.loc[0, "n"])
    if n == 0:
        return {"count": 0}
    return {"count": n, "centroid": (float(st.loc[0, "y_mean"]), float(st.loc[0, "z_mean"])),
            "rms_radius": float(st.loc[0, "rms_radius"])}


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (one process per GPU).

    Pinned host buffers are then allocated on the local node (first touch), which matters when
    eight ranks each copy a 36 GB frame to the host at the same time.  Returns the node or None.
    """
    import glob
    import os

    try:
        import torch

        prop = torch.cuda.get_device_properties(device_index)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        nodes = glob.glob("/sys/devices/system/node/node[0-9]*")
        if node < 0 or len(nodes) < 2:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def measure_host_copy_peak(device, nbytes: int = 4 << 30, world: int = 1, repeats: int = 3) -> dict:
    """The host-side ceiling of the end-to-end path: pinned device -> host copy rate with every rank
    copying at the same time (plain cudaMemcpyAsync through ``Tensor.copy_``), and the host -> device
    rate the same way.  Timed on the device (CUDA events), max over ranks.  Returns GB/s per rank and
    in aggregate; with a process group the start is aligned by a barrier."""
    import torch
    import torch.distributed as dist

    multi = world > 1 and dist.is_available() and dist.is_initialized()
    n = max(1, nbytes // 8)
    d = torch.empty(n, dtype=torch.float64, device=device)
    d.fill_(1.0)
    h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    h.fill_(0.0)  # first touch: the pages exist before anything is timed

    def timed(dst, src):
        best = float("inf")
        for _ in range(repeats):
            if multi:
                dist.barrier()
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            if multi:
                t = torch.tensor([ms], dtype=torch.float64, device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            best = min(best, ms)
        return best

    d2h_ms = timed(h, d)
    h2d_ms = timed(d, h)
    gb = n * 8 / 1e9
    return {"d2h_gbs_per_rank": gb / (d2h_ms * 1e-3), "d2h_gbs_aggregate": world * gb / (d2h_ms * 1e-3),
            "h2d_gbs_per_rank": gb / (h2d_ms * 1e-3), "h2d_gbs_aggregate": world * gb / (h2d_ms * 1e-3),
            "bytes_per_rank": n * 8, "how": "pinned cudaMemcpyAsync, all ranks at once, best of %d, max over ranks" % repeats}
