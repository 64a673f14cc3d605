"""Multi-GPU host logic on CPU: world_size 2 over gloo (the N > 1 path without GPUs)."""
import os
import socket

import numpy as np
import pytest

from oracle import oracle
from pyrayt_b200 import dist as pdist
from tests.helpers import load_case


def test_shard_ranges_tile_the_ray_index_space():
    for n, w in ((10, 3), (1 << 24, 8), (7, 8), (0, 2)):
        spans = [pdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_global_offsets_interleave_ranks_inside_each_generation():
    counts = np.array([[3, 2, 0], [1, 1, 1]])
    assert np.array_equal(pdist.global_row_offsets(counts), [[0, 4, 7], [3, 6, 7]])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene, rays, _, gl = load_case("thick_lens_zoo")
        b, e = pdist.shard_range(rays.shape[1], rank, world)
        # each rank traces its own ray-index range (the oracle stands in for the GPU kernel here)
        frame, _ = oracle.trace(scene, np.ascontiguousarray(rays[:, b:e]), gl)
        gen_counts = np.bincount(frame[0].astype(np.int64), minlength=gl)[:gl]
        all_counts = pdist.exchange_counts(gen_counts)                       # C1
        det = int(scene.leaf_sid[-1])
        parts = pdist.gather_rows(torch.from_numpy(np.ascontiguousarray(frame[:, frame[5] == det])))  # C2
        # per-rank moment table as prt_spot_moments lays it out (count, sums, min / max of y1, z1)
        hit = frame[:, frame[5] == det]
        sums = torch.zeros((1, 16), dtype=torch.float64)
        sums[0, 0], sums[0, 1], sums[0, 2] = hit.shape[1], hit[10].sum(), hit[11].sum()
        sums[0, 6], sums[0, 7] = hit[10].min(initial=np.inf), hit[10].max(initial=-np.inf)
        sums[0, 8], sums[0, 9] = hit[11].min(initial=np.inf), hit[11].max(initial=-np.inf)
        pdist.reduce_spot_sums(sums)
        frames = pdist.gather_rows(torch.from_numpy(frame))
        if rank == 0:
            whole, _ = oracle.trace(scene, rays, gl)
            glob = pdist.assemble_global_frame([f.numpy() for f in frames], all_counts)
            det_rows = np.hstack([p.numpy() for p in parts])
            want_det = whole[:, whole[5] == det]
            order = np.lexsort((det_rows[4], det_rows[0]))
            out.put((bool(np.array_equal(glob, whole)), bool(np.array_equal(det_rows[:, order], want_det)),
                     int(sums[0, 0]) == want_det.shape[1] and float(sums[0, 6]) == want_det[10].min()
                     and float(sums[0, 9]) == want_det[11].max()
                     and abs(float(sums[0, 1]) - want_det[10].sum()) <= 1e-9 * max(1.0, np.abs(want_det[10]).sum()),
                     all_counts.shape))
    finally:
        dist.destroy_process_group()


def test_two_rank_trace_reassembles_the_global_frame():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0], "frames assembled from 2 ranks differ from the monolithic frame"
    assert res[1], "gathered detector rows differ"
    assert res[2]
    assert res[3][0] == 2
