"""Multi-GPU path on real GPUs: one process per GPU over NCCL (skipped with fewer than 2 GPUs)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out):
    import torch
    import torch.distributed as dist

    import pyrayt_b200
    from pyrayt_b200 import dist as pdist
    from pyrayt_b200 import workloads

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        wl = workloads.WORKLOADS["config4"]
        scene = wl.scene()
        eng = pyrayt_b200.Engine(scene, device=rank)
        b, e = pdist.shard_range(n_total, rank, world)
        rays = wl.source.generate(e - b, device=rank, first_index=b)      # ids stay global
        res = eng.trace(rays, generation_limit=wl.generation_limit)
        counts = pdist.exchange_counts(res.gen_counts, device=torch.device("cuda", rank))            # C1
        det = int(scene.leaf_sid[-1])
        det_rows = pdist.gather_rows(res.frame[:, res.frame[5] == float(det)], device=torch.device("cuda", rank))  # C2
        summary = pdist.detector_summary(res.frame, det, device=torch.device("cuda", rank))
        frames = pdist.gather_rows(res.frame, device=torch.device("cuda", rank))
        if rank == 0:
            whole = eng.trace(wl.source.generate(n_total, device=0), generation_limit=wl.generation_limit)
            w = whole.frame.cpu().numpy()
            glob = pdist.assemble_global_frame([f.cpu().numpy() for f in frames], counts)
            dr = np.hstack([d.cpu().numpy() for d in det_rows])
            order = np.lexsort((dr[4], dr[0]))
            out.put((bool(np.array_equal(glob, w)), bool(np.array_equal(dr[:, order], w[:, w[5] == det])),
                     summary["count"] == int((w[5] == det).sum()), counts.shape))
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_trace_matches_single_gpu(cuda_device):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1 << 18, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0], "frames assembled from 2 GPUs differ from the single-GPU frame"
    assert res[1] and res[2] and res[3][0] == 2
