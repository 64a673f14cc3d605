"""Host-side ceiling of the end-to-end path: aggregate pinned device-to-host copy bandwidth at N ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/d2h_peak.py

Every rank copies a device buffer into its own pinned host buffer (plain cudaMemcpyAsync through
torch's copy_) at the same time as all the others; rank 0 prints one JSON line with the aggregate
GB/s (max-over-ranks time) and the per-rank figures.  bench.py's `e2e.host_peak` is the same
measurement taken inside the bench run itself; this script exists to sweep N quickly.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pyrayt_b200 import dist as pdist  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    gib = float(os.environ.get("D2H_GIB", "4"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    numa = pdist.bind_to_gpu_numa_node(local)
    res = pdist.measure_host_copy_peak(torch.device("cuda", local), int(gib * (1 << 30)), world)
    if rank == 0:
        res.update({"n_gpus": world, "gib_per_rank": gib, "numa_node_rank0": numa, "host_cores": os.cpu_count()})
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
