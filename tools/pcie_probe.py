"""Dev probe: pinned-memory PCIe bandwidth (H2D, D2H, both at once) -- the floor of bench.py's e2e.  Run under
torchrun to measure N ranks at once (every rank its own GPU, all copying in the same interval):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py [--bind]
Rank 0 prints per-rank and aggregate figures as one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bind = None
if "--bind" in sys.argv:
    from posggym_baselines_b200.hostbind import bind_to_gpu_node

    bind = bind_to_gpu_node(local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = 1 << 27  # 512 MiB of int32
h_a = torch.empty(n, dtype=torch.int32).pin_memory()
h_b = torch.empty(n, dtype=torch.int32).pin_memory()
h_a.zero_(); h_b.zero_()
d_a = torch.empty(n, dtype=torch.int32, device="cuda")
d_b = torch.empty(n, dtype=torch.int32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def t(fn, reps=4):
    fn(); barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    barrier()
    return (time.perf_counter() - t0) / reps


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


gb = n * 4 / 1e9
res = {"h2d": gb / t(lambda: d_a.copy_(h_a, non_blocking=True)), "d2h": gb / t(lambda: h_b.copy_(d_b, non_blocking=True)),
       "duplex_each_way": gb / t(both)}
if rank == 0:
    # the barriers make every figure "all ranks at once": aggregate = world x the (slowest-rank) per-rank rate
    print(json.dumps({"ranks": world, "bind": bind, "per_rank_gbs": res,
                      "aggregate_gbs": {k: v * world for k, v in res.items()}, "cores": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
