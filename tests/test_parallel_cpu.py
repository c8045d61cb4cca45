"""Multi-rank host logic on CPU: world_size-2 gloo run of the root-parallel merge (the one
collective of the path) and the sharding arithmetic."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from posggym_baselines_b200.parallel import shard_range

    for n in (0, 1, 7, 65536, 65537):
        for ws in (1, 2, 3, 8):
            blocks = [shard_range(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(65536, 3, 8) == (3 * 8192, 4 * 8192)


def _worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    import torch.distributed as dist

    from posggym_baselines_b200 import parallel

    dist.init_process_group("gloo", rank=rank, world_size=ws)
    assert parallel.world() == (rank, ws)
    # rank-specific root statistics of a 5-action node
    visits = np.array([10, 20, 30, 0, 5], np.float64) * (rank + 1)
    total = np.array([1.0, -2.0, 6.0, 0.0, 0.5]) * (rank + 1) + rank
    v, t = parallel.merge_root_statistics(visits, total)
    a_ucb = parallel.select_root_action(v, t, range(5), "ucb", tie_seed=123)
    a_pucb = parallel.select_root_action(v, t, range(5), "pucb", tie_seed=123)
    # exact tie: every rank must break it the same way
    a_tie = parallel.select_root_action(np.ones(5), np.ones(5), range(5), "ucb", tie_seed=77)
    q.put((rank, v.tolist(), t.tolist(), a_ucb, a_pucb, a_tie))
    dist.barrier()
    dist.destroy_process_group()


def test_root_parallel_merge_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, v0, t0, a0, b0, c0), (_, v1, t1, a1, b1, c1) = res
    assert v0 == v1 == [30.0, 60.0, 90.0, 0.0, 15.0]
    assert t0 == t1 == [1.0 * 3 + 1, -2.0 * 3 + 1, 6.0 * 3 + 1, 1.0, 0.5 * 3 + 1]
    # values: 4/30, -5/60, 19/90, 0 (unvisited holds 0.0), 2.5/15 -> action 2; visits max -> action 2
    assert a0 == a1 == 2 and b0 == b1 == 2 and c0 == c1
