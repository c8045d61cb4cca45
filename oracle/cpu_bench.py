"""CPU baseline workloads (oracle; test infrastructure only): the reference's own
scale-out pattern is one single-threaded Python process per core
(baseline_exps/run_planning_exps.py:378), so that is what is timed here."""
from __future__ import annotations

import os
import random
import time
from multiprocessing import get_context


def _driving_sweep_worker(args):
    """Random-action sweep of Driving-v1: `episodes` episodes of up to `depth` steps,
    fresh initial state per episode.  Returns (executed env-steps, padded env-steps, seconds)."""
    seed, budget_s, depth = args
    from oracle import shims

    shims.install()
    from oracle.driving import DrivingModel

    m = DrivingModel()
    m.seed(seed)
    rng = random.Random(seed ^ 0x5EED)
    agents = m.possible_agents
    live = padded = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        s = m.sample_initial_state()
        for _ in range(depth):
            ts = m.step(s, {i: rng.randrange(5) for i in agents})
            s = ts.state
            live += 1
            if ts.all_done:
                break
        padded += depth
    return live, padded, time.perf_counter() - t0


def driving_sweep(budget_s: float = 10.0, depth: int = 50, procs: int | None = None, seed: int = 0):
    """Returns dict(env_steps, seconds, env_steps_per_s, cores)."""
    procs = procs or os.cpu_count() or 1
    if procs == 1:
        res = [_driving_sweep_worker((seed, budget_s, depth))]
    else:
        ctx = get_context("fork")
        with ctx.Pool(procs) as pool:
            res = pool.map(_driving_sweep_worker, [(seed + k, budget_s, depth) for k in range(procs)])
    live = sum(r[0] for r in res)
    wall = max(r[2] for r in res)
    return {"env_steps": live, "seconds": wall, "env_steps_per_s": live / wall, "cores": procs}


def driving_sweep_compiled(budget_s: float = 5.0, depth: int = 50, threads: int | None = None, seed: int = 0):
    """The same sweep through the C restatement (oracle/c/driving_oracle.c), one thread per
    core (ctypes releases the GIL).  A compiled-code CPU baseline beside the pure-Python one."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import native

    h = native.init_driving()
    threads = threads or os.cpu_count() or 1
    # calibrate: steps per second per thread
    t0 = time.perf_counter()
    h.drv_sweep(seed, depth, 200_000)
    rate = 200_000 / (time.perf_counter() - t0)
    budget = int(rate * budget_s)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        done = sum(ex.map(lambda k: h.drv_sweep(seed + 1 + k, depth, budget), range(threads)))
    wall = time.perf_counter() - t0
    return {"env_steps": int(done), "seconds": wall, "env_steps_per_s": done / wall, "cores": threads}
