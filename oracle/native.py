"""Builds and loads the C restatements of the oracle models (oracle/c/*_oracle.c).
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(HERE, "c")
OUT_DIR = os.path.join(HERE, "_build")
NAMES = ("driving", "pp", "lbf")

_libs = {}


def _paths(name: str):
    return os.path.join(SRC_DIR, f"{name}_oracle.c"), os.path.join(OUT_DIR, f"lib{name}_oracle.so")


def build(force: bool = False) -> str:
    """Compiles every C oracle; returns the Driving library's path."""
    os.makedirs(OUT_DIR, exist_ok=True)
    hdr = os.path.join(SRC_DIR, "oracle_common.h")
    for name in NAMES:
        src, lib_path = _paths(name)
        newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
        if force or not os.path.exists(lib_path) or os.path.getmtime(lib_path) < newest:
            gcc = shutil.which("gcc") or "gcc"
            # -ffp-contract=off: no fused multiply-add, rewards / returns must round like CPython's
            subprocess.run([gcc, "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", lib_path, src, "-lm"],
                           check=True)
    return _paths("driving")[1]


_TRACE_ARGS = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int] + [C.c_void_p] * 6
_ROLLOUT_ARGS = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]


def _load(name: str) -> C.CDLL:
    if name not in _libs:
        build()
        _libs[name] = C.CDLL(_paths(name)[1])
    return _libs[name]


def lib() -> C.CDLL:
    h = _load("driving")
    if not getattr(h, "_typed", False):
        h.drv_init.argtypes = [C.c_char_p, C.c_int]
        h.drv_init.restype = C.c_int
        h.drv_step.argtypes = [C.c_int64] + [C.c_void_p] * 7
        h.drv_step.restype = None
        h.drv_rollouts.argtypes = _ROLLOUT_ARGS
        h.drv_rollouts.restype = C.c_int64
        h.drv_sweep.argtypes = [C.c_uint64, C.c_int, C.c_int64]
        h.drv_sweep.restype = C.c_int64
        h.drv_trace.argtypes = _TRACE_ARGS
        h.drv_trace.restype = C.c_int64
        h._typed = True
    return h


def init_driving(grid: str = "14x14RoundAbout", obstacle_collisions: bool = False) -> C.CDLL:
    from oracle.driving import SUPPORTED_GRIDS

    h = lib()
    h.drv_init(SUPPORTED_GRIDS[grid][0].encode(), int(obstacle_collisions))
    return h


def init_pp(size: int = 10, cooperative: bool = True, prey_strength: int = 2, obs_dim: int = 2) -> C.CDLL:
    h = _load("pp")
    h.pp_init.argtypes = [C.c_int] * 4
    h.pp_init.restype = C.c_int
    h.pp_step.argtypes = [C.c_int64] + [C.c_void_p] * 7
    h.pp_step.restype = None
    h.pp_trace.argtypes = _TRACE_ARGS
    h.pp_trace.restype = C.c_int64
    h.pp_rollouts.argtypes = _ROLLOUT_ARGS
    h.pp_rollouts.restype = C.c_int64
    h.pp_init(size, int(cooperative), prey_strength, obs_dim)
    return h


def init_lbf(size: int = 10, sight: int = 2, max_agent_level: int = 3, force_coop: bool = False,
             normalize_reward: bool = True) -> C.CDLL:
    h = _load("lbf")
    h.lbf_init.argtypes = [C.c_int] * 5
    h.lbf_init.restype = C.c_int
    h.lbf_step.argtypes = [C.c_int64] + [C.c_void_p] * 6
    h.lbf_step.restype = None
    h.lbf_trace.argtypes = _TRACE_ARGS
    h.lbf_trace.restype = C.c_int64
    h.lbf_rollouts.argtypes = _ROLLOUT_ARGS
    h.lbf_rollouts.restype = C.c_int64
    h.lbf_init(size, sight, max_agent_level, int(force_coop), int(normalize_reward))
    return h


#: (trace function name, state words, observation words per env, agents) per model kind
TRACE_SHAPES = {"driving": ("drv_trace", 2, 2, 2), "pp": ("pp_trace", 2, 2, 2), "lbf": ("lbf_trace", 2, 4, 2)}


def trace(h: C.CDLL, kind: str, seed: int, p0: int, n: int, depth: int, full: bool = False, threads: int = 0):
    """The random-action ``model.step`` sweep of BASELINE.json configs[1] on the C oracle for particles
    [p0, p0 + n): returns ``(hash [n] u64, live env-steps, trace)`` where ``trace`` is None or -- with
    ``full`` -- a dict of per-step packed arrays ``state [depth, n, 2] u32``, ``obs [depth, n, OW] u64``,
    ``rew [depth, n, 2] f64``, ``flags [depth, n] u8``, ``act [depth, n, 2] u8``.  Multi-threaded over
    particle ranges (ctypes releases the GIL)."""
    fn_name, sw, ow, nag = TRACE_SHAPES[kind]
    fn = getattr(h, fn_name)
    hashes = np.zeros(n, np.uint64)
    threads = threads or os.cpu_count() or 1
    if not full:
        step = max(1, -(-n // (threads * 4)))
        chunks = [(a, min(n, a + step)) for a in range(0, n, step)]

        def run(c):
            a, b = c
            return fn(seed, p0 + a, b - a, depth, hashes[a:].ctypes.data, None, None, None, None, None)

        with ThreadPoolExecutor(threads) as ex:
            live = sum(ex.map(run, chunks))
        return hashes, int(live), None
    # full trace: [depth][n][..] rows are strided by n, so one call covers the whole range
    tr = dict(state=np.zeros((depth, n, sw), np.uint32), obs=np.zeros((depth, n, ow), np.uint64),
              rew=np.zeros((depth, n, nag), np.float64), flags=np.zeros((depth, n), np.uint8),
              act=np.zeros((depth, n, nag), np.uint8))
    live = fn(seed, p0, n, depth, hashes.ctypes.data, tr["state"].ctypes.data, tr["obs"].ctypes.data,
              tr["rew"].ctypes.data, tr["flags"].ctypes.data, tr["act"].ctypes.data)
    return hashes, int(live), tr
