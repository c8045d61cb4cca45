"""Builds and loads the C restatement of the Driving-v1 oracle (oracle/c/driving_oracle.c).
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "driving_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libdriving_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        gcc = shutil.which("gcc") or "gcc"
        # -ffp-contract=off: no fused multiply-add, rewards / returns must round like CPython's
        subprocess.run([gcc, "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        h = C.CDLL(build())
        h.drv_init.argtypes = [C.c_char_p, C.c_int]
        h.drv_init.restype = C.c_int
        h.drv_step.argtypes = [C.c_int64] + [C.c_void_p] * 7
        h.drv_step.restype = None
        h.drv_rollouts.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_double,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
        h.drv_rollouts.restype = C.c_int64
        h.drv_sweep.argtypes = [C.c_uint64, C.c_int, C.c_int64]
        h.drv_sweep.restype = C.c_int64
        _lib = h
    return _lib


def init_driving(grid: str = "14x14RoundAbout", obstacle_collisions: bool = False) -> C.CDLL:
    from oracle.driving import SUPPORTED_GRIDS

    h = lib()
    h.drv_init(SUPPORTED_GRIDS[grid][0].encode(), int(obstacle_collisions))
    return h
