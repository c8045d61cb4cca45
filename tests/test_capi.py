"""The C-ABI library: builds, loads, and exports every symbol include/posg_b200.h
declares; the product never imports the oracle; no compute without a GPU.  CPU only."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from posggym_baselines_b200 import build

    return build.build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "posg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(posg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    syms = header_symbols()
    assert len(syms) >= 12
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (posg_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, missing


def test_ctypes_signatures_cover_header(built_lib):
    from posggym_baselines_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()
    handle = _lib.lib()
    assert handle.posg_version() >= 100


def test_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out


def test_tma_bulk_copy_in_sass(built_lib):
    """The model blob is staged by a TMA bulk copy (UBLKCP) + mbarrier wait (SYNCS)."""
    out = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UBLKCP" in out and "SYNCS" in out


def test_policy_kernel_runs_on_tcgen05(built_lib):
    """The collector's policy step is a tcgen05 kernel: its SASS has the fifth-generation tensor-core MMA
    (UTCHMMA), accumulator loads from tensor memory (LDTM) and the MMA-completion barrier arrive (UTCBAR) -- and
    no warp-level HMMA fallback."""
    out = subprocess.run(["cuobjdump", "-sass", "-fun", "policy_step_kernel", built_lib], capture_output=True, text=True).stdout
    if "UTCHMMA" not in out:   # (-fun wants the mangled name on some toolkits: fall back to the whole library)
        out = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out and "LDTM" in out and "UTCBAR" in out
    kernel = out[out.find("policy_step_kernel"):] if "policy_step_kernel" in out else out
    assert "HMMA.16816" not in kernel.split("Function :")[0]


def test_no_gpu_fails_loudly(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from posggym_baselines_b200 import PosgB200Error, _lib
    from posggym_baselines_b200.envs.driving import BatchedDriving

    assert _lib.lib().posg_device_count() == -4
    with pytest.raises(PosgB200Error):
        BatchedDriving()


def test_product_never_touches_oracle():
    bad = []
    pkg = os.path.join(ROOT, "posggym_baselines_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                # docstrings cite /root/reference paths; what must never happen is importing the
                # oracle or putting the reference tree on the import path
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or re.search(
                        r"sys\.path[^\n]*reference|import_module\([^\n]*oracle", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
